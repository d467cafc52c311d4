// Exercises include/canny_b200_compat.hpp exactly the way the reference's callers use src/cuda.h:
// reference-to-pointer outputs allocated by the callee, freed here with delete[].
// usage: compat_main <in.u8> <height> <width> <sigma> <lo> <hi> <out_prefix>
// writes <out_prefix>.{blur,mag,ang,nms,edges,edges2}.i16
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "canny_b200_compat.hpp"

static void dump(const std::string& path, const short* p, size_t n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f || fwrite(p, sizeof(short), n, f) != n) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc != 8) { fprintf(stderr, "usage\n"); return 2; }
    const int h = atoi(argv[2]), w = atoi(argv[3]);
    const float sigma = (float)atof(argv[4]);
    const int lo = atoi(argv[5]), hi = atoi(argv[6]);
    const std::string out = argv[7];
    const size_t n = (size_t)h * w;
    std::vector<unsigned char> img(n);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(img.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);
    try {
        unsigned char* p = img.data();
        short *blur = nullptr, *mag = nullptr, *ang = nullptr, *nms = nullptr;
        cuda_gaussian(p, sigma, h, w, blur);
        cuda_sobel(blur, h, w, mag, ang);
        cuda_nonmaixmal_suppression(mag, ang, h, w, nms);
        dump(out + ".blur.i16", blur, n);
        dump(out + ".mag.i16", mag, n);
        dump(out + ".ang.i16", ang, n);
        dump(out + ".nms.i16", nms, n);
        cuda_hysteresis(nms, h, w, lo, hi);  // in place, like hysteresis() in src/utils.cpp:322
        dump(out + ".edges.i16", nms, n);
        short* e2 = cuda_canny_edges(p, sigma, lo, hi, h, w);
        dump(out + ".edges2.i16", e2, n);
        cuda_canny(p, sigma, lo, hi, h, w, false);  // the call src/main.cpp:128 makes
        delete[] blur; delete[] mag; delete[] ang; delete[] nms; delete[] e2;
    } catch (const std::exception& e) {
        fprintf(stderr, "compat_main: %s\n", e.what());
        return 1;
    }
    return 0;
}
