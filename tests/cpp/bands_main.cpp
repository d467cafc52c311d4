// A C++ caller in the position of src/main.cpp:128 that owns several GPUs (or, here, several bands on whatever GPUs there are):
// one image split into row bands through the C handle of include/canny_b200.h (b200_bands_*), no Python, no torch, no NCCL
// (an in-process group: peers are plain pointers).  Bands are spread round-robin over the visible devices.
// usage: bands_main <in.u8> <height> <width> <sigma> <lo> <hi> <n_bands> <steps> <out.u8>
// writes the assembled 0/255 map; prints the transport and the per-stage times of band 0.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "canny_b200.h"

#define CHECK(expr)                                                                             \
    do {                                                                                        \
        int st_ = (expr);                                                                       \
        if (st_ != B200_OK) { fprintf(stderr, "bands_main: %s -> %d: %s\n", #expr, st_, b200_last_error()); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    if (argc != 10) { fprintf(stderr, "usage\n"); return 2; }
    const int h = atoi(argv[2]), w = atoi(argv[3]);
    const float sigma = (float)atof(argv[4]);
    const int lo = atoi(argv[5]), hi = atoi(argv[6]), n_bands = atoi(argv[7]), steps = atoi(argv[8]);
    std::vector<unsigned char> img((size_t)h * w), out((size_t)h * w);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(img.data(), 1, img.size(), f) != img.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);

    std::vector<b200_ctx*> ctxs((size_t)n_bands, nullptr);
    int n_dev = 1;
    for (int b = 0; b < n_bands; ++b) {
        int st = b200_ctx_create(b % n_dev, &ctxs[(size_t)b]);
        if (st != B200_OK && b == 0) { fprintf(stderr, "bands_main: %s\n", b200_last_error()); return 1; }   // no device: fail loudly
        CHECK(st);
        if (b == 0) {                       // how many devices are there?  (probe by creating contexts until it fails)
            b200_ctx* probe = nullptr;
            while (n_dev < 64 && b200_ctx_create(n_dev, &probe) == B200_OK) { b200_ctx_destroy(probe); ++n_dev; }
        }
    }
    std::vector<b200_bands*> bands((size_t)n_bands, nullptr);
    CHECK(b200_bands_create_group(ctxs.data(), n_bands, h, w, sigma, lo, hi, bands.data()));
    std::vector<uint8_t*> d_edges((size_t)n_bands, nullptr);
    std::vector<int> row0((size_t)n_bands), rows((size_t)n_bands);
    int transport = -1;
    for (int b = 0; b < n_bands; ++b) {
        CHECK(b200_bands_info(bands[(size_t)b], &row0[(size_t)b], &rows[(size_t)b], nullptr, &transport));
        uint8_t* d_in = nullptr;
        CHECK(b200_bands_input(bands[(size_t)b], &d_in));
        CHECK(b200_memcpy_h2d(ctxs[(size_t)b], d_in, img.data() + (size_t)row0[(size_t)b] * w, (size_t)rows[(size_t)b] * w));
        void* p = nullptr;
        CHECK(b200_device_alloc(ctxs[(size_t)b], (size_t)rows[(size_t)b] * w, &p));
        d_edges[(size_t)b] = static_cast<uint8_t*>(p);
    }
    CHECK(b200_bands_set_timing(bands[0], 1));
    for (int s = 0; s < steps; ++s) CHECK(b200_bands_run_group(bands.data(), n_bands, d_edges.data()));
    for (int b = 0; b < n_bands; ++b) {
        CHECK(b200_bands_check(bands[(size_t)b]));
        CHECK(b200_memcpy_d2h(ctxs[(size_t)b], out.data() + (size_t)row0[(size_t)b] * w, d_edges[(size_t)b], (size_t)rows[(size_t)b] * w));
    }
    float ms[6];
    CHECK(b200_bands_stage_ms(bands[0], ms));
    printf("bands=%d devices=%d transport=%d stage_ms=%.3f,%.3f,%.3f,%.3f,%.3f,%.3f\n", n_bands, n_dev, transport, ms[0], ms[1], ms[2], ms[3], ms[4], ms[5]);
    f = fopen(argv[9], "wb");
    if (!f || fwrite(out.data(), 1, out.size(), f) != out.size()) { fprintf(stderr, "cannot write %s\n", argv[9]); return 2; }
    fclose(f);
    for (int b = 0; b < n_bands; ++b) {
        CHECK(b200_device_free(ctxs[(size_t)b], d_edges[(size_t)b]));
        CHECK(b200_bands_destroy(bands[(size_t)b]));
    }
    for (auto* c : ctxs) CHECK(b200_ctx_destroy(c));
    return 0;
}
