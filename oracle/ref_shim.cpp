// TEST INFRASTRUCTURE ONLY (oracle/): C-callable wrapper around the UNMODIFIED reference CPU path.
//
// Linked against /root/reference/src/utils.cpp compiled where it lies (see oracle/Makefile), output
// oracle/_ref/libcanny_ref.so.  The reference's stage functions take reference-to-pointer arguments,
// allocate their outputs with new[] and (sobelOperator, nonmaximalSuppression) delete[] their inputs
// (utils.cpp:235,306-307); this shim gives them plain borrowed-in / caller-allocated-out semantics so
// ctypes can drive them.  It contains no arithmetic of its own.
#include <chrono>
#include <cstdint>
#include <cstring>

#include "utils.h"  // /root/reference/src/utils.h (on the include path, not copied)

namespace {
short* clone16(const int16_t* src, size_t n) {
    short* p = new short[n];
    std::memcpy(p, src, n * sizeof(short));
    return p;
}
}  // namespace

extern "C" {

// createGaussianKernel (utils.cpp:77-95).  w must hold >= 1+2*ceil(3*sigma) floats.
int ref_gaussian_kernel(float sigma, float* w, int* window) {
    float* k = nullptr;
    createGaussianKernel(k, sigma, window);
    std::memcpy(w, k, sizeof(float) * (size_t)(*window));
    delete[] k;
    return 0;
}

// gaussian (utils.cpp:26-68)
int ref_gaussian(const uint8_t* img, float sigma, int h, int w, int16_t* out) {
    unsigned char* in = const_cast<unsigned char*>(img);
    short* res = nullptr;
    gaussian(in, sigma, h, w, res);
    std::memcpy(out, res, sizeof(short) * (size_t)h * w);
    delete[] res;
    return 0;
}

// calculateXYGradient (utils.cpp:106-187)
int ref_xy_gradient(const int16_t* blur, int h, int w, int16_t* gx, int16_t* gy) {
    short* in = clone16(blur, (size_t)h * w);
    short *x = nullptr, *y = nullptr;
    calculateXYGradient(in, h, w, x, y);
    std::memcpy(gx, x, sizeof(short) * (size_t)h * w);
    std::memcpy(gy, y, sizeof(short) * (size_t)h * w);
    delete[] x;
    delete[] y;
    delete[] in;
    return 0;
}

// sobelOperator (utils.cpp:201-236); it frees its input, so it gets a private copy.
int ref_sobel(const int16_t* blur, int h, int w, int16_t* mag, int16_t* ang) {
    short* in = clone16(blur, (size_t)h * w);
    short *m = nullptr, *a = nullptr;
    sobelOperator(in, h, w, m, a);
    std::memcpy(mag, m, sizeof(short) * (size_t)h * w);
    std::memcpy(ang, a, sizeof(short) * (size_t)h * w);
    delete[] m;
    delete[] a;
    return 0;
}

// nonmaximalSuppression (utils.cpp:248-308); frees both inputs.
int ref_nonmaximal(const int16_t* mag, const int16_t* ang, int h, int w, int16_t* out) {
    short* m = clone16(mag, (size_t)h * w);
    short* a = clone16(ang, (size_t)h * w);
    short* r = nullptr;
    nonmaximalSuppression(m, a, h, w, r);
    std::memcpy(out, r, sizeof(short) * (size_t)h * w);
    delete[] r;
    return 0;
}

// hysteresis (utils.cpp:322-342), in place.
int ref_hysteresis(int16_t* nms, int h, int w, int lo, int hi) {
    short* p = nms;
    hysteresis(p, h, w, lo, hi);
    return 0;
}

// findEdgePixels (utils.cpp:360-427), in place; visited is h*w bytes of 0/1.
int ref_find_edge_pixels(int16_t* nms, uint8_t* visited, int start, int lo, int hi, int h, int w) {
    short* p = nms;
    bool* v = reinterpret_cast<bool*>(visited);
    findEdgePixels(p, v, start, lo, hi, h, w);
    return 0;
}

// The four calls canny() makes (utils.cpp:438,452,465,478) without its display code; edges is the
// 0/255 int16 map canny() would show.  Returns the seconds spent in the four calls, measured with
// the same clock canny() uses (utils.cpp:435,479).
double ref_canny(const uint8_t* img, float sigma, int lo, int hi, int h, int w, int16_t* edges) {
    unsigned char* in = const_cast<unsigned char*>(img);
    short *sm = nullptr, *m = nullptr, *a = nullptr, *n = nullptr;
    auto t0 = std::chrono::high_resolution_clock::now();
    gaussian(in, sigma, h, w, sm);
    sobelOperator(sm, h, w, m, a);
    nonmaximalSuppression(m, a, h, w, n);
    hysteresis(n, h, w, lo, hi);
    auto t1 = std::chrono::high_resolution_clock::now();
    if (edges) std::memcpy(edges, n, sizeof(short) * (size_t)h * w);
    delete[] n;
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
