// canny_b200_compat.hpp — the reference's C++ GPU entry points, implemented on libcanny_b200.so.
//
// Drop-in for the reference's src/cuda.h:4-10 (same names, same reference-to-pointer parameters, same
// ownership: results are allocated with new[] and handed back through the reference, the caller
// delete[]s them — src/cuda.cu:81,225-226,370).  src/main.cpp:128 calls cuda_canny() unchanged when this
// header is what `#include "cuda.h"` resolves to and the program links against
// libcanny_b200.so instead of the reference's `cuda` library.
//
// Differences from the reference, all deliberate:
//   * semantics are those of the CPU path (src/utils.cpp) bit for bit — the reference's CUDA kernels flip
//     grad_y (src/cuda.cu:189-191) and mis-handle borders / sizes that are not multiples of 32;
//   * hysteresis runs on the GPU (the reference calls the CPU one, src/cuda.cu:436) and is also exported
//     as cuda_hysteresis();
//   * failures are reported: a non-zero status from the C ABI throws std::runtime_error carrying
//     b200_last_error() (the reference checks no CUDA status, src/cuda.cu:83-101).
//   * display stays on the host: define CANNY_B200_WITH_OPENCV before including this header to get the
//     reference's imshow()/waitKey() behaviour in cuda_canny (src/cuda.cu:400-444); without it cuda_canny
//     computes the edge map and returns (use cuda_canny_edges() to obtain it).
#ifndef CANNY_B200_COMPAT_HPP
#define CANNY_B200_COMPAT_HPP

#include <stdexcept>
#include <string>

#include "canny_b200.h"

#ifdef CANNY_B200_WITH_OPENCV
#include <opencv2/opencv.hpp>
#endif

namespace canny_b200_detail {
inline void check(int status, const char* what) {
    if (status != B200_OK) throw std::runtime_error(std::string(what) + ": " + b200_last_error());
}
#ifdef CANNY_B200_WITH_OPENCV
inline void show_plane(const char* title, short* plane, int height, int width) {
    cv::Mat as16(height, width, CV_16S, plane), as8;
    cv::normalize(as16, as8, 0, 255, cv::NORM_MINMAX, CV_8U);  // min-max stretch to 8 bit, as src/cuda.cu:404-405
    cv::imshow(title, as8);
    cv::waitKey(0);
}
#endif
}  // namespace canny_b200_detail

// src/cuda.h:4 — Gaussian blur with a kernel generated from sigma (window = 1 + 2*ceil(3*sigma)).
inline void cuda_gaussian(unsigned char*& img_h, float sigma, int height, int width, short int*& result_h) {
    result_h = new short[(size_t)height * width];
    const int st = b200_gaussian(nullptr, img_h, sigma, height, width, result_h);
    if (st != B200_OK) { delete[] result_h; result_h = nullptr; }   // no leak, no dangling output on failure
    canny_b200_detail::check(st, "cuda_gaussian");
}

// src/cuda.h:6 — Sobel magnitude and quantised angle (0/45/90/135). The input is NOT freed (as in src/cuda.cu:220-246).
inline void cuda_sobel(short int*& img_h, int height, int width, short int*& magnitude_h, short int*& angle_h) {
    magnitude_h = new short[(size_t)height * width];
    angle_h = new short[(size_t)height * width];
    const int st = b200_sobel(nullptr, img_h, height, width, magnitude_h, angle_h);
    if (st != B200_OK) { delete[] magnitude_h; delete[] angle_h; magnitude_h = nullptr; angle_h = nullptr; }
    canny_b200_detail::check(st, "cuda_sobel");
}

// src/cuda.h:8 — non-maximal suppression (the misspelling is the reference's symbol name).
inline void cuda_nonmaixmal_suppression(short int*& magnitude_h, short int*& angle_h, int height, int width,
                                        short int*& result_h) {
    result_h = new short[(size_t)height * width];
    const int st = b200_nonmaximal(nullptr, magnitude_h, angle_h, height, width, result_h);
    if (st != B200_OK) { delete[] result_h; result_h = nullptr; }
    canny_b200_detail::check(st, "cuda_nonmaixmal_suppression");
}

// GPU twin of hysteresis() (src/utils.h:18, src/utils.cpp:322-342): in place, 0 / 255 on return.
inline void cuda_hysteresis(short int*& edge_candidates, int height, int width, int min_val, int max_val) {
    canny_b200_detail::check(b200_hysteresis(nullptr, edge_candidates, height, width, min_val, max_val), "cuda_hysteresis");
}

// The whole pipeline, returning the 0/255 map the reference only displays.  Caller delete[]s the result.
inline short* cuda_canny_edges(unsigned char* img, float sigma, int min_val, int max_val, int height, int width) {
    short* edges = new short[(size_t)height * width];
    const int st = b200_canny(nullptr, img, sigma, min_val, max_val, height, width, edges);
    if (st != B200_OK) { delete[] edges; canny_b200_detail::check(st, "cuda_canny"); }
    return edges;
}

// src/cuda.h:10 — cuda_canny(img, sigma, min_val, max_val, height, width, steps).
inline void cuda_canny(unsigned char* img, float sigma, int min_val, int max_val, int height, int width, bool steps) {
#ifdef CANNY_B200_WITH_OPENCV
    const size_t n = (size_t)height * width;
    short *blur = nullptr, *mag = nullptr, *nms = nullptr;
    if (steps) { blur = new short[n]; mag = new short[n]; nms = new short[n]; }
    short* edges = new short[n];
    const int st = b200_canny_steps(nullptr, img, sigma, min_val, max_val, height, width, blur, mag, nullptr, nms, edges);
    if (st == B200_OK) {
        if (steps) {
            // window titles of the reference's GPU driver (src/cuda.cu:407,420,432,443)
            canny_b200_detail::show_plane("CudaGaussian Visual Test", blur, height, width);
            canny_b200_detail::show_plane("Sobel Visual Test", mag, height, width);
            canny_b200_detail::show_plane("Nonmaximal Visual Test", nms, height, width);
        }
        canny_b200_detail::show_plane("Final Image", edges, height, width);
    }
    delete[] blur; delete[] mag; delete[] nms; delete[] edges;
    canny_b200_detail::check(st, "cuda_canny");
#else
    (void)steps;
    delete[] cuda_canny_edges(img, sigma, min_val, max_val, height, width);
#endif
}

#endif  // CANNY_B200_COMPAT_HPP
