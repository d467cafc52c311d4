// TEST INFRASTRUCTURE ONLY (oracle/): a do-nothing stand-in for <opencv2/opencv.hpp>.
//
// OpenCV's C++ headers are not installed in this image.  The reference's CPU path
// (/root/reference/src/utils.cpp) only touches cv:: inside canny()'s display code
// (utils.cpp:440-486); the stage functions are plain C++.  With this header on the include
// path the UNMODIFIED reference file compiles, and oracle/ref_shim.cpp calls its stage
// functions directly.  Nothing here computes anything.
#pragma once
#include <string>
namespace cv {
struct Mat {
    int rows = 0, cols = 0;
    void* data = nullptr;
    Mat() {}
    Mat(int r, int c, int /*type*/, void* d) : rows(r), cols(c), data(d) {}
    void convertTo(Mat&, int) const {}
};
enum { NORM_MINMAX = 32 };
inline void normalize(const Mat&, Mat&, double, double, int) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int) { return 0; }
}  // namespace cv
#define CV_16S 3
#define CV_8U 0
