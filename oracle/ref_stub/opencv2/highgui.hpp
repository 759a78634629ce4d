// Stand-in for <opencv2/highgui.hpp>, used ONLY to compile the reference's own GPU sources
// (guided_filter_d.cu includes the header without using it; guided_filter.cpp touches cv::Mat
// and cv::imwrite only inside `if (false)` debug blocks, guided_filter.cpp:32-55).
// OpenCV's C++ headers are not installed in this image.  Nothing here is ever executed.
#pragma once
#include <string>
#define CV_8U 0
#define CV_32F 5
#define CV_32FC(n) (CV_32F + (((n) - 1) << 3))
namespace cv {
struct Mat {
    unsigned char* data = nullptr;
    Mat() {}
    Mat(int, int, int) {}
    void convertTo(Mat&, int, double = 1.0, double = 0.0) const {}
};
inline bool imwrite(const std::string&, const Mat&) { return false; }
}  // namespace cv
