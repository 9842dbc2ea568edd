// Minimal stand-in for <opencv2/core.hpp>: just enough of cv::Mat / Size / Point2f / CV_Assert for the
// SFMGMS_WITH_OPENCV adaptor in sfm_gms_b200/cxx/sfmgms.hpp to COMPILE where OpenCV's C++ headers are not installed
// (tests/test_cxx_opencv_adaptor.py).  Field layouts follow OpenCV 4.x; nothing here is used at run time by the product.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CV_8U 0
#define CV_32F 5
#define CV_Assert(expr) do { if (!(expr)) std::abort(); } while (0)
namespace cv {
struct Point2f { float x, y; };
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
class Mat {
   public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type, void* data) : rows(r), cols(c), type_(type), data_((uint8_t*)data) {}
    int type() const { return type_; }
    bool isContinuous() const { return true; }
    template <typename T> const T* ptr(int row = 0) const { return reinterpret_cast<const T*>(data_) + (size_t)row * cols; }
   private:
    int type_ = CV_8U;
    uint8_t* data_ = nullptr;
};
}  // namespace cv
