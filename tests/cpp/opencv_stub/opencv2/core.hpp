// Minimal stand-in for <opencv2/core.hpp> (TEST INFRASTRUCTURE): just enough of cv::Mat / cv::Point2f for the
// DV_SHIM_WITH_OPENCV overloads of csrc/shim/deep_net_shim.h to be compiled and exercised in an image without OpenCV
// C++ headers.  Field and method names follow OpenCV 3.4 (the reference's version, README.md:22).
#pragma once
#include <stddef.h>
#include <stdint.h>
#define CV_8UC1 0
#define CV_8UC3 16
namespace cv {
struct Point2f {
  float x = 0.f, y = 0.f;
  Point2f() = default;
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct Mat {
  int rows = 0, cols = 0, type_ = CV_8UC1;
  uint8_t* data = nullptr;
  size_t step = 0;
  Mat() = default;
  Mat(int r, int c, int type, void* d, size_t s = 0)
      : rows(r), cols(c), type_(type), data(static_cast<uint8_t*>(d)), step(s ? s : (size_t)c * (type == CV_8UC3 ? 3 : 1)) {}
  int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
};
}  // namespace cv
