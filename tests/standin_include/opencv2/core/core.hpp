// TEST DOUBLE, not OpenCV: the handful of cv::Mat members the adapter overloads in rgbd_odometry_b200/host/ touch
// (rows, cols, data, type(), channels(), isContinuous(), create()).  OpenCV is not installed in this image; this stand-in only
// lets the `#ifdef DVO_HAVE_OPENCV` code be compiled and executed by tests/test_gpu_host_classes.py.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#define CV_8U 0
#define CV_16U 2
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
namespace cv {
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(int r, int c, int t, void* ext) : rows(r), cols(c), data((unsigned char*)ext), type_(t) {}      // external data, not owned
    void create(int r, int c, int t) {
        rows = r; cols = c; type_ = t;
        store_.reset(new unsigned char[(size_t)r * c * elemSize()], std::default_delete<unsigned char[]>());
        data = store_.get();
    }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    size_t elemSize() const { return (size_t)channels() * ((type_ & 7) == CV_16U ? 2 : 1); }
    bool isContinuous() const { return true; }
private:
    int type_ = 0;
    std::shared_ptr<unsigned char> store_;
};
}  // namespace cv
