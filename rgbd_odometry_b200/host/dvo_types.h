// dvo_types.h -- POD stand-ins for the cv::Mat / Eigen / geometry_msgs types of the reference's class API.
// Neither OpenCV nor Eigen is installed in this image; these views carry the same information (pointer, shape,
// element type) so the class signatures keep the reference's names, arity and argument meaning.
//
// When OpenCV / Eigen headers ARE present at the caller's site (they are where the reference builds: include/SolveDVO.h:17-52,
// include/EPoseEstimator.h:21-29), the same types convert implicitly from / to cv::Mat, Eigen::Matrix<T,3,3>, Eigen::Matrix<T,3,1>,
// Eigen::VectorXf / MatrixXf / ArrayXXd / MatrixXd, and the classes gain overloads with the reference's exact signatures
// (guarded by DVO_HAVE_OPENCV / DVO_HAVE_EIGEN; define DVO_NO_OPENCV / DVO_NO_EIGEN to suppress).  Eigen dense types are
// column-major, these PODs row-major: the converters go element by element.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(__has_include)
#if !defined(DVO_NO_OPENCV) && __has_include(<opencv2/core/core.hpp>)
#include <opencv2/core/core.hpp>
#define DVO_HAVE_OPENCV 1
#endif
#if !defined(DVO_NO_EIGEN) && __has_include(<Eigen/Dense>)
#include <Eigen/Dense>
#define DVO_HAVE_EIGEN 1
#endif
#endif

namespace dvo {

enum ImageType { U8C1 = 0, U8C3 = 1, U16C1 = 2 };

// a view of a tightly packed row-major image (what cv::Mat::data / rows / cols / type() expose)
struct ImageView {
    const void* data = nullptr;
    int rows = 0, cols = 0, type = U8C1;
    ImageView() {}
    ImageView(const void* d, int r, int c, int t) : data(d), rows(r), cols(c), type(t) {}
#ifdef DVO_HAVE_OPENCV
    // a continuous CV_8UC1 / CV_8UC3 / CV_16UC1 cv::Mat (what the reference passes: src/dvo.cpp, src/EPoseEstimator.cpp:68-73)
    ImageView(const cv::Mat& m) : data(m.data), rows(m.rows), cols(m.cols), type(m.type() == CV_8UC3 ? U8C3 : (m.type() == CV_16UC1 ? U16C1 : U8C1)) {
        assert(m.isContinuous() && (m.type() == CV_8UC1 || m.type() == CV_8UC3 || m.type() == CV_16UC1));
    }
#endif
    int channels() const { return type == U8C3 ? 3 : 1; }
    size_t bytes() const { return (size_t)rows * cols * (type == U8C1 ? 1 : (type == U8C3 ? 3 : 2)); }
};

// Eigen::Matrix<T,3,3> / Matrix<T,3,1> replacements, row-major storage, (r,c) access
template <typename T> struct Mat3 {
    T m[9];
    Mat3() { for (int i = 0; i < 9; ++i) m[i] = (i % 4 == 0) ? T(1) : T(0); }
#ifdef DVO_HAVE_EIGEN
    Mat3(const Eigen::Matrix<T, 3, 3>& e) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[3 * r + c] = e(r, c); }
    operator Eigen::Matrix<T, 3, 3>() const { Eigen::Matrix<T, 3, 3> e; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) e(r, c) = m[3 * r + c]; return e; }
#endif
    static Mat3 Identity() { return Mat3(); }
    T& operator()(int r, int c) { return m[3 * r + c]; }
    const T& operator()(int r, int c) const { return m[3 * r + c]; }
    Mat3 operator*(const Mat3& b) const {
        Mat3 o;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.m[3 * i + j] = m[3 * i] * b.m[j] + m[3 * i + 1] * b.m[3 + j] + m[3 * i + 2] * b.m[6 + j];
        return o;
    }
};
template <typename T> struct Vec3 {
    T v[3];
    Vec3() { v[0] = v[1] = v[2] = T(0); }
    Vec3(T a, T b, T c) { v[0] = a; v[1] = b; v[2] = c; }
#ifdef DVO_HAVE_EIGEN
    Vec3(const Eigen::Matrix<T, 3, 1>& e) { for (int i = 0; i < 3; ++i) v[i] = e(i); }
    operator Eigen::Matrix<T, 3, 1>() const { Eigen::Matrix<T, 3, 1> e; for (int i = 0; i < 3; ++i) e(i) = v[i]; return e; }
#endif
    static Vec3 Zero() { return Vec3(); }
    T& operator()(int i) { return v[i]; }
    const T& operator()(int i) const { return v[i]; }
    Vec3 operator+(const Vec3& b) const { return Vec3(v[0] + b.v[0], v[1] + b.v[1], v[2] + b.v[2]); }
};
template <typename T> Vec3<T> operator*(const Mat3<T>& a, const Vec3<T>& b) {
    return Vec3<T>(a.m[0] * b.v[0] + a.m[1] * b.v[1] + a.m[2] * b.v[2], a.m[3] * b.v[0] + a.m[4] * b.v[1] + a.m[5] * b.v[2],
                   a.m[6] * b.v[0] + a.m[7] * b.v[1] + a.m[8] * b.v[2]);
}
typedef Mat3<double> Matrix3d;
typedef Vec3<double> Vector3d;
typedef Mat3<float> Matrix3f;
typedef Vec3<float> Vector3f;

// geometry_msgs::Pose replacement
struct Pose { struct { double x, y, z; } position; struct { double x, y, z, w; } orientation; };

typedef std::vector<float> VectorXf;
struct MatrixXf { int rows = 0, cols = 0; std::vector<float> data; float& operator()(int r, int c) { return data[(size_t)r * cols + c]; } };
struct ArrayXXd { int rows = 0, cols = 0; std::vector<double> data; double& operator()(int r, int c) { return data[(size_t)r * cols + c]; } };
typedef ArrayXXd MatrixXd;

#ifdef DVO_HAVE_EIGEN
// row-major PODs <-> Eigen's (column-major) dynamic types, element by element
template <typename E> inline void toEigen(const ArrayXXd& a, E& e) { e.resize(a.rows, a.cols); for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) e(r, c) = a.data[(size_t)r * a.cols + c]; }
template <typename E> inline void fromEigen(const E& e, ArrayXXd& a) { a.rows = (int)e.rows(); a.cols = (int)e.cols(); a.data.resize((size_t)a.rows * a.cols); for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) a.data[(size_t)r * a.cols + c] = e(r, c); }
inline void toEigen(const MatrixXf& a, Eigen::MatrixXf& e) { e.resize(a.rows, a.cols); for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) e(r, c) = a.data[(size_t)r * a.cols + c]; }
inline void toEigen(const VectorXf& a, Eigen::VectorXf& e) { e.resize((int)a.size()); for (size_t i = 0; i < a.size(); ++i) e((int)i) = a[i]; }
#endif
#ifdef DVO_HAVE_OPENCV
inline void toMat(const void* src, int rows, int cols, int cvtype, size_t bytes, cv::Mat& out) { out.create(rows, cols, cvtype); assert(out.isContinuous()); std::memcpy(out.data, src, bytes); }
#endif

}  // namespace dvo
