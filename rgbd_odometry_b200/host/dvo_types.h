// dvo_types.h -- POD stand-ins for the cv::Mat / Eigen / geometry_msgs types of the reference's class API.
// Neither OpenCV nor Eigen is installed in this image; these views carry the same information (pointer, shape,
// element type) so the class signatures keep the reference's names, arity and argument meaning.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

namespace dvo {

enum ImageType { U8C1 = 0, U8C3 = 1, U16C1 = 2 };

// a view of a tightly packed row-major image (what cv::Mat::data / rows / cols / type() expose)
struct ImageView {
    const void* data = nullptr;
    int rows = 0, cols = 0, type = U8C1;
    ImageView() {}
    ImageView(const void* d, int r, int c, int t) : data(d), rows(r), cols(c), type(t) {}
    int channels() const { return type == U8C3 ? 3 : 1; }
    size_t bytes() const { return (size_t)rows * cols * (type == U8C1 ? 1 : (type == U8C3 ? 3 : 2)); }
};

// Eigen::Matrix<T,3,3> / Matrix<T,3,1> replacements, row-major storage, (r,c) access
template <typename T> struct Mat3 {
    T m[9];
    Mat3() { for (int i = 0; i < 9; ++i) m[i] = (i % 4 == 0) ? T(1) : T(0); }
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

}  // namespace dvo
