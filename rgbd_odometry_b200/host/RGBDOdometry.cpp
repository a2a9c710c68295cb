#include "RGBDOdometry.h"

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

RGBDOdometry::RGBDOdometry()
    : cameraIntrinsicsReady(false), fx(0), fy(0), cx(0), cy(0), isFrameAvailable(false), isRefFrameAvailable(false), isNowFrameAvailable(false),
      isPyramidalRefFrameAvailable(false), isPyramidalNowFrameAvailable(false), isJacobiansAvailable(false), nFrame(0), ctx_(nullptr), width_(0), height_(0) {
    const_gradientThreshold = 5; const_maxJacobianSize = 50000; const_minimumRequiredPts = 100;       // src/RGBDOdometry.cpp:32-34
    std::memset(&lastInfo, 0, sizeof(lastInfo));
}
RGBDOdometry::~RGBDOdometry() { if (ctx_) dvo_rgbd_destroy(ctx_); }

void RGBDOdometry::check(int rc, const char* what) {
    if (rc != DVO_OK) { std::fprintf(stderr, "[RGBDOdometry::%s] %s\n", what, dvo_last_error()); std::abort(); }   // the reference asserts
}

void RGBDOdometry::ensureContext(int width, int height) {
    if (ctx_ && width == width_ && height == height_) return;
    if (ctx_) { dvo_rgbd_destroy(ctx_); ctx_ = nullptr; }
    dvo_rgbd_config cfg = {width, height, 4, 1, 0};                                                    // 4 levels (:343)
    check(dvo_rgbd_create(&cfg, &ctx_), "RGBDOdometry");
    width_ = width; height_ = height;
    if (cameraIntrinsicsReady) check(dvo_rgbd_set_intrinsics(ctx_, fx, fy, cx, cy), "setCameraMatrix");
}

void RGBDOdometry::setCameraMatrix(char* calibFile) {
    std::ifstream f(calibFile);
    if (!f.is_open()) { std::fprintf(stderr, "[RGBDOdometry::setCameraMatrix] Error opening camera params file : %s\n", calibFile); return; }   // :41-46
    std::stringstream ss; ss << f.rdbuf(); const std::string s = ss.str();
    size_t p = s.find("<cameraMatrix"); if (p == std::string::npos) return;
    p = s.find("<data>", p); if (p == std::string::npos) return;
    std::stringstream d(s.substr(p + 6));
    double v[9]; for (int i = 0; i < 9; ++i) if (!(d >> v[i])) return;
    setCameraMatrix(v[0], v[4], v[2], v[5]);                                                           // :51-54
}
void RGBDOdometry::setCameraMatrix(double fx_, double fy_, double cx_, double cy_) {
    fx = fx_; fy = fy_; cx = cx_; cy = cy_; cameraIntrinsicsReady = true;
    if (ctx_) check(dvo_rgbd_set_intrinsics(ctx_, fx, fy, cx, cy), "setCameraMatrix");
}

void RGBDOdometry::setRcvdFrame(const dvo::ImageView& frame, const dvo::ImageView& dframe) {
    assert(frame.type == dvo::U8C3 && dframe.type == dvo::U16C1 && frame.rows == dframe.rows && frame.cols == dframe.cols);
    rcvd_frame_.assign((const uint8_t*)frame.data, (const uint8_t*)frame.data + frame.bytes());
    rcvd_dframe_.assign((const uint16_t*)dframe.data, (const uint16_t*)dframe.data + (size_t)dframe.rows * dframe.cols);
    ensureContext(frame.cols, frame.rows);
    isFrameAvailable = true;
}

void RGBDOdometry::setRefFrame(dvo::ImageView rgb, dvo::ImageView depth) {
    assert(isFrameAvailable && "Frame not retrived to set in refFrame");
    ensureContext(rgb.cols, rgb.rows);
    isPyramidalRefFrameAvailable = false;
    check(dvo_rgbd_set_frames(ctx_, 0, 0, 1, (const uint8_t*)rgb.data, (const uint16_t*)depth.data, DVO_MEM_HOST), "setRefFrame");
    isRefFrameAvailable = true; isPyramidalRefFrameAvailable = true;
}
void RGBDOdometry::setNowFrame(dvo::ImageView rgb, dvo::ImageView depth) {
    assert(isFrameAvailable && "Frame not retrived to set in nowFrame");
    ensureContext(rgb.cols, rgb.rows);
    isPyramidalNowFrameAvailable = false;
    check(dvo_rgbd_set_frames(ctx_, 1, 0, 1, (const uint8_t*)rgb.data, (const uint16_t*)depth.data, DVO_MEM_HOST), "setNowFrame");
    isNowFrameAvailable = true; isPyramidalNowFrameAvailable = true;
}

void RGBDOdometry::computeJacobianAllLevels() {
    assert(isPyramidalRefFrameAvailable && isRefFrameAvailable && cameraIntrinsicsReady);
    isJacobiansAvailable = false;
    check(dvo_rgbd_compute_jacobians(ctx_, 0, 1, const_gradientThreshold), "computeJacobianAllLevels");
    _A.clear();
    dvo::MatrixXd ph; ph.rows = 5; ph.cols = 5; ph.data.assign(25, 0.0); _A.push_back(ph);              // placeholder for level 0 (:400-402)
    for (int l = 1; l <= 3; ++l) {
        dvo::MatrixXd A; A.rows = 6; A.cols = 6; A.data.assign(36, 0.0);
        int n = 0;
        check(dvo_rgbd_get_A(ctx_, 0, l, A.data.data(), &n), "computeJacobianAllLevels");
        assert(n > const_minimumRequiredPts && "Bad image....too few points with good texture");       // :497
        assert(n < const_maxJacobianSize);                                                              // :463
        _A.push_back(A);
    }
    isJacobiansAvailable = true;
}

void RGBDOdometry::computeJacobian(int level, dvo::MatrixXd& J, MatrixXi& semiDenseMarkings) {
    assert(level >= 0 && level < 4 && "Level has to be between 0 and 4");
    assert(isPyramidalRefFrameAvailable && isRefFrameAvailable && cameraIntrinsicsReady);
    int rows = 0, cols = 0; check(dvo_rgbd_level_dims(ctx_, level, &rows, &cols), "computeJacobian");
    const int cap = rows * cols;
    std::vector<int> ij((size_t)2 * cap); std::vector<double> Jd((size_t)6 * cap);
    TransformRep I; int n = 0;
    check(dvo_rgbd_eval(ctx_, 0, level, I.m, const_gradientThreshold, cap, &n, ij.data(), Jd.data(), nullptr, nullptr, nullptr, nullptr, nullptr), "computeJacobian");
    J.rows = n; J.cols = 6; J.data.assign(Jd.begin(), Jd.begin() + (size_t)6 * n);
    semiDenseMarkings.rows = rows; semiDenseMarkings.cols = cols; semiDenseMarkings.data.assign((size_t)cap, 0);
    for (int k = 0; k < n; ++k) semiDenseMarkings(ij[2 * k], ij[2 * k + 1]) = 1;                       // :468
}

void RGBDOdometry::gaussNewtonIterations(int level, TransformRep& T_) {
    assert(isPyramidalRefFrameAvailable && isPyramidalNowFrameAvailable && isJacobiansAvailable);
    assert(level >= 0 && level < 4 && "Level has to be between 0 and 4");
    assert((level != 0) && "Critical error, jacobians at level-0 (base) are not computed for complexity reasons");
    dvo_rgbd_params p = {3, const_gradientThreshold, const_minimumRequiredPts, const_maxJacobianSize, 200.0};   // :541, :556
    check(dvo_rgbd_set_pose(ctx_, 0, 1, T_.m), "gaussNewtonIterations");
    check(dvo_rgbd_gauss_newton(ctx_, 0, 1, level, &p), "gaussNewtonIterations");
    check(dvo_rgbd_get_poses(ctx_, 0, 1, T_.m, &lastInfo), "gaussNewtonIterations");
}

void RGBDOdometry::computeEpsilon(int level, TransformRep T_, std::vector<double>& epsilon, MatrixXi& newroimask) {
    assert(isPyramidalRefFrameAvailable && isPyramidalNowFrameAvailable && isJacobiansAvailable);
    assert(level >= 0 && level < 4 && (level != 0));
    int rows = 0, cols = 0; check(dvo_rgbd_level_dims(ctx_, level, &rows, &cols), "computeEpsilon");
    const int cap = rows * cols;
    std::vector<int> uv((size_t)2 * cap); epsilon.assign((size_t)cap, 0.0);
    int n = 0;
    check(dvo_rgbd_eval(ctx_, 0, level, T_.m, const_gradientThreshold, cap, &n, nullptr, nullptr, epsilon.data(), uv.data(), nullptr, nullptr, nullptr), "computeEpsilon");
    epsilon.resize((size_t)n);
    newroimask.rows = rows; newroimask.cols = cols; newroimask.data.assign((size_t)cap, 0);
    for (int k = 0; k < n; ++k) if (uv[2 * k] >= 0) newroimask(uv[2 * k], uv[2 * k + 1]) = 1;          // :667
}

void RGBDOdometry::to_se_3(const double* w, double* wx) {
    for (int i = 0; i < 9; ++i) wx[i] = 0.0;
    wx[1 * 3 + 2] = -w[0]; wx[0 * 3 + 2] = w[1]; wx[0 * 3 + 1] = -w[2]; wx[2 * 3 + 1] = w[0]; wx[2 * 3 + 0] = -w[1]; wx[1 * 3 + 0] = w[2];
}
void RGBDOdometry::exponentialMap(const double* psi, double* outTr) {
    const double* t = psi; const double* w = psi + 3;
    double wx[9]; to_se_3(w, wx);
    const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    for (int k = 0; k < 16; ++k) outTr[k] = (k % 5 == 0) ? 1.0 : 0.0;
    if (theta < 1E-12) return;
    double wx2[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) wx2[3 * r + c] = wx[3 * r] * wx[c] + wx[3 * r + 1] * wx[3 + c] + wx[3 * r + 2] * wx[6 + c];
    const double a = std::sin(theta) / theta, b = (1.0 - std::cos(theta)) / (theta * theta), c3 = (theta - std::sin(theta)) / (theta * theta * theta);
    double V[9];
    for (int k = 0; k < 9; ++k) { const double I = (k % 4 == 0) ? 1.0 : 0.0; outTr[4 * (k / 3) + (k % 3)] = I + a * wx[k] + b * wx2[k]; V[k] = I + b * wx[k] + c3 * wx2[k]; }
    for (int r = 0; r < 3; ++r) outTr[4 * r + 3] = V[3 * r] * t[0] + V[3 * r + 1] * t[1] + V[3 * r + 2] * t[2];
}

TransformRep RGBDOdometry::processFrame(int refEvery) {
    assert(isFrameAvailable);
    const dvo::ImageView frame(rcvd_frame_.data(), height_, width_, dvo::U8C3), dframe(rcvd_dframe_.data(), height_, width_, dvo::U16C1);
    if ((nFrame % refEvery) == 0) {                                                                    // :147-154
        base = base * T;
        setRefFrame(frame, dframe);
        T = TransformRep();
        computeJacobianAllLevels();
    }
    setNowFrame(frame, dframe);
    gaussNewtonIterations(3, T);                                                                       // :157-158
    gaussNewtonIterations(2, T);
    ++nFrame;
    isFrameAvailable = false;
    return base * T;                                                                                   // :170
}
