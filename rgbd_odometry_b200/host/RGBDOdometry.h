// RGBDOdometry.h -- the semi-dense photometric Gauss-Newton odometry (reference include/RGBDOdometry.h:36-131,
// src/RGBDOdometry.cpp).  Same method names, arity and member names, made public so the class can be driven without
// ROS; cv::Mat -> dvo::ImageView, Eigen -> dvo::ArrayXXd / plain arrays, TransformRep (Eigen Affine3d) -> row-major
// 4x4.  All compute runs in libdvo_b200.so (dvo_rgbd_*); there is no host fallback.
#pragma once
#include <vector>

#include "dvo_b200.h"
#include "dvo_types.h"

struct TransformRep {                                   // Eigen::Transform<double, 3, Eigen::Affine> (include/RGBDOdometry.h:33)
    double m[16];
    TransformRep() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }
    double& operator()(int r, int c) { return m[4 * r + c]; }
    const double& operator()(int r, int c) const { return m[4 * r + c]; }
    TransformRep operator*(const TransformRep& b) const {
        TransformRep o;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) o.m[4 * i + j] = ((m[4 * i] * b.m[j] + m[4 * i + 1] * b.m[4 + j]) + m[4 * i + 2] * b.m[8 + j]) + m[4 * i + 3] * b.m[12 + j];
        return o;
    }
};

struct MatrixXi { int rows = 0, cols = 0; std::vector<int> data; int& operator()(int r, int c) { return data[(size_t)r * cols + c]; } };

class RGBDOdometry {
public:
    RGBDOdometry();
    virtual ~RGBDOdometry();

    void setCameraMatrix(char* calibFile);                             // OpenCV-XML "cameraMatrix" (src/RGBDOdometry.cpp:38-60)
    void setCameraMatrix(double fx, double fy, double cx, double cy);
    void setRcvdFrame(const dvo::ImageView& frame, const dvo::ImageView& dframe);    // replaces imageArrivedCallBack (:212-238)
    void setRefFrame(dvo::ImageView rgb, dvo::ImageView depth);        // :330-360
    void setNowFrame(dvo::ImageView rgb, dvo::ImageView depth);        // :362-390
    void computeJacobianAllLevels();                                   // :393-415
    void computeJacobian(int level, dvo::MatrixXd& J, MatrixXi& semiDenseMarkings);   // :407-505 (materialised from the device)
    void gaussNewtonIterations(int level, TransformRep& T);            // :514-597
    void computeEpsilon(int level, TransformRep T, std::vector<double>& epsilon, MatrixXi& newroimask);   // :602-700
    void exponentialMap(const double* psi6, double* outTr16);          // :707-745 (host helper, same formula)
    void to_se_3(const double* w3, double* wx9);                       // :752-763
    // one body of eventLoop (:139-206) for the frame last given to setRcvdFrame: reference switch every `refEvery` frames,
    // gaussNewtonIterations(3, T), gaussNewtonIterations(2, T); returns base * T
    TransformRep processFrame(int refEvery = 10000);

    bool cameraIntrinsicsReady;
    double fx, fy, cx, cy;
    bool isFrameAvailable, isRefFrameAvailable, isNowFrameAvailable, isPyramidalRefFrameAvailable, isPyramidalNowFrameAvailable, isJacobiansAvailable;
    int const_gradientThreshold, const_maxJacobianSize, const_minimumRequiredPts;
    std::vector<dvo::MatrixXd> _A;                                     // A := J' * J per level (index 0 is the reference's 5x5 placeholder)
    dvo_rgbd_info lastInfo;
    int nFrame;
    TransformRep T, base;
private:
    dvo_rgbd_ctx* ctx_; int width_, height_;
    std::vector<uint8_t> rcvd_frame_; std::vector<uint16_t> rcvd_dframe_;
    void ensureContext(int width, int height);
    void check(int rc, const char* what);
};
