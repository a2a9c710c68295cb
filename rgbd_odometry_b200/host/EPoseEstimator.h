// EPoseEstimator.h -- dense photometric pose estimator (reference include/EPoseEstimator.h:33-106,
// src/EPoseEstimator.cpp).  Same public method names and arity; cv::Mat -> dvo::ImageView, Eigen -> dvo::Matrix3d /
// Vector3d.  All compute runs in libdvo_b200.so (dvo_photo_*); there is no host fallback.
#pragma once
#include "PyramidalStorage.h"
#include "dvo_b200.h"
#include "dvo_types.h"

class EPoseEstimator {
public:
    // bug_compat = true: the reference arithmetic bug for bug; estimate() then only produces the
    // (singular) normal equations.  false: corrected formulation with Huber weights / LM damping.
    explicit EPoseEstimator(bool bug_compat = true);
    virtual ~EPoseEstimator();

    void setCameraMatrix(char* calibFile);                          // OpenCV-XML "cameraMatrix" (src/EPoseEstimator.cpp:35-59)
    void setCameraMatrix(double fx, double fy, double cx, double cy);
    void setRefFrame(dvo::ImageView& rgb, dvo::ImageView& depth);  // :68-108
    void setNowFrame(dvo::ImageView& rgb, dvo::ImageView& depth);  // :117-126
    void setPyramidalImages(int level);                             // :216-260
    void setRefPyramidalImages(int level);                          // :266-290
    float estimate(dvo::Matrix3d& initR, dvo::Vector3d& initT);     // :135-209
    void xdebug();
#ifdef DVO_HAVE_OPENCV
    // the reference's signatures (include/EPoseEstimator.h:38-39)
    void setRefFrame(cv::Mat& rgb, cv::Mat& depth) { dvo::ImageView r(rgb), d(depth); setRefFrame(r, d); }
    void setNowFrame(cv::Mat& rgb, cv::Mat& depth) { dvo::ImageView r(rgb), d(depth); setNowFrame(r, d); }
#endif
#ifdef DVO_HAVE_EIGEN
    float estimate(Eigen::Matrix3d& initR, Eigen::Vector3d& initT) {       // include/EPoseEstimator.h:45
        dvo::Matrix3d R(initR); dvo::Vector3d T(initT);
        const float v = estimate(R, T);
        initR = static_cast<Eigen::Matrix3d>(R); initT = static_cast<Eigen::Vector3d>(T);
        return v;
    }
#endif

    // public members, as in the reference ("//private:" is commented out, include/EPoseEstimator.h:54)
    bool cameraIntrinsicsReady;
    double fx, fy, cx, cy;
    bool isRefFrameAvailable, isNowFrameAvailable, isPydImageAvailableRef, isPydImageAvailableNow, isJEvaluated, is3dCordsReady;
    int pydLevel; double scaleFactor;
    PyramidalStorageStruct pydStore;
    double A[36];                      // J' * J of the current level (:229)
    int iterations;                    // the reference hard-codes 3 (:163)
    double huber_k, lm_lambda0;        // corrected mode only
    dvo_photo_info lastInfo;
private:
    bool compat_; dvo_photo_ctx* ctx_; int width_, height_;
    void ensureContext(int width, int height);
};
