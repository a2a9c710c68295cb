#include "EPoseEstimator.h"

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>

EPoseEstimator::EPoseEstimator(bool bug_compat)
    : cameraIntrinsicsReady(false), fx(0), fy(0), cx(0), cy(0), isRefFrameAvailable(false), isNowFrameAvailable(false),
      isPydImageAvailableRef(false), isPydImageAvailableNow(false), isJEvaluated(false), is3dCordsReady(false), pydLevel(0),
      scaleFactor(1.0), iterations(3), huber_k(0.0), lm_lambda0(0.0), compat_(bug_compat), ctx_(nullptr), width_(0), height_(0) {
    for (int i = 0; i < 36; ++i) A[i] = 0.0;
    std::memset(&lastInfo, 0, sizeof(lastInfo));
}
EPoseEstimator::~EPoseEstimator() { if (ctx_) dvo_photo_destroy(ctx_); }

// reads the <cameraMatrix> ... <data> fx 0 cx 0 fy cy 0 0 1 </data> block of an OpenCV FileStorage XML (:35-59);
// failure logs and returns with cameraIntrinsicsReady == false, like the reference
void EPoseEstimator::setCameraMatrix(char* calibFile) {
    std::ifstream f(calibFile);
    if (!f.is_open()) { std::fprintf(stderr, "[PoseEstimator::setCameraMatrix] Error opening camera params file : %s\n", calibFile); return; }
    std::stringstream ss; ss << f.rdbuf(); const std::string s = ss.str();
    size_t p = s.find("<cameraMatrix"); if (p == std::string::npos) return;
    p = s.find("<data>", p); if (p == std::string::npos) return;
    std::stringstream d(s.substr(p + 6));
    double v[9]; for (int i = 0; i < 9; ++i) if (!(d >> v[i])) return;
    setCameraMatrix(v[0], v[4], v[2], v[5]);
}
void EPoseEstimator::setCameraMatrix(double fx_, double fy_, double cx_, double cy_) {
    fx = fx_; fy = fy_; cx = cx_; cy = cy_; cameraIntrinsicsReady = true;
    if (ctx_) dvo_photo_set_intrinsics(ctx_, fx, fy, cx, cy);
}

void EPoseEstimator::ensureContext(int width, int height) {
    if (ctx_ && width == width_ && height == height_) return;
    if (ctx_) { dvo_photo_destroy(ctx_); ctx_ = nullptr; }
    int levels = 5;                                            // the reference always builds levels 0..4 (:88)
    while (levels > 1 && ((width % (1 << (levels - 1))) || (height % (1 << (levels - 1))))) --levels;
    dvo_photo_config cfg = {width, height, levels, 1, 0};
    const int rc = dvo_photo_create(&cfg, &ctx_);
    if (rc != DVO_OK) { std::fprintf(stderr, "EPoseEstimator: %s\n", dvo_last_error()); std::abort(); }   // no CPU fallback
    width_ = width; height_ = height;
    if (cameraIntrinsicsReady) dvo_photo_set_intrinsics(ctx_, fx, fy, cx, cy);
    pydStore.bind(ctx_, 0, compat_ ? 1 : 0); pydStore.width_ = width; pydStore.height_ = height;
}

void EPoseEstimator::setRefFrame(dvo::ImageView& rgb, dvo::ImageView& depth) {
    assert(rgb.channels() == 3);                               // :69
    assert(cameraIntrinsicsReady);
    ensureContext(rgb.cols, rgb.rows);
    int rc = dvo_photo_set_frames(ctx_, DVO_FRAME_REF, 0, 1, (const uint8_t*)rgb.data, (const uint16_t*)depth.data, DVO_MEM_HOST);
    assert(rc == 0);
    isRefFrameAvailable = true; isPydImageAvailableRef = false; isJEvaluated = false; is3dCordsReady = false;
    pydStore.clearPyramid();                                   // :87
    dvo_photo_config cfg; cfg.levels = 5; while (cfg.levels > 1 && ((rgb.cols % (1 << (cfg.levels - 1))) || (rgb.rows % (1 << (cfg.levels - 1))))) --cfg.levels;
    rc = dvo_photo_prepare_ref(ctx_, 0, 1, compat_ ? 1 : 0);   // evaluateJacobian for every level (:88-102)
    assert(rc == 0);
    for (int lvl = 0; lvl < cfg.levels; ++lvl) { setRefPyramidalImages(lvl); pydStore.addLevelFromDevice(lvl, lvl); }   // :88-102, no host copies
    isJEvaluated = true; is3dCordsReady = true;
    (void)rc;
}

void EPoseEstimator::setNowFrame(dvo::ImageView& rgb, dvo::ImageView& depth) {
    assert(rgb.channels() == 3);                               // :118
    ensureContext(rgb.cols, rgb.rows);
    const int rc = dvo_photo_set_frames(ctx_, DVO_FRAME_NOW, 0, 1, (const uint8_t*)rgb.data, (const uint16_t*)depth.data, DVO_MEM_HOST);
    assert(rc == 0); (void)rc;
    isNowFrameAvailable = true; isPydImageAvailableNow = false;
}

void EPoseEstimator::setRefPyramidalImages(int level) {
    assert(isRefFrameAvailable);
    assert(level >= 0 && level <= 4);                          // :269
    pydLevel = level; scaleFactor = std::pow(2.0, -level); isPydImageAvailableRef = true;   // the INTER_AREA levels already exist on the device
}

void EPoseEstimator::setPyramidalImages(int level) {
    assert(isRefFrameAvailable); assert(isNowFrameAvailable);
    assert(level >= 0 && level <= 4);                          // :220
    pydLevel = level; scaleFactor = std::pow(2.0, -level);
    const int rc = dvo_photo_get_A(ctx_, 0, pydStore.deviceLevelAt(level), A);   // A = J.transpose()*J (:229)
    assert(rc == 0); (void)rc;
    isPydImageAvailableRef = true; isPydImageAvailableNow = true; isJEvaluated = true;
}

float EPoseEstimator::estimate(dvo::Matrix3d& R, dvo::Vector3d& T) {
    assert(cameraIntrinsicsReady); assert(isRefFrameAvailable); assert(isNowFrameAvailable);   // :136-141
    assert(isPydImageAvailableRef); assert(isPydImageAvailableNow); assert(isJEvaluated && "J not evaluated\n");
    double pose[12];
    for (int i = 0; i < 9; ++i) pose[i] = R.m[i];
    for (int i = 0; i < 3; ++i) pose[9 + i] = T.v[i];
    int rc = dvo_photo_set_pose(ctx_, 0, 1, pose);
    rc |= dvo_photo_estimate(ctx_, 0, 1, pydStore.deviceLevelAt(pydLevel), iterations, compat_ ? 1 : 0, huber_k, lm_lambda0);
    rc |= dvo_photo_get_poses(ctx_, 0, 1, pose, &lastInfo);
    assert(rc == 0); (void)rc;
    for (int i = 0; i < 9; ++i) R.m[i] = pose[i];
    for (int i = 0; i < 3; ++i) T.v[i] = pose[9 + i];
    // the reference returns nReporj / (rows*cols) with integer division, i.e. 0 or 1 (:171); same here in compat mode
    if (compat_) { const int rows = height_ >> pydLevel, cols = width_ >> pydLevel; return (float)(lastInfo.nreproj / (rows * cols)); }
    return (float)lastInfo.visible;
}

void EPoseEstimator::xdebug() { assert(isPydImageAvailableNow && isPydImageAvailableRef); pydStore.printSize(); }
