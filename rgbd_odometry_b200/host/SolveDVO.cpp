#include "SolveDVO.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>

SolveDVO::SolveDVO(int width, int height, int levels)
    : fx(0), fy(0), cx(0), cy(0), isCameraIntrinsicsAvailable(false), isFrameAvailable(false), isRefFrameAvailable(false),
      isNowFrameAvailable(false), isPrevFrameAvailable(false), nFrame(0), lastRefFrame(0), ctx_(nullptr), width_(width), height_(height),
      levels_(levels) {
    ratio_of_visible_pts_thresh = 0.8f; laplacianThreshExitCond = 3.0f;                 // src/SolveDVO.cpp:22-23
    psiNormTerminationThreshold = 1.0E-7f; trustRegionHyperSphereRadius = 0.003f;       // :24-25
    for (int i = 0; i < levels; ++i) iterationsConfig.push_back(50);                    // :30-33
    std::memset(&solver, 0, sizeof(solver));
    solver.solver = DVO_SOLVER_SUBGRAD_REF; solver.jacobian = DVO_JAC_REFERENCE; solver.weight = DVO_WEIGHT_REF_CAUCHY;
    solver.arithmetic = DVO_ARITH_EXACT; solver.huber_k = 1.345f; solver.lm_lambda0 = 1e-3;
    solver.residual = DVO_RESIDUAL_DT_FLOOR;               // set to DVO_RESIDUAL_DT_INTERP for the __INTERPOLATE_DISTANCE_TRANSFORM build (:443)
    std::memset(&lastInfo, 0, sizeof(lastInfo));
    casualFolder = "TUM_RGBD/fr1_rpy"; casualRefIndex = 80; casualNowIndex = 85; casualIterations = 100;   // :2397, :2414, :2428
    fileLoopFolder = "TUM_RGBD/fr2_desk"; fileLoopStart = 300; fileLoopEnd = 3000; fileLoopIterations = 500;   // :2460-2463, :2513
    dryFrames = 0;
    // no per-iteration trace: energyAtEachIteration comes from dvo_get_energies, so the class API runs the production kernels
    dvo_config cfg = {width, height, levels, 1, 0, /*keep_now_depth*/ 1, /*trace_iters*/ 0};
    const int rc = dvo_create(&cfg, &ctx_);
    if (rc != DVO_OK) { std::fprintf(stderr, "SolveDVO: %s\n", dvo_last_error()); std::abort(); }   // no CPU fallback
}
SolveDVO::~SolveDVO() { if (ctx_) dvo_destroy(ctx_); }

void SolveDVO::check(int rc, const char* what) {
    if (rc != DVO_OK) { std::fprintf(stderr, "SolveDVO::%s: %s\n", what, dvo_last_error()); assert(rc == DVO_OK); std::abort(); }
}

void SolveDVO::setCameraMatrix(const char* calibFile) {
    std::ifstream f(calibFile);
    if (!f.is_open()) { std::fprintf(stderr, "[SolveDVO::setCameraMatrix] Error opening camera params file : %s\n", calibFile); return; }   // :94-99
    std::stringstream ss; ss << f.rdbuf(); const std::string s = ss.str();
    size_t p = s.find("<cameraMatrix"); if (p == std::string::npos) return;
    p = s.find("<data>", p); if (p == std::string::npos) return;
    std::stringstream d(s.substr(p + 6));
    double v[9]; for (int i = 0; i < 9; ++i) if (!(d >> v[i])) return;
    setIntrinsics((float)v[0], (float)v[4], (float)v[2], (float)v[5]);                  // :103-106
}
void SolveDVO::setIntrinsics(float fx_, float fy_, float cx_, float cy_) {
    fx = fx_; fy = fy_; cx = cx_; cy = cy_;
    check(dvo_set_intrinsics(ctx_, fx, fy, cx, cy), "setIntrinsics");
    isCameraIntrinsicsAvailable = true;
}

void SolveDVO::setRcvdFrame(const dvo::ImageView& framemono, const dvo::ImageView& dframe) {
    assert(framemono.rows == height_ && framemono.cols == width_ && framemono.type == dvo::U8C1);
    assert(dframe.rows == height_ && dframe.cols == width_ && dframe.type == dvo::U16C1);
    isFrameAvailable = false;
    rcvd_gray_.assign((const uint8_t*)framemono.data, (const uint8_t*)framemono.data + framemono.bytes());
    rcvd_depth_.assign((const uint16_t*)dframe.data, (const uint16_t*)dframe.data + (size_t)dframe.rows * dframe.cols);
    isFrameAvailable = true;
}

// SolveDVO::loadFromFile (src/SolveDVO.cpp:154-190).  The dump holds every pyramid level; the device rebuilds levels
// 1.. from level 0 with the publisher's own rule (NEAREST, src/camTopic2PublisherPyD.cpp:338-348), so only mono_0 /
// depth_0 are uploaded.  A dump whose stored levels disagree with that rule is reported (it was not written by the
// reference's publisher) but still loaded from its level 0.
bool SolveDVO::loadFromFile(const char* xmlFileName) {
    dvo::RGBDFramePyd pyd;
    if (!dvo::loadFrameXml(xmlFileName, pyd, levels_)) return false;                     // :158-162: logs, leaves isFrameAvailable untouched
    const dvo::Image& m0 = pyd.framemono[0]; const dvo::Image& d0 = pyd.dframe[0];
    if (m0.elem != dvo::ELEM_U8 || m0.channels != 1 || d0.elem != dvo::ELEM_U16 || d0.channels != 1 || m0.rows != height_ || m0.cols != width_ ||
        d0.rows != height_ || d0.cols != width_) {
        std::fprintf(stderr, "[SolveDVO::loadFromFile] %s: mono_0 / depth_0 are not %dx%d u8 / u16\n", xmlFileName, width_, height_);
        return false;
    }
    for (int l = 1; l < levels_; ++l) {
        const dvo::Image& m = pyd.framemono[l];
        int lw = 0, lh = 0; dvo_level_dims(ctx_, l, &lw, &lh);
        bool same = (m.rows == lh && m.cols == lw && m.elem == dvo::ELEM_U8);
        const int s = 1 << l;
        for (int y = 0; same && y < lh; ++y) for (int x = 0; x < lw; ++x) if (m.data[(size_t)y * lw + x] != m0.data[(size_t)(y * s) * width_ + x * s]) { same = false; break; }
        if (!same) std::fprintf(stderr, "[SolveDVO::loadFromFile] %s: stored mono_%d is not the NEAREST subsample of mono_0; rebuilt on the device\n", xmlFileName, l);
    }
    setRcvdFrame(m0.view(), d0.view());
    return true;
}

void SolveDVO::setRcvdFrameAsRefFrame() {
    assert(isFrameAvailable);
    isRefFrameAvailable = false;
    check(dvo_set_frames(ctx_, DVO_FRAME_REF, 0, 1, rcvd_gray_.data(), rcvd_depth_.data(), DVO_MEM_HOST), "setRcvdFrameAsRefFrame");
    check(dvo_build_pyramids(ctx_, 0, 1, 1), "setRcvdFrameAsRefFrame");
    isRefFrameAvailable = true;
    computeDistTransfrmOfRef();
}
void SolveDVO::setPrevFrameAsRefFrame() {
    assert(isPrevFrameAvailable);
    isRefFrameAvailable = false;
    check(dvo_promote_now_to_ref(ctx_, 0, 1), "setPrevFrameAsRefFrame");
    check(dvo_build_pyramids(ctx_, 0, 1, 1), "setPrevFrameAsRefFrame");
    isRefFrameAvailable = true;
    computeDistTransfrmOfRef();
}
void SolveDVO::setRcvdFrameAsNowFrame() {
    assert(isFrameAvailable && "FRAME NOT AVAILABLE");
    if (isNowFrameAvailable) isPrevFrameAvailable = true;       // p_now_* are kept on the device by dvo_set_frames (:594-600)
    isNowFrameAvailable = false;
    check(dvo_set_frames(ctx_, DVO_FRAME_NOW, 0, 1, rcvd_gray_.data(), rcvd_depth_.data(), DVO_MEM_HOST), "setRcvdFrameAsNowFrame");
    check(dvo_build_pyramids(ctx_, 0, 1, 2), "setRcvdFrameAsNowFrame");
    isNowFrameAvailable = true;
    computeDistTransfrmOfNow();
}
// Canny + (ref) edge-point list; the reference's ref DT / gradients are computed and never read and are skipped
void SolveDVO::computeDistTransfrmOfRef() { assert(isRefFrameAvailable); check(dvo_prepare(ctx_, 0, 1, 1), "computeDistTransfrmOfRef"); }
void SolveDVO::computeDistTransfrmOfNow() { assert(isNowFrameAvailable); check(dvo_prepare(ctx_, 0, 1, 2), "computeDistTransfrmOfNow"); }

void SolveDVO::preProcessRefFrame() {
    assert(isRefFrameAvailable);
    // selectedPts + enlistRefEdgePts ran on the device inside dvo_prepare; keep the reference's invariant (:282)
    for (int l = 0; l < levels_; ++l) { int n = 0; check(dvo_get_points(ctx_, 0, l, nullptr, nullptr, nullptr, 0, &n), "preProcessRefFrame"); assert(n > 0 && "nSelectedPts > 0"); }
}

void SolveDVO::runIterations(int level, int maxIterations, dvo::Matrix3d& cR, dvo::Vector3d& cT, dvo::VectorXf& energyAtEachIteration,
                             dvo::VectorXf& finalEpsilons, dvo::MatrixXf& finalReprojections, int& bestEnergyIndex, float& finalVisibleRatio) {
    assert(level >= 0 && level < levels_);                      // :625 (the reference asserts level <= 3)
    assert(maxIterations > 0);                                 // energies beyond the first 128 iterations are not recorded (left 0)
    assert(isRefFrameAvailable && isNowFrameAvailable && isCameraIntrinsicsAvailable);
    dvo_solver_params p = solver;
    for (int l = 0; l < DVO_MAX_LEVELS; ++l) p.iters[l] = 0;
    p.iters[level] = maxIterations;
    double pose[12];
    for (int i = 0; i < 9; ++i) pose[i] = cR.m[i];
    for (int i = 0; i < 3; ++i) pose[9 + i] = cT.v[i];
    check(dvo_set_initial_pose(ctx_, 0, 1, pose, DVO_MEM_HOST), "runIterations");
    check(dvo_run(ctx_, 0, 1, &p), "runIterations");
    check(dvo_get_poses(ctx_, 0, 1, pose, &lastInfo, DVO_MEM_HOST), "runIterations");
    for (int i = 0; i < 9; ++i) cR.m[i] = pose[i];
    for (int i = 0; i < 3; ++i) cT.v[i] = pose[9 + i];
    float en[128];
    check(dvo_get_energies(ctx_, 0, level, en, 128), "runIterations");
    energyAtEachIteration.assign(maxIterations, 0.f);           // :634
    for (int k = 0; k < maxIterations && k < 128 && k < lastInfo.iterations_run[level]; ++k) energyAtEachIteration[k] = en[k];
    const int n = lastInfo.npts[level];
    finalEpsilons.assign(n, 0.f); finalReprojections.rows = 3; finalReprojections.cols = n; finalReprojections.data.assign((size_t)3 * n, 1.f);
    if (n > 0) {
        double H[36], g[6], ss; int nvis;
        check(dvo_eval_normal_equations_ex(ctx_, 0, level, pose, &p, H, g, &ss, &nvis, finalEpsilons.data(),
                                           nullptr, finalReprojections.data.data(), finalReprojections.data.data() + n, nullptr), "runIterations");
    }
    bestEnergyIndex = lastInfo.best_index[level]; finalVisibleRatio = lastInfo.visible_ratio[level];
}

float SolveDVO::processResidueHistogram(dvo::VectorXf& residi, bool /*quite*/) {
    float b_cap = 0; for (size_t i = 0; i < residi.size(); ++i) b_cap += residi[i];     // :1466-1472
    return residi.empty() ? 0.f : b_cap / (float)residi.size();
}

int SolveDVO::processFrame() {
    assert(isFrameAvailable && isCameraIntrinsicsAvailable);
    dvo::VectorXf energy, eps; dvo::MatrixXf reproj; int bestIdx = -1; float vis = 0.f;
    if (!isRefFrameAvailable) {                                 // phase A (:2009-2030): first frame is the reference / key frame
        setRcvdFrameAsRefFrame(); preProcessRefFrame(); lastRefFrame = 0;
        gop.pushAsKeyFrame((int)nFrame, 1, cR_64, cT_64);
        isFrameAvailable = false; nFrame++;
        return gop.size() - 1;
    }
    setRcvdFrameAsNowFrame();                                   // phase B (:2060-2240)
    bool signalGetNewRefImage = (nFrame - lastRefFrame) == 5;   // :2155-2160
    const bool switch_ref = signalGetNewRefImage && lastRefFrame != (nFrame - 1);
    if (!switch_ref) {
        for (int f = (int)iterationsConfig.size() - 1; f >= 0; --f)                    // :2097-2104
            if (iterationsConfig[f] > 0) runIterations(f, iterationsConfig[f], cR_64, cT_64, energy, eps, reproj, bestIdx, vis);
        gop.pushAsOrdinaryFrame((int)nFrame, cR_64, cT_64);                              // :2239
    } else {
        // the reference first solves against the outgoing key frame and discards the result (:2097-2104, :2210-2211)
        lastRefFrame = nFrame - 1;                                                       // :2201
        setPrevFrameAsRefFrame(); preProcessRefFrame();
        gop.updateMostRecentToKeyFrame(5);                                               // :2207
        cR_64 = dvo::Matrix3d::Identity(); cT_64 = dvo::Vector3d::Zero();                // :2210-2211
        for (int f = (int)iterationsConfig.size() - 1; f >= 0; --f)                    // :2220-2227
            if (iterationsConfig[f] > 0) runIterations(f, iterationsConfig[f], cR_64, cT_64, energy, eps, reproj, bestIdx, vis);
        gop.pushAsOrdinaryFrame((int)nFrame, cR_64, cT_64);                              // :2232
    }
    isFrameAvailable = false; nFrame++;
    return gop.size() - 1;
}

void SolveDVO::loopFromFrames(const uint8_t* gray, const uint16_t* depth, int nframes) {
    const size_t P = (size_t)width_ * height_;
    for (int t = 0; t < nframes; ++t) {
        dvo::ImageView g(gray + P * t, height_, width_, dvo::U8C1), d(depth + P * t, height_, width_, dvo::U16C1);
        setRcvdFrame(g, d);
        processFrame();
    }
}

// ---------------------------------------------------------------------------------------------------- the reference's loops
void SolveDVO::setXmlFrameSource(const char* folder, int start, int end) {
    const std::string dir(folder);
    int next = start;
    frameSource_ = [dir, next, end](SolveDVO& self) mutable -> bool {
        if (next > end) { std::fprintf(stderr, "Done with all files...Quitting...\n"); return false; }          // :1955-1959
        char name[1000];
        std::snprintf(name, sizeof(name), "%s/framemono_%04d.xml", dir.c_str(), next);                            // :1961
        if (!self.loadFromFile(name)) { std::fprintf(stderr, "No More files, Quitting..\n"); return false; }     // :1963-1966
        ++next;
        return true;
    };
}

// loop() (:1896-2373): the first available frame becomes the reference / first key frame (:1950-2030), every later frame goes
// through setRcvdFrameAsNowFrame, the coarse-to-fine runIterations and the key-frame rule -- processFrame() is one turn.
void SolveDVO::loop() {
    assert(frameSource_ && "setFrameSource / setXmlFrameSource first (replaces the ROS subscription)");
    while (frameSource_(*this)) {
        if (!isFrameAvailable) continue;                                                                          // :1971, :2055
        processFrame();
    }
}

// loopDry() (:1803-1893): frames are consumed and set as the now frame (pyramid, edges, distance transform), nothing is solved.
void SolveDVO::loopDry() {
    assert(frameSource_ && "setFrameSource / setXmlFrameSource first");
    dryFrames = 0;
    while (frameSource_(*this)) {
        if (!isFrameAvailable) continue;                                                                          // :1843
        setRcvdFrameAsNowFrame();                                                                                 // :1848
        isFrameAvailable = false;                                                                                 // :1886
        ++dryFrames;                                                                                              // nFrame++ (:1888)
    }
}

// casualTestFunction() (:2377-2442): reference frame from one dump, now frame from another, runIterations(0, 100, ...) from
// identity; the energies it prints to stdout are kept in casualEnergies as well.
void SolveDVO::casualTestFunction() {
    char frameFileName[1000];
    cR_64 = dvo::Matrix3d::Identity(); cT_64 = dvo::Vector3d::Zero();                                             // :2383-2384
    dvo::VectorXf epsilonVec; dvo::MatrixXf reprojections; int bestEnergyIndex = -1; float visibleRatio = 0.0f;
    std::snprintf(frameFileName, sizeof(frameFileName), "%s/framemono_%04d.xml", casualFolder.c_str(), casualRefIndex);
    if (!loadFromFile(frameFileName)) { std::fprintf(stderr, "Cannot open file1\n"); return; }                   // :2399-2401 (the reference carries on and asserts)
    setRcvdFrameAsRefFrame(); preProcessRefFrame();                                                               // :2405-2406
    std::snprintf(frameFileName, sizeof(frameFileName), "%s/framemono_%04d.xml", casualFolder.c_str(), casualNowIndex);
    if (!loadFromFile(frameFileName)) { std::fprintf(stderr, "Cannot open file1\n"); return; }                   // :2416-2418
    setRcvdFrameAsNowFrame();                                                                                     // :2422
    runIterations(0, casualIterations, cR_64, cT_64, casualEnergies, epsilonVec, reprojections, bestEnergyIndex, visibleRatio);   // :2428
    for (size_t i = 0; i < casualEnergies.size(); ++i) std::printf("%g\n", casualEnergies[i]);                    // :2438-2441
}

// loopFromFile() (:2448-2600): dumps START..END-1.  A file whose index is a multiple of 5 becomes the reference -- after its pose
// against the outgoing reference has been estimated when it is not the first (:2507-2531) -- and then EVERY file, the new
// reference included (the `else` is commented out, :2533-2544), is set as the now frame and aligned with one
// runIterations(0, 500) at level 0.  nT = keyT + keyR cT, nR = keyR cR (:2547-2548) is kept per file in fileLoopR / fileLoopT
// (the reference hands it to its RViz publisher).
void SolveDVO::loopFromFile() {
    char frameFileName[1000];
    dvo::VectorXf energy, eps; dvo::MatrixXf reproj; int bestIdx = -1; float vis = 0.f;
    dvo::Matrix3d cR = dvo::Matrix3d::Identity(), keyR = dvo::Matrix3d::Identity(), nR = dvo::Matrix3d::Identity();
    dvo::Vector3d cT = dvo::Vector3d::Zero(), keyT = dvo::Vector3d::Zero(), nT = dvo::Vector3d::Zero();
    fileLoopR.clear(); fileLoopT.clear();
    for (int iFrameNum = fileLoopStart; iFrameNum < fileLoopEnd; ++iFrameNum) {
        std::snprintf(frameFileName, sizeof(frameFileName), "%s/framemono_%04d.xml", fileLoopFolder.c_str(), iFrameNum);
        if (!loadFromFile(frameFileName)) { std::fprintf(stderr, "No More files, Quitting..\n"); break; }         // :2500-2503
        if (iFrameNum % 5 == 0) {                                                                                 // :2507
            if (iFrameNum > fileLoopStart) {
                setRcvdFrameAsNowFrame();                                                                         // :2513
                runIterations(0, fileLoopIterations, cR, cT, energy, eps, reproj, bestIdx, vis);                  // :2517
            }
            keyR = nR; keyT = nT;                                                                                 // :2519-2520
            lastRefFrame = iFrameNum;
            setRcvdFrameAsRefFrame(); preProcessRefFrame();                                                       // :2523-2525
            cR = dvo::Matrix3d::Identity(); cT = dvo::Vector3d::Zero();                                           // :2527-2528
        }
        setRcvdFrameAsNowFrame();                                                                                 // :2535
        runIterations(0, fileLoopIterations, cR, cT, energy, eps, reproj, bestIdx, vis);                          // :2538
        nT = keyT + keyR * cT; nR = keyR * cR;                                                                    // :2547-2548
        fileLoopR.push_back(nR); fileLoopT.push_back(nT);
        isFrameAvailable = false;
    }
}
