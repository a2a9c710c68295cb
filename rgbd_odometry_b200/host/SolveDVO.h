// SolveDVO.h -- edge distance-transform direct odometry (reference include/SolveDVO.h:148-360, src/SolveDVO.cpp).
// Same class and member names for the hot path; the members the reference keeps private (:199-315) are public here so the
// class can be driven without ROS.  Ingest replaces the ROS callback (imageArrivedCallBack, :490-534): setRcvdFrame takes
// the full-resolution mono8 / depth16 frame and the NEAREST pyramid the publisher node would have sent
// (src/camTopic2PublisherPyD.cpp:338-348) is built on the device.  All compute runs in libdvo_b200.so; no host fallback.
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "FrameIO.h"
#include "GOP.h"
#include "dvo_b200.h"
#include "dvo_types.h"

class SolveDVO {
public:
    SolveDVO(int width = 640, int height = 480, int levels = 4);
    ~SolveDVO();

    // ---- the reference's public surface (include/SolveDVO.h:154-168) -------------------------------------------------
    // loop() / loopDry() poll a frame source where the reference polls ros::spinOnce() (:1823, :1952, :2037) or, in its
    // __DATA_FROM_XML_FILES__ build, loads "<dir>/framemono_%04d.xml" (:1826-1840, :1954-1968): the source is called once per
    // turn of the loop; it may hand over a frame (setRcvdFrame / loadFromFile -> isFrameAvailable) and returns false when
    // there is nothing more to come (ros::ok() false / "No More files, Quitting..").
    typedef std::function<bool(SolveDVO&)> FrameSource;
    void setFrameSource(FrameSource src) { frameSource_ = src; }
    void setXmlFrameSource(const char* folder, int start, int end);    // __DATA_FROM_XML_FILES__, ..START, ..END (:119-121)
    void loop();                                                       // :1896-2373
    void loopDry();                                                    // :1803-1893: every frame becomes the now frame; nothing is solved
    void loopFromFile();                                               // :2448-2560: XML dumps, new reference every 5th file, runIterations(0, 500)
    void casualTestFunction();                                         // :2377-2442: two XML dumps, runIterations(0, 100, ...)
    // casualTestFunction / loopFromFile read fixed paths in the reference ("TUM_RGBD/fr1_rpy" files 80 and 85; "TUM_RGBD/fr2_desk"
    // files 300..3000): same defaults, settable
    std::string casualFolder; int casualRefIndex, casualNowIndex, casualIterations;
    std::string fileLoopFolder; int fileLoopStart, fileLoopEnd, fileLoopIterations;
    std::vector<dvo::Matrix3d> fileLoopR; std::vector<dvo::Vector3d> fileLoopT;   // nR / nT of every file loopFromFile() processed
    dvo::VectorXf casualEnergies;                                      // what casualTestFunction prints (:2438-2441)
    long dryFrames;                                                    // frames loopDry() consumed

    void setCameraMatrix(const char* calibFile);                       // :88-126 (OpenCV-XML cameraMatrix)
    void setIntrinsics(float fx, float fy, float cx, float cy);

    // ---- ingest (replaces imageArrivedCallBack :490-534): level-0 mono8 + depth16 (mm); zeros become 1 (:512)
    void setRcvdFrame(const dvo::ImageView& framemono, const dvo::ImageView& dframe);
    bool loadFromFile(const char* xmlFileName);                        // :154-190 (OpenCV XML frame dump, mono_i / depth_i)
    void setRcvdFrameAsRefFrame();                                     // :537-557  (+ computeDistTransfrmOfRef)
    void setRcvdFrameAsNowFrame();                                     // :588-614  (+ computeDistTransfrmOfNow, keeps p_now_*)
    void setPrevFrameAsRefFrame();                                     // :561-584
    void preProcessRefFrame();                                         // :269-303  (asserts nSelectedPts > 0 per level)
    void computeDistTransfrmOfRef();                                   // :1679-1738
    void computeDistTransfrmOfNow();                                   // :1740-1799

    // :619-1017 -- in/out pose, per-iteration energies, residuals / reprojections of the best iterate, its index, visible ratio
    void runIterations(int level, int maxIterations, dvo::Matrix3d& cR, dvo::Vector3d& cT, dvo::VectorXf& energyAtEachIteration,
                       dvo::VectorXf& finalEpsilons, dvo::MatrixXf& finalReprojections, int& bestEnergyIndex, float& finalVisibleRatio);
    float processResidueHistogram(dvo::VectorXf& residi, bool quite = true);   // :1398-1483 (quiet path: mean residual)
#ifdef DVO_HAVE_EIGEN
    // the reference's signature (include/SolveDVO.h:262-266)
    void runIterations(int level, int maxIterations, Eigen::Matrix3d& cR, Eigen::Vector3d& cT, Eigen::VectorXf& energyAtEachIteration,
                       Eigen::VectorXf& finalEpsilons, Eigen::MatrixXf& finalReprojections, int& bestEnergyIndex, float& finalVisibleRatio) {
        dvo::Matrix3d R(cR); dvo::Vector3d T(cT); dvo::VectorXf en, ep; dvo::MatrixXf rp;
        runIterations(level, maxIterations, R, T, en, ep, rp, bestEnergyIndex, finalVisibleRatio);
        cR = static_cast<Eigen::Matrix3d>(R); cT = static_cast<Eigen::Vector3d>(T); dvo::toEigen(en, energyAtEachIteration); dvo::toEigen(ep, finalEpsilons); dvo::toEigen(rp, finalReprojections);
    }
    float processResidueHistogram(Eigen::VectorXf& residi, bool quite = true) {
        dvo::VectorXf r((size_t)residi.rows()); for (int i = 0; i < (int)residi.rows(); ++i) r[(size_t)i] = residi(i);
        return processResidueHistogram(r, quite);
    }
#endif

    // one body of loop() (:2017, :2083-2240) for the frame last given to setRcvdFrame; returns the GOP index it pushed
    int processFrame();
    void loopFromFrames(const uint8_t* gray, const uint16_t* depth, int nframes);   // loop() over an in-memory sequence

    // constants / state with the reference's names (:21-33, :171-197)
    float fx, fy, cx, cy;
    bool isCameraIntrinsicsAvailable, isFrameAvailable, isRefFrameAvailable, isNowFrameAvailable, isPrevFrameAvailable;
    float trustRegionHyperSphereRadius, psiNormTerminationThreshold, laplacianThreshExitCond, ratio_of_visible_pts_thresh;
    std::vector<int> iterationsConfig;
    GOP<double> gop;
    dvo::Matrix3d cR_64; dvo::Vector3d cT_64;
    long nFrame, lastRefFrame;
    dvo_solver_params solver;                                          // SUBGRAD_REF / REFERENCE Jacobian / REF_CAUCHY by default
    dvo_pair_info lastInfo;
    dvo_ctx* context() { return ctx_; }
private:
    FrameSource frameSource_;
    dvo_ctx* ctx_; int width_, height_, levels_;
    std::vector<uint8_t> rcvd_gray_; std::vector<uint16_t> rcvd_depth_;
    void check(int rc, const char* what);
};
