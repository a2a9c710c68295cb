// Test shim compiled WITH the stand-in <opencv2/core/core.hpp> / <Eigen/Dense> of tests/standin_include (neither library exists in
// this image), so that the `#ifdef DVO_HAVE_OPENCV / DVO_HAVE_EIGEN` overloads of the host classes -- the reference's exact
// cv::Mat / Eigen signatures -- are compiled and executed.  Links against libdvo_host.so.
#include <cstring>

#include "EPoseEstimator.h"
#include "GOP.h"
#include "SolveDVO.h"

#if !defined(DVO_HAVE_OPENCV) || !defined(DVO_HAVE_EIGEN)
#error "compile with -I tests/standin_include (or real OpenCV / Eigen include paths)"
#endif

extern "C" {

// EPoseEstimator through setRefFrame(cv::Mat&, cv::Mat&), setNowFrame(cv::Mat&, cv::Mat&), estimate(Eigen::Matrix3d&, Eigen::Vector3d&),
// then PyramidalStorageStruct::getLevel / addLevel with cv::Mat& / Eigen:: arguments (column-major) round-tripped through a second storage
int cvapi_eposeestimator(uint8_t* ref_bgr, uint16_t* ref_depth, uint8_t* now_bgr, uint16_t* now_depth, int W, int H, double fx, double fy, double cx,
                         double cy, int level, int iters, double huber_k, double lambda0, double* R9, double* T3, double* J_colmajor, int* roundtrip_ok) {
    cv::Mat rgb(H, W, CV_8UC3, ref_bgr), dep(H, W, CV_16UC1, ref_depth), nrgb(H, W, CV_8UC3, now_bgr), ndep(H, W, CV_16UC1, now_depth);
    EPoseEstimator e(false);
    e.setCameraMatrix(fx, fy, cx, cy);
    e.iterations = iters; e.huber_k = huber_k; e.lm_lambda0 = lambda0;
    e.setRefFrame(rgb, dep);
    e.setNowFrame(nrgb, ndep);
    e.setPyramidalImages(level);
    Eigen::Matrix3d R = Eigen::Matrix3d::Identity(); Eigen::Vector3d T = Eigen::Vector3d::Zero();
    e.estimate(R, T);
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R9[3 * r + c] = R(r, c); T3[r] = T(r); }
    cv::Mat c3, g, d; Eigen::ArrayXXd X, Y, Z, gv, rv, gr, bv; Eigen::MatrixXd J;
    e.pydStore.getLevel(level, c3, g, d, X, Y, Z, J, gv, rv, gr, bv);
    std::memcpy(J_colmajor, J.data(), sizeof(double) * (size_t)J.rows() * 6);
    PyramidalStorageStruct st;
    st.addLevel(level, c3, g, d, X, Y, Z, J, gv, rv, gr, bv);
    cv::Mat c3b, gb, db; Eigen::ArrayXXd Xb, Yb, Zb, gvb, rvb, grb, bvb; Eigen::MatrixXd Jb;
    st.getLevel(0, c3b, gb, db, Xb, Yb, Zb, Jb, gvb, rvb, grb, bvb);
    const size_t P = (size_t)g.rows * g.cols;
    *roundtrip_ok = gb.rows == g.rows && gb.cols == g.cols && db.type() == CV_16UC1 && c3b.type() == CV_8UC3 &&
                    !std::memcmp(gb.data, g.data, P) && !std::memcmp(db.data, d.data, P * 2) && !std::memcmp(c3b.data, c3.data, P * 3) &&
                    !std::memcmp(Jb.data(), J.data(), P * 48) && !std::memcmp(Xb.data(), X.data(), P * 8) && !std::memcmp(bvb.data(), bv.data(), P * 8) &&
                    Xb.rows() == g.rows && Xb.cols() == g.cols && Jb.rows() == (long)P && Jb.cols() == 6;
    return e.pydStore.size();
}

// SolveDVO::runIterations with the reference's Eigen signature + GOP<double> fed with Eigen matrices
int cvapi_solvedvo_run_iterations(uint8_t* ref_gray, uint16_t* ref_depth, uint8_t* now_gray, uint16_t* now_depth, int W, int H, int levels, float fx,
                                  float fy, float cx, float cy, int level, int maxIter, double* R9, double* T3, float* energies, float* eps, float* reproj_u,
                                  int* best_index, float* visible_ratio, double* gop19) {
    SolveDVO s(W, H, levels);
    s.setIntrinsics(fx, fy, cx, cy);
    cv::Mat g(H, W, CV_8UC1, ref_gray), d(H, W, CV_16UC1, ref_depth), g2(H, W, CV_8UC1, now_gray), d2(H, W, CV_16UC1, now_depth);
    s.setRcvdFrame(g, d); s.setRcvdFrameAsRefFrame(); s.preProcessRefFrame();          // cv::Mat -> dvo::ImageView implicitly
    s.setRcvdFrame(g2, d2); s.setRcvdFrameAsNowFrame();
    Eigen::Matrix3d cR = Eigen::Matrix3d::Identity(); Eigen::Vector3d cT = Eigen::Vector3d::Zero();
    Eigen::VectorXf en, ep; Eigen::MatrixXf rp; int bi = -1; float vr = 0.f;
    s.runIterations(level, maxIter, cR, cT, en, ep, rp, bi, vr);
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R9[3 * r + c] = cR(r, c); T3[r] = cT(r); }
    for (int k = 0; k < maxIter; ++k) energies[k] = en(k);
    for (int i = 0; i < (int)ep.rows(); ++i) { eps[i] = ep(i); reproj_u[i] = rp(0, i); }
    *best_index = bi; *visible_ratio = vr;
    GOP<double> gop;
    gop.pushAsKeyFrame(0, 1, Eigen::Matrix3d::Identity(), Eigen::Vector3d::Zero());
    gop.pushAsOrdinaryFrame(1, cR, cT);
    Eigen::Matrix3d gR = static_cast<Eigen::Matrix3d>(gop.getGlobalRAt(1)); Eigen::Vector3d gT = static_cast<Eigen::Vector3d>(gop.getGlobalTAt(1));
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) gop19[3 * r + c] = gR(r, c); gop19[9 + r] = gT(r); }
    return (int)ep.rows();
}

}  // extern "C"
