// host_capi.cpp -- C shim so the pytest suite can drive the C++ classes (SolveDVO, EPoseEstimator, PyramidalStorageStruct,
// GOP) through ctypes exactly as a C++ caller of the reference would.
#include <cstring>

#include <fstream>
#include <sstream>

#include "EPoseEstimator.h"
#include "FrameIO.h"
#include "GOP.h"
#include "RGBDOdometry.h"
#include "SolveDVO.h"

extern "C" {

// SolveDVO::loop over one in-memory sequence; out: nframes * 19 doubles (R, T, pose), is_key, reason
int hostapi_solvedvo_sequence(const uint8_t* gray, const uint16_t* depth, int nframes, int W, int H, int levels, float fx, float fy, float cx,
                              float cy, const int* iters, double* out19, int* is_key, int* reason) {
    SolveDVO s(W, H, levels);
    s.setIntrinsics(fx, fy, cx, cy);
    for (int l = 0; l < levels; ++l) s.iterationsConfig[l] = iters[l];
    s.loopFromFrames(gray, depth, nframes);
    if (s.gop.size() != nframes) return -1;
    for (int i = 0; i < nframes; ++i) {
        const dvo::Matrix3d& R = s.gop.getGlobalRAt(i); const dvo::Vector3d& T = s.gop.getGlobalTAt(i); const dvo::Pose& p = s.gop.getGlobalPoseAt(i);
        double* o = out19 + 19 * (size_t)i;
        for (int k = 0; k < 9; ++k) o[k] = R.m[k];
        for (int k = 0; k < 3; ++k) o[9 + k] = T.v[k];
        o[12] = p.position.x; o[13] = p.position.y; o[14] = p.position.z;
        o[15] = p.orientation.x; o[16] = p.orientation.y; o[17] = p.orientation.z; o[18] = p.orientation.w;
        is_key[i] = s.gop.isKeyFrameAt(i); reason[i] = s.gop.getReasonAt(i);
    }
    return 0;
}

// SolveDVO::runIterations at one level of one pair; outputs as the reference's out-parameters
int hostapi_solvedvo_run_iterations(const uint8_t* ref_gray, const uint16_t* ref_depth, const uint8_t* now_gray, const uint16_t* now_depth, int W,
                                    int H, int levels, float fx, float fy, float cx, float cy, int level, int maxIter, double* R9, double* T3,
                                    float* energies, float* eps, float* reproj_u, float* reproj_v, int* best_index, float* visible_ratio,
                                    int* npts, float* b_cap) {
    SolveDVO s(W, H, levels);
    s.setIntrinsics(fx, fy, cx, cy);
    dvo::ImageView g(ref_gray, H, W, dvo::U8C1), d(ref_depth, H, W, dvo::U16C1), g2(now_gray, H, W, dvo::U8C1), d2(now_depth, H, W, dvo::U16C1);
    s.setRcvdFrame(g, d); s.setRcvdFrameAsRefFrame(); s.preProcessRefFrame();
    s.setRcvdFrame(g2, d2); s.setRcvdFrameAsNowFrame();
    dvo::Matrix3d cR; dvo::Vector3d cT;
    std::memcpy(cR.m, R9, 72); std::memcpy(cT.v, T3, 24);
    dvo::VectorXf en, ep; dvo::MatrixXf rp; int bi; float vr;
    s.runIterations(level, maxIter, cR, cT, en, ep, rp, bi, vr);
    std::memcpy(R9, cR.m, 72); std::memcpy(T3, cT.v, 24);
    for (int k = 0; k < maxIter; ++k) energies[k] = en[k];
    *npts = (int)ep.size();
    for (size_t i = 0; i < ep.size(); ++i) { eps[i] = ep[i]; reproj_u[i] = rp(0, (int)i); reproj_v[i] = rp(1, (int)i); }
    *best_index = bi; *visible_ratio = vr; *b_cap = s.processResidueHistogram(ep, true);
    return 0;
}

// EPoseEstimator: setRefFrame / setNowFrame / setPyramidalImages(level) / estimate, plus one PyramidalStorage::getLevel
int hostapi_eposeestimator(const uint8_t* ref_bgr, const uint16_t* ref_depth, const uint8_t* now_bgr, const uint16_t* now_depth, int W, int H,
                           double fx, double fy, double cx, double cy, int compat, int level, int iters, double huber_k, double lambda0, double* R9,
                           double* T3, double* A36, float* visible, int* status, double* J_level, double* X_level, uint8_t* gray_level) {
    EPoseEstimator e(compat != 0);
    e.setCameraMatrix(fx, fy, cx, cy);
    e.iterations = iters; e.huber_k = huber_k; e.lm_lambda0 = lambda0;
    dvo::ImageView rb(ref_bgr, H, W, dvo::U8C3), rd(ref_depth, H, W, dvo::U16C1), nb(now_bgr, H, W, dvo::U8C3), nd(now_depth, H, W, dvo::U16C1);
    e.setRefFrame(rb, rd); e.setNowFrame(nb, nd);
    e.setPyramidalImages(level);
    std::memcpy(A36, e.A, 288);
    dvo::Matrix3d R; dvo::Vector3d T; std::memcpy(R.m, R9, 72); std::memcpy(T.v, T3, 24);
    *visible = e.estimate(R, T);
    std::memcpy(R9, R.m, 72); std::memcpy(T3, T.v, 24);
    *status = e.lastInfo.status;
    if (J_level || X_level || gray_level) {
        std::vector<uint8_t> c, g; std::vector<uint16_t> d; dvo::ArrayXXd X, Y, Z, gv, rv, gr, bv; dvo::MatrixXd J;
        e.pydStore.getLevel(level, c, g, d, X, Y, Z, J, gv, rv, gr, bv);
        if (J_level) std::memcpy(J_level, J.data.data(), J.data.size() * 8);
        if (X_level) std::memcpy(X_level, X.data.data(), X.data.size() * 8);
        if (gray_level) std::memcpy(gray_level, g.data(), g.size());
    }
    return e.pydStore.size();
}

// PyramidalStorageStruct::addLevel with the reference's eleven-argument signature (include/PyramidalStorage.h:42-48).
// (1) an UNBOUND storage is pure storage: what is pushed comes back from getLevel, entry by push order, `level` ignored;
// (2) a storage bound to an estimator that was given reference frame A receives the eleven arrays of reference frame B (taken from
//     a second estimator): estimate() must then behave exactly like an estimator whose reference frame is B.
// Returns 0 when every storage round trip is exact; R9 / T3 / A36 = pose and normal matrix of (2).
int hostapi_pydstore_add_level(const uint8_t* bgrA, const uint16_t* depthA, const uint8_t* bgrB, const uint16_t* depthB, const uint8_t* now_bgr,
                               const uint16_t* now_depth, int W, int H, double fx, double fy, double cx, double cy, int level, int iters,
                               double huber_k, double lambda0, double* R9, double* T3, double* A36, int* sizes) {
    dvo::ImageView a_rgb(bgrA, H, W, dvo::U8C3), a_d(depthA, H, W, dvo::U16C1), b_rgb(bgrB, H, W, dvo::U8C3), b_d(depthB, H, W, dvo::U16C1);
    dvo::ImageView n_rgb(now_bgr, H, W, dvo::U8C3), n_d(now_depth, H, W, dvo::U16C1);
    EPoseEstimator eA(false), eB(false);
    eA.setCameraMatrix(fx, fy, cx, cy); eB.setCameraMatrix(fx, fy, cx, cy);
    eA.iterations = iters; eA.huber_k = huber_k; eA.lm_lambda0 = lambda0;
    eA.setRefFrame(a_rgb, a_d); eA.setNowFrame(n_rgb, n_d);
    eB.setRefFrame(b_rgb, b_d);
    const int nlev = eB.pydStore.size();
    struct Lvl { std::vector<uint8_t> c, g; std::vector<uint16_t> d; dvo::ArrayXXd X, Y, Z, gv, rv, gr, bv; dvo::MatrixXd J; int rows, cols; };
    std::vector<Lvl> lv(nlev);
    for (int l = 0; l < nlev; ++l) { Lvl& q = lv[l]; eB.pydStore.getLevel(l, q.c, q.g, q.d, q.X, q.Y, q.Z, q.J, q.gv, q.rv, q.gr, q.bv); q.rows = H >> l; q.cols = W >> l; }
    int bad = 0;
    auto same = [&](PyramidalStorageStruct& st, int idx, const Lvl& q) {
        Lvl r; st.getLevel(idx, r.c, r.g, r.d, r.X, r.Y, r.Z, r.J, r.gv, r.rv, r.gr, r.bv);
        return r.c == q.c && r.g == q.g && r.d == q.d && r.X.data == q.X.data && r.Y.data == q.Y.data && r.Z.data == q.Z.data && r.J.data == q.J.data &&
               r.gv.data == q.gv.data && r.rv.data == q.rv.data && r.gr.data == q.gr.data && r.bv.data == q.bv.data && r.J.rows == q.J.rows && r.X.rows == q.X.rows;
    };
    {   // (1) unbound: push levels 2 and 1 in that order, with "level" arguments that do not match the push order
        PyramidalStorageStruct st;
        for (int k = 0; k < 2; ++k) {
            Lvl& q = lv[2 - k];
            st.addLevel(k == 0 ? 4 : 0, dvo::ImageView(q.c.data(), q.rows, q.cols, dvo::U8C3), dvo::ImageView(q.g.data(), q.rows, q.cols, dvo::U8C1),
                        dvo::ImageView(q.d.data(), q.rows, q.cols, dvo::U16C1), q.X, q.Y, q.Z, q.J, q.gv, q.rv, q.gr, q.bv);
        }
        sizes[0] = st.size();
        if (!same(st, 0, lv[2]) || !same(st, 1, lv[1])) bad |= 1;
        st.clearPyramid();
        sizes[1] = st.size();
    }
    // (2) bound: replace A's pyramid by B's arrays
    eA.pydStore.clearPyramid();
    for (int l = 0; l < nlev; ++l) {
        Lvl& q = lv[l];
        eA.pydStore.addLevel(l, dvo::ImageView(q.c.data(), q.rows, q.cols, dvo::U8C3), dvo::ImageView(q.g.data(), q.rows, q.cols, dvo::U8C1),
                             dvo::ImageView(q.d.data(), q.rows, q.cols, dvo::U16C1), q.X, q.Y, q.Z, q.J, q.gv, q.rv, q.gr, q.bv);
    }
    sizes[2] = eA.pydStore.size();
    for (int l = 0; l < nlev; ++l) if (!same(eA.pydStore, l, lv[l])) bad |= 2;
    eA.setPyramidalImages(level);
    std::memcpy(A36, eA.A, 288);
    dvo::Matrix3d R; dvo::Vector3d T;
    eA.estimate(R, T);
    std::memcpy(R9, R.m, 72); std::memcpy(T3, T.v, 24);
    return bad;
}

// GOP<float> and GOP<double> replay (explicit instantiations, src/GOP.cpp:244-245)
int hostapi_gop_replay(int n, const int* kind, const int* reason, const double* rel, double* out19, int* is_key, int* reason_out, int use_float) {
    if (use_float) {
        GOP<float> g;
        for (int i = 0; i < n; ++i) {
            dvo::Matrix3f R; dvo::Vector3f T;
            for (int k = 0; k < 9; ++k) R.m[k] = (float)rel[12 * (size_t)i + k];
            for (int k = 0; k < 3; ++k) T.v[k] = (float)rel[12 * (size_t)i + 9 + k];
            if (kind[i] == 1) g.pushAsKeyFrame(i, reason[i], R, T); else g.pushAsOrdinaryFrame(i, R, T);
            if (kind[i] == 2) g.updateMostRecentToKeyFrame(reason[i]);
        }
        for (int i = 0; i < n; ++i) {
            double* o = out19 + 19 * (size_t)i; const dvo::Pose& p = g.getGlobalPoseAt(i);
            for (int k = 0; k < 9; ++k) o[k] = g.getGlobalRAt(i).m[k];
            for (int k = 0; k < 3; ++k) o[9 + k] = g.getGlobalTAt(i).v[k];
            o[12] = p.position.x; o[13] = p.position.y; o[14] = p.position.z; o[15] = p.orientation.x; o[16] = p.orientation.y; o[17] = p.orientation.z; o[18] = p.orientation.w;
            is_key[i] = g.isKeyFrameAt(i); reason_out[i] = g.getReasonAt(i);
        }
        return g.size();
    }
    GOP<double> g;
    for (int i = 0; i < n; ++i) {
        dvo::Matrix3d R; dvo::Vector3d T;
        std::memcpy(R.m, rel + 12 * (size_t)i, 72); std::memcpy(T.v, rel + 12 * (size_t)i + 9, 24);
        if (kind[i] == 1) g.pushAsKeyFrame(i, reason[i], R, T); else g.pushAsOrdinaryFrame(i, R, T);
        if (kind[i] == 2) g.updateMostRecentToKeyFrame(reason[i]);
    }
    for (int i = 0; i < n; ++i) {
        double* o = out19 + 19 * (size_t)i; const dvo::Pose& p = g.getGlobalPoseAt(i);
        std::memcpy(o, g.getGlobalRAt(i).m, 72); std::memcpy(o + 9, g.getGlobalTAt(i).v, 24);
        o[12] = p.position.x; o[13] = p.position.y; o[14] = p.position.z; o[15] = p.orientation.x; o[16] = p.orientation.y; o[17] = p.orientation.z; o[18] = p.orientation.w;
        is_key[i] = g.isKeyFrameAt(i); reason_out[i] = g.getReasonAt(i);
    }
    return g.size();
}

// ---- FrameIO (formats either side of the path) ----
// write `levels` mono (u8) / depth (u16) images, level l of size (H >> l) x (W >> l), packed one after the other
int hostapi_store_frame_xml(const char* file, const uint8_t* mono, const uint16_t* depth, int W, int H, int levels) {
    dvo::RGBDFramePyd p;
    size_t off = 0;
    for (int l = 0; l < levels; ++l) {
        const int w = W >> l, h = H >> l;
        dvo::Image m; m.rows = h; m.cols = w; m.elem = dvo::ELEM_U8; m.data.assign(mono + off, mono + off + (size_t)w * h);
        dvo::Image d; d.rows = h; d.cols = w; d.elem = dvo::ELEM_U16; d.data.resize((size_t)w * h * 2); std::memcpy(d.data.data(), depth + off, (size_t)w * h * 2);
        p.framemono.push_back(m); p.dframe.push_back(d);
        off += (size_t)w * h;
    }
    return dvo::storeFrameXml(file, p) ? 0 : -1;
}
// read them back into the same packed layout; dims[2*l], dims[2*l+1] = rows, cols of level l
int hostapi_load_frame_xml(const char* file, int levels, uint8_t* mono, uint16_t* depth, size_t capacity, int* dims) {
    dvo::RGBDFramePyd p;
    if (!dvo::loadFrameXml(file, p, levels)) return -1;
    size_t off = 0;
    for (int l = 0; l < levels; ++l) {
        const dvo::Image& m = p.framemono[l]; const dvo::Image& d = p.dframe[l];
        if (m.elem != dvo::ELEM_U8 || d.elem != dvo::ELEM_U16 || m.rows != d.rows || m.cols != d.cols) return -2;
        const size_t n = (size_t)m.rows * m.cols;
        if (off + n > capacity) return -3;
        std::memcpy(mono + off, m.data.data(), n); std::memcpy(depth + off, d.data.data(), n * 2);
        dims[2 * l] = m.rows; dims[2 * l + 1] = m.cols;
        off += n;
    }
    return 0;
}
// one generic node (any dt) as doubles, for checking the reader against files written by cv::FileStorage itself
int hostapi_read_xml_matrix(const char* file, const char* name, double* out, size_t capacity, int* rows, int* cols, int* channels, int* elem) {
    std::ifstream f(file, std::ios::binary);
    if (!f.is_open()) return -1;
    std::stringstream ss; ss << f.rdbuf();
    dvo::Image m;
    if (!dvo::readXmlMatrix(ss.str(), name, m)) return -2;
    const size_t n = (size_t)m.rows * m.cols * m.channels;
    if (n > capacity) return -3;
    for (size_t i = 0; i < n; ++i) {
        switch (m.elem) {
            case dvo::ELEM_U8: out[i] = m.data[i]; break;
            case dvo::ELEM_U16: out[i] = reinterpret_cast<const uint16_t*>(m.data.data())[i]; break;
            case dvo::ELEM_S16: out[i] = reinterpret_cast<const int16_t*>(m.data.data())[i]; break;
            case dvo::ELEM_S32: out[i] = reinterpret_cast<const int32_t*>(m.data.data())[i]; break;
            case dvo::ELEM_F32: out[i] = reinterpret_cast<const float*>(m.data.data())[i]; break;
            default: out[i] = reinterpret_cast<const double*>(m.data.data())[i]; break;
        }
    }
    *rows = m.rows; *cols = m.cols; *channels = m.channels; *elem = m.elem;
    return 0;
}
// pose text: write n poses (7 doubles each: qx qy qz qw tx ty tz) with SolveDVO::printPose, read them back
int hostapi_write_pose_file(const char* file, const double* q7, int n) {
    std::ofstream f(file);
    if (!f.is_open()) return -1;
    for (int i = 0; i < n; ++i) {
        dvo::Pose p;
        p.orientation.x = q7[7 * i]; p.orientation.y = q7[7 * i + 1]; p.orientation.z = q7[7 * i + 2]; p.orientation.w = q7[7 * i + 3];
        p.position.x = q7[7 * i + 4]; p.position.y = q7[7 * i + 5]; p.position.z = q7[7 * i + 6];
        dvo::printPose(p, f);
    }
    return 0;
}
static int poses_out(const std::vector<dvo::Pose>& v, double* q7, int capacity) {
    const int n = (int)v.size() < capacity ? (int)v.size() : capacity;
    for (int i = 0; i < n; ++i) {
        q7[7 * i] = v[i].orientation.x; q7[7 * i + 1] = v[i].orientation.y; q7[7 * i + 2] = v[i].orientation.z; q7[7 * i + 3] = v[i].orientation.w;
        q7[7 * i + 4] = v[i].position.x; q7[7 * i + 5] = v[i].position.y; q7[7 * i + 6] = v[i].position.z;
    }
    return (int)v.size();
}
int hostapi_read_pose_file(const char* file, double* q7, int capacity) {
    std::vector<dvo::Pose> v;
    if (!dvo::readPoseFile(file, v)) return -1;
    return poses_out(v, q7, capacity);
}
int hostapi_load_gt_path(const char* file, int skip, double* q7, int capacity) {
    std::vector<dvo::Pose> v;
    if (!dvo::loadGTPath(file, v, skip)) return -1;
    return poses_out(v, q7, capacity);
}
// SolveDVO::loadFromFile + one alignment of two dumps (ref, now): the replay path of src/SolveDVO.cpp:154-190
int hostapi_solvedvo_from_files(const char* ref_xml, const char* now_xml, int W, int H, int levels, float fx, float fy, float cx, float cy,
                                const int* iters, double* R9, double* T3) {
    SolveDVO s(W, H, levels);
    s.setIntrinsics(fx, fy, cx, cy);
    s.iterationsConfig.assign(iters, iters + levels);
    if (!s.loadFromFile(ref_xml)) return -1;
    s.setRcvdFrameAsRefFrame(); s.preProcessRefFrame();
    if (!s.loadFromFile(now_xml)) return -2;
    s.setRcvdFrameAsNowFrame();
    dvo::Matrix3d cR; dvo::Vector3d cT;
    dvo::VectorXf energy, eps; dvo::MatrixXf reproj; int best; float vis;
    for (int l = levels - 1; l >= 0; --l)
        if (s.iterationsConfig[l] > 0) s.runIterations(l, s.iterationsConfig[l], cR, cT, energy, eps, reproj, best, vis);
    for (int i = 0; i < 9; ++i) R9[i] = cR.m[i];
    for (int i = 0; i < 3; ++i) T3[i] = cT.v[i];
    return 0;
}

// SolveDVO::loop() / loopDry() over the XML dumps "<folder>/framemono_%04d.xml", start..end (the reference's
// __DATA_FROM_XML_FILES__ build); out as hostapi_solvedvo_sequence.  Returns the number of frames processed.
int hostapi_solvedvo_loop_xml(const char* folder, int start, int end, int dry, int W, int H, int levels, float fx, float fy, float cx, float cy,
                              const int* iters, double* out19, int* is_key, int* reason, int capacity) {
    SolveDVO s(W, H, levels);
    s.setIntrinsics(fx, fy, cx, cy);
    for (int l = 0; l < levels; ++l) s.iterationsConfig[l] = iters[l];
    s.setXmlFrameSource(folder, start, end);
    if (dry) { s.loopDry(); return (int)s.dryFrames + (s.gop.size() == 0 && s.isNowFrameAvailable ? 0 : 100000); }
    s.loop();
    const int n = s.gop.size() < capacity ? s.gop.size() : capacity;
    for (int i = 0; i < n; ++i) {
        const dvo::Pose& p = s.gop.getGlobalPoseAt(i);
        double* o = out19 + 19 * (size_t)i;
        std::memcpy(o, s.gop.getGlobalRAt(i).m, 72); std::memcpy(o + 9, s.gop.getGlobalTAt(i).v, 24);
        o[12] = p.position.x; o[13] = p.position.y; o[14] = p.position.z;
        o[15] = p.orientation.x; o[16] = p.orientation.y; o[17] = p.orientation.z; o[18] = p.orientation.w;
        is_key[i] = s.gop.isKeyFrameAt(i); reason[i] = s.gop.getReasonAt(i);
    }
    return s.gop.size();
}
// SolveDVO::loop() fed by a callback that hands over in-memory frames, skipping every `skip_every`-th turn without a frame
// (a ros::spinOnce() that delivered nothing): the loop must `continue` on those turns
int hostapi_solvedvo_loop_callback(const uint8_t* gray, const uint16_t* depth, int nframes, int skip_every, int W, int H, int levels, float fx, float fy,
                                   float cx, float cy, const int* iters, double* out19) {
    SolveDVO s(W, H, levels);
    s.setIntrinsics(fx, fy, cx, cy);
    for (int l = 0; l < levels; ++l) s.iterationsConfig[l] = iters[l];
    const size_t P = (size_t)W * H;
    int t = 0, turn = 0;
    s.setFrameSource([&](SolveDVO& self) -> bool {
        if (t >= nframes) return false;
        if (skip_every > 0 && (++turn % skip_every) == 0) return true;            // nothing arrived this turn
        self.setRcvdFrame(dvo::ImageView(gray + P * t, H, W, dvo::U8C1), dvo::ImageView(depth + P * t, H, W, dvo::U16C1));
        ++t;
        return true;
    });
    s.loop();
    for (int i = 0; i < s.gop.size() && i < nframes; ++i) { double* o = out19 + 19 * (size_t)i; std::memcpy(o, s.gop.getGlobalRAt(i).m, 72); std::memcpy(o + 9, s.gop.getGlobalTAt(i).v, 24); }
    return s.gop.size();
}
// SolveDVO::casualTestFunction(): folder/framemono_<ref>.xml vs folder/framemono_<now>.xml, runIterations(0, iters)
int hostapi_solvedvo_casual(const char* folder, int ref_index, int now_index, int iterations, int W, int H, int levels, float fx, float fy, float cx,
                            float cy, double* R9, double* T3, float* energies, int capacity) {
    SolveDVO s(W, H, levels);
    s.setIntrinsics(fx, fy, cx, cy);
    s.casualFolder = folder; s.casualRefIndex = ref_index; s.casualNowIndex = now_index; s.casualIterations = iterations;
    s.casualTestFunction();
    std::memcpy(R9, s.cR_64.m, 72); std::memcpy(T3, s.cT_64.v, 24);
    const int n = (int)s.casualEnergies.size() < capacity ? (int)s.casualEnergies.size() : capacity;
    for (int i = 0; i < n; ++i) energies[i] = s.casualEnergies[i];
    return (int)s.casualEnergies.size();
}
// SolveDVO::loopFromFile(): out = 12 doubles (nR, nT) per processed file
int hostapi_solvedvo_loop_from_file(const char* folder, int start, int end, int iterations, int W, int H, int levels, float fx, float fy, float cx,
                                    float cy, double* out12, int capacity) {
    SolveDVO s(W, H, levels);
    s.setIntrinsics(fx, fy, cx, cy);
    s.fileLoopFolder = folder; s.fileLoopStart = start; s.fileLoopEnd = end; s.fileLoopIterations = iterations;
    s.loopFromFile();
    const int n = (int)s.fileLoopR.size() < capacity ? (int)s.fileLoopR.size() : capacity;
    for (int i = 0; i < n; ++i) { std::memcpy(out12 + 12 * (size_t)i, s.fileLoopR[i].m, 72); std::memcpy(out12 + 12 * (size_t)i + 9, s.fileLoopT[i].v, 24); }
    return (int)s.fileLoopR.size();
}

// RGBDOdometry::eventLoop body over an in-memory sequence of BGR / depth frames: out = nframes * 16 doubles (base * T)
int hostapi_rgbdodometry_sequence(const uint8_t* bgr, const uint16_t* depth, int nframes, int W, int H, double fx, double fy, double cx, double cy,
                                  int ref_every, double* out16, int* npts_l2, double* eps_last_l2) {
    RGBDOdometry o;
    o.setCameraMatrix(fx, fy, cx, cy);
    const size_t P = (size_t)W * H;
    for (int t = 0; t < nframes; ++t) {
        o.setRcvdFrame(dvo::ImageView(bgr + 3 * P * t, H, W, dvo::U8C3), dvo::ImageView(depth + P * t, H, W, dvo::U16C1));
        const TransformRep g = o.processFrame(ref_every);
        std::memcpy(out16 + 16 * (size_t)t, g.m, sizeof(g.m));
        if (npts_l2) npts_l2[t] = o.lastInfo.npts[2];
        if (eps_last_l2) eps_last_l2[t] = o.lastInfo.eps_norm_last[2];
    }
    return 0;
}
// computeJacobian / computeEpsilon materialised through the class (level, T) for one pair
int hostapi_rgbdodometry_materialise(const uint8_t* ref_bgr, const uint16_t* ref_depth, const uint8_t* now_bgr, const uint16_t* now_depth, int W, int H,
                                     double fx, double fy, double cx, double cy, int level, const double* T16, double* J, int* marks, double* eps,
                                     int* newroi, int capacity) {
    RGBDOdometry o;
    o.setCameraMatrix(fx, fy, cx, cy);
    const dvo::ImageView rf(ref_bgr, H, W, dvo::U8C3), rd(ref_depth, H, W, dvo::U16C1), nf(now_bgr, H, W, dvo::U8C3), nd(now_depth, H, W, dvo::U16C1);
    o.setRcvdFrame(rf, rd); o.setRefFrame(rf, rd); o.computeJacobianAllLevels();
    o.setRcvdFrame(nf, nd); o.setNowFrame(nf, nd);
    dvo::MatrixXd Jm; MatrixXi mk;
    o.computeJacobian(level, Jm, mk);
    if (Jm.rows > capacity) return -1;
    std::memcpy(J, Jm.data.data(), sizeof(double) * Jm.data.size());
    std::memcpy(marks, mk.data.data(), sizeof(int) * mk.data.size());
    TransformRep T; std::memcpy(T.m, T16, sizeof(T.m));
    std::vector<double> e; MatrixXi roi;
    o.computeEpsilon(level, T, e, roi);
    std::memcpy(eps, e.data(), sizeof(double) * e.size());
    std::memcpy(newroi, roi.data.data(), sizeof(int) * roi.data.size());
    return Jm.rows;
}

}  // extern "C"
