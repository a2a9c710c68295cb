// C shim over dvo_oracle.hpp so tests / bench can call the oracle through ctypes.
// TEST INFRASTRUCTURE ONLY -- see the header of dvo_oracle.hpp ("parity unpinned" by the reference's own tests).
#include "dvo_oracle.hpp"

#include <chrono>
#include <thread>

using namespace orc;

extern "C" {

struct orc_solver_cfg {
    int solver;        // 0 SUBGRAD_REF, 1 GN, 2 LM
    int jacobian;      // 0 REFERENCE, 1 EXACT
    int weight;        // 0 REF_CAUCHY, 1 HUBER, 2 NONE
    float huber_k;
    double lm_lambda0;
    int residual;      // 0 DT at floor (shipped), 1 interpolated DT (compiled-out reference variant)
};

static SolverConfig to_cfg(const orc_solver_cfg* c) {
    SolverConfig s;
    if (c) {
        s.solver = (SolverMode)c->solver; s.jac = (JacobianMode)c->jacobian; s.weight = (WeightMode)c->weight;
        s.huber_k = c->huber_k; s.lm_lambda0 = c->lm_lambda0; s.residual = (ResidualMode)c->residual;
    }
    return s;
}

int orc_level_dim(int dim, int level) { return level_dim(dim, level); }
void orc_pyr_nearest_u8(const uint8_t* s, int W, int H, int level, uint8_t* d, int ch) { pyr_nearest(s, W, H, level, d, ch); }
void orc_pyr_nearest_u16(const uint16_t* s, int W, int H, int level, uint16_t* d) { pyr_nearest(s, W, H, level, d, 1); }
void orc_pyr_area_u8(const uint8_t* s, int W, int H, int level, uint8_t* d, int ch) { pyr_area(s, W, H, level, d, ch); }
void orc_pyr_area_u16(const uint16_t* s, int W, int H, int level, uint16_t* d) { pyr_area(s, W, H, level, d, 1); }
void orc_bgr2gray(const uint8_t* bgr, size_t n, uint8_t* g) { bgr2gray(bgr, n, g); }
void orc_depth_m_to_mm(const float* m, size_t n, uint16_t* mm) { depth_m_to_mm(m, n, mm); }

void orc_canny(const uint8_t* img, int W, int H, uint8_t* edge) { canny(img, W, H, edge); }
size_t orc_edt_d2(const uint8_t* edge, int W, int H, int32_t* d2) { return edt_d2(edge, W, H, d2); }
float orc_dt_normalize(const int32_t* d2, size_t n, float* dtn) { float s; dt_normalize(d2, n, dtn, &s); return s; }
void orc_gradient(const float* img, int W, int H, float* gx, float* gy) { gradient(img, W, H, gx, gy); }

// Brute-force EDT for validating edt_d2 on small images.
void orc_edt_d2_brute(const uint8_t* edge, int W, int H, int32_t* d2) {
    std::vector<int> ex, ey;
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) if (edge[(size_t)y * W + x]) { ex.push_back(x); ey.push_back(y); }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            long long best = EDT_INF;
            for (size_t k = 0; k < ex.size(); ++k) {
                long long dx = x - ex[k], dy = y - ey[k], d = dx * dx + dy * dy;
                if (d < best) best = d;
            }
            d2[(size_t)y * W + x] = (int32_t)best;
        }
}

int orc_select_points(const uint8_t* edge, const uint16_t* depth, int W, int H, int level, float fx, float fy,
                      float cx, float cy, float* X, float* Y, float* Z, float* u, float* v, int capacity) {
    PointList p; select_points(edge, depth, W, H, level, Intrinsics{fx, fy, cx, cy}, p);
    int n = (int)p.size();
    for (int i = 0; i < n && i < capacity; ++i) {
        if (X) X[i] = p.X[i]; if (Y) Y[i] = p.Y[i]; if (Z) Z[i] = p.Z[i]; if (u) u[i] = p.u[i]; if (v) v[i] = p.v[i];
    }
    return n;
}

// One evaluation of the normal equations at pose (R,T).  H36 row-major.  Per-point outputs optional.
void orc_evaluate(const float* X, const float* Y, const float* Z, int N, const float* dtn, const float* gx,
                  const float* gy, int W, int H, int level, float fx, float fy, float cx, float cy, const double* R,
                  const double* T, int jac, int weight, float huber_k, double* g6, double* H36, double* sumsq,
                  int* nvis, float* eps, float* w, float* ru, float* rv, float* J, int residual) {
    PointList p; p.X.assign(X, X + N); p.Y.assign(Y, Y + N); p.Z.assign(Z, Z + N);
    LevelImages li; li.W = W; li.H = H; size_t P = (size_t)W * H;
    li.dtn.assign(dtn, dtn + P); li.gx.assign(gx, gx + P); li.gy.assign(gy, gy + P);
    EvalOut ev;
    bool pp = eps || w || ru || rv || J;
    evaluate(p, li, level, Intrinsics{fx, fy, cx, cy}, R, T, (JacobianMode)jac, (WeightMode)weight, huber_k, pp, ev, (ResidualMode)residual);
    std::memcpy(g6, ev.g, sizeof(ev.g)); std::memcpy(H36, ev.H, sizeof(ev.H)); *sumsq = ev.sumsq; *nvis = ev.nvis;
    if (eps) std::memcpy(eps, ev.eps.data(), sizeof(float) * N);
    if (w) std::memcpy(w, ev.w.data(), sizeof(float) * N);
    if (ru) std::memcpy(ru, ev.reproj_u.data(), sizeof(float) * N);
    if (rv) std::memcpy(rv, ev.reproj_v.data(), sizeof(float) * N);
    if (J) std::memcpy(J, ev.J.data(), sizeof(float) * 6 * N);
}

// runIterations at one level on explicit buffers.  R9/T3 in-out.  energies has maxIter entries.
// trace (optional): per executed iteration 6 g + 36 H + energy + nvis + 9 R + 3 T = 56 doubles.
int orc_run_iterations(const float* X, const float* Y, const float* Z, int N, const float* dtn, const float* gx,
                       const float* gy, int W, int H, int level, int maxIter, float fx, float fy, float cx, float cy,
                       const orc_solver_cfg* cfg, double* R9, double* T3, float* energies, int* best_index,
                       float* visible_ratio, int* iterations_run, float* best_eps, float* best_u, float* best_v,
                       double* trace) {
    PointList p; p.X.assign(X, X + N); p.Y.assign(Y, Y + N); p.Z.assign(Z, Z + N);
    LevelImages li; li.W = W; li.H = H; size_t P = (size_t)W * H;
    li.dtn.assign(dtn, dtn + P); li.gx.assign(gx, gx + P); li.gy.assign(gy, gy + P);
    LevelResult res;
    run_iterations(p, li, level, maxIter, Intrinsics{fx, fy, cx, cy}, to_cfg(cfg), R9, T3, res, trace != nullptr, true);
    if (energies) std::memcpy(energies, res.energies.data(), sizeof(float) * maxIter);
    if (best_index) *best_index = res.best_index;
    if (visible_ratio) *visible_ratio = res.visible_ratio;
    if (iterations_run) *iterations_run = res.iterations_run;
    if (best_eps && !res.best_eps.empty()) std::memcpy(best_eps, res.best_eps.data(), sizeof(float) * N);
    if (best_u && !res.best_u.empty()) std::memcpy(best_u, res.best_u.data(), sizeof(float) * N);
    if (best_v && !res.best_v.empty()) std::memcpy(best_v, res.best_v.data(), sizeof(float) * N);
    if (trace)
        for (size_t k = 0; k < res.trace.size(); ++k) {
            double* t = trace + 56 * k; const IterTrace& it = res.trace[k];
            std::memcpy(t, it.g, 48); std::memcpy(t + 6, it.H, 288); t[42] = it.energy; t[43] = it.nvis;
            std::memcpy(t + 44, it.R, 72); std::memcpy(t + 53, it.T, 24);
        }
    return 0;
}

// Full pipeline for one pair (SolveDVO edge path): NEAREST pyramids, Canny, EDT, normalise, gradient, point list,
// levels (L-1)..0.  Outputs: R9,T3; per level (arrays of `levels`): npts, best_index, iterations_run, best_energy,
// visible_ratio.  trace (optional): levels * max_iter * 56 doubles, level-major, zero where not executed.
int orc_align_pair(const uint8_t* ref_gray, const uint16_t* ref_depth, const uint8_t* now_gray, int W, int H,
                   int levels, float fx, float fy, float cx, float cy, const int* iters, const orc_solver_cfg* cfg,
                   const double* R0, const double* T0, double* R9, double* T3, int* npts, int* best_index,
                   int* iterations_run, float* best_energy, float* visible_ratio, double* trace, int trace_max_iter, float* b_cap) {
    PairResult pr;
    align_pair(ref_gray, ref_depth, now_gray, W, H, levels, Intrinsics{fx, fy, cx, cy}, iters, to_cfg(cfg), R0, T0, pr,
               trace != nullptr, b_cap != nullptr);
    if (b_cap) {          // processResidueHistogram of the last runIterations call, i.e. the finest level run (src/SolveDVO.cpp:2117-2120)
        *b_cap = 0.f;
        for (int l = 0; l < levels; ++l) if (iters[l] > 0) { *b_cap = laplacian_b(pr.levels[l].best_eps); break; }
    }
    std::memcpy(R9, pr.R, sizeof(pr.R)); std::memcpy(T3, pr.T, sizeof(pr.T));
    for (int l = 0; l < levels; ++l) {
        if (npts) npts[l] = (int)pr.npts[l];
        if (best_index) best_index[l] = pr.levels[l].best_index;
        if (iterations_run) iterations_run[l] = pr.levels[l].iterations_run;
        if (best_energy) best_energy[l] = pr.levels[l].best_energy;
        if (visible_ratio) visible_ratio[l] = pr.levels[l].visible_ratio;
        if (trace)
            for (size_t k = 0; k < pr.levels[l].trace.size() && (int)k < trace_max_iter; ++k) {
                double* t = trace + 56 * ((size_t)l * trace_max_iter + k); const IterTrace& it = pr.levels[l].trace[k];
                std::memcpy(t, it.g, 48); std::memcpy(t + 6, it.H, 288); t[42] = it.energy; t[43] = it.nvis;
                std::memcpy(t + 44, it.R, 72); std::memcpy(t + 53, it.T, 24);
            }
    }
    return pr.status;
}

// Batch of independent pairs on `nthreads` host threads (the CPU baseline of SURVEY §8d).  Buffers pair-major.
// poses: count * 12 doubles (R row-major, then T).  Returns wall seconds.
double orc_align_batch(const uint8_t* ref_gray, const uint16_t* ref_depth, const uint8_t* now_gray, int count, int W,
                       int H, int levels, float fx, float fy, float cx, float cy, const int* iters,
                       const orc_solver_cfg* cfg, double* poses, int* status, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    const size_t P = (size_t)W * H;
    SolverConfig sc = to_cfg(cfg);
    auto t0 = std::chrono::steady_clock::now();
    auto work = [&](int tid) {
        for (int i = tid; i < count; i += nthreads) {
            PairResult pr;
            align_pair(ref_gray + P * i, ref_depth + P * i, now_gray + P * i, W, H, levels, Intrinsics{fx, fy, cx, cy}, iters,
                       sc, nullptr, nullptr, pr);
            if (poses) { std::memcpy(poses + 12 * (size_t)i, pr.R, 72); std::memcpy(poses + 12 * (size_t)i + 9, pr.T, 24); }
            if (status) status[i] = pr.status;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Preprocess one frame for all levels; copies out level `level`'s buffers (any pointer may be NULL).
void orc_preprocess_level(const uint8_t* gray, const uint16_t* depth, int W, int H, int level, uint8_t* gray_l,
                          uint16_t* depth_l, uint8_t* edge, int32_t* d2, float* dtn, float* gx, float* gy) {
    int w = level_dim(W, level), h = level_dim(H, level); size_t P = (size_t)w * h;
    std::vector<uint8_t> g(P), e(P);
    pyr_nearest(gray, W, H, level, g.data());
    if (gray_l) std::memcpy(gray_l, g.data(), P);
    if (depth && depth_l) { pyr_nearest(depth, W, H, level, depth_l); depth_fix_zero(depth_l, P); }
    canny(g.data(), w, h, e.data());
    if (edge) std::memcpy(edge, e.data(), P);
    if (d2 || dtn || gx || gy) {
        std::vector<int32_t> dd(P); edt_d2(e.data(), w, h, dd.data());
        if (d2) std::memcpy(d2, dd.data(), P * 4);
        if (dtn || gx || gy) {
            std::vector<float> n(P), a(P), b(P);
            dt_normalize(dd.data(), P, n.data()); gradient(n.data(), w, h, a.data(), b.data());
            if (dtn) std::memcpy(dtn, n.data(), P * 4); if (gx) std::memcpy(gx, a.data(), P * 4); if (gy) std::memcpy(gy, b.data(), P * 4);
        }
    }
}

void orc_se3_exp(const double* psi, double* R9, double* t3) { se3_exp(psi, R9, t3); }
void orc_se3_log(const double* R9, const double* t3, double* psi) { se3_log(R9, t3, psi); }
void orc_rotationize(double* R9) { rotationize(R9); }
void orc_rot_to_quat(const double* R9, double* q4) { rot_to_quat(R9, q4); }
int orc_chol6_solve(const double* H36, const double* b6, double* x6) { return chol6_solve(H36, b6, x6) ? 0 : -1; }
float orc_weight_ref(float r) { return weight_ref(r); }
float orc_laplacian_b(const float* eps, int n) { return laplacian_b(std::vector<float>(eps, eps + n)); }

// GOP replay: n frames; kind[i]: 0 ordinary, 1 keyframe, 2 ordinary followed by updateMostRecentToKeyFrame.
// rel: n*12 (R,T relative to last keyframe); out: n*(12+7) doubles (R,T,pose px py pz qx qy qz qw); key/reason out.
void orc_gop_replay(int n, const int* kind, const int* reason, const double* rel, double* out, int* is_key, int* reason_out) {
    Gop g;
    for (int i = 0; i < n; ++i) {
        g.push(i, kind[i] == 1, reason[i], rel + 12 * (size_t)i, rel + 12 * (size_t)i + 9);
        if (kind[i] == 2) g.update_most_recent_to_key(reason[i]);
    }
    for (int i = 0; i < n; ++i) {
        std::memcpy(out + 19 * (size_t)i, g.v[i].R, 72); std::memcpy(out + 19 * (size_t)i + 9, g.v[i].T, 24);
        std::memcpy(out + 19 * (size_t)i + 12, g.v[i].pose, 56);
        is_key[i] = g.v[i].key; reason_out[i] = g.v[i].key ? g.v[i].reason : -1;
    }
}


// ---- EPoseEstimator (photometric) ----
static photo::RefLevel g_ref_level;   // scratch shared by the calls below (tests are single threaded)

int orc_photo_build_ref_level(const uint8_t* bgr, const uint16_t* depth, int W, int H, int level, const double* K4, int compat,
                              uint8_t* gray, uint8_t* bgr_l, uint16_t* depth_l, double* X, double* Y, double* Z, double* gx, double* gy,
                              double* J, double* A36) {
    photo::RefLevel& L = g_ref_level;
    photo::build_ref_level(bgr, depth, W, H, level, photo::Cam{K4[0], K4[1], K4[2], K4[3]}, compat != 0, L);
    const size_t N = (size_t)L.rows * L.cols;
    if (gray) std::memcpy(gray, L.gray.data(), N);
    if (bgr_l) std::memcpy(bgr_l, L.bgr.data(), N * 3);
    if (depth_l) std::memcpy(depth_l, L.depth.data(), N * 2);
    if (X) std::memcpy(X, L.X.data(), N * 8); if (Y) std::memcpy(Y, L.Y.data(), N * 8); if (Z) std::memcpy(Z, L.Z.data(), N * 8);
    if (gx) std::memcpy(gx, L.gx.data(), N * 8); if (gy) std::memcpy(gy, L.gy.data(), N * 8);
    if (J) std::memcpy(J, L.J.data(), N * 48);
    if (A36) std::memcpy(A36, L.A, 288);
    return (int)N;
}

void orc_photo_now_level(const uint8_t* bgr, int W, int H, int level, uint8_t* gray_l) {
    std::vector<uint8_t> g; photo::build_level_images(bgr, nullptr, W, H, level, g, nullptr, nullptr);
    std::memcpy(gray_l, g.data(), g.size());
}

// uses the level built by the last orc_photo_build_ref_level call
void orc_photo_evaluate(const uint8_t* now_gray_l, const double* K4, const double* Tr16, int compat, double huber_k, double* b6, double* A36,
                        double* sumsq, int* nreproj, int* nused, double* canvas) {
    photo::IterOut o; std::vector<double> cv;
    photo::evaluate(g_ref_level, now_gray_l, photo::Cam{K4[0], K4[1], K4[2], K4[3]}, Tr16, compat != 0, huber_k, o, canvas ? &cv : nullptr);
    std::memcpy(b6, o.b, 48); std::memcpy(A36, o.A, 288); *sumsq = o.sumsq; *nreproj = o.nreproj; *nused = o.nused;
    if (canvas) std::memcpy(canvas, cv.data(), cv.size() * 8);
}

void orc_photo_estimate(const uint8_t* now_gray_l, const double* K4, const double* R9, const double* T3, int iters, int compat, double huber_k,
                        double lambda0, double* Rout, double* Tout, double* sumsq_first, double* sumsq_last, int* iters_run, int* status,
                        double* visible) {
    photo::EstimateOut o;
    photo::estimate(g_ref_level, now_gray_l, photo::Cam{K4[0], K4[1], K4[2], K4[3]}, R9, T3, iters, compat != 0, huber_k, lambda0, o);
    std::memcpy(Rout, o.R, 72); std::memcpy(Tout, o.T, 24); *sumsq_first = o.sumsq_first; *sumsq_last = o.sumsq_last; *iters_run = o.iters_run;
    *status = o.status; *visible = o.visible;
}

void orc_photo_exp_map(const double* psi, int compat, double* T16) { photo::exp_map(psi, compat != 0, T16); }

// ---- RGBDOdometry (semi-dense photometric GN, src/RGBDOdometry.cpp) ----
void orc_rgbd_level(const uint8_t* bgr, const uint16_t* depth, int W, int H, int level, uint8_t* gray_l, uint16_t* depth_l) {
    std::vector<rgbd::Level> pyr; rgbd::build_pyramid(bgr, depth, W, H, level + 1, pyr);
    std::memcpy(gray_l, pyr[level].gray.data(), pyr[level].gray.size());
    std::memcpy(depth_l, pyr[level].depth.data(), pyr[level].depth.size() * 2);
}
// computeJacobian at one level; returns the number of selected points (J: cap x 6, ij: cap x 2 as (row, col))
int orc_rgbd_jacobian(const uint8_t* bgr, const uint16_t* depth, int W, int H, int level, const double* K4, int thresh, int maxJ, int minPts,
                      double* J, int* ij, int capacity, double* A36, int* status) {
    std::vector<rgbd::Level> pyr; rgbd::build_pyramid(bgr, depth, W, H, level + 1, pyr);
    rgbd::RefJacobian jac; rgbd::compute_jacobian(pyr[level], rgbd::Cam{K4[0], K4[1], K4[2], K4[3]}, thresh, maxJ, minPts, jac);
    const int n = (int)jac.pi.size();
    for (int k = 0; k < n && k < capacity; ++k) { if (ij) { ij[2 * k] = jac.pi[k]; ij[2 * k + 1] = jac.pj[k]; } if (J) std::memcpy(J + 6 * (size_t)k, &jac.J[6 * (size_t)k], 48); }
    if (A36) std::memcpy(A36, jac.A, sizeof(jac.A));
    if (status) *status = jac.status;
    return n;
}
// computeEpsilon at one level and pose T (row-major 4x4)
int orc_rgbd_epsilon(const uint8_t* ref_bgr, const uint16_t* ref_depth, const uint8_t* now_bgr, const uint16_t* now_depth, int W, int H, int level,
                     const double* K4, int thresh, const double* T16, double* eps, int* uv, int capacity, double* b6, double* sumsq, int* nvis) {
    std::vector<rgbd::Level> rp, np; rgbd::build_pyramid(ref_bgr, ref_depth, W, H, level + 1, rp); rgbd::build_pyramid(now_bgr, now_depth, W, H, level + 1, np);
    const rgbd::Cam K{K4[0], K4[1], K4[2], K4[3]};
    rgbd::RefJacobian jac; rgbd::compute_jacobian(rp[level], K, thresh, 1 << 30, 0, jac);
    rgbd::EpsOut e; rgbd::compute_epsilon(rp[level], jac, np[level], K, T16, e);
    const int n = (int)jac.pi.size();
    for (int k = 0; k < n && k < capacity; ++k) { if (eps) eps[k] = e.eps[k]; if (uv) { uv[2 * k] = e.u[k]; uv[2 * k + 1] = e.v[k]; } }
    if (b6) std::memcpy(b6, e.b, sizeof(e.b));
    if (sumsq) *sumsq = e.sumsq;
    if (nvis) *nvis = e.nvis;
    return n;
}
void orc_rgbd_exponential_map(const double* psi6, double* out16) { rgbd::exponential_map(psi6, out16); }
void orc_rgbd_qr_solve6(const double* A36, const double* b6, double* x6) { rgbd::colpiv_qr_solve6(A36, b6, x6); }
// eventLoop body (:147-160): gaussNewtonIterations over `levels[]` in order, T carried; info: per level npts, iters_run, updates, nvis_last, eps_first, eps_last
void orc_rgbd_gauss_newton(const uint8_t* ref_bgr, const uint16_t* ref_depth, const uint8_t* now_bgr, const uint16_t* now_depth, int W, int H,
                           const double* K4, int thresh, const int* levels, int nlevels, int iters, double eps_exit, double* T16, double* info6) {
    int maxl = 0; for (int k = 0; k < nlevels; ++k) if (levels[k] > maxl) maxl = levels[k];
    std::vector<rgbd::Level> rp, np; rgbd::build_pyramid(ref_bgr, ref_depth, W, H, maxl + 1, rp); rgbd::build_pyramid(now_bgr, now_depth, W, H, maxl + 1, np);
    const rgbd::Cam K{K4[0], K4[1], K4[2], K4[3]};
    for (int k = 0; k < nlevels; ++k) {
        rgbd::RefJacobian jac; rgbd::compute_jacobian(rp[levels[k]], K, thresh, 1 << 30, 0, jac);
        rgbd::GnInfo gi; rgbd::gauss_newton(rp[levels[k]], jac, np[levels[k]], K, T16, iters, eps_exit, gi);
        if (info6) { double* o = info6 + 6 * k; o[0] = gi.npts; o[1] = gi.iters_run; o[2] = gi.updates; o[3] = gi.nvis_last; o[4] = gi.eps_norm_first; o[5] = gi.eps_norm_last; }
    }
}

// cv::undistort restatement (publisher front-end): type 0 = u8 (cn channels), 1 = u16 (1 channel)
void orc_undistort(const void* src, int W, int H, int cn, int type, const double* K4, const double* D5, void* dst) {
    if (type == 0) undistort<uint8_t>((const uint8_t*)src, W, H, cn, K4, D5, (uint8_t*)dst);
    else undistort<uint16_t>((const uint16_t*)src, W, H, cn, K4, D5, (uint16_t*)dst);
}

}  // extern "C"
