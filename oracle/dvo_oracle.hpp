// dvo_oracle.hpp -- CPU restatement of the reference's direct-alignment hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under rgbd_odometry_b200/ (the product) includes, links or calls this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// PARITY STATUS: "parity unpinned" by the reference's own tests -- mpkuse/rgbd_odometry ships no tests, no
// golden vectors and no data files, and it cannot be compiled here (needs ROS, Eigen, OpenCV
// C++, Sophus, libigl).  This file restates the reference's arithmetic line by line, each
// function citing the reference file:line it follows (paths relative to the reference checkout).  The
// OpenCV-defined stages (Canny, distanceTransform, normalize, filter2D, resize, cvtColor) are pinned against
// cv2 4.13 by tests/test_oracle_vs_cv2.py and by the committed fixtures under tests/golden/.
//
// Build: g++ -O3 -ffp-contract=off (no -ffast-math): every float expression below is evaluated exactly in
// the written order with one IEEE rounding per operation, which is what the CUDA "exact" arithmetic mirrors.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------------------
// Pyramid construction (R1, R12)
// ------------------------------------------------------------------------------------------------------

// cv::resize(..., Size(), s, s, ...) output size: cvRound(dim * s), round-half-to-even
// (reference call sites: src/camTopic2PublisherPyD.cpp:344-345).
inline int level_dim(int dim, int level) { return (int)std::nearbyint((double)dim * std::ldexp(1.0, -level)); }

// INTER_NEAREST from full resolution: dst(y,x) = src(min(y*2^k, H-1), min(x*2^k, W-1))
// (src/camTopic2PublisherPyD.cpp:338-345, src/publisherPyD.cpp:230-239).
template <typename T>
void pyr_nearest(const T* src, int W, int H, int level, T* dst, int channels = 1) {
    const int w = level_dim(W, level), h = level_dim(H, level), s = 1 << level;
    for (int y = 0; y < h; ++y) {
        int sy = y * s; if (sy > H - 1) sy = H - 1;
        for (int x = 0; x < w; ++x) {
            int sx = x * s; if (sx > W - 1) sx = W - 1;
            for (int c = 0; c < channels; ++c)
                dst[((size_t)y * w + x) * channels + c] = src[((size_t)sy * W + sx) * channels + c];
        }
    }
}

// INTER_AREA from full resolution with integer scale 2^k (src/EPoseEstimator.cpp:251-253, 284-286):
// k=1 -> (sum4 + 2) >> 2 ; k>=2 -> rint_half_even((float)sum * (float)(1/area))  (SURVEY Appendix B.5).
// Requires W, H divisible by 2^k (true for every BASELINE config level that EPoseEstimator builds).
template <typename T>
void pyr_area(const T* src, int W, int H, int level, T* dst, int channels = 1) {
    const int s = 1 << level, w = W / s, h = H / s;
    if (level == 0) { std::memcpy(dst, src, sizeof(T) * (size_t)W * H * channels); return; }
    const float inv_area = (float)(1.0 / (double)(s * s));
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            for (int c = 0; c < channels; ++c) {
                int sum = 0;
                for (int dy = 0; dy < s; ++dy)
                    for (int dx = 0; dx < s; ++dx)
                        sum += (int)src[((size_t)(y * s + dy) * W + (x * s + dx)) * channels + c];
                T out;
                if (level == 1) out = (T)((sum + 2) >> 2);
                else out = (T)(int)std::nearbyintf((float)sum * inv_area);
                dst[((size_t)y * w + x) * channels + c] = out;
            }
}

// cvtColor BGR2GRAY, u8, OpenCV 4.x 15-bit fixed point (src/camTopic2PublisherPyD.cpp:347,
// src/EPoseEstimator.cpp:71,120).
inline void bgr2gray(const uint8_t* bgr, size_t npix, uint8_t* gray) {
    for (size_t i = 0; i < npix; ++i)
        gray[i] = (uint8_t)((bgr[3 * i] * 3735 + bgr[3 * i + 1] * 19235 + bgr[3 * i + 2] * 9798 + (1 << 14)) >> 15);
}

// Depth ingest: metres f32 -> millimetres u16 (src/camTopic2PublisherPyD.cpp:72-80): fp32 multiply by 1000,
// saturate_cast<ushort>(cvRound) (half-even), then 0 -> 1.
inline void depth_m_to_mm(const float* m, size_t n, uint16_t* mm) {
    for (size_t i = 0; i < n; ++i) {
        float v = m[i] * 1000.0f;
        int r = std::isfinite(v) ? (int)std::nearbyintf(v > 1e9f ? 1e9f : (v < -1e9f ? -1e9f : v)) : 0;
        r = r < 0 ? 0 : (r > 65535 ? 65535 : r);
        mm[i] = (uint16_t)(r == 0 ? 1 : r);
    }
}
// SolveDVO ingest: dframe.setTo(1, dframe==0) (src/SolveDVO.cpp:512).
inline void depth_fix_zero(uint16_t* d, size_t n) { for (size_t i = 0; i < n; ++i) if (d[i] == 0) d[i] = 1; }

// ------------------------------------------------------------------------------------------------------
// Canny (R3; cv::Canny(img, 150, 100, 3, true) at src/SolveDVO.cpp:1705,1767)
// ------------------------------------------------------------------------------------------------------
inline void canny(const uint8_t* img, int W, int H, uint8_t* edge, double t1 = 150.0, double t2 = 100.0) {
    double lo_d = t1 < t2 ? t1 : t2, hi_d = t1 < t2 ? t2 : t1;
    lo_d = lo_d > 32767.0 ? 32767.0 : lo_d; hi_d = hi_d > 32767.0 ? 32767.0 : hi_d;
    if (lo_d > 0) lo_d *= lo_d; if (hi_d > 0) hi_d *= hi_d;           // L2gradient: thresholds are squared
    const int low = (int)std::floor(lo_d), high = (int)std::floor(hi_d);
    const size_t P = (size_t)W * H;
    std::vector<int> dx(P), dy(P);
    // magnitude with a 1-pixel zero border
    const int MW = W + 2;
    std::vector<int> mag((size_t)MW * (H + 2), 0);
    auto px = [&](int y, int x) -> int {                               // BORDER_REPLICATE
        y = y < 0 ? 0 : (y >= H ? H - 1 : y); x = x < 0 ? 0 : (x >= W ? W - 1 : x);
        return (int)img[(size_t)y * W + x];
    };
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int gx = (px(y - 1, x + 1) + 2 * px(y, x + 1) + px(y + 1, x + 1)) -
                     (px(y - 1, x - 1) + 2 * px(y, x - 1) + px(y + 1, x - 1));
            int gy = (px(y + 1, x - 1) + 2 * px(y + 1, x) + px(y + 1, x + 1)) -
                     (px(y - 1, x - 1) + 2 * px(y - 1, x) + px(y - 1, x + 1));
            dx[(size_t)y * W + x] = gx; dy[(size_t)y * W + x] = gy;
            mag[(size_t)(y + 1) * MW + (x + 1)] = gx * gx + gy * gy;
        }
    // non-maximum suppression: 0 = none, 1 = weak candidate, 2 = strong
    const int CANNY_SHIFT = 15;
    const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << CANNY_SHIFT) + 0.5);
    std::vector<uint8_t> cls(P, 0);
    std::vector<int> stack;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int* m0 = &mag[(size_t)(y + 1) * MW + (x + 1)];
            const int m = *m0;
            if (m <= low) continue;
            int xs = dx[(size_t)y * W + x], ys = dy[(size_t)y * W + x];
            int ax = std::abs(xs), ay = std::abs(ys) << CANNY_SHIFT;
            int tg22x = ax * TG22;
            bool keep;
            if (ay < tg22x) keep = (m > m0[-1]) && (m >= m0[1]);
            else {
                int tg67x = tg22x + (ax << (CANNY_SHIFT + 1));
                if (ay > tg67x) keep = (m > m0[-MW]) && (m >= m0[MW]);
                else { int s = ((xs ^ ys) < 0) ? -1 : 1; keep = (m > m0[-MW - s]) && (m > m0[MW + s]); }
            }
            if (!keep) continue;
            if (m > high) { cls[(size_t)y * W + x] = 2; stack.push_back(y * W + x); }
            else cls[(size_t)y * W + x] = 1;
        }
    // hysteresis: grow strong pixels through 8-connected weak candidates
    while (!stack.empty()) {
        int p = stack.back(); stack.pop_back();
        int y = p / W, x = p % W;
        for (int dy2 = -1; dy2 <= 1; ++dy2)
            for (int dx2 = -1; dx2 <= 1; ++dx2) {
                int yy = y + dy2, xx = x + dx2;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                size_t q = (size_t)yy * W + xx;
                if (cls[q] == 1) { cls[q] = 2; stack.push_back((int)q); }
            }
    }
    for (size_t i = 0; i < P; ++i) edge[i] = cls[i] == 2 ? 255 : 0;
}

// ------------------------------------------------------------------------------------------------------
// Exact squared Euclidean distance transform (R3; cv::distanceTransform(255-E, L2, PRECISE) at
// src/SolveDVO.cpp:1709,1771).  Integer arithmetic only (Meijster et al. 2000).
// d2(y,x) = min over edge pixels (edge != 0) of dy^2 + dx^2.  Returns the number of edge pixels; if there
// are none every d2 is EDT_INF (the reference's DT is undefined there and it later asserts, :282).
// ------------------------------------------------------------------------------------------------------
constexpr int EDT_INF_1D = 1 << 14;                   // "no edge in this column" marker (> any image dim)
constexpr int EDT_INF = EDT_INF_1D * EDT_INF_1D;      // 2^28

inline size_t edt_d2(const uint8_t* edge, int W, int H, int32_t* d2) {
    std::vector<int> g((size_t)W * H);
    size_t nedge = 0;
    for (int x = 0; x < W; ++x) {                      // phase 1: vertical distance to nearest edge in column
        int d = EDT_INF_1D;
        for (int y = 0; y < H; ++y) {
            if (edge[(size_t)y * W + x]) { d = 0; ++nedge; } else if (d < EDT_INF_1D) ++d;
            g[(size_t)y * W + x] = d;
        }
        d = EDT_INF_1D;
        for (int y = H - 1; y >= 0; --y) {
            if (edge[(size_t)y * W + x]) d = 0; else if (d < EDT_INF_1D) ++d;
            if (d < g[(size_t)y * W + x]) g[(size_t)y * W + x] = d;
        }
    }
    std::vector<int> s(W), t(W);
    for (int y = 0; y < H; ++y) {                      // phase 2: lower envelope of parabolas per row
        const int* gr = &g[(size_t)y * W];
        auto f = [&](int x, int i) -> long long { long long dx = x - i; return dx * dx + (long long)gr[i] * gr[i]; };
        auto sep = [&](int i, int u) -> long long {   // first x where parabola u beats parabola i (i < u)
            long long num = (long long)u * u - (long long)i * i + (long long)gr[u] * gr[u] - (long long)gr[i] * gr[i];
            long long den = 2LL * (u - i);
            // floor division (num may be negative)
            long long q = num / den; if ((num % den != 0) && ((num < 0) != (den < 0))) --q;
            return q;
        };
        int q = 0; s[0] = 0; t[0] = 0;
        for (int u = 1; u < W; ++u) {
            while (q >= 0 && f(t[q], s[q]) > f(t[q], u)) --q;
            if (q < 0) { q = 0; s[0] = u; }
            else {
                long long w = 1 + sep(s[q], u);
                if (w < W) { ++q; s[q] = u; t[q] = (int)w; }
            }
        }
        for (int u = W - 1; u >= 0; --u) {
            long long v = f(u, s[q]);
            d2[(size_t)y * W + u] = (int32_t)(v > EDT_INF ? EDT_INF : v);
            if (u == t[q]) --q;
        }
    }
    return nedge;
}

// DT := correctly rounded sqrtf(d2); normalize(DT, 0, 255, NORM_MINMAX) (src/SolveDVO.cpp:1712,1774):
// OpenCV computes scale = (dmax-dmin) * (1/(smax-smin)) in double, narrows it to float for a 32F output, and
// shift = (float)dmin - (float)(smin*scale); dst = src*scale + shift in fp32.
inline void dt_normalize(const int32_t* d2, size_t n, float* dtn, float* scale_out = nullptr) {
    float mn = INFINITY, mx = -INFINITY;
    for (size_t i = 0; i < n; ++i) { float v = std::sqrt((float)d2[i]); if (v < mn) mn = v; if (v > mx) mx = v; }
    double smin = mn, smax = mx;
    double scale = (255.0 - 0.0) * ((smax - smin > 2.220446049250313e-16) ? 1.0 / (smax - smin) : 0.0);
    float fscale = (float)scale;
    float shift = (float)0.0 - (float)(smin * (double)fscale);
    if (scale_out) *scale_out = fscale;
    for (size_t i = 0; i < n; ++i) { float v = std::sqrt((float)d2[i]); float p = v * fscale; dtn[i] = p + shift; }
}

// imageGradient (R4; src/SolveDVO.cpp:1063-1098): filter2D with [-.5 0 .5] and its transpose,
// BORDER_REFLECT_101 => exactly 0 on the first/last column (gx) and row (gy)  (SURVEY Appendix B.6).
inline void gradient(const float* img, int W, int H, float* gx, float* gy) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int xl = x - 1 < 0 ? 1 : x - 1, xr = x + 1 >= W ? W - 2 : x + 1;
            int yu = y - 1 < 0 ? 1 : y - 1, yd = y + 1 >= H ? H - 2 : y + 1;
            if (W == 1) xl = xr = 0; if (H == 1) yu = yd = 0;
            float a = 0.5f * img[(size_t)y * W + xr], b = -0.5f * img[(size_t)y * W + xl];
            gx[(size_t)y * W + x] = b + a;
            float c = 0.5f * img[(size_t)yd * W + x], d = -0.5f * img[(size_t)yu * W + x];
            gy[(size_t)y * W + x] = d + c;
        }
}

// ------------------------------------------------------------------------------------------------------
// Camera + reference point list (R5)
// ------------------------------------------------------------------------------------------------------
struct Intrinsics { float fx, fy, cx, cy; };   // of level 0 (src/SolveDVO.cpp:103-106, include/SolveDVO.h:179)

struct PointList { std::vector<float> X, Y, Z, u, v; size_t size() const { return X.size(); } };

// selectedPts (src/SolveDVO.cpp:1230-1264) + enlistRefEdgePts (:224-264): column-major enumeration
// (xx outer, yy inner); mask = edge > 0 && depth > 100.0f (mm).
inline void select_points(const uint8_t* edge, const uint16_t* depth_mm, int W, int H, int level, Intrinsics K,
                          PointList& out) {
    out.X.clear(); out.Y.clear(); out.Z.clear(); out.u.clear(); out.v.clear();
    const float scaleFac = (float)std::pow(2.0, -level);            // :231
    const float tmpfx = (float)(1. / (double)(scaleFac * K.fx));    // :232 (double division, narrowed)
    const float tmpfy = (float)(1. / (double)(scaleFac * K.fy));    // :233
    const float tmpcx = scaleFac * K.cx, tmpcy = scaleFac * K.cy;   // :234-235
    for (int xx = 0; xx < W; ++xx)
        for (int yy = 0; yy < H; ++yy) {
            float d = (float)depth_mm[(size_t)yy * W + xx];
            if (edge[(size_t)yy * W + xx] > 0 && d > 100.0f) {
                float Z = d / 1000.0f;                               // :248
                float X = Z * ((float)xx - tmpcx) * tmpfx;           // :249  ((Z*(xx-cx))*tmpfx)
                float Y = Z * ((float)yy - tmpcy) * tmpfy;           // :250
                out.X.push_back(X); out.Y.push_back(Y); out.Z.push_back(Z);
                out.u.push_back((float)xx); out.v.push_back((float)yy);
            }
        }
}

// ------------------------------------------------------------------------------------------------------
// Small fp64 linear algebra / Lie group helpers (Eigen + Sophus are absent; closed forms)
// All 3x3 matrices are row-major double[9].
// ------------------------------------------------------------------------------------------------------
inline void mat3_mul(const double* A, const double* B, double* C) {
    double r[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    std::memcpy(C, r, sizeof(r));
}
inline void mat3_vec(const double* A, const double* v, double* o) {
    double r[3];
    for (int i = 0; i < 3; ++i) r[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
    o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
}
inline void mat3_identity(double* A) { for (int i = 0; i < 9; ++i) A[i] = (i % 4 == 0) ? 1.0 : 0.0; }

// Eigen::Quaternion(Matrix3) (src/GOP.cpp:105).  q = (x, y, z, w).
inline void rot_to_quat(const double* R, double* q) {
    double t = R[0] + R[4] + R[8];
    if (t > 0) {
        double s = std::sqrt(t + 1.0);
        q[3] = 0.5 * s; s = 0.5 / s;
        q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
    } else {
        int i = 0; if (R[4] > R[0]) i = 1; if (R[8] > R[4 * i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        double s = std::sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = 0.5 * s; s = 0.5 / s;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * s;
        q[j] = (R[3 * j + i] + R[3 * i + j]) * s;
        q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
    }
}

// Sophus::SE3d::exp(psi) (src/SolveDVO.cpp:905-907): psi = (upsilon, omega), translation part first.
inline void se3_exp(const double* psi, double* R, double* t) {
    const double* u = psi; const double* w = psi + 3;
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
    double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}, O2[9];
    mat3_mul(O, O, O2);
    double A, B, C;                                        // sin(th)/th, (1-cos)/th^2, (th-sin)/th^3
    if (th < 1e-5) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0; }
    else { A = std::sin(th) / th; B = (1.0 - std::cos(th)) / th2; C = (th - std::sin(th)) / (th2 * th); }
    double V[9];
    for (int i = 0; i < 9; ++i) {
        double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + A * O[i] + B * O2[i];
        V[i] = I + B * O[i] + C * O2[i];
    }
    mat3_vec(V, u, t);
}

// Sophus::SE3d::log(SE3(R,t)) (src/SolveDVO.cpp:736-739).  Rotation goes through a unit quaternion
// (setRotationMatrix), theta = 2*atan2(|q_v|, q_w).
inline void se3_log(const double* R, const double* t, double* psi) {
    double q[4]; rot_to_quat(R, q);
    double qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= qn;
    if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
    double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], n = std::sqrt(n2), w = q[3];
    double k;                                              // omega = k * q_v
    if (n < 1e-10) k = 2.0 / w - 2.0 * n2 / (3.0 * w * w * w);
    else k = 2.0 * std::atan2(n, w) / n;
    double om[3] = {k * q[0], k * q[1], k * q[2]};
    double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2], th = std::sqrt(th2);
    double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0}, O2[9];
    mat3_mul(O, O, O2);
    double D;                                              // (1 - th*sin/(2(1-cos)))/th^2
    if (th < 1e-5) D = 1.0 / 12.0 + th2 / 720.0;
    else D = (1.0 - th * std::sin(th) / (2.0 * (1.0 - std::cos(th)))) / th2;
    double Vi[9];
    for (int i = 0; i < 9; ++i) Vi[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + D * O2[i];
    mat3_vec(Vi, t, psi);
    psi[3] = om[0]; psi[4] = om[1]; psi[5] = om[2];
}

// 3x3 SVD by two-sided Jacobi on A^T A (symmetric eigen-decomposition), U from A*V.  Sufficient for
// rotationize(): A is within ~1e-15 of orthonormal, so all singular values are ~1 and well separated from 0.
inline void svd3(const double* A, double* U, double* S, double* V) {
    double B[9];                                           // B = A^T A
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) B[3 * i + j] = A[i] * A[j] + A[3 + i] * A[3 + j] + A[6 + i] * A[6 + j];
    mat3_identity(V);
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = std::fabs(B[1]) + std::fabs(B[2]) + std::fabs(B[5]);
        if (off < 1e-300) break;
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double bpq = B[3 * p + q];
                if (std::fabs(bpq) <= 1e-17 * std::sqrt(std::fabs(B[4 * p] * B[4 * q]))) continue;
                rotated = true;
                double tau = (B[4 * q] - B[4 * p]) / (2.0 * bpq);
                double tt = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                double c = 1.0 / std::sqrt(1.0 + tt * tt), s = tt * c;
                for (int k = 0; k < 3; ++k) {              // B <- B * J
                    double bkp = B[3 * k + p], bkq = B[3 * k + q];
                    B[3 * k + p] = c * bkp - s * bkq; B[3 * k + q] = s * bkp + c * bkq;
                }
                for (int k = 0; k < 3; ++k) {              // B <- J^T * B
                    double bpk = B[3 * p + k], bqk = B[3 * q + k];
                    B[3 * p + k] = c * bpk - s * bqk; B[3 * q + k] = s * bpk + c * bqk;
                }
                for (int k = 0; k < 3; ++k) {              // V <- V * J
                    double vkp = V[3 * k + p], vkq = V[3 * k + q];
                    V[3 * k + p] = c * vkp - s * vkq; V[3 * k + q] = s * vkp + c * vkq;
                }
            }
        if (!rotated) break;
    }
    for (int j = 0; j < 3; ++j) {                          // U_j = A v_j / |A v_j|
        double col[3];
        for (int i = 0; i < 3; ++i) col[i] = A[3 * i] * V[j] + A[3 * i + 1] * V[3 + j] + A[3 * i + 2] * V[6 + j];
        double n = std::sqrt(col[0] * col[0] + col[1] * col[1] + col[2] * col[2]);
        S[j] = n;
        for (int i = 0; i < 3; ++i) U[3 * i + j] = n > 0 ? col[i] / n : (i == j ? 1.0 : 0.0);
    }
}

// SolveDVO::rotationize (src/SolveDVO.cpp:1269-1282): R <- U * diag(sigma>0 ? 1 : -1) * V^T.
inline void rotationize(double* R) {
    double U[9], S[3], V[9], US[9];
    svd3(R, U, S, V);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) US[3 * i + j] = U[3 * i + j] * (S[j] > 0 ? 1.0 : -1.0);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = US[3 * i] * V[3 * j] + US[3 * i + 1] * V[3 * j + 1] + US[3 * i + 2] * V[3 * j + 2];
}

// 6x6 SPD solve (Cholesky, fp64); reads the upper triangle of Hin.  Returns false if not positive definite.
inline bool chol6_solve(const double* Hin, const double* b, double* x) {
    double L[36];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = Hin[6 * j + i];
            for (int k = 0; k < j; ++k) s -= L[6 * i + k] * L[6 * j + k];
            if (i == j) { if (!(s > 0)) return false; L[6 * i + i] = std::sqrt(s); }
            else L[6 * i + j] = s / L[6 * j + j];
        }
    double y[6];
    for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[6 * i + k] * y[k]; y[i] = s / L[6 * i + i]; }
    for (int i = 5; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 6; ++k) s -= L[6 * k + i] * x[k]; x[i] = s / L[6 * i + i]; }
    return true;
}

// ------------------------------------------------------------------------------------------------------
// The per-level "now" buffers the solver gathers from, and the per-iteration evaluation (R6, R7)
// ------------------------------------------------------------------------------------------------------
struct LevelImages { int W = 0, H = 0; std::vector<float> dtn, gx, gy; };

enum JacobianMode { JAC_REFERENCE = 0, JAC_EXACT = 1 };
enum WeightMode { W_REF_CAUCHY = 0, W_HUBER = 1, W_NONE = 2 };
enum SolverMode { SOLVER_SUBGRAD_REF = 0, SOLVER_GN = 1, SOLVER_LM = 2 };
// RES_DT_FLOOR: eps = DT(floor(v), floor(u)), the shipped build (src/SolveDVO.cpp:446).
// RES_DT_INTERP: the compiled-out __INTERPOLATE_DISTANCE_TRANSFORM variant (:443-444, interpolate() :1285-1308).
enum ResidualMode { RES_DT_FLOOR = 0, RES_DT_INTERP = 1 };

struct EvalOut {
    double g[6]; double H[36]; double sumsq; int nvis;
    std::vector<float> eps, w, reproj_u, reproj_v, J;    // per point (J: N x 6 row-major), optional
};

// getWeightOf (src/SolveDVO.cpp:1047-1053): 6.0 / (6.0 + r*r/.25), r*r in float, the rest in double.
inline float weight_ref(float r) { float rr = r * r; return (float)(6.0 / (6.0 + (double)rr / .25)); }
// Huber weight (extension; north_star "Huber/sub-gradient weights"): w = 1 if |r|<=k else k/|r|.
inline float weight_huber(float r, float k) { float a = std::fabs(r); return a <= k ? 1.0f : k / a; }

// SolveDVO::interpolate (src/SolveDVO.cpp:1285-1308): "squared-bilinear" lookup sqrt((1-a) F0^2 + a F1^2) along x on
// the floor and ceil rows, then the same along y; all in fp32, products evaluated left to right.  The reference
// indexes F(ceil(ry), ceil(rx)) without a bound check although its visibility test admits ry up to rows; with its
// forced-on asserts (include/SolveDVO.h:124-125) that aborts -- here the ceil indices are clamped to the last
// row / column (documented deviation, only reachable within one pixel of the bottom / right border).
inline float interpolate_dt(const float* F, int W, int H, float ry, float rx) {
    int ry_d = (int)std::floor(ry), rx_d = (int)std::floor(rx);
    int ry_u = (int)std::ceil(ry), rx_u = (int)std::ceil(rx);
    float inc_x = rx - (float)rx_d, inc_y = ry - (float)ry_d;
    if (ry_u > H - 1) ry_u = H - 1;
    if (rx_u > W - 1) rx_u = W - 1;
    float Fdd = F[(size_t)ry_d * W + rx_d], Fdu = F[(size_t)ry_d * W + rx_u];
    float Fud = F[(size_t)ry_u * W + rx_d], Fuu = F[(size_t)ry_u * W + rx_u];
    float f1 = std::sqrt(((1.0f - inc_x) * Fdd) * Fdd + (inc_x * Fdu) * Fdu);
    float f2 = std::sqrt(((1.0f - inc_x) * Fud) * Fud + (inc_x * Fuu) * Fuu);
    return std::sqrt(((1.0f - inc_y) * f1) * f1 + (inc_y * f2) * f2);
}

// computeJacobianOfNowFrame (src/SolveDVO.cpp:306-414) + getReprojectedEpsilons (:425-462) + the normal
// equation sums of runIterations (:714-720, :777).  cR, cT are the fp64 pose; they are narrowed to fp32
// exactly as :673-674 does.  H = sum (double)(J_i w_i)^T (double)J_i is the oracle's GN extension.
inline void evaluate(const PointList& pts, const LevelImages& now, int level, Intrinsics K, const double* cR,
                     const double* cT, JacobianMode jm, WeightMode wm, float huber_k, bool keep_per_point,
                     EvalOut& out, ResidualMode rm = RES_DT_FLOOR) {
    const size_t N = pts.size();
    float R[9], T[3];
    for (int i = 0; i < 9; ++i) R[i] = (float)cR[i];
    for (int i = 0; i < 3; ++i) T[i] = (float)cT[i];
    const float scaleFac = (float)std::pow(2.0, -level);            // :334
    // scaleMatrix * K (3x3 product evaluated first, :344): M00 = sf*fx, M02 = sf*cx, M11 = sf*fy, M12 = sf*cy
    const float M00 = scaleFac * K.fx, M02 = scaleFac * K.cx, M11 = scaleFac * K.fy, M12 = scaleFac * K.cy;
    const int nCols = now.W, nRows = now.H;
    for (int i = 0; i < 6; ++i) out.g[i] = 0; for (int i = 0; i < 36; ++i) out.H[i] = 0;
    out.sumsq = 0; out.nvis = 0;
    if (keep_per_point) {
        out.eps.assign(N, 0.f); out.w.assign(N, 0.f); out.reproj_u.assign(N, 0.f); out.reproj_v.assign(N, 0.f);
        out.J.assign(N * 6, 0.f);
    }
    for (size_t i = 0; i < N; ++i) {
        // p' = cR^T (P - cT)   (:330), accumulated k = 0,1,2
        float d0 = pts.X[i] - T[0], d1 = pts.Y[i] - T[1], d2 = pts.Z[i] - T[2];
        float px = (R[0] * d0 + R[3] * d1) + R[6] * d2;
        float py = (R[1] * d0 + R[4] * d1) + R[7] * d2;
        float pz = (R[2] * d0 + R[5] * d1) + R[8] * d2;
        float inv = 1.0f / pz;                                       // :339
        float X = px * inv, Y = py * inv, Z = pz * inv;              // :340-341 (all three rows)
        // (scaleMatrix*K) * p_n  (:344): row 0 = (M00*X + 0*Y) + M02*Z ; the 0*Y term is an exact no-op
        float u = M00 * X + M02 * Z, v = M11 * Y + M12 * Z;
        if (keep_per_point) { out.reproj_u[i] = u; out.reproj_v[i] = v; }
        // visibility (:371, :435): '>' bounds; u == nCols / v == nRows would index out of range in the
        // reference (UB) and NaN would pass its test -- both treated as invisible here (measure zero).
        if (!(u >= 0.0f && u < (float)nCols && v >= 0.0f && v < (float)nRows)) continue;
        int xx = (int)u, yy = (int)v;                                // :376-377
        size_t idx = (size_t)yy * nCols + xx;
        float G0 = now.gx[idx], G1 = now.gy[idx];                    // :384-385
        float Jr[6];
        if (jm == JAC_REFERENCE) {
            float A00 = M00 / Z, A02 = -((M00 * X) / (Z * Z));       // :388-390  (scaleFac*fx == M00)
            float A11 = M11 / Z, A12 = -((M11 * Y) / (Z * Z));       // :392-393
            // tmp = cR^T * p_n (:399)
            float w0 = (R[0] * X + R[3] * Y) + R[6] * Z;
            float w1 = (R[1] * X + R[4] * Y) + R[7] * Z;
            float w2 = (R[2] * X + R[5] * Y) + R[8] * Z;
            // G*A1 (1x3): the products with A1's structural zeros are exact no-ops
            float a = G0 * A00, b = G1 * A11, c = G0 * A02 + G1 * A12;
            // (G*A1) * A2, A2 = [ -cR^T | hat(w) ] (:397-405, to_se_3 :1104-1114)
            Jr[0] = -((a * R[0] + b * R[1]) + c * R[2]);
            Jr[1] = -((a * R[3] + b * R[4]) + c * R[5]);
            Jr[2] = -((a * R[6] + b * R[7]) + c * R[8]);
            Jr[3] = b * w2 - c * w1;                                 // a*0 + b*w2 + c*(-w1)
            Jr[4] = c * w0 - a * w2;                                 // a*(-w2) + b*0 + c*w0  == (-(a*w2)) + c*w0
            Jr[5] = a * w1 - b * w0;                                 // a*w1 + b*(-w0) + c*0
        } else {
            // Exact Jacobian of eps wrt the right-multiplicative update (T += R*ups, R <- R*exp(om)):
            // d p'/d ups = -I, d p'/d om = hat(p'), projection Jacobian at the un-normalised p'.
            float iz = 1.0f / pz;
            float a = (G0 * M00) * iz, b = (G1 * M11) * iz;
            float c = -((a * px + b * py) * iz);
            Jr[0] = -a; Jr[1] = -b; Jr[2] = -c;
            Jr[3] = b * pz - c * py; Jr[4] = c * px - a * pz; Jr[5] = a * py - b * px;
        }
        float e = (rm == RES_DT_INTERP) ? interpolate_dt(now.dtn.data(), nCols, nRows, v, u) : now.dtn[idx];   // :443-446
        float w;
        if (wm == W_REF_CAUCHY) w = weight_ref(e); else if (wm == W_HUBER) w = weight_huber(e, huber_k); else w = 1.0f;
        out.nvis++;
        out.sumsq += (double)e * (double)e;
        double de = (double)e;
        double Jw[6];
        for (int k = 0; k < 6; ++k) { float jw = Jr[k] * w; Jw[k] = (double)jw; out.g[k] += Jw[k] * de; }   // :714-720,:777
        for (int r = 0; r < 6; ++r)
            for (int c2 = 0; c2 < 6; ++c2) out.H[6 * r + c2] += Jw[r] * (double)Jr[c2];
        if (keep_per_point) { out.eps[i] = e; out.w[i] = w; for (int k = 0; k < 6; ++k) out.J[6 * i + k] = Jr[k]; }
    }
}

// ------------------------------------------------------------------------------------------------------
// runIterations (R8; src/SolveDVO.cpp:619-1017) and the GN / LM extension
// ------------------------------------------------------------------------------------------------------
struct SolverConfig {
    SolverMode solver = SOLVER_SUBGRAD_REF;
    JacobianMode jac = JAC_REFERENCE;
    WeightMode weight = W_REF_CAUCHY;
    float huber_k = 1.345f;
    double lm_lambda0 = 1e-3;
    float trust_radius = 0.003f;            // trustRegionHyperSphereRadius (src/SolveDVO.cpp:25)
    float psi_term = 1.0E-7f;               // psiNormTerminationThreshold (:24)
    ResidualMode residual = RES_DT_FLOOR;
};

struct IterTrace { double g[6]; double H[36]; float energy; int nvis; double R[9], T[3]; };

struct LevelResult {
    int best_index = -1; float best_energy = 1.0E10f; float visible_ratio = 1.0f; int iterations_run = 0;
    std::vector<float> energies;            // energyAtEachIteration
    std::vector<float> best_eps, best_u, best_v;
    std::vector<IterTrace> trace;
};

inline void run_iterations(const PointList& pts, const LevelImages& now, int level, int maxIterations,
                           Intrinsics K, const SolverConfig& cfg, double* cR, double* cT, LevelResult& res,
                           bool keep_trace = false, bool keep_per_point = true) {
    res = LevelResult();
    res.energies.assign(maxIterations, 0.f);
    float bestTotalEpsilon = 1.0E10f; float bestRatio = 1.0f;
    double bestcR[9], bestcT[3] = {0, 0, 0}; mat3_identity(bestcR);
    int bestItr = -1;
    double descent[6] = {0, 0, 0, 0, 0, 0};
    const double BETA = 0.5;
    // LM state: last accepted linearisation
    double accR[9] = {0}, accT[3] = {0}, accH[36] = {0}, accg[6] = {0}; float accE = 0; bool have_acc = false; double lambda = cfg.lm_lambda0;
    EvalOut ev;
    const size_t N = pts.size();
    for (int itr = 0; itr < maxIterations; ++itr) {
        evaluate(pts, now, level, K, cR, cT, cfg.jac, cfg.weight, cfg.huber_k, keep_per_point, ev, cfg.residual);
        float ratio = N ? (float)ev.nvis / (float)N : 0.f;            // :457
        float energy = (float)std::sqrt(ev.sumsq);                    // aggregateEpsilons :1310-1312
        res.energies[itr] = energy; res.iterations_run = itr + 1;
        if (keep_trace) {
            IterTrace t; std::memcpy(t.g, ev.g, sizeof(t.g)); std::memcpy(t.H, ev.H, sizeof(t.H));
            t.energy = energy; t.nvis = ev.nvis; std::memcpy(t.R, cR, sizeof(t.R)); std::memcpy(t.T, cT, sizeof(t.T));
            res.trace.push_back(t);
        }
        if (energy <= bestTotalEpsilon) {                             // :696-705
            bestTotalEpsilon = energy; bestRatio = ratio; std::memcpy(bestcR, cR, sizeof(bestcR));
            std::memcpy(bestcT, cT, sizeof(bestcT)); bestItr = itr;
            if (keep_per_point) { res.best_eps = ev.eps; res.best_u = ev.reproj_u; res.best_v = ev.reproj_v; }
        }
        double psi[6];
        if (cfg.solver == SOLVER_SUBGRAD_REF) {
            double g[6]; std::memcpy(g, ev.g, sizeof(g));
            // L2 regulariser (:736-743, :796)
            double cPsi[6]; se3_log(cR, cT, cPsi);
            double cn = 0; for (int k = 0; k < 6; ++k) cn += cPsi[k] * cPsi[k]; cn = std::sqrt(cn);
            if (cn > 0) for (int k = 0; k < 6; ++k) cPsi[k] = cPsi[k] / cn;
            double stepLength = 9.0 * 1.0E-2 / ((itr > 5) ? (double)(itr - 4) : 1.0);     // :773
            for (int k = 0; k < 6; ++k) g[k] += 0.05 * cPsi[k];                            // :796
            for (int k = 0; k < 6; ++k) descent[k] = (1.0 - BETA) * g[k] + BETA * descent[k];   // :799
            const double P[6] = {1.0, 1.0, 1.0, 0.5, 0.5, 0.5};                             // :729
            for (int k = 0; k < 6; ++k) psi[k] = -stepLength * P[k] * descent[k];           // :821
            double norm = 0; for (int k = 0; k < 6; ++k) norm += psi[k] * psi[k]; norm = std::sqrt(norm);
            if (norm > (double)cfg.trust_radius) {                                          // :835-839
                for (int k = 0; k < 6; ++k) psi[k] = psi[k] / norm * (double)cfg.trust_radius;
            } else {
                double n2 = 0; for (int k = 0; k < 6; ++k) n2 += psi[k] * psi[k];
                if (std::sqrt(n2) < cfg.psi_term) break;                                    // :840 -> :872 (dangling else)
            }
        } else {
            // GN / LM extension on H = J^T W J, g = J^T W eps (not in the reference)
            bool accept = (cfg.solver == SOLVER_GN) || !have_acc || (energy <= accE);
            if (accept) {
                std::memcpy(accR, cR, sizeof(accR)); std::memcpy(accT, cT, sizeof(accT));
                std::memcpy(accH, ev.H, sizeof(accH)); std::memcpy(accg, ev.g, sizeof(accg)); accE = energy;
                if (cfg.solver == SOLVER_LM && have_acc) lambda = lambda * 0.1 < 1e-9 ? 1e-9 : lambda * 0.1;
                have_acc = true;
            } else {
                lambda *= 10.0;
                if (lambda > 1e8) break;
                std::memcpy(cR, accR, sizeof(accR)); std::memcpy(cT, accT, sizeof(accT));
            }
            double A[36], nb[6];
            for (int k = 0; k < 36; ++k) A[k] = accH[k];
            double lam = (cfg.solver == SOLVER_LM) ? lambda : 0.0;
            for (int k = 0; k < 6; ++k) { A[7 * k] += lam * accH[7 * k] + 1e-9; nb[k] = -accg[k]; }
            if (!chol6_solve(A, nb, psi)) break;
            double n2 = 0; for (int k = 0; k < 6; ++k) n2 += psi[k] * psi[k];
            if (std::sqrt(n2) < cfg.psi_term) break;
        }
        // pose update (:905-920)
        double xR[9], xt[3], Rt[3];
        se3_exp(psi, xR, xt);
        mat3_vec(cR, xt, Rt);
        cT[0] += Rt[0]; cT[1] += Rt[1]; cT[2] += Rt[2];
        mat3_mul(cR, xR, cR);
        rotationize(cR);
    }
    std::memcpy(cR, bestcR, sizeof(bestcR)); rotationize(cR);          // :997-1001
    std::memcpy(cT, bestcT, sizeof(bestcT));
    res.best_index = bestItr; res.best_energy = bestTotalEpsilon; res.visible_ratio = bestRatio;
}

// processResidueHistogram quiet path (R10; src/SolveDVO.cpp:1398-1483): b_cap = mean(eps) in fp32.
inline float laplacian_b(const std::vector<float>& eps) {
    float b = 0; for (float e : eps) b += e; return eps.empty() ? 0.f : b / (float)eps.size();
}

// ------------------------------------------------------------------------------------------------------
// Whole-frame helpers: computeDistTransfrmOfNow/Ref (R3/R4; src/SolveDVO.cpp:1679-1799) for all levels
// ------------------------------------------------------------------------------------------------------
struct FramePyramid {
    int levels = 0;
    std::vector<int> W, H;
    std::vector<std::vector<uint8_t>> gray, edge;
    std::vector<std::vector<uint16_t>> depth;
    std::vector<std::vector<int32_t>> d2;
    std::vector<LevelImages> img;
    std::vector<size_t> nedge;
};

inline void build_frame(const uint8_t* gray, const uint16_t* depth, int W, int H, int levels, bool want_dt,
                        FramePyramid& f) {
    f.levels = levels; f.W.resize(levels); f.H.resize(levels); f.gray.resize(levels); f.edge.resize(levels);
    f.depth.resize(levels); f.d2.resize(levels); f.img.resize(levels); f.nedge.assign(levels, 0);
    for (int l = 0; l < levels; ++l) {
        int w = level_dim(W, l), h = level_dim(H, l);
        f.W[l] = w; f.H[l] = h;
        f.gray[l].resize((size_t)w * h); pyr_nearest(gray, W, H, l, f.gray[l].data());
        if (depth) { f.depth[l].resize((size_t)w * h); pyr_nearest(depth, W, H, l, f.depth[l].data()); depth_fix_zero(f.depth[l].data(), f.depth[l].size()); }
        f.edge[l].resize((size_t)w * h); canny(f.gray[l].data(), w, h, f.edge[l].data());
        if (want_dt) {
            f.d2[l].resize((size_t)w * h); f.nedge[l] = edt_d2(f.edge[l].data(), w, h, f.d2[l].data());
            LevelImages& li = f.img[l]; li.W = w; li.H = h;
            li.dtn.resize((size_t)w * h); li.gx.resize((size_t)w * h); li.gy.resize((size_t)w * h);
            dt_normalize(f.d2[l].data(), (size_t)w * h, li.dtn.data());
            gradient(li.dtn.data(), w, h, li.gx.data(), li.gy.data());
        } else {
            size_t n = 0; for (uint8_t e : f.edge[l]) n += e ? 1 : 0; f.nedge[l] = n;
        }
    }
}

struct PairResult {
    double R[9], T[3];
    std::vector<LevelResult> levels;       // indexed by pyramid level
    std::vector<size_t> npts;
    int status = 0;                          // 0 ok; 1 = some level had no reference points / no now edges
};

// One frame pair through the shipped schedule (R9; src/SolveDVO.cpp:2097-2104): levels (L-1)..0, pose carried.
inline void align_pair(const uint8_t* ref_gray, const uint16_t* ref_depth, const uint8_t* now_gray, int W, int H,
                       int levels, Intrinsics K, const int* iters, const SolverConfig& cfg, const double* R0,
                       const double* T0, PairResult& out, bool keep_trace = false, bool keep_per_point = false) {
    FramePyramid ref, now;
    build_frame(ref_gray, ref_depth, W, H, levels, false, ref);
    build_frame(now_gray, nullptr, W, H, levels, true, now);
    if (R0) std::memcpy(out.R, R0, sizeof(out.R)); else mat3_identity(out.R);
    if (T0) std::memcpy(out.T, T0, sizeof(out.T)); else out.T[0] = out.T[1] = out.T[2] = 0;
    out.levels.assign(levels, LevelResult()); out.npts.assign(levels, 0); out.status = 0;
    for (int l = levels - 1; l >= 0; --l) {
        if (iters[l] <= 0) continue;
        PointList pts;
        select_points(ref.edge[l].data(), ref.depth[l].data(), ref.W[l], ref.H[l], l, K, pts);
        out.npts[l] = pts.size();
        if (pts.size() == 0 || now.nedge[l] == 0) { out.status = 1; continue; }   // reference asserts (:282)
        run_iterations(pts, now.img[l], l, iters[l], K, cfg, out.R, out.T, out.levels[l], keep_trace, keep_per_point);
    }
}

// ------------------------------------------------------------------------------------------------------
// GOP (R11; src/GOP.cpp:138-196)
// ------------------------------------------------------------------------------------------------------
struct GopElement { bool key; int frame, reason; double R[9], T[3], pose[7]; };   // pose = px py pz qx qy qz qw
struct Gop {
    std::vector<GopElement> v; double kR[9], kT[3];
    Gop() { mat3_identity(kR); kT[0] = kT[1] = kT[2] = 0; }
    void push(int frame, bool key, int reason, const double* cR, const double* cT) {
        GopElement e; e.key = key; e.frame = frame; e.reason = key ? reason : -1;
        double RT[3]; mat3_vec(kR, cT, RT);
        for (int i = 0; i < 3; ++i) e.T[i] = kT[i] + RT[i];                 // :144, :172
        mat3_mul(kR, cR, e.R);                                                // :145, :173
        e.pose[0] = e.T[0]; e.pose[1] = e.T[1]; e.pose[2] = e.T[2]; rot_to_quat(e.R, e.pose + 3);
        v.push_back(e);
        if (key) { std::memcpy(kR, e.R, sizeof(kR)); std::memcpy(kT, e.T, sizeof(kT)); }   // :184-185
    }
    void update_most_recent_to_key(int reason) {                              // :189-196
        GopElement& e = v.back(); std::memcpy(kR, e.R, sizeof(kR)); std::memcpy(kT, e.T, sizeof(kT));
        e.key = true; e.reason = reason;
    }
};


// ======================================================================================================
// EPoseEstimator (R12-R17): dense photometric estimator with a reference-frame Jacobian cache.
// `compat` = true restates the reference bug for bug; `compat` = false is the
// corrected formulation the north_star's "photometric + Huber + LM" configuration runs on.
// Arrays are row-major [rows][cols]; the reference's column-major flattening (pixel k = c*rows + r,
// src/EPoseEstimator.cpp:398-410) is applied where it matters (J row order, splat order, quirk 5).
// ======================================================================================================
namespace photo {

struct Cam { double fx, fy, cx, cy; };        // double members (include/EPoseEstimator.h:56)

struct RefLevel {
    int rows = 0, cols = 0, level = 0;
    std::vector<uint8_t> gray, bgr; std::vector<uint16_t> depth;
    std::vector<double> X, Y, Z, gx, gy;      // row-major
    std::vector<double> J;                    // [rows*cols][6], row k = column-major pixel k = c*rows + r
    double A[36];
};

// setRefPyramidalImages (src/EPoseEstimator.cpp:266-290): INTER_AREA from full resolution; gray is converted at full
// resolution first (:71) and then resized.
inline void build_level_images(const uint8_t* bgr, const uint16_t* depth, int W, int H, int level, std::vector<uint8_t>& gray_l,
                               std::vector<uint8_t>* bgr_l, std::vector<uint16_t>* depth_l) {
    const int s = 1 << level, w = W / s, h = H / s;
    std::vector<uint8_t> gray_full((size_t)W * H);
    bgr2gray(bgr, (size_t)W * H, gray_full.data());
    gray_l.resize((size_t)w * h); pyr_area(gray_full.data(), W, H, level, gray_l.data(), 1);
    if (bgr_l) { bgr_l->resize((size_t)w * h * 3); pyr_area(bgr, W, H, level, bgr_l->data(), 3); }
    if (depth_l && depth) { depth_l->resize((size_t)w * h); pyr_area(depth, W, H, level, depth_l->data(), 1); }
}

// filter2D with [0 -1 1] / its transpose, CV_64F, REFLECT_101 (:331-339): forward differences.
inline void forward_gradients(const uint8_t* g, int rows, int cols, double* gx, double* gy) {
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const int c1 = (c + 1 < cols) ? c + 1 : cols - 2, r1 = (r + 1 < rows) ? r + 1 : rows - 2;
            gx[(size_t)r * cols + c] = (double)g[(size_t)r * cols + (cols > 1 ? c1 : c)] - (double)g[(size_t)r * cols + c];
            gy[(size_t)r * cols + c] = (double)g[(size_t)(rows > 1 ? r1 : r) * cols + c] - (double)g[(size_t)r * cols + c];
        }
}

// evaluate3d (:439-477) + evaluateJacobian (:320-430) + A = J^T J (:229) for one level.
inline void build_ref_level(const uint8_t* bgr, const uint16_t* depth, int W, int H, int level, Cam K, bool compat, RefLevel& L) {
    build_level_images(bgr, depth, W, H, level, L.gray, &L.bgr, &L.depth);
    const int s = 1 << level; L.rows = H / s; L.cols = W / s; L.level = level;
    const int rows = L.rows, cols = L.cols; const size_t N = (size_t)rows * cols;
    L.gx.resize(N); L.gy.resize(N); L.X.resize(N); L.Y.resize(N); L.Z.resize(N); L.J.assign(N * 6, 0.0);
    forward_gradients(L.gray.data(), rows, cols, L.gx.data(), L.gy.data());
    const double sf = std::pow(2.0, -level);
    const double fx = K.fx, fy = K.fy, cx = K.cx, cy = K.cy;
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const size_t i = (size_t)r * cols + c;
            const double Zmm = (double)L.depth[i];
            double X, Y, Z;
            if (compat) {                                   // :466-472: row index pairs with cx, fx is not scaled
                X = Zmm / fx * ((double)r - sf * cx); Y = Zmm / fy * ((double)c - sf * cy);
                X = X / 1000.; Y = Y / 1000.; Z = Zmm / 1000.;
            } else {
                // corrected formulation (an extension, defined here): reciprocal focal lengths once per level, metres by a multiply --
                // one rounding per operation in this order; the device evaluates exactly the same expressions
                const double ifx = 1.0 / (sf * fx), ify = 1.0 / (sf * fy);
                Z = Zmm * 1.0e-3;
                X = (Z * ((double)c - sf * cx)) * ifx; Y = (Z * ((double)r - sf * cy)) * ify;
            }
            L.X[i] = X; L.Y[i] = Y; L.Z[i] = Z;
            const double egx = L.gx[i], egy = L.gy[i];
            double J1, J2, J3, J4, J5, J6;
            if (compat) {
                const double Z_inv = 1.0 / Z, Z2_inv = 1.0 / (Z * Z);                                   // :355-356
                J1 = fx * egx * Z_inv;                                                                    // :390
                J2 = fy * egy * Z_inv;
                J3 = -fy * egy * Y * Z2_inv - fx * egx * X * Z2_inv;
                J4 = egy * (-fy * Y * Y * Z2_inv - fy) - fx * egx * X * Y * Z2_inv;
                J5 = egx * (fx * X * X * Z2_inv + fx) + fy * egy * X * Y * Z2_inv;
                J6 = fy * egy * X * Z_inv - fx * egy * Y * Z_inv;                                         // quirk 2: egy twice
            } else if (Zmm > 0.0) {
                const double fxs = sf * fx, fys = sf * fy, Z_inv = 1.0 / Z;
                J1 = fxs * egx * Z_inv; J2 = fys * egy * Z_inv; J3 = -(J1 * X + J2 * Y) * Z_inv;
                J4 = J3 * Y - J2 * Z; J5 = J1 * Z - J3 * X; J6 = J2 * X - J1 * Y;
            } else { J1 = J2 = J3 = J4 = J5 = J6 = 0.0; }
            double* Jr = &L.J[((size_t)c * rows + r) * 6];                                               // column-major pixel order
            Jr[0] = J1; Jr[1] = J2; Jr[2] = J3; Jr[3] = J4; Jr[4] = compat ? J4 : J5; Jr[5] = J6;        // quirk 1: col 5 = col 4
        }
    for (int k = 0; k < 36; ++k) L.A[k] = 0.0;
    for (size_t k = 0; k < N; ++k) {                                                                      // A = J^T J (:229)
        const double* Jr = &L.J[k * 6];
        for (int a = 0; a < 6; ++a) for (int b2 = 0; b2 < 6; ++b2) L.A[6 * a + b2] += Jr[a] * Jr[b2];
    }
}

// 4x4 rigid inverse as a general inverse would give for [R t; 0 1] (Tr.inverse(), :515): R^T, -R^T t.
inline void rigid_inverse(const double* Tr, double* Ti) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Ti[4 * i + j] = Tr[4 * j + i];
    for (int i = 0; i < 3; ++i) Ti[4 * i + 3] = -((Ti[4 * i] * Tr[3] + Ti[4 * i + 1] * Tr[7]) + Ti[4 * i + 2] * Tr[11]);
    Ti[12] = Ti[13] = Ti[14] = 0.0; Ti[15] = 1.0;
}

// warpImage (:490-553): forward splat of the reference gray values, 2x2 block, last writer (largest column-major
// pixel index) wins, holes stay 0.  `winner` (optional) receives the source pixel index per cell or -1.
inline int warp_image(const RefLevel& L, Cam K, const double* Tr, bool compat, std::vector<double>& canvas, std::vector<int>* winner) {
    const int rows = L.rows, cols = L.cols; const size_t N = (size_t)rows * cols;
    double Ti[16]; rigid_inverse(Tr, Ti);
    const double sf = std::pow(2.0, -L.level);
    canvas.assign(N, 0.0); if (winner) winner->assign(N, -1);
    int n = 0;
    for (int c = 0; c < cols; ++c)
        for (int r = 0; r < rows; ++r) {                      // column-major pixel order k = c*rows + r
            const size_t i = (size_t)r * cols + c;
            const double X = L.X[i], Y = L.Y[i], Z = L.Z[i];
            if (!compat && !(L.depth[i] > 0)) continue;
            const double px = ((Ti[0] * X + Ti[1] * Y) + Ti[2] * Z) + Ti[3] * 1.0;
            const double py = ((Ti[4] * X + Ti[5] * Y) + Ti[6] * Z) + Ti[7] * 1.0;
            const double pz = ((Ti[8] * X + Ti[9] * Y) + Ti[10] * Z) + Ti[11] * 1.0;
            int tR, tC;                                       // destination row / column
            if (compat) {                                     // :526-527: u is a ROW coordinate, fx unscaled
                const double u = K.fx * px / pz + sf * K.cx, v = K.fy * py / pz + sf * K.cy;
                tR = (int)std::floor(u); tC = (int)std::floor(v);
            } else {
                if (!(pz > 0.0)) continue;
                const double ipz = 1.0 / pz;                  // one division per point (corrected formulation)
                const double uc = ((sf * K.fx) * px) * ipz + sf * K.cx, vr = ((sf * K.fy) * py) * ipz + sf * K.cy;
                tC = (int)std::floor(uc); tR = (int)std::floor(vr);
            }
            if (tR >= 0 && tR < rows - 1 && tC >= 0 && tC < cols - 1) {
                ++n;
                const double gv = (double)L.gray[i]; const int k = c * rows + r;
                for (int dr = 0; dr < 2; ++dr) for (int dc = 0; dc < 2; ++dc) {
                    canvas[(size_t)(tR + dr) * cols + tC + dc] = gv; if (winner) (*winner)[(size_t)(tR + dr) * cols + tC + dc] = k;
                }
            }
        }
    return n;
}

struct IterOut { double b[6]; double A[36]; double sumsq; int nreproj; int nused; };

inline double huber_w(double r, double k) { const double a = std::fabs(r); return a <= k ? 1.0 : k / a; }

// One evaluation of estimate()'s loop body (:163-188): warp, eps = warped - now, b = J^T eps.
// compat: eps flattened row-major against column-major J rows (quirk 5), holes count (eps = -I_now), unweighted, A = J^T J.
// fixed : eps paired with its own pixel, holes skipped, optional Huber weights, A = J^T W J recomputed.
inline void evaluate(const RefLevel& L, const uint8_t* now_gray, Cam K, const double* Tr, bool compat, double huber_k, IterOut& o,
                     std::vector<double>* canvas_out = nullptr) {
    const int rows = L.rows, cols = L.cols; const size_t N = (size_t)rows * cols;
    std::vector<double> canvas; std::vector<int> win;
    o.nreproj = warp_image(L, K, Tr, compat, canvas, &win);
    for (int k = 0; k < 6; ++k) o.b[k] = 0.0; for (int k = 0; k < 36; ++k) o.A[k] = 0.0; o.sumsq = 0.0; o.nused = 0;
    if (compat) {
        for (size_t k = 0; k < N; ++k) {                      // flattened index k: eps row-major, J column-major
            const double e = canvas[k] - (double)now_gray[k];
            const double* Jr = &L.J[k * 6];
            for (int a = 0; a < 6; ++a) o.b[a] += Jr[a] * e;
            o.sumsq += e * e; ++o.nused;
        }
        std::memcpy(o.A, L.A, sizeof(o.A));
    } else {
        for (int c = 0; c < cols; ++c)
            for (int r = 0; r < rows; ++r) {
                const size_t i = (size_t)r * cols + c;
                if (win[i] < 0) continue;
                const double e = canvas[i] - (double)now_gray[i];
                const double w = huber_k > 0 ? huber_w(e, huber_k) : 1.0;
                const double* Jr = &L.J[(size_t)win[i] * 6];      // Jacobian row of the source pixel that owns this cell
                for (int a = 0; a < 6; ++a) { const double jw = Jr[a] * w; o.b[a] += jw * e; for (int b2 = 0; b2 < 6; ++b2) o.A[6 * a + b2] += jw * Jr[b2]; }
                o.sumsq += e * e; ++o.nused;
            }
    }
    if (canvas_out) *canvas_out = canvas;
}

// exponentialMap (:570-597) -- no small-angle guard in the reference (quirk 8); guarded when !compat.
inline void exp_map(const double* psi, bool compat, double* T4) {
    const double* t = psi; const double* w = psi + 3;
    const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}; double wx2[9]; mat3_mul(wx, wx, wx2);
    double a, b, c;
    if (!compat && th < 1e-8) { a = 1.0; b = 0.5; c = 1.0 / 6.0; }
    else { a = std::sin(th) / th; b = (1.0 - std::cos(th)) / (th * th); c = (th - std::sin(th)) / (th * th * th); }
    double R[9], V[9];
    for (int i = 0; i < 9; ++i) { const double I = (i % 4 == 0) ? 1.0 : 0.0; R[i] = I + a * wx[i] + b * wx2[i]; V[i] = I + b * wx[i] + c * wx2[i]; }
    double vt[3]; mat3_vec(V, t, vt);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T4[4 * i + j] = R[3 * i + j]; T4[4 * i + 3] = vt[i]; }
    T4[12] = T4[13] = T4[14] = 0.0; T4[15] = 1.0;
}

inline void mat4_mul(const double* A, const double* B, double* C) {
    double r[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r[4 * i + j] = ((A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j]) + A[4 * i + 3] * B[12 + j];
    std::memcpy(C, r, sizeof(r));
}

// estimate() (:135-209) in the corrected formulation: `iters` iterations of Tr <- Tr * exp(delta), delta = (A + lambda diag A)^-1 b,
// LM accept/reject on sum eps^2 (lambda0 = 0 -> plain Gauss-Newton as the reference intends).  Returns visible fraction.
// In compat mode A is singular by construction (quirk 1): the normal equations are produced but no step is taken.
struct EstimateOut { double R[9], T[3]; double sumsq_first, sumsq_last; int iters_run; int status; double visible; };
inline void estimate(const RefLevel& L, const uint8_t* now_gray, Cam K, const double* R0, const double* T0, int iters, bool compat,
                     double huber_k, double lambda0, EstimateOut& out) {
    double Tr[16] = {0};
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Tr[4 * i + j] = R0[3 * i + j]; Tr[4 * i + 3] = T0[i]; }
    Tr[15] = 1.0;
    out.status = 0; out.iters_run = 0; out.sumsq_first = out.sumsq_last = 0; out.visible = 0;
    double lambda = lambda0, accTr[16] = {0}, accA[36] = {0}, accb[6] = {0}, accE = 0; bool have = false;
    for (int itr = 0; itr < iters; ++itr) {
        IterOut ev; evaluate(L, now_gray, K, Tr, compat, huber_k, ev);
        out.visible = (double)ev.nreproj / ((double)L.rows * L.cols);
        if (itr == 0) out.sumsq_first = ev.sumsq;
        out.sumsq_last = ev.sumsq; out.iters_run = itr + 1;
        if (compat) { out.status = 2; break; }               // singular normal equations (duplicated Jacobian column)
        const bool accept = !have || lambda0 <= 0.0 || ev.sumsq <= accE;
        if (accept) { std::memcpy(accTr, Tr, sizeof(accTr)); std::memcpy(accA, ev.A, sizeof(accA)); std::memcpy(accb, ev.b, sizeof(accb)); accE = ev.sumsq;
                      if (have && lambda0 > 0.0) lambda = lambda * 0.1 < 1e-9 ? 1e-9 : lambda * 0.1; have = true; }
        else { lambda *= 10.0; if (lambda > 1e8) break; std::memcpy(Tr, accTr, sizeof(accTr)); }
        double Am[36], rhs[6], d[6];
        for (int k = 0; k < 36; ++k) Am[k] = accA[k];
        for (int k = 0; k < 6; ++k) { Am[7 * k] += (lambda0 > 0.0 ? lambda : 0.0) * accA[7 * k] + 1e-9; rhs[k] = accb[k]; }
        if (!chol6_solve(Am, rhs, d)) { out.status = 1; break; }
        // eps = warped - now with J = d(I_ref o warp)/d(psi): the update that decreases ||eps|| is Tr <- Tr * exp(-delta)
        for (int k = 0; k < 6; ++k) d[k] = -d[k];
        double E4[16]; exp_map(d, false, E4); mat4_mul(Tr, E4, Tr);
    }
    if (have && !compat) std::memcpy(Tr, accTr, sizeof(accTr));
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) out.R[3 * i + j] = Tr[4 * i + j]; out.T[i] = Tr[4 * i + 3]; }
}

}  // namespace photo

// ======================================================================================================
// Undistortion front-end of the publisher (src/camTopic2PublisherPyD.cpp:86-117: cv::undistort(src, dst, K, D) for the
// BGR frame and for the 16-bit depth frame).  cv::undistort = initUndistortRectifyMap (fp64 forward distortion model,
// map quantised to 1/32 pixel: CV_16SC2 + 5-bit fractions) + remap(INTER_LINEAR, BORDER_CONSTANT 0).  8-bit images use
// 15-bit fixed-point weights, w = (32-a)(32-b)*32 etc., value = (sum + 2^14) >> 15; 16-bit images use fp32 weights
// (32-a)(32-b)/1024 and saturate_cast<ushort>(cvRound(sum)) with the four products added left to right.  Pinned
// against cv2.undistort 4.13 (tests/golden/undistort_*.npz, tests/test_oracle_golden.py).  D = (k1, k2, p1, p2, k3).
// ======================================================================================================
template <typename T>
void undistort(const T* src, int W, int H, int cn, const double* K4, const double* D5, T* dst) {
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double k1 = D5[0], k2 = D5[1], p1 = D5[2], p2 = D5[3], k3 = D5[4];
    const double ir0 = 1.0 / fx, ir2 = -cx / fx, ir4 = 1.0 / fy, ir5 = -cy / fy;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            const double x = (double)j * ir0 + ir2, y = (double)i * ir4 + ir5;
            const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
            const double kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2;
            const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2), yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
            const double u = fx * xd + cx, v = fy * yd + cy;
            const long long iu = (long long)std::nearbyint(u * 32.0), iv = (long long)std::nearbyint(v * 32.0);
            const long long sx = iu >> 5, sy = iv >> 5; const int a = (int)(iu & 31), b = (int)(iv & 31);
            auto at = [&](long long yy, long long xx, int c) -> T {
                return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? src[((size_t)yy * W + (size_t)xx) * cn + c] : (T)0;
            };
            for (int c = 0; c < cn; ++c) {
                const T s00 = at(sy, sx, c), s01 = at(sy, sx + 1, c), s10 = at(sy + 1, sx, c), s11 = at(sy + 1, sx + 1, c);
                if (sizeof(T) == 1) {
                    const int w00 = (32 - a) * (32 - b) * 32, w01 = a * (32 - b) * 32, w10 = (32 - a) * b * 32, w11 = a * b * 32;
                    int r = ((int)s00 * w00 + (int)s01 * w01 + (int)s10 * w10 + (int)s11 * w11 + (1 << 14)) >> 15;
                    dst[((size_t)i * W + j) * cn + c] = (T)(r < 0 ? 0 : (r > 255 ? 255 : r));
                } else {
                    const float w00 = (float)((32 - a) * (32 - b)) / 1024.f, w01 = (float)(a * (32 - b)) / 1024.f;
                    const float w10 = (float)((32 - a) * b) / 1024.f, w11 = (float)(a * b) / 1024.f;
                    const float f = (((float)s00 * w00 + (float)s01 * w01) + (float)s10 * w10) + (float)s11 * w11;
                    int r = (int)std::nearbyintf(f);
                    dst[((size_t)i * W + j) * cn + c] = (T)(r < 0 ? 0 : (r > 65535 ? 65535 : r));
                }
            }
        }
}

// ======================================================================================================
// RGBDOdometry: the semi-dense photometric Gauss-Newton (src/RGBDOdometry.cpp).  fp64 like the
// reference; quirks kept as written and listed where they occur.  Eigen-specific internals that cannot be restated
// operation for operation (Affine inverse, colPivHouseholderQr) are replaced by the closed forms noted below.
// ======================================================================================================
namespace rgbd {

struct Cam { double fx, fy, cx, cy; };                 // NOT scaled per pyramid level (src/RGBDOdometry.cpp:476-478, :606-612)

struct Level { int rows = 0, cols = 0; std::vector<uint8_t> gray; std::vector<uint16_t> depth; };

// setRefFrame / setNowFrame (:330-390): gray = BGR2GRAY at full resolution, then level i = resize(INTER_NEAREST, 2^-i)
// of gray and depth (the colour pyramid is only drawn on).
inline void build_pyramid(const uint8_t* bgr, const uint16_t* depth, int W, int H, int levels, std::vector<Level>& out) {
    std::vector<uint8_t> gray((size_t)W * H);
    bgr2gray(bgr, (size_t)W * H, gray.data());
    out.assign(levels, Level());
    for (int l = 0; l < levels; ++l) {
        Level& L = out[l];
        L.cols = level_dim(W, l); L.rows = level_dim(H, l);
        L.gray.resize((size_t)L.rows * L.cols); L.depth.resize((size_t)L.rows * L.cols);
        pyr_nearest(gray.data(), W, H, l, L.gray.data(), 1);
        pyr_nearest(depth, W, H, l, L.depth.data(), 1);
        for (auto& d : L.depth) if (d == 0) d = 1;        // imageArrivedCallBack: dframe.setTo(1, dframe == 0) (:232), "to avoid zero depth"
    }
}

struct RefJacobian {
    std::vector<int> pi, pj;          // selected pixels (row i, column j) in the reference's order: j outer, i inner (:459-462)
    std::vector<double> J;            // N x 6
    double A[36];                     // J^T J (:412)
    int status = 0;                   // bit0: fewer than const_minimumRequiredPts + 1 points (:497); bit1: more than const_maxJacobianSize (:463)
};

// computeJacobian (:407-505).  Gradients: filter2D(CV_64F) with [0 -1 1] and its transpose (:425-437), REFLECT_101.
// Selection: egx >= 5 -- the signed x-gradient only (:465).  X pairs the ROW index with cx and Y the column with cy,
// depth stays in raw sensor units (mm), column 0 is fx*fx/Z (:474-488): all as written.
inline void compute_jacobian(const Level& L, Cam K, int gradThresh, int maxJ, int minPts, RefJacobian& out) {
    const int rows = L.rows, cols = L.cols;
    out.pi.clear(); out.pj.clear(); out.J.clear(); out.status = 0;
    for (int k = 0; k < 36; ++k) out.A[k] = 0.0;
    const double fx = K.fx, fy = K.fy, cx = K.cx, cy = K.cy;
    for (int j = 0; j < cols; ++j)
        for (int i = 0; i < rows; ++i) {
            const size_t p = (size_t)i * cols + j;
            const int j1 = (j + 1 < cols) ? j + 1 : (cols > 1 ? cols - 2 : j), i1 = (i + 1 < rows) ? i + 1 : (rows > 1 ? rows - 2 : i);
            const double gx = (double)L.gray[(size_t)i * cols + j1] - (double)L.gray[p];
            const double gy = (double)L.gray[(size_t)i1 * cols + j] - (double)L.gray[p];
            if (gx < (double)gradThresh) continue;
            if ((int)out.pi.size() >= maxJ) { out.status |= 2; continue; }        // the reference asserts (:463)
            const double Z = (double)L.depth[p];
            const double X = Z * ((double)i - cx) / fx, Y = Z * ((double)j - cy) / fy;
            const double invZ = 1 / Z, invZ2 = 1 / (Z * Z);
            double r[6];
            r[0] = fx * fx * invZ;
            r[1] = fy * gy * invZ;
            r[2] = -fy * gy * Y * invZ2 - fx * gx * X * invZ2;
            r[3] = gy * (-fy * Y * Y * invZ2 - fy) - fx * gx * X * Y * invZ2;
            r[4] = gx * (fx * X * X * invZ2 + fx) + fx * gy * X * Y * invZ2;
            r[5] = fy * gy * X * invZ - fx * gy * Y * invZ;
            out.pi.push_back(i); out.pj.push_back(j);
            for (int k = 0; k < 6; ++k) out.J.push_back(r[k]);
            for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) out.A[6 * a + b] += r[a] * r[b];
        }
    if ((int)out.pi.size() <= minPts) out.status |= 1;
}

// 3x3 inverse by the adjugate (Transform<double,3,Affine>::inverse() inverts the linear part as a general matrix).
inline void inv3(const double* M, double* Mi) {
    const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
    const double det = M[0] * c00 + M[1] * c01 + M[2] * c02, id = 1.0 / det;
    Mi[0] = c00 * id; Mi[1] = (M[2] * M[7] - M[1] * M[8]) * id; Mi[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    Mi[3] = c01 * id; Mi[4] = (M[0] * M[8] - M[2] * M[6]) * id; Mi[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    Mi[6] = c02 * id; Mi[7] = (M[1] * M[6] - M[0] * M[7]) * id; Mi[8] = (M[0] * M[4] - M[1] * M[3]) * id;
}
// inverse of the affine transform T (row-major 4x4, last row 0 0 0 1): [L^-1, -L^-1 t]
inline void affine_inverse(const double* T, double* Ti) {
    double L[9] = {T[0], T[1], T[2], T[4], T[5], T[6], T[8], T[9], T[10]}, Li[9];
    inv3(L, Li);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Ti[4 * r + c] = Li[3 * r + c];
        Ti[4 * r + 3] = -((Li[3 * r] * T[3] + Li[3 * r + 1] * T[7]) + Li[3 * r + 2] * T[11]);
    }
    Ti[12] = Ti[13] = Ti[14] = 0.0; Ti[15] = 1.0;
}

struct EpsOut { std::vector<double> eps; std::vector<int> u, v; double sumsq = 0; int nvis = 0; double b[6]; };

// computeEpsilon (:602-700): eps_k = ref_gray(i,j) - now_gray(floor(outu), floor(outv)) where (outu, outv) projects
// T^-1 [X Y Z]; outu is a ROW coordinate (it came from i) and is tested against rows; unseen points keep eps = 0.
// b = -J^T eps (:562) is accumulated alongside.
inline void compute_epsilon(const Level& ref, const RefJacobian& jac, const Level& now, Cam K, const double* T, EpsOut& o) {
    const size_t N = jac.pi.size();
    o.eps.assign(N, 0.0); o.u.assign(N, -1); o.v.assign(N, -1); o.sumsq = 0; o.nvis = 0;
    for (int k = 0; k < 6; ++k) o.b[k] = 0.0;
    double Ti[16]; affine_inverse(T, Ti);
    const double fx = K.fx, fy = K.fy, cx = K.cx, cy = K.cy;
    for (size_t k = 0; k < N; ++k) {
        const int i = jac.pi[k], j = jac.pj[k];
        const double Z = (double)ref.depth[(size_t)i * ref.cols + j];
        const double X = Z * ((double)i - cx) / fx, Y = Z * ((double)j - cy) / fy;
        const double o0 = ((Ti[0] * X + Ti[1] * Y) + Ti[2] * Z) + Ti[3];
        const double o1 = ((Ti[4] * X + Ti[5] * Y) + Ti[6] * Z) + Ti[7];
        const double o2 = ((Ti[8] * X + Ti[9] * Y) + Ti[10] * Z) + Ti[11];
        const double outu = o0 * fx / o2 + cx, outv = o1 * fy / o2 + cy;
        if (outu >= 0 && outu < (double)now.rows && outv >= 0 && outv < (double)now.cols) {
            const int fu = (int)std::floor(outu), fv = (int)std::floor(outv);
            const double e = (double)ref.gray[(size_t)i * ref.cols + j] - (double)now.gray[(size_t)fu * now.cols + fv];
            o.eps[k] = e; o.u[k] = fu; o.v[k] = fv; o.sumsq += e * e; o.nvis++;
            for (int c = 0; c < 6; ++c) o.b[c] -= jac.J[6 * k + c] * e;
        }
    }
}

// exponentialMap (:707-745): identity below theta = 1e-12, otherwise Rodrigues + V (no small-angle series).
inline void exponential_map(const double* psi, double* out16) {
    const double* t = psi; const double* w = psi + 3;
    const double theta = std::sqrt((w[0] * w[0] + w[1] * w[1]) + w[2] * w[2]);
    for (int k = 0; k < 16; ++k) out16[k] = (k % 5 == 0) ? 1.0 : 0.0;
    if (theta < 1E-12) return;
    const double wx[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};           // to_se_3 (:752-763)
    double wx2[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) wx2[3 * r + c] = (wx[3 * r] * wx[c] + wx[3 * r + 1] * wx[3 + c]) + wx[3 * r + 2] * wx[6 + c];
    const double a = std::sin(theta) / theta, b = (1.0 - std::cos(theta)) / (theta * theta), c3 = (theta - std::sin(theta)) / (theta * theta * theta);
    double V[9];
    for (int k = 0; k < 9; ++k) {
        const double I = (k % 4 == 0) ? 1.0 : 0.0;
        out16[4 * (k / 3) + (k % 3)] = (I + a * wx[k]) + b * wx2[k];
        V[k] = (I + b * wx[k]) + c3 * wx2[k];
    }
    for (int r = 0; r < 3; ++r) out16[4 * r + 3] = (V[3 * r] * t[0] + V[3 * r + 1] * t[1]) + V[3 * r + 2] * t[2];
}

// A.colPivHouseholderQr().solve(b) (:563) for 6x6: Householder QR with column pivoting (largest remaining column norm).
inline void colpiv_qr_solve6(const double* Ain, const double* bin, double* x) {
    double A[36], b[6]; int perm[6];
    std::memcpy(A, Ain, sizeof(A)); std::memcpy(b, bin, sizeof(b));
    for (int k = 0; k < 6; ++k) perm[k] = k;
    for (int k = 0; k < 6; ++k) {
        int piv = k; double best = -1.0;
        for (int c = k; c < 6; ++c) { double n2 = 0; for (int r = k; r < 6; ++r) n2 += A[6 * r + c] * A[6 * r + c]; if (n2 > best) { best = n2; piv = c; } }
        if (piv != k) { for (int r = 0; r < 6; ++r) std::swap(A[6 * r + k], A[6 * r + piv]); std::swap(perm[k], perm[piv]); }
        const double nrm = std::sqrt(best);
        if (nrm == 0.0) continue;
        const double alpha = (A[6 * k + k] > 0) ? -nrm : nrm;
        double v[6] = {0, 0, 0, 0, 0, 0};
        for (int r = k; r < 6; ++r) v[r] = A[6 * r + k];
        v[k] -= alpha;
        double vv = 0; for (int r = k; r < 6; ++r) vv += v[r] * v[r];
        if (vv == 0.0) continue;
        for (int c = k; c < 6; ++c) { double d = 0; for (int r = k; r < 6; ++r) d += v[r] * A[6 * r + c]; d = 2.0 * d / vv; for (int r = k; r < 6; ++r) A[6 * r + c] -= d * v[r]; }
        { double d = 0; for (int r = k; r < 6; ++r) d += v[r] * b[r]; d = 2.0 * d / vv; for (int r = k; r < 6; ++r) b[r] -= d * v[r]; }
    }
    double y[6];
    for (int r = 5; r >= 0; --r) {
        double sacc = b[r]; for (int c = r + 1; c < 6; ++c) sacc -= A[6 * r + c] * y[c];
        y[r] = (A[6 * r + r] != 0.0) ? sacc / A[6 * r + r] : 0.0;
    }
    for (int k = 0; k < 6; ++k) x[perm[k]] = y[k];
}

inline void mat4_mul_affine(const double* A, const double* B, double* C) {
    double r[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r[4 * i + j] = ((A[4 * i] * B[j] + A[4 * i + 1] * B[4 + j]) + A[4 * i + 2] * B[8 + j]) + A[4 * i + 3] * B[12 + j];
    std::memcpy(C, r, sizeof(r));
}

struct GnInfo { int npts = 0, iters_run = 0, updates = 0, nvis_last = 0; double eps_norm_first = 0, eps_norm_last = 0; };

// gaussNewtonIterations (:514-597): up to 3 x { eps; stop if ||eps|| < 200; b = -J^T eps; psi = A^-1 b; T <- T * exp(psi)^-1 }.
inline void gauss_newton(const Level& ref, const RefJacobian& jac, const Level& now, Cam K, double* T, int iters, double epsExit, GnInfo& info) {
    info = GnInfo(); info.npts = (int)jac.pi.size();
    for (int itr = 0; itr < iters; ++itr) {
        EpsOut e; compute_epsilon(ref, jac, now, K, T, e);
        const double nrm = std::sqrt(e.sumsq);
        if (itr == 0) info.eps_norm_first = nrm;
        info.eps_norm_last = nrm; info.nvis_last = e.nvis; info.iters_run = itr + 1;
        if (nrm < epsExit) break;                                                    // :556
        double psi[6]; colpiv_qr_solve6(jac.A, e.b, psi);
        double E[16], Ei[16]; exponential_map(psi, E); affine_inverse(E, Ei);
        mat4_mul_affine(T, Ei, T);                                                   // :579
        info.updates++;
    }
}

}  // namespace rgbd

}  // namespace orc
