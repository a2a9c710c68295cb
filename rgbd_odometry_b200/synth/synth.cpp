// Synthetic RGB-D frame-pair renderer (test/bench infrastructure, host only).
//
// Produces the inputs SURVEY.md §8(d) specifies for the BASELINE configs: a seeded scene of textured
// planes (one tilted back wall + a few finite boards) ray-cast from two camera poses related by a known
// small SE(3) motion.  Output per frame: BGR u8 (HWC), gray u8 (OpenCV 4.x BGR2GRAY fixed point, SURVEY
// Appendix B.7), depth u16 in millimetres (> 100 everywhere, no zeros).
//
// Pose convention follows the reference solver (src/SolveDVO.cpp:330): a reference-frame point P maps into
// the now frame as p' = R^T (P - T), i.e. (R, T) is the now camera's pose expressed in the reference frame.
//
// Piecewise-constant two-scale textures guarantee Canny edges at every NEAREST pyramid level (the reference
// asserts nSelectedPts > 0 per level, src/SolveDVO.cpp:282).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    double range(double a, double b) { return a + (b - a) * uni(); }
};

inline uint32_t hash3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = a * 0x9E3779B1u;
    h ^= b + 0x7F4A7C15u + (h << 6) + (h >> 2);
    h *= 0x85EBCA6Bu;
    h ^= c + 0x165667B1u + (h << 6) + (h >> 2);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
    return h;
}

struct Vec3 { double x, y, z; };
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline Vec3 normalize(Vec3 a) { double n = std::sqrt(dot(a, a)); return (1.0 / n) * a; }

struct Surface {
    Vec3 c, n, a, b;     // centre, unit normal, in-plane orthonormal axes
    double hu, hv;       // half extents along a, b (<= 0: infinite)
    double cell1, cell2; // coarse / fine texture cell size (metres)
    uint32_t id;
};

struct Scene {
    std::vector<Surface> surf;
    uint32_t tex_seed;
};

Vec3 random_unit(Rng& r) {
    for (;;) {
        Vec3 v{r.range(-1, 1), r.range(-1, 1), r.range(-1, 1)};
        double n2 = dot(v, v);
        if (n2 > 1e-4 && n2 <= 1.0) return (1.0 / std::sqrt(n2)) * v;
    }
}

void make_basis(Vec3 n, Vec3& a, Vec3& b) {
    Vec3 up = std::fabs(n.y) < 0.9 ? Vec3{0, 1, 0} : Vec3{1, 0, 0};
    a = normalize(cross(up, n));
    b = cross(n, a);
}

Scene make_scene(Rng& r) {
    Scene sc;
    sc.tex_seed = (uint32_t)r.next();
    // back wall: infinite, tilted by a few degrees, 2.6 .. 3.8 m away
    {
        Surface s{};
        Vec3 n = normalize(Vec3{r.range(-0.12, 0.12), r.range(-0.12, 0.12), -1.0});
        s.n = n; s.c = Vec3{0, 0, r.range(2.6, 3.8)};
        make_basis(n, s.a, s.b);
        s.hu = s.hv = -1.0;
        s.cell1 = r.range(0.30, 0.50); s.cell2 = s.cell1 / 3.0;
        s.id = 0;
        sc.surf.push_back(s);
    }
    int nboards = 3 + (int)(r.next() % 3);
    for (int k = 0; k < nboards; ++k) {
        Surface s{};
        double z = r.range(0.9, 2.3);
        Vec3 n = normalize(Vec3{r.range(-0.25, 0.25), r.range(-0.25, 0.25), -1.0});
        s.n = n;
        s.c = Vec3{r.range(-0.55, 0.55) * z, r.range(-0.40, 0.40) * z, z};
        make_basis(n, s.a, s.b);
        s.hu = r.range(0.15, 0.40) * z * 0.6; s.hv = r.range(0.15, 0.40) * z * 0.6;
        s.cell1 = r.range(0.14, 0.26) * (z / 1.5); s.cell2 = s.cell1 / 3.0;
        s.id = (uint32_t)(k + 1);
        sc.surf.push_back(s);
    }
    return sc;
}

// Piecewise-constant two-scale BGR texture.
inline void texture(const Scene& sc, const Surface& s, double u, double v, int bgr[3]) {
    int32_t i1 = (int32_t)std::floor(u / s.cell1), j1 = (int32_t)std::floor(v / s.cell1);
    uint32_t h1 = hash3(sc.tex_seed + s.id * 7919u, (uint32_t)i1, (uint32_t)j1);
    int base = 40 + (int)(h1 % 176u);                       // 40..215
    int tint_b = (int)((h1 >> 8) % 31u) - 15, tint_r = (int)((h1 >> 16) % 31u) - 15;
    int fine = 0;
    if ((h1 >> 24) & 1u) {                                   // half of the coarse cells carry fine detail
        int32_t i2 = (int32_t)std::floor(u / s.cell2), j2 = (int32_t)std::floor(v / s.cell2);
        uint32_t h2 = hash3(sc.tex_seed ^ 0xA5A5A5A5u, (uint32_t)i2 + s.id * 131u, (uint32_t)j2);
        fine = (int)(h2 % 81u) - 40;                         // -40..40
    }
    int g = base + fine;
    bgr[0] = g + tint_b; bgr[1] = g; bgr[2] = g + tint_r;
}

inline uint8_t clamp_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

struct Pose { double R[9]; double T[3]; };  // row-major R

void render(const Scene& sc, const Pose& pose, uint32_t noise_seed, int W, int H, double fx, double fy,
            double cx, double cy, uint8_t* bgr_out, uint8_t* gray_out, uint16_t* depth_out) {
    const double* R = pose.R;
    Vec3 o{pose.T[0], pose.T[1], pose.T[2]};
    const size_t ns = sc.surf.size();
    std::vector<double> on(ns);  // n . (c - o)
    for (size_t k = 0; k < ns; ++k) on[k] = dot(sc.surf[k].n, sc.surf[k].c - o);
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            double dxc = (x - cx) / fx, dyc = (y - cy) / fy;  // camera-frame ray (dxc, dyc, 1)
            Vec3 d{R[0] * dxc + R[1] * dyc + R[2], R[3] * dxc + R[4] * dyc + R[5], R[6] * dxc + R[7] * dyc + R[8]};
            double best_t = 1e30; int best_k = -1; double bu = 0, bv = 0;
            for (size_t k = 0; k < ns; ++k) {
                const Surface& s = sc.surf[k];
                double den = dot(s.n, d);
                if (den > -1e-6) continue;                    // back-facing or parallel
                double t = on[k] / den;
                if (t <= 0.1 || t >= best_t) continue;
                Vec3 p = o + t * d;
                Vec3 q = p - s.c;
                double u = dot(q, s.a), v = dot(q, s.b);
                if (s.hu > 0 && (std::fabs(u) > s.hu || std::fabs(v) > s.hv)) continue;
                best_t = t; best_k = (int)k; bu = u; bv = v;
            }
            int c[3] = {128, 128, 128};
            double depth_m = 5.0;
            if (best_k >= 0) { texture(sc, sc.surf[best_k], bu, bv, c); depth_m = best_t; }
            uint32_t hn = hash3(noise_seed, (uint32_t)x, (uint32_t)y);
            int nz = (int)(hn % 5u) - 2;                      // -2..2 sensor noise, same on all channels
            uint8_t b8 = clamp_u8(c[0] + nz), g8 = clamp_u8(c[1] + nz), r8 = clamp_u8(c[2] + nz);
            size_t i = (size_t)y * W + x;
            if (bgr_out) { bgr_out[3 * i] = b8; bgr_out[3 * i + 1] = g8; bgr_out[3 * i + 2] = r8; }
            if (gray_out) gray_out[i] = (uint8_t)((b8 * 3735 + g8 * 19235 + r8 * 9798 + (1 << 14)) >> 15);
            if (depth_out) {
                double mm = std::floor(depth_m * 1000.0 + 0.5);
                if (mm < 500) mm = 500; if (mm > 5000) mm = 5000;
                depth_out[i] = (uint16_t)mm;
            }
        }
    }
}

void rodrigues(Vec3 axis, double ang, double R[9]) {
    double c = std::cos(ang), s = std::sin(ang), t = 1 - c;
    double x = axis.x, y = axis.y, z = axis.z;
    R[0] = t * x * x + c;     R[1] = t * x * y - s * z; R[2] = t * x * z + s * y;
    R[3] = t * x * y + s * z; R[4] = t * y * y + c;     R[5] = t * y * z - s * x;
    R[6] = t * x * z - s * y; R[7] = t * y * z + s * x; R[8] = t * z * z + c;
}

}  // namespace

extern "C" {

// Render one frame pair.  Any output pointer may be NULL.  max_angle_deg / max_trans_m bound the true motion
// (1 degree, 0.02 m for the 640x480 configs).
void dvo_synth_pair(uint64_t seed, int W, int H, double fx, double fy, double cx, double cy, double max_angle_deg,
                    double max_trans_m, uint8_t* ref_bgr, uint8_t* ref_gray, uint16_t* ref_depth, uint8_t* now_bgr,
                    uint8_t* now_gray, uint16_t* now_depth, double* R_true9, double* T_true3) {
    Rng r(seed * 0x2545F4914F6CDD1Dull + 0x1234567ull);
    Scene sc = make_scene(r);
    Pose ref{}; ref.R[0] = ref.R[4] = ref.R[8] = 1.0;
    Pose now{};
    Vec3 axis = random_unit(r);
    double ang = r.range(0.0, max_angle_deg) * M_PI / 180.0;
    rodrigues(axis, ang, now.R);
    Vec3 tdir = random_unit(r);
    double tl = r.range(0.0, max_trans_m);
    now.T[0] = tl * tdir.x; now.T[1] = tl * tdir.y; now.T[2] = tl * tdir.z;
    uint32_t nseed = (uint32_t)r.next();
    render(sc, ref, nseed, W, H, fx, fy, cx, cy, ref_bgr, ref_gray, ref_depth);
    render(sc, now, nseed ^ 0x5bd1e995u, W, H, fx, fy, cx, cy, now_bgr, now_gray, now_depth);
    if (R_true9) std::memcpy(R_true9, now.R, sizeof(double) * 9);
    if (T_true3) std::memcpy(T_true3, now.T, sizeof(double) * 3);
}

// Render `count` pairs with seeds seed0 .. seed0+count-1 into contiguous batch buffers (pair-major).
void dvo_synth_batch(uint64_t seed0, int count, int W, int H, double fx, double fy, double cx, double cy,
                     double max_angle_deg, double max_trans_m, uint8_t* ref_bgr, uint8_t* ref_gray,
                     uint16_t* ref_depth, uint8_t* now_bgr, uint8_t* now_gray, uint16_t* now_depth, double* R_true,
                     double* T_true, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    const size_t P = (size_t)W * H;
    auto work = [&](int tid) {
        for (int i = tid; i < count; i += nthreads) {
            dvo_synth_pair(seed0 + (uint64_t)i, W, H, fx, fy, cx, cy, max_angle_deg, max_trans_m,
                           ref_bgr ? ref_bgr + 3 * P * i : nullptr, ref_gray ? ref_gray + P * i : nullptr,
                           ref_depth ? ref_depth + P * i : nullptr, now_bgr ? now_bgr + 3 * P * i : nullptr,
                           now_gray ? now_gray + P * i : nullptr, now_depth ? now_depth + P * i : nullptr,
                           R_true ? R_true + 9 * (size_t)i : nullptr, T_true ? T_true + 3 * (size_t)i : nullptr);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
}

// Render a trajectory of `nframes` frames of one scene (consecutive-pair odometry, BASELINE config 4).
// Frame 0 is at identity; frame k's pose (now camera in frame-0 coordinates) accumulates seeded small motions.
void dvo_synth_sequence(uint64_t seed, int nframes, int W, int H, double fx, double fy, double cx, double cy,
                        double max_angle_deg, double max_trans_m, uint8_t* gray, uint16_t* depth, uint8_t* bgr,
                        double* R_world, double* T_world) {
    Rng r(seed * 0x2545F4914F6CDD1Dull + 0x7654321ull);
    Scene sc = make_scene(r);
    Pose cur{}; cur.R[0] = cur.R[4] = cur.R[8] = 1.0;
    const size_t P = (size_t)W * H;
    // smooth motion: a fixed axis / direction with slowly varying magnitude
    Vec3 axis = random_unit(r), tdir = random_unit(r);
    uint32_t nseed = (uint32_t)r.next();
    for (int k = 0; k < nframes; ++k) {
        render(sc, cur, nseed + 977u * (uint32_t)k, W, H, fx, fy, cx, cy, bgr ? bgr + 3 * P * k : nullptr,
               gray ? gray + P * k : nullptr, depth ? depth + P * k : nullptr);
        if (R_world) std::memcpy(R_world + 9 * (size_t)k, cur.R, sizeof(double) * 9);
        if (T_world) std::memcpy(T_world + 3 * (size_t)k, cur.T, sizeof(double) * 3);
        double ang = r.range(0.3, 1.0) * max_angle_deg * M_PI / 180.0;
        double tl = r.range(0.3, 1.0) * max_trans_m;
        double dR[9]; rodrigues(axis, ang, dR);
        // cur <- cur * (dR, dT): T += R*dT ; R = R*dR
        Vec3 dT = tl * tdir;
        double nR[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                nR[3 * i + j] = cur.R[3 * i] * dR[j] + cur.R[3 * i + 1] * dR[3 + j] + cur.R[3 * i + 2] * dR[6 + j];
        cur.T[0] += cur.R[0] * dT.x + cur.R[1] * dT.y + cur.R[2] * dT.z;
        cur.T[1] += cur.R[3] * dT.x + cur.R[4] * dT.y + cur.R[5] * dT.z;
        cur.T[2] += cur.R[6] * dT.x + cur.R[7] * dT.y + cur.R[8] * dT.z;
        std::memcpy(cur.R, nR, sizeof(nR));
        // slowly drift the axis so the path is not a pure screw motion
        Vec3 j = random_unit(r);
        axis = normalize(axis + 0.15 * j);
        tdir = normalize(tdir + 0.15 * random_unit(r));
    }
}

}  // extern "C"
