// preprocess.cu -- pyramid, Canny (+ fused EDT column pass / reference point list), EDT row pass,
// normalise + gradient.  All integer stages are bit-exact against oracle/dvo_oracle.hpp.
#include "common.cuh"

// =====================================================================================================
// NEAREST pyramid (src/camTopic2PublisherPyD.cpp:338-348): level l pixel (y,x) = level-0 pixel (y<<l, x<<l),
// always taken from full resolution.  Depth zeros become 1 on levels >= 1 here; level 0 keeps the caller's
// bytes and the zero fix (src/SolveDVO.cpp:512) is applied where depth is consumed / read back.
// =====================================================================================================
struct PyrArgs {
    PyrGeom g;
    uint8_t* gray;
    uint16_t* depth;   // may be null
    int first;
    int sub_total;     // sum_{l>=1} P[l]
};

__global__ void __launch_bounds__(256) pyramid_nearest_kernel(PyrArgs a) {
    const int b = a.first + blockIdx.y;
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.sub_total) return;
    int l = 1;
    while (l < a.g.L - 1 && q >= a.g.P[l]) { q -= a.g.P[l]; ++l; }
    const int w = a.g.w[l], W0 = a.g.w[0], H0 = a.g.h[0];
    const int y = q / w, x = q - y * w;
    const int sy = min(y << l, H0 - 1), sx = min(x << l, W0 - 1);
    const long long src = lvl_at(a.g, 0, b) + (long long)sy * W0 + sx;
    const long long dst = lvl_at(a.g, l, b) + q;
    a.gray[dst] = a.gray[src];
    if (a.depth) { uint16_t d = a.depth[src]; a.depth[dst] = d == 0 ? (uint16_t)1 : d; }
}

int launch_pyramid(dvo_ctx* c, int first, int count, int frames_mask) {
    if (c->geom.L < 2) return DVO_OK;
    int sub = 0;
    for (int l = 1; l < c->geom.L; ++l) sub += c->geom.P[l];
    for (int f = 0; f < 2; ++f) {
        if (!(frames_mask & (1 << f))) continue;
        PyrArgs a; a.g = c->geom; a.gray = c->gray[f]; a.depth = c->depth[f]; a.first = first; a.sub_total = sub;
        dim3 grid((sub + 255) / 256, count);
        pyramid_nearest_kernel<<<grid, 256, 0, c->stream>>>(a);
        c->launches++;
    }
    DVO_CUDA(cudaGetLastError());
    return DVO_OK;
}

// =====================================================================================================
// Canny (cv::Canny(img,150,100,3,true), src/SolveDVO.cpp:1705,1767).
// One CTA per (slot, frame) image at one level.  Candidate / edge sets live as bitmaps (1 bit per pixel,
// bit i of word wx <-> x = 32*wx + i) with a one-word / one-row zero border, in shared memory when they fit
// (640x480: 2 x 42 KB) or in a global scratch otherwise.  Hysteresis is a monotone fixed-point iteration
// E <- C & dilate3x3(E) from E = strong, whose fixed point is the traversal-order-independent closure the
// reference computes with a stack -- hence bit-exact.
// Fused epilogues (the bitmap is already on chip):
//   now frame: EDT phase 1 -- per-column distance to the nearest edge pixel above/below (u16)
//   ref frame: selectedPts + enlistRefEdgePts (src/SolveDVO.cpp:1230-1264, 224-264): column-major stable
//              compaction of edge && depth > 100 pixels, back-projected to 3-D.
// =====================================================================================================
struct CannyArgs {
    const uint8_t* gray;   // level region of this frame (slot 0)
    uint8_t* edge;
    const uint16_t* depth; // ref
    uint16_t* gcol;        // now
    float *X, *Y, *Z;      // ref, level region
    int* npts;             // + level, stride L
    unsigned* nedge;       // + level, stride L (this frame)
    int w, h, P, L;
    int do_points, do_cols;
    float tmpfx, tmpfy, tmpcx, tmpcy;
    uint32_t* gscratch;    // null -> shared memory bitmaps
    long long gscratch_stride;
    int first;
    int low, high;         // squared thresholds (10000, 22500)
};

__device__ __forceinline__ void sobel3(const uint8_t* __restrict__ g, int w, int h, int y, int x, int& dx, int& dy) {
    const int ym = max(y - 1, 0), yp = min(y + 1, h - 1), xm = max(x - 1, 0), xp = min(x + 1, w - 1);
    const uint8_t* r0 = g + ym * w; const uint8_t* r1 = g + y * w; const uint8_t* r2 = g + yp * w;
    const int a = r0[xm], b = r0[x], c = r0[xp], d = r1[xm], f = r1[xp], p = r2[xm], q = r2[x], r = r2[xp];
    dx = (c + 2 * f + r) - (a + 2 * d + p);
    dy = (p + 2 * q + r) - (a + 2 * b + c);
}
__device__ __forceinline__ int mag_at(const uint8_t* __restrict__ g, int w, int h, int y, int x) {
    if ((unsigned)y >= (unsigned)h || (unsigned)x >= (unsigned)w) return 0;   // zero-padded magnitude border
    int dx, dy; sobel3(g, w, h, y, x, dx, dy);
    return dx * dx + dy * dy;
}

__device__ __forceinline__ uint32_t expand4(uint32_t nib) { return (((nib & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu; }

template <int THREADS>
__global__ void __launch_bounds__(THREADS) canny_kernel(CannyArgs a) {
    extern __shared__ uint32_t smem_u32[];
    __shared__ int s_scan[THREADS / 32 + 1];
    __shared__ int s_base;
    __shared__ unsigned s_cnt;
    const int tid = threadIdx.x, lane = tid & 31;
    const int b = a.first + blockIdx.x;
    const int w = a.w, h = a.h;
    const int wd = (w + 31) >> 5, pitch = wd + 2, rows = h + 2;
    const int nwords = pitch * rows;
    uint32_t* C = a.gscratch ? a.gscratch + (long long)blockIdx.x * a.gscratch_stride : smem_u32;
    uint32_t* E = C + nwords;
    const uint8_t* __restrict__ g = a.gray + (long long)b * a.P;

    for (int i = tid; i < 2 * nwords; i += THREADS) C[i] = 0u;
    if (tid == 0) { s_cnt = 0u; s_base = 0; }
    __syncthreads();

    // ---- phase 1: Sobel + non-maximum suppression, one warp per 32-pixel word ----
    const int TG22 = 13573;   // (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
    const int total = h * wd * 32;
    for (int q = tid; q < total; q += THREADS) {
        const int y = q / (wd * 32);
        const int x = q - y * (wd * 32);
        bool cand = false, strong = false;
        if (x < w) {
            int dx, dy; sobel3(g, w, h, y, x, dx, dy);
            const int m = dx * dx + dy * dy;
            if (m > a.low) {
                const int ax = abs(dx), ay = abs(dy) << 15;
                const int tg22x = ax * TG22;
                bool keep;
                if (ay < tg22x) keep = (m > mag_at(g, w, h, y, x - 1)) && (m >= mag_at(g, w, h, y, x + 1));
                else {
                    const int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x) keep = (m > mag_at(g, w, h, y - 1, x)) && (m >= mag_at(g, w, h, y + 1, x));
                    else { const int s = ((dx ^ dy) < 0) ? -1 : 1; keep = (m > mag_at(g, w, h, y - 1, x - s)) && (m > mag_at(g, w, h, y + 1, x + s)); }
                }
                cand = keep; strong = keep && (m > a.high);
            }
        }
        const uint32_t cw = __ballot_sync(0xffffffffu, cand), sw = __ballot_sync(0xffffffffu, strong);
        if (lane == 0) { const int idx = (y + 1) * pitch + (x >> 5) + 1; C[idx] = cw; E[idx] = sw; }
    }
    __syncthreads();

    // ---- phase 2: hysteresis closure ----
    const int nw = h * wd;
    for (;;) {
        int changed = 0;
        for (int q = tid; q < nw; q += THREADS) {
            const int y = q / wd, wx = q - y * wd;
            const int idx = (y + 1) * pitch + wx + 1;
            const uint32_t c = C[idx];
            if (c == 0u) continue;
            const uint32_t e = E[idx];
            if (e == c) continue;
            uint32_t n = 0u;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int j = idx + dy * pitch;
                const uint32_t ec = E[j], el = E[j - 1], er = E[j + 1];
                n |= ec | (ec << 1) | (ec >> 1) | (el >> 31) | (er << 31);
            }
            uint32_t f = e | (c & n);
            for (;;) { const uint32_t f2 = f | (c & ((f << 1) | (f >> 1))); if (f2 == f) break; f = f2; }
            if (f != e) { E[idx] = f; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }

    // ---- phase 3a: edge bytes (0/255) + edge count ----
    uint8_t* eo = a.edge + (long long)b * a.P;
    unsigned cnt = 0;
    const bool vec_ok = (w % 16 == 0) && ((reinterpret_cast<uintptr_t>(eo) & 15) == 0);
    for (int q = tid; q < nw; q += THREADS) {
        const int y = q / wd, wx = q - y * wd;
        const uint32_t e = E[(y + 1) * pitch + wx + 1];
        cnt += __popc(e);
        const int x0 = wx << 5;
        if (vec_ok && x0 + 32 <= w) {
            uint4 v0, v1;
            v0.x = expand4(e); v0.y = expand4(e >> 4); v0.z = expand4(e >> 8); v0.w = expand4(e >> 12);
            v1.x = expand4(e >> 16); v1.y = expand4(e >> 20); v1.z = expand4(e >> 24); v1.w = expand4(e >> 28);
            uint4* dst = reinterpret_cast<uint4*>(eo + (long long)y * w + x0);
            dst[0] = v0; dst[1] = v1;
        } else {
            for (int i = 0; i < 32 && x0 + i < w; ++i) eo[(long long)y * w + x0 + i] = ((e >> i) & 1u) ? 255 : 0;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (tid == 0) a.nedge[(long long)b * a.L] = s_cnt;

    // ---- phase 3b (now): EDT phase 1, per-column distance to the nearest edge pixel ----
    if (a.do_cols) {
        uint16_t* gc = a.gcol + (long long)b * a.P;
        for (int x = tid; x < w; x += THREADS) {
            const int wi = (x >> 5) + 1, sh = x & 31;
            int d = DVO_EDT_INF_1D;
            for (int y = 0; y < h; ++y) {
                const uint32_t bit = (E[(y + 1) * pitch + wi] >> sh) & 1u;
                d = bit ? 0 : min(d + 1, DVO_EDT_INF_1D);
                gc[(long long)y * w + x] = (uint16_t)d;
            }
            d = DVO_EDT_INF_1D;
            for (int y = h - 1; y >= 0; --y) {
                const uint32_t bit = (E[(y + 1) * pitch + wi] >> sh) & 1u;
                d = bit ? 0 : min(d + 1, DVO_EDT_INF_1D);
                const int prev = gc[(long long)y * w + x];
                if (d < prev) gc[(long long)y * w + x] = (uint16_t)d;
            }
        }
    }

    // ---- phase 3c (ref): column-major stable compaction + back-projection ----
    if (a.do_points) {
        const uint16_t* __restrict__ dep = a.depth + (long long)b * a.P;
        float* X = a.X + (long long)b * a.P; float* Y = a.Y + (long long)b * a.P; float* Z = a.Z + (long long)b * a.P;
        for (int x0 = 0; x0 < w; x0 += THREADS) {
            const int x = x0 + tid;
            int mycnt = 0;
            int wi = 0, sh = 0;
            if (x < w) {
                wi = (x >> 5) + 1; sh = x & 31;
                for (int y = 0; y < h; ++y)
                    if ((E[(y + 1) * pitch + wi] >> sh) & 1u) mycnt += (dep[(long long)y * w + x] > 100) ? 1 : 0;
            }
            // block exclusive scan of mycnt
            int incl = mycnt;
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) s_scan[tid >> 5] = incl;
            __syncthreads();
            if (tid < 32) {
                int v = (tid < THREADS / 32) ? s_scan[tid] : 0;
                int iv = v;
                for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
                if (tid < THREADS / 32) s_scan[tid] = iv - v;       // exclusive warp offsets
                if (tid == 31) s_scan[THREADS / 32] = iv;            // chunk total
            }
            __syncthreads();
            int off = s_base + s_scan[tid >> 5] + (incl - mycnt);
            if (x < w && mycnt) {
                for (int y = 0; y < h; ++y)
                    if ((E[(y + 1) * pitch + wi] >> sh) & 1u) {
                        const float d = (float)dep[(long long)y * w + x];
                        if (d > 100.0f) {
                            const float z = __fdiv_rn(d, 1000.0f);                                      // :248
                            X[off] = __fmul_rn(__fmul_rn(z, __fsub_rn((float)x, a.tmpcx)), a.tmpfx);      // :249
                            Y[off] = __fmul_rn(__fmul_rn(z, __fsub_rn((float)y, a.tmpcy)), a.tmpfy);      // :250
                            Z[off] = z;
                            ++off;
                        }
                    }
            }
            __syncthreads();
            if (tid == 0) s_base += s_scan[THREADS / 32];
            __syncthreads();
        }
        if (tid == 0) a.npts[(long long)b * a.L] = s_base;
    }
}

static size_t canny_smem_bytes(int w, int h) { return (size_t)2 * (((w + 31) >> 5) + 2) * (h + 2) * sizeof(uint32_t); }

int launch_canny(dvo_ctx* c, int first, int count, int frames_mask) {
    const PyrGeom& g = c->geom;
    for (int l = 0; l < g.L; ++l) {
        const size_t smem = canny_smem_bytes(g.w[l], g.h[l]);
        const bool use_global = smem + 1024 > c->smem_optin;
        if (use_global && !c->bitmap_scratch) { dvo_set_error("canny: bitmap scratch missing for %dx%d", g.w[l], g.h[l]); return DVO_ERR_STATE; }
        for (int f = 0; f < 2; ++f) {
            if (!(frames_mask & (1 << f))) continue;
            CannyArgs a;
            a.gray = c->gray[f] + g.off[l]; a.edge = c->edge[f] + g.off[l];
            a.depth = c->depth[f] ? c->depth[f] + g.off[l] : nullptr;
            a.gcol = c->gcol + g.off[l];
            a.X = c->ptsX + g.off[l]; a.Y = c->ptsY + g.off[l]; a.Z = c->ptsZ + g.off[l];
            a.npts = c->npts + l; a.nedge = c->nedge + (size_t)f * g.Bmax * g.L + l;
            a.w = g.w[l]; a.h = g.h[l]; a.P = g.P[l]; a.L = g.L;
            a.do_points = (f == DVO_FRAME_REF); a.do_cols = (f == DVO_FRAME_NOW);
            if (a.do_points && !a.depth) { dvo_set_error("canny: reference depth missing"); return DVO_ERR_STATE; }
            const float scaleFac = (float)ldexp(1.0, -l);                                  // src/SolveDVO.cpp:231
            a.tmpfx = (float)(1. / (double)(scaleFac * c->K.fx));                          // :232
            a.tmpfy = (float)(1. / (double)(scaleFac * c->K.fy));                          // :233
            a.tmpcx = scaleFac * c->K.cx; a.tmpcy = scaleFac * c->K.cy;                    // :234-235
            a.gscratch = use_global ? c->bitmap_scratch : nullptr;
            a.gscratch_stride = (long long)c->bitmap_scratch_words;
            a.first = first; a.low = 10000; a.high = 22500;
            const size_t dyn = use_global ? 0 : smem;
            if (g.P[l] >= 64 * 1024) {
                if (dyn > 48 * 1024) DVO_CUDA(cudaFuncSetAttribute(canny_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                canny_kernel<512><<<count, 512, dyn, c->stream>>>(a);
            } else if (g.P[l] >= 8 * 1024) {
                if (dyn > 48 * 1024) DVO_CUDA(cudaFuncSetAttribute(canny_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
                canny_kernel<256><<<count, 256, dyn, c->stream>>>(a);
            } else {
                canny_kernel<128><<<count, 128, dyn, c->stream>>>(a);
            }
            c->launches++;
        }
    }
    DVO_CUDA(cudaGetLastError());
    return DVO_OK;
}

// =====================================================================================================
// EDT phase 2 (rows): d2(y,x) = min_x' ( gcol(y,x')^2 + (x-x')^2 ), exact in integers.
// The cost matrix M[x][x'] = f(x') + (x-x')^2 is Monge, so the leftmost argmin is non-decreasing in x.  One
// warp per row resolves positions in breadth-first bisection order: at bisection step s the positions
// p = s*(2j+1) (1-based) are searched only between the argmins of their already-solved neighbours p-s and p+s,
// so each level costs O(w) evaluations and the whole row O(w log w), with no sequential stack as in the
// Meijster/Felzenszwalb scan.  Result is independent of evaluation order (pure min) -> bit-exact.
// =====================================================================================================
struct EdtArgs {
    const uint16_t* gcol; int32_t* d2; unsigned* maxd2;   // level regions; maxd2 + level, stride L
    int w, h, P, L, first;
};

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) edt_rows_kernel(EdtArgs a) {
    extern __shared__ int smem_i32[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = a.w;
    const int row = blockIdx.x * WARPS + warp;
    const int b = a.first + blockIdx.y;
    if (row >= a.h) return;                      // whole warp exits together; only __syncwarp below
    int* g2 = smem_i32 + (size_t)warp * (3 * w);
    int* dv = g2 + w;
    int* arg = dv + w;
    const uint16_t* __restrict__ gr = a.gcol + (long long)b * a.P + (long long)row * w;
    for (int x = lane; x < w; x += 32) { const int gv = gr[x]; g2[x] = gv * gv; }
    __syncwarp();
    int n = 1; while (n < w + 1) n <<= 1;
    for (int step = n >> 1; step >= 1; step >>= 1) {
        const int cnt = ((w / step) + 1) >> 1;   // positions p = step*(2j+1) <= w, j = 0..cnt-1
        if (cnt == 0) continue;
        if (cnt >= 32) {
            for (int j = lane; j < cnt; j += 32) {
                const int p = step * (2 * j + 1);
                const int lo = (p - step >= 1) ? arg[p - step - 1] : 0;
                const int hi = (p + step <= w) ? arg[p + step - 1] : w - 1;
                const int m = p - 1;
                int best = 0x7fffffff, bx = lo;
                for (int xq = lo; xq <= hi; ++xq) {
                    const int dxx = m - xq; const int v = g2[xq] + dxx * dxx;
                    if (v < best) { best = v; bx = xq; }
                }
                dv[m] = best; arg[m] = bx;
            }
        } else {
            int c2 = 1; while (c2 < cnt) c2 <<= 1;
            const int G = 32 / c2;               // lanes per position
            const int grp = lane / G, sub = lane - grp * G;
            unsigned long long key = 0xffffffffffffffffull;
            int m = 0;
            if (grp < cnt) {
                const int p = step * (2 * grp + 1);
                const int lo = (p - step >= 1) ? arg[p - step - 1] : 0;
                const int hi = (p + step <= w) ? arg[p + step - 1] : w - 1;
                m = p - 1;
                int best = 0x7fffffff, bx = lo;
                for (int xq = lo + sub; xq <= hi; xq += G) {
                    const int dxx = m - xq; const int v = g2[xq] + dxx * dxx;
                    if (v < best) { best = v; bx = xq; }
                }
                key = ((unsigned long long)(unsigned)best << 32) | (unsigned)bx;
            }
            for (int o = G >> 1; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other < key ? other : key;
            }
            if (grp < cnt && sub == 0) { dv[m] = (int)(key >> 32); arg[m] = (int)(key & 0xffffffffu); }
        }
        __syncwarp();
    }
    int32_t* out = a.d2 + (long long)b * a.P + (long long)row * w;
    int mx = 0;
    for (int x = lane; x < w; x += 32) { const int v = dv[x]; out[x] = v; mx = max(mx, v); }
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if (lane == 0) atomicMax(&a.maxd2[(long long)b * a.L], (unsigned)mx);
}

int launch_edt_rows(dvo_ctx* c, int first, int count) {
    const PyrGeom& g = c->geom;
    constexpr int WARPS = 8;
    DVO_CUDA(cudaMemsetAsync(c->maxd2 + (size_t)first * g.L, 0, sizeof(unsigned) * (size_t)count * g.L, c->stream));
    for (int l = 0; l < g.L; ++l) {
        EdtArgs a; a.gcol = c->gcol + g.off[l]; a.d2 = c->d2 + g.off[l]; a.maxd2 = c->maxd2 + l;
        a.w = g.w[l]; a.h = g.h[l]; a.P = g.P[l]; a.L = g.L; a.first = first;
        const size_t smem = (size_t)WARPS * 3 * g.w[l] * sizeof(int);
        if (smem > 48 * 1024) DVO_CUDA(cudaFuncSetAttribute(edt_rows_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((g.h[l] + WARPS - 1) / WARPS, count);
        edt_rows_kernel<WARPS><<<grid, WARPS * 32, smem, c->stream>>>(a);
        c->launches++;
    }
    DVO_CUDA(cudaGetLastError());
    return DVO_OK;
}

// =====================================================================================================
// normalise + gradient (src/SolveDVO.cpp:1774, 1790, 1063-1098):
//   DT = sqrtf(d2) (IEEE), DTn = DT * (float)(255 * (1/(max-min))) + 0, gx/gy = 0.5*(next) - 0.5*(prev) with
//   REFLECT_101 (exactly 0 on the border).  Output packed as one 16-byte texel {DTn, gx, gy, 0} so that the
//   solver does a single 128-bit gather per reprojected point.
// =====================================================================================================
struct NormArgs { const int32_t* d2; float4* texel; const unsigned* maxd2; const unsigned* nedge; int w, h, P, L, first; };

__global__ void __launch_bounds__(256) normgrad_kernel(NormArgs a) {
    const int b = a.first + blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.P) return;
    const int w = a.w, h = a.h;
    const int y = q / w, x = q - y * w;
    const unsigned mx2 = a.maxd2[(long long)b * a.L];
    const unsigned ne = a.nedge[(long long)b * a.L];
    float scale = 0.0f;
    if (ne > 0u) {
        const double smax = (double)__fsqrt_rn((float)mx2);          // smin = 0 whenever an edge exists
        const double sc = 255.0 * ((smax > 2.220446049250313e-16) ? __ddiv_rn(1.0, smax) : 0.0);
        scale = __double2float_rn(sc);
    }
    const int32_t* __restrict__ d = a.d2 + (long long)b * a.P;
    auto dtn = [&](int yy, int xx) -> float { return __fmul_rn(__fsqrt_rn((float)d[yy * w + xx]), scale); };
    int xl = x - 1 < 0 ? 1 : x - 1, xr = x + 1 >= w ? w - 2 : x + 1;
    int yu = y - 1 < 0 ? 1 : y - 1, yd = y + 1 >= h ? h - 2 : y + 1;
    if (w == 1) xl = xr = 0;
    if (h == 1) yu = yd = 0;
    float4 t;
    t.x = dtn(y, x);
    t.y = __fadd_rn(__fmul_rn(-0.5f, dtn(y, xl)), __fmul_rn(0.5f, dtn(y, xr)));
    t.z = __fadd_rn(__fmul_rn(-0.5f, dtn(yu, x)), __fmul_rn(0.5f, dtn(yd, x)));
    t.w = 0.0f;
    a.texel[(long long)b * a.P + q] = t;
}

int launch_normgrad(dvo_ctx* c, int first, int count) {
    const PyrGeom& g = c->geom;
    for (int l = 0; l < g.L; ++l) {
        NormArgs a; a.d2 = c->d2 + g.off[l]; a.texel = c->texel + g.off[l]; a.maxd2 = c->maxd2 + l;
        a.nedge = c->nedge + (size_t)DVO_FRAME_NOW * g.Bmax * g.L + l;
        a.w = g.w[l]; a.h = g.h[l]; a.P = g.P[l]; a.L = g.L; a.first = first;
        dim3 grid((g.P[l] + 255) / 256, count);
        normgrad_kernel<<<grid, 256, 0, c->stream>>>(a);
        c->launches++;
    }
    DVO_CUDA(cudaGetLastError());
    return DVO_OK;
}

// =====================================================================================================
// setPrevFrameAsRefFrame (src/SolveDVO.cpp:561-584): move the now frame's pyramid into the reference slot.
// =====================================================================================================
int launch_promote(dvo_ctx* c, int first, int count) {
    const PyrGeom& g = c->geom;
    if (!c->depth[DVO_FRAME_NOW]) { dvo_set_error("promote: context was created without keep_now_depth"); return DVO_ERR_STATE; }
    for (int l = 0; l < g.L; ++l) {
        const long long o = lvl_at(g, l, first); const size_t n = (size_t)count * g.P[l];
        DVO_CUDA(cudaMemcpyAsync(c->gray[0] + o, c->gray[1] + o, n, cudaMemcpyDeviceToDevice, c->stream));
        DVO_CUDA(cudaMemcpyAsync(c->depth[0] + o, c->depth[1] + o, n * 2, cudaMemcpyDeviceToDevice, c->stream));
    }
    return DVO_OK;
}
