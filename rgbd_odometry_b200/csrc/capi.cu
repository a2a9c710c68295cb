// capi.cu -- the extern "C" boundary declared in include/dvo_b200.h: context lifetime, device memory arena,
// host<->device staging and the stage sequencing.  No CPU compute path exists here: every entry point either
// launches the sm_100a kernels or fails.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"

static thread_local char g_err[512] = "";

void dvo_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

// cv::resize output size: cvRound(dim * 2^-level), round half to even
static int level_dim(int dim, int level) { return (int)nearbyint((double)dim * ldexp(1.0, -level)); }

namespace {
struct StageTimer {
    dvo_ctx* c; int stage; bool on;
    StageTimer(dvo_ctx* c_, int s) : c(c_), stage(s), on(c_->timing) { if (on) cudaEventRecord(c->ev_a, c->stream); }
    ~StageTimer() {
        if (!on) return;
        cudaEventRecord(c->ev_b, c->stream); cudaEventSynchronize(c->ev_b);
        float ms = 0.f; cudaEventElapsedTime(&ms, c->ev_a, c->ev_b); c->stage_ms[stage] += ms;
    }
};
template <typename T> int dalloc(T** p, size_t n) {
    if (n == 0) { *p = nullptr; return DVO_OK; }
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) { dvo_set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e)); return DVO_ERR_NOMEM; }
    return DVO_OK;
}
bool range_ok(dvo_ctx* c, int first, int count) { return c && first >= 0 && count >= 0 && first + count <= c->cfg.max_batch; }
// make the context stream wait for whatever dvo_process left running on the internal streams
int join_aux(dvo_ctx* c) {
    if (!c || !c->aux_pending) return DVO_OK;
    c->aux_pending = false;
    for (int k = 0; k < 2; ++k) DVO_CUDA(cudaStreamWaitEvent(c->stream, c->ev_aux_done[k], 0));
    return DVO_OK;
}
}  // namespace

extern "C" {

const char* dvo_last_error(void) { return g_err; }

int dvo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// inside dvo_create: any CUDA failure after the context object exists releases everything acquired so far
#define CREATE_CUDA(call)                                                                       \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            dvo_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            dvo_destroy(c);                                                                     \
            return DVO_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

int dvo_create(const dvo_config* cfg, dvo_ctx** out) {
    if (!cfg || !out) { dvo_set_error("dvo_create: null argument"); return DVO_ERR_ARG; }
    // max_batch <= 65535: several kernels carry the slot index in gridDim.y / gridDim.z
    if (cfg->width < 8 || cfg->height < 8 || cfg->levels < 1 || cfg->levels > DVO_MAX_LEVELS || cfg->max_batch < 1 || cfg->max_batch > 65535 ||
        cfg->width >= DVO_EDT_INF_1D || cfg->height >= DVO_EDT_INF_1D) {
        dvo_set_error("dvo_create: bad config %dx%d levels=%d max_batch=%d", cfg->width, cfg->height, cfg->levels, cfg->max_batch);
        return DVO_ERR_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        dvo_set_error("dvo_create: no usable CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
        cudaGetLastError();
        return DVO_ERR_CUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { dvo_set_error("dvo_create: device %d out of range (%d devices)", cfg->device, ndev); return DVO_ERR_ARG; }
    DVO_CUDA(cudaSetDevice(cfg->device));
    if (const char* e = getenv("DVO_L2_FETCH")) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)); cudaGetLastError(); }   // experiment knob
    dvo_ctx* c = new (std::nothrow) dvo_ctx();
    if (!c) return DVO_ERR_NOMEM;
    memset(c, 0, sizeof(*c));
    c->cfg = *cfg;
    PyrGeom& g = c->geom;
    g.L = cfg->levels; g.Bmax = cfg->max_batch;
    long long acc = 0; int maxP = 0;
    for (int l = 0; l < g.L; ++l) {
        g.w[l] = level_dim(cfg->width, l); g.h[l] = level_dim(cfg->height, l);
        if (g.w[l] < 2 || g.h[l] < 2) { dvo_set_error("dvo_create: level %d is %dx%d (too small)", l, g.w[l], g.h[l]); dvo_destroy(c); return DVO_ERR_ARG; }
        g.P[l] = g.w[l] * g.h[l]; g.off[l] = acc * g.Bmax; acc += g.P[l];
        if (g.P[l] > maxP) maxP = g.P[l];
    }
    g.total = acc * g.Bmax;
    long long acct = 0;
    for (int l = 0; l < g.L; ++l) {
        g.tw[l] = (g.w[l] + 3) >> 2; g.Pt[l] = g.tw[l] * ((g.h[l] + 3) >> 2) * 16;
        g.offt[l] = acct * g.Bmax; acct += g.Pt[l];
    }
    g.total_t = acct * g.Bmax;
    cudaDeviceProp prop;
    CREATE_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    c->solve_shape = (cfg->max_batch <= c->sm_count) ? 512 : 256;
    if (const char* e = getenv("DVO_SOLVE_SHAPE")) { const int v = atoi(e); if (v == 256 || v == 512) c->solve_shape = v; }   // experiment knob
    c->texel_mode = 1;
    c->edt_band = 0;
    if (const char* e = getenv("DVO_EDT_BAND")) c->edt_band = atoi(e);                                                       // experiment knob
    if (const char* e = getenv("DVO_TEXEL_MODE")) c->texel_mode = atoi(e) ? 1 : 0;                                           // experiment knob (A/B)
    CREATE_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CREATE_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 16; ++i) CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_entry, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) {
        CREATE_CUDA(cudaStreamCreateWithFlags(&c->aux[k], cudaStreamNonBlocking));
        CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_pre[k], cudaEventDisableTiming));
        CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_aux_done[k], cudaEventDisableTiming));
    }
    CREATE_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    c->e2e_chunk = 128;      // 8 pipeline stages per 1024 pairs: 25.7 ms per pass against 27.2 (256), 30.2 (512), 33.3 (64)
    if (const char* e = getenv("DVO_E2E_CHUNK")) { const int v = atoi(e); if (v > 0) c->e2e_chunk = v; }
    CREATE_CUDA(cudaEventCreate(&c->ev_a));
    CREATE_CUDA(cudaEventCreate(&c->ev_b));

    const size_t T = (size_t)g.total, B = (size_t)g.Bmax;
    int rc = DVO_OK;
    auto A = [&](int r) { if (rc == DVO_OK) rc = r; };
    for (int f = 0; f < 2; ++f) { A(dalloc(&c->gray[f], T)); A(dalloc(&c->edge[f], T)); }
    A(dalloc(&c->depth[0], T));
    if (cfg->keep_now_depth) { A(dalloc(&c->depth[1], T)); A(dalloc(&c->prev_gray, B * (size_t)g.P[0])); A(dalloc(&c->prev_depth, B * (size_t)g.P[0])); }
    c->now_valid = (unsigned char*)calloc(B, 1); c->prev_valid = (unsigned char*)calloc(B, 1);
    A(dalloc(&c->gcol, T)); A(dalloc(&c->d2, T));
    if (c->texel_mode) A(dalloc(&c->tex8, (size_t)g.total_t)); else A(dalloc(&c->texel, T));
    A(dalloc(&c->ptsX, T)); A(dalloc(&c->ptsY, T)); A(dalloc(&c->ptsZ, T)); A(dalloc(&c->ptsPix, T));
    A(dalloc(&c->npts, B * g.L)); A(dalloc(&c->solve_order, B)); A(dalloc(&c->seq_mask, B)); A(dalloc(&c->seq_state, B)); A(dalloc(&c->nedge, 2 * B * g.L)); A(dalloc(&c->maxd2, B * g.L));
    A(dalloc(&c->pose0, B * 12)); A(dalloc(&c->pose, B * 12)); A(dalloc(&c->info, B));
    if (cfg->trace_iters > 0) A(dalloc(&c->trace, B * g.L * cfg->trace_iters * DVO_TRACE_DOUBLES));
    A(dalloc(&c->energy, B * g.L * DVO_ENERGY_ITERS));
    // candidate / edge bitmaps: sobel_nms_kernel -> canny_kernel (which keeps working in them when they do not fit in shared memory)
    canny_scratch_layout(g, c->bm_off, c->bm_words, &c->bitmap_scratch_words);
    A(dalloc(&c->bitmap_scratch, c->bitmap_scratch_words * B * 2));
    if (rc != DVO_OK) { dvo_destroy(c); return rc; }
    CREATE_CUDA(cudaMemsetAsync(c->npts, 0, sizeof(int) * B * g.L, c->stream));
    CREATE_CUDA(cudaMemsetAsync(c->bitmap_scratch, 0, sizeof(uint32_t) * c->bitmap_scratch_words * B * 2, c->stream));   // the bitmaps' zero borders are never written
    CREATE_CUDA(cudaMemsetAsync(c->nedge, 0, sizeof(unsigned) * 2 * B * g.L, c->stream));
    CREATE_CUDA(cudaMemsetAsync(c->maxd2, 0, sizeof(unsigned) * B * g.L, c->stream));
    CREATE_CUDA(cudaMemsetAsync(c->info, 0, sizeof(dvo_pair_info) * B, c->stream));
    CREATE_CUDA(cudaMallocHost((void**)&c->h_pose, sizeof(double) * 12 * B));
    CREATE_CUDA(cudaMallocHost((void**)&c->h_info, sizeof(dvo_pair_info) * B));
    dvo_set_initial_pose(c, 0, g.Bmax, nullptr, DVO_MEM_HOST);
    CREATE_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
    return DVO_OK;
}

int dvo_destroy(dvo_ctx* c) {
    if (!c) return DVO_OK;
    cudaSetDevice(c->cfg.device);
    if (c->own_stream) cudaStreamSynchronize(c->own_stream);
    for (int f = 0; f < 2; ++f) { cudaFree(c->gray[f]); cudaFree(c->depth[f]); cudaFree(c->edge[f]); }
    cudaFree(c->gcol); cudaFree(c->d2); cudaFree(c->texel); cudaFree(c->tex8); cudaFree(c->ptsX); cudaFree(c->ptsY); cudaFree(c->ptsZ); cudaFree(c->ptsPix);
    cudaFree(c->npts); cudaFree(c->solve_order); cudaFree(c->seq_mask); cudaFree(c->seq_state); cudaFree(c->nedge); cudaFree(c->maxd2); cudaFree(c->pose0); cudaFree(c->pose); cudaFree(c->info);
    cudaFree(c->seq_rel); cudaFree(c->seq_kind); cudaFree(c->seq_reason); cudaFree(c->seq_glob);
    for (int k = 0; k < 2; ++k) {
        cudaFree(c->seq_stage_g[k]); cudaFree(c->seq_stage_d[k]);
        if (c->seq_ev_ready[k]) cudaEventDestroy(c->seq_ev_ready[k]);
        if (c->seq_ev_free[k]) cudaEventDestroy(c->seq_ev_free[k]);
    }
    cudaFree(c->trace); cudaFree(c->energy); cudaFree(c->bitmap_scratch); cudaFree(c->prev_gray); cudaFree(c->prev_depth); cudaFree(c->raw_bgr); cudaFree(c->raw_depth);
    free(c->now_valid); free(c->prev_valid);
    if (c->h_pose) cudaFreeHost(c->h_pose);
    if (c->h_info) cudaFreeHost(c->h_info);
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    for (int k = 0; k < 2; ++k) {
        if (c->aux[k]) { cudaStreamSynchronize(c->aux[k]); cudaStreamDestroy(c->aux[k]); }
        if (c->ev_pre[k]) cudaEventDestroy(c->ev_pre[k]);
        if (c->ev_aux_done[k]) cudaEventDestroy(c->ev_aux_done[k]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int i = 0; i < 16; ++i) if (c->ev_chunk[i]) cudaEventDestroy(c->ev_chunk[i]);
    if (c->ev_entry) cudaEventDestroy(c->ev_entry);
    delete c;
    return DVO_OK;
}

int dvo_set_stream(dvo_ctx* c, void* s) {
    if (!c) return DVO_ERR_ARG;
    { int rc = join_aux(c); if (rc) return rc; }
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return DVO_OK;
}

int dvo_synchronize(dvo_ctx* c) {
    if (!c) return DVO_ERR_ARG;
    { int rc = join_aux(c); if (rc) return rc; }
    DVO_CUDA(cudaStreamSynchronize(c->stream));
    return DVO_OK;
}

int dvo_set_intrinsics(dvo_ctx* c, float fx, float fy, float cx, float cy) {
    if (!c) return DVO_ERR_ARG;
    c->K.fx = fx; c->K.fy = fy; c->K.cx = cx; c->K.cy = cy; c->haveK = true;
    return DVO_OK;
}

}  // extern "C"

// Validation and p_now_* bookkeeping shared by dvo_set_frames / dvo_set_frames_raw.  Every argument is checked before any
// copy is enqueued or any flag changes, so an error return leaves the context as it was.
static int begin_set_frames(dvo_ctx* c, const char* who, int frame, int first, int count, bool have_gray, bool have_depth) {
    if (!range_ok(c, first, count) || (frame != DVO_FRAME_REF && frame != DVO_FRAME_NOW) || !have_gray) { dvo_set_error("%s: bad argument", who); return DVO_ERR_ARG; }
    if (frame == DVO_FRAME_REF && !have_depth) { dvo_set_error("%s: the reference frame needs depth", who); return DVO_ERR_ARG; }
    bool rotate = false;
    if (frame == DVO_FRAME_NOW && c->prev_gray) {
        if (!have_depth) { dvo_set_error("%s: a context with keep_now_depth needs the now frame's depth", who); return DVO_ERR_ARG; }
        bool any = false, all = true;
        for (int i = first; i < first + count; ++i) { any |= c->now_valid[i] != 0; all &= c->now_valid[i] != 0; }
        if (any && !all) { dvo_set_error("%s: slots [%d,%d) mix first and subsequent now frames", who, first, first + count); return DVO_ERR_STATE; }
        rotate = all && count > 0;
    }
    const size_t P0 = c->geom.P[0];
    if (rotate) {
        // setRcvdFrameAsNowFrame keeps the outgoing now frame as p_now_* (src/SolveDVO.cpp:594-600)
        DVO_CUDA(cudaMemcpyAsync(c->prev_gray + P0 * first, c->gray[1] + lvl_at(c->geom, 0, first), P0 * count, cudaMemcpyDeviceToDevice, c->stream));
        DVO_CUDA(cudaMemcpyAsync(c->prev_depth + P0 * first, c->depth[1] + lvl_at(c->geom, 0, first), P0 * count * 2, cudaMemcpyDeviceToDevice, c->stream));
        for (int i = first; i < first + count; ++i) c->prev_valid[i] = 1;
    }
    if (frame == DVO_FRAME_NOW) for (int i = first; i < first + count; ++i) c->now_valid[i] = 1;
    return DVO_OK;
}

extern "C" {

int dvo_set_frames(dvo_ctx* c, int frame, int first, int count, const uint8_t* gray, const uint16_t* depth, int mem) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    { const int rc = begin_set_frames(c, "dvo_set_frames", frame, first, count, gray != nullptr, depth != nullptr); if (rc) return rc; }
    StageTimer t(c, DVO_STAGE_H2D);
    const size_t P0 = c->geom.P[0];
    const cudaMemcpyKind k = mem == DVO_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    DVO_CUDA(cudaMemcpyAsync(c->gray[frame] + lvl_at(c->geom, 0, first), gray, P0 * count, k, c->stream));
    if (depth && c->depth[frame])
        DVO_CUDA(cudaMemcpyAsync(c->depth[frame] + lvl_at(c->geom, 0, first), depth, P0 * count * sizeof(uint16_t), k, c->stream));
    return DVO_OK;
}

int dvo_set_frames_raw(dvo_ctx* c, int frame, int first, int count, const uint8_t* bgr, const float* depth_m, int mem) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    { const int rc = begin_set_frames(c, "dvo_set_frames_raw", frame, first, count, bgr != nullptr, depth_m != nullptr); if (rc) return rc; }
    if (count == 0) return DVO_OK;
    StageTimer t(c, DVO_STAGE_H2D);
    const size_t P0 = c->geom.P[0];
    const uint8_t* d_bgr = bgr; const float* d_dm = depth_m;
    if (mem != DVO_MEM_DEVICE) {
        if (c->raw_capacity < (size_t)count) {           // staging sized for the largest range seen so far
            DVO_CUDA(cudaStreamSynchronize(c->stream));
            cudaFree(c->raw_bgr); cudaFree(c->raw_depth); c->raw_bgr = nullptr; c->raw_depth = nullptr; c->raw_capacity = 0;
            int rc = dalloc(&c->raw_bgr, (size_t)count * P0 * 3); if (rc) return rc;
            rc = dalloc(&c->raw_depth, (size_t)count * P0); if (rc) return rc;
            c->raw_capacity = (size_t)count;
        }
        DVO_CUDA(cudaMemcpyAsync(c->raw_bgr, bgr, P0 * count * 3, cudaMemcpyHostToDevice, c->stream));
        if (depth_m) DVO_CUDA(cudaMemcpyAsync(c->raw_depth, depth_m, P0 * count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        d_bgr = c->raw_bgr; d_dm = depth_m ? c->raw_depth : nullptr;
    }
    return launch_ingest_raw(c, frame, first, count, d_bgr, d_dm);
}

int dvo_promote_now_to_ref(dvo_ctx* c, int first, int count) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, first, count)) return DVO_ERR_ARG;
    return launch_promote(c, first, count);
}

int dvo_build_pyramids(dvo_ctx* c, int first, int count, int frames_mask) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, first, count)) { dvo_set_error("dvo_build_pyramids: bad range"); return DVO_ERR_ARG; }
    if (count == 0) return DVO_OK;
    StageTimer t(c, DVO_STAGE_PYRAMID);
    return launch_pyramid(c, first, count, frames_mask);
}

int dvo_prepare(dvo_ctx* c, int first, int count, int frames_mask) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, first, count)) { dvo_set_error("dvo_prepare: bad range"); return DVO_ERR_ARG; }
    if (!c->haveK) { dvo_set_error("dvo_prepare: intrinsics not set (SolveDVO asserts isCameraIntrinsicsAvailable)"); return DVO_ERR_STATE; }
    if (count == 0) return DVO_OK;
    int rc;
    { StageTimer t(c, DVO_STAGE_CANNY); rc = launch_canny(c, first, count, frames_mask); if (rc) return rc; }
    if (frames_mask & 2) {
        if (c->texel_mode && c->edt_band >= 0) {       // fused: the stage slot of the row pass times both
            StageTimer t(c, DVO_STAGE_EDT_ROWS); rc = launch_edt_pack(c, first, count); if (rc) return rc;
        } else {
            { StageTimer t(c, DVO_STAGE_EDT_ROWS); rc = launch_edt_rows(c, first, count); if (rc) return rc; }
            { StageTimer t(c, DVO_STAGE_NORMGRAD); rc = c->texel_mode ? launch_pack(c, first, count) : launch_normgrad(c, first, count); if (rc) return rc; }
        }
    }
    return DVO_OK;
}

int dvo_set_initial_pose(dvo_ctx* c, int first, int count, const double* R9T3, int mem) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, first, count)) return DVO_ERR_ARG;
    if (R9T3) {
        DVO_CUDA(cudaMemcpyAsync(c->pose0 + 12 * (size_t)first, R9T3, sizeof(double) * 12 * count,
                                 mem == DVO_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
        if (mem != DVO_MEM_DEVICE) DVO_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        std::vector<double> id((size_t)12 * count, 0.0);
        for (int i = 0; i < count; ++i) { id[12 * (size_t)i] = 1.0; id[12 * (size_t)i + 4] = 1.0; id[12 * (size_t)i + 8] = 1.0; }
        DVO_CUDA(cudaMemcpyAsync(c->pose0 + 12 * (size_t)first, id.data(), sizeof(double) * 12 * count, cudaMemcpyHostToDevice, c->stream));
        DVO_CUDA(cudaStreamSynchronize(c->stream));
    }
    return DVO_OK;
}

int dvo_run(dvo_ctx* c, int first, int count, const dvo_solver_params* p) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, first, count) || !p) { dvo_set_error("dvo_run: bad argument"); return DVO_ERR_ARG; }
    if (!c->haveK) { dvo_set_error("dvo_run: intrinsics not set"); return DVO_ERR_STATE; }
    if (count == 0) return DVO_OK;
    StageTimer t(c, DVO_STAGE_SOLVE);
    return launch_solve(c, first, count, p);
}

// pyramids + Canny / EDT / normalise + solve of one slot range on the CURRENT c->stream, no stage timers, no join
static int process_range(dvo_ctx* c, int first, int count, const dvo_solver_params* p, cudaEvent_t pre_done) {
    int rc;
    if ((rc = launch_pyramid(c, first, count, 3))) return rc;
    if ((rc = launch_canny(c, first, count, 3))) return rc;
    if (c->texel_mode && c->edt_band >= 0) { if ((rc = launch_edt_pack(c, first, count))) return rc; }
    else {
        if ((rc = launch_edt_rows(c, first, count))) return rc;
        if ((rc = c->texel_mode ? launch_pack(c, first, count) : launch_normgrad(c, first, count))) return rc;
    }
    if (pre_done) DVO_CUDA(cudaEventRecord(pre_done, c->stream));
    return launch_solve(c, first, count, p);
}

int dvo_process(dvo_ctx* c, int first, int count, const dvo_solver_params* p, double* d_poses) {
    if (!range_ok(c, first, count) || !p) { dvo_set_error("dvo_process: bad argument"); return DVO_ERR_ARG; }
    if (!c->haveK) { dvo_set_error("dvo_process: intrinsics not set"); return DVO_ERR_STATE; }
    if (count == 0) return DVO_OK;
    // Per-stage timers synchronise, and a half batch must still fill the chip for the overlap to pay: otherwise run staged.
    if (c->timing || count < 4 * c->sm_count) {
        int rc = join_aux(c); if (rc) return rc;
        if ((rc = dvo_build_pyramids(c, first, count, 3))) return rc;
        if ((rc = dvo_prepare(c, first, count, 3))) return rc;
        if ((rc = dvo_run(c, first, count, p))) return rc;
        if (d_poses) DVO_CUDA(cudaMemcpyAsync(d_poses, c->pose + 12 * (size_t)first, sizeof(double) * 12 * count, cudaMemcpyDeviceToDevice, c->stream));
        return DVO_OK;
    }
    // Fork: both internal streams start after everything issued on the context stream so far.  Half 1's preprocessing
    // additionally waits for half 0's, so it runs beside half 0's solve; each internal stream is in order with its own
    // work of the previous call, which is all the dependency there is (the halves own disjoint slots).
    // A call whose slot split differs from the one still in flight could touch slots the OTHER internal stream is working
    // on: join first (device-side) in that case.  With the same (first, count) every stream only meets its own slots.
    if (c->aux_pending && (c->proc_first != first || c->proc_count != count)) { const int rc = join_aux(c); if (rc) return rc; }
    c->proc_first = first; c->proc_count = count;
    DVO_CUDA(cudaEventRecord(c->ev_fork, c->stream));
    const cudaStream_t user = c->stream;
    const int n0 = count / 2;
    const int starts[2] = {first, first + n0}, counts[2] = {n0, count - n0};
    int rc = DVO_OK;
    for (int k = 0; k < 2 && rc == DVO_OK; ++k) {
        cudaError_t e = cudaStreamWaitEvent(c->aux[k], c->ev_fork, 0);
        if (e == cudaSuccess && k == 1) e = cudaStreamWaitEvent(c->aux[1], c->ev_pre[0], 0);
        if (e != cudaSuccess) { dvo_set_error("dvo_process: %s", cudaGetErrorString(e)); rc = DVO_ERR_CUDA; break; }
        c->stream = c->aux[k];
        rc = process_range(c, starts[k], counts[k], p, c->ev_pre[k]);
        if (rc == DVO_OK && d_poses && cudaMemcpyAsync(d_poses + 12 * (size_t)(starts[k] - first), c->pose + 12 * (size_t)starts[k], sizeof(double) * 12 * counts[k],
                                                       cudaMemcpyDeviceToDevice, c->aux[k]) != cudaSuccess) rc = DVO_ERR_CUDA;
        if (rc == DVO_OK && cudaEventRecord(c->ev_aux_done[k], c->aux[k]) != cudaSuccess) rc = DVO_ERR_CUDA;
    }
    c->stream = user;
    c->aux_pending = true;
    if (rc != DVO_OK) join_aux(c);
    return rc;
}

int dvo_join(dvo_ctx* c) {
    if (!c) return DVO_ERR_ARG;
    return join_aux(c);
}

int dvo_join_stream(dvo_ctx* c, void* s) {
    if (!c) return DVO_ERR_ARG;
    if (!c->aux_pending) return DVO_OK;
    for (int k = 0; k < 2; ++k) DVO_CUDA(cudaStreamWaitEvent((cudaStream_t)s, c->ev_aux_done[k], 0));
    return DVO_OK;
}

int dvo_get_poses(dvo_ctx* c, int first, int count, double* R9T3, dvo_pair_info* info, int mem) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, first, count)) return DVO_ERR_ARG;
    StageTimer t(c, DVO_STAGE_D2H);
    const cudaMemcpyKind k = mem == DVO_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (R9T3) DVO_CUDA(cudaMemcpyAsync(R9T3, c->pose + 12 * (size_t)first, sizeof(double) * 12 * count, k, c->stream));
    if (info) DVO_CUDA(cudaMemcpyAsync(info, c->info + first, sizeof(dvo_pair_info) * count, k, c->stream));
    if (mem != DVO_MEM_DEVICE) DVO_CUDA(cudaStreamSynchronize(c->stream));
    return DVO_OK;
}

int dvo_align_batch(dvo_ctx* c, int count, const uint8_t* ref_gray, const uint16_t* ref_depth, const uint8_t* now_gray,
                    const uint16_t* now_depth, const dvo_solver_params* p, double* R9T3, dvo_pair_info* info) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!c || count < 0 || !ref_gray || !ref_depth || !now_gray || !p) { dvo_set_error("dvo_align_batch: bad argument"); return DVO_ERR_ARG; }
    if (!c->haveK) { dvo_set_error("dvo_align_batch: intrinsics not set"); return DVO_ERR_STATE; }
    const size_t P0 = c->geom.P[0];
    const int B = c->cfg.max_batch;
    const bool timing = c->timing;
    c->timing = false;                                   // the per-stage timers synchronise; keep the pipeline asynchronous
    int rc = DVO_OK;
    for (int done = 0; done < count && rc == DVO_OK; done += B) {
        const int n = (count - done < B) ? count - done : B;
        if ((rc = dvo_set_initial_pose(c, 0, n, nullptr, DVO_MEM_HOST))) break;
        // Uploads run on their own stream, chunk by chunk, so that the copy of chunk k+1 overlaps the kernels of chunk k.
        // Every chunk owns its slots, so no buffer is reused inside one pass over the context.
        cudaError_t e = cudaEventRecord(c->ev_entry, c->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(c->copy_stream, c->ev_entry, 0);
        if (e != cudaSuccess) { dvo_set_error("dvo_align_batch: %s", cudaGetErrorString(e)); rc = DVO_ERR_CUDA; break; }
        int nchunks = (n + c->e2e_chunk - 1) / c->e2e_chunk;
        if (nchunks > 16) nchunks = 16;
        const int chunk = (n + nchunks - 1) / nchunks;
        int k = 0;
        for (int c0 = 0; c0 < n && rc == DVO_OK; c0 += chunk, ++k) {
            const int m = (n - c0 < chunk) ? n - c0 : chunk;
            const size_t src = P0 * (size_t)(done + c0);
            const long long dst = lvl_at(c->geom, 0, c0);
            e = cudaMemcpyAsync(c->gray[0] + dst, ref_gray + src, P0 * m, cudaMemcpyHostToDevice, c->copy_stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(c->depth[0] + dst, ref_depth + src, P0 * m * 2, cudaMemcpyHostToDevice, c->copy_stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(c->gray[1] + dst, now_gray + src, P0 * m, cudaMemcpyHostToDevice, c->copy_stream);
            if (e == cudaSuccess && now_depth && c->depth[1])
                e = cudaMemcpyAsync(c->depth[1] + dst, now_depth + src, P0 * m * 2, cudaMemcpyHostToDevice, c->copy_stream);
            if (e == cudaSuccess) e = cudaEventRecord(c->ev_chunk[k], c->copy_stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ev_chunk[k], 0);
            if (e != cudaSuccess) { dvo_set_error("dvo_align_batch: %s", cudaGetErrorString(e)); rc = DVO_ERR_CUDA; break; }
            if ((rc = dvo_build_pyramids(c, c0, m, 3))) break;
            if ((rc = dvo_prepare(c, c0, m, 3))) break;
            if ((rc = dvo_run(c, c0, m, p))) break;
        }
        if (rc == DVO_OK)
            rc = dvo_get_poses(c, 0, n, R9T3 ? R9T3 + 12 * (size_t)done : nullptr, info ? info + done : nullptr, DVO_MEM_HOST);
    }
    c->timing = timing;
    return rc;
}

// ---- sequence mode: SolveDVO::loop (src/SolveDVO.cpp:1896-2373) for `nseq` independent sequences in lock step ----
}  // extern "C"

namespace {
__global__ void seq_gate_kernel(int nseq, int nframes, int t, dvo_keyframe_policy pol, int finest, const dvo_pair_info* __restrict__ info,
                                const double* __restrict__ pose, double* __restrict__ pose0, int* __restrict__ lastRef,
                                unsigned char* __restrict__ mask, double* __restrict__ rel, int* __restrict__ kind, int* __restrict__ reason) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseq) return;
    const dvo_pair_info& I = info[s];
    int signal = 0, why = 0;
    if (pol.use_quality_gates) {                                                             // :2129-2151, in source order
        if (I.laplacian_b > pol.laplacian_thresh) { signal = 1; why = 2; }
        if (I.visible_ratio[finest] < pol.visible_ratio_thresh) { signal = 1; why = 3; }
        if (I.npts[finest] < pol.min_reprojections) { signal = 1; why = 4; }
    }
    if (pol.keyframe_every > 0 && (t - lastRef[s]) == pol.keyframe_every) { signal = 1; why = 5; }   // :2155-2160
    const long long cur = (long long)s * nframes + t;
    if (signal && lastRef[s] != t - 1) {                                                     // :2192
        mask[s] = 1; lastRef[s] = t - 1;
        kind[cur - 1] = 2; reason[cur - 1] = why;                                            // gop.updateMostRecentToKeyFrame(reasonForChange)
        for (int k = 0; k < 12; ++k) pose0[12 * (long long)s + k] = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0;   // cR_64 = I, cT_64 = 0 (:2203-2204)
    } else {
        mask[s] = 0;
        for (int k = 0; k < 12; ++k) { const double v = pose[12 * (long long)s + k]; rel[12 * cur + k] = v; pose0[12 * (long long)s + k] = v; }
        kind[cur] = 0; reason[cur] = 0;                                                      // gop.pushAsOrdinaryFrame (:2232)
    }
}
__global__ void seq_finish_kernel(int nseq, int nframes, int t, const unsigned char* __restrict__ mask, const double* __restrict__ pose,
                                  double* __restrict__ pose0, double* __restrict__ rel, int* __restrict__ kind, int* __restrict__ reason) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseq || !mask[s]) return;
    const long long cur = (long long)s * nframes + t;
    for (int k = 0; k < 12; ++k) { const double v = pose[12 * (long long)s + k]; rel[12 * cur + k] = v; pose0[12 * (long long)s + k] = v; }
    kind[cur] = 0; reason[cur] = 0;                                                          // pushAsOrdinaryFrame after the re-run (:2224)
}
__global__ void seq_init_kernel(int nseq, int nframes, double* __restrict__ pose0, int* __restrict__ lastRef, double* __restrict__ rel,
                                int* __restrict__ kind, int* __restrict__ reason) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseq) return;
    lastRef[s] = 0;
    for (int k = 0; k < 12; ++k) { const double v = (k == 0 || k == 4 || k == 8) ? 1.0 : 0.0; pose0[12 * (long long)s + k] = v; rel[12 * (long long)s * nframes + k] = v; }
    kind[(long long)s * nframes] = 1; reason[(long long)s * nframes] = 1;                    // gop.pushAsKeyFrame(nFrame, 1, I, 0) (:2016)
}

// One loop for both entry points.  Frame 0 is the first reference / key frame (:2013-2017); every later frame is aligned against
// the current reference with the previous frame's pose as the initial guess (:1931-1932, :2097-2104); a key-frame signal makes the
// PREVIOUS frame the reference, resets the pose and solves the frame again (:2155-2233).  Everything between frames stays on the
// device: poses, statistics, the per-slot decision (seq_gate_kernel) and the record of relative poses / kinds / reasons.
//
// Uploads are pipelined: frame t+1 of every sequence travels on the copy stream into one of two staging buffers (one strided
// 2-D copy per image plane instead of nseq small ones) while frame t is being solved; the compute stream then moves it into
// the level-0 regions with a device-to-device copy (after saving the outgoing now frame as p_now_*, :594-600).
//
// With the periodic rule alone (the shipped policy) the host knows the switch frames, so (a) the masked switch pass is only
// launched at those frames and (b) the solve against the outgoing key frame, whose result the reference discards
// (:2210-2227), is skipped.  With the quality gates the decision depends on that solve, so it runs and every frame from t = 2
// on is followed by the masked pass (its kernels return at once for unflagged slots).
int run_sequences_core(dvo_ctx* c, int nseq, int nframes, const uint8_t* gray, const uint16_t* depth, int mem, const dvo_solver_params* p,
                       const dvo_keyframe_policy* pol, double* rel_poses, int* kind, int* reason, double* global_poses) {
    const char* who = "dvo_run_sequences";
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!c || nseq < 1 || nseq > c->cfg.max_batch || nframes < 1 || !gray || !depth || !p || !pol || !rel_poses || !kind) { dvo_set_error("%s: bad argument", who); return DVO_ERR_ARG; }
    if (!c->prev_gray) { dvo_set_error("%s: create the context with keep_now_depth = 1", who); return DVO_ERR_STATE; }
    if (!c->haveK) { dvo_set_error("%s: intrinsics not set", who); return DVO_ERR_STATE; }
    int finest = -1;
    for (int l = 0; l < c->geom.L; ++l) if (p->iters[l] > 0) { finest = l; break; }
    if (finest < 0) { dvo_set_error("%s: no level has iterations", who); return DVO_ERR_ARG; }
    const size_t P0 = c->geom.P[0], n = (size_t)nseq * nframes;
    const bool host_in = (mem != DVO_MEM_DEVICE);
    int rc = DVO_OK;
    auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == DVO_OK) { dvo_set_error("%s: %s", who, cudaGetErrorString(e)); rc = DVO_ERR_CUDA; } return rc != DVO_OK; };

    // ---- buffers: record of the run (device), staging for the uploads -- owned by the context, grown on demand, reused across calls
    if (c->seq_rec_n < n) {
        cudaFree(c->seq_rel); cudaFree(c->seq_kind); cudaFree(c->seq_reason); c->seq_rel = nullptr; c->seq_kind = nullptr; c->seq_reason = nullptr; c->seq_rec_n = 0;
        fail(cudaMalloc((void**)&c->seq_rel, sizeof(double) * 12 * n)); fail(cudaMalloc((void**)&c->seq_kind, sizeof(int) * n)); fail(cudaMalloc((void**)&c->seq_reason, sizeof(int) * n));
        if (rc == DVO_OK) c->seq_rec_n = n;
    }
    if (host_in && c->seq_stage_slots < (size_t)nseq) {
        for (int k = 0; k < 2; ++k) { cudaFree(c->seq_stage_g[k]); cudaFree(c->seq_stage_d[k]); c->seq_stage_g[k] = nullptr; c->seq_stage_d[k] = nullptr; }
        c->seq_stage_slots = 0;
        for (int k = 0; k < 2 && rc == DVO_OK; ++k) {
            fail(cudaMalloc((void**)&c->seq_stage_g[k], P0 * nseq)); fail(cudaMalloc((void**)&c->seq_stage_d[k], P0 * nseq * sizeof(uint16_t)));
            if (!c->seq_ev_ready[k]) fail(cudaEventCreateWithFlags(&c->seq_ev_ready[k], cudaEventDisableTiming));
            if (!c->seq_ev_free[k]) fail(cudaEventCreateWithFlags(&c->seq_ev_free[k], cudaEventDisableTiming));
        }
        if (rc == DVO_OK) c->seq_stage_slots = (size_t)nseq;
    }
    if (rc != DVO_OK) return rc;
    double* d_rel = c->seq_rel; int* d_kind = c->seq_kind; int* d_reason = c->seq_reason; double* d_glob = nullptr;
    uint8_t* stage_g[2] = {c->seq_stage_g[0], c->seq_stage_g[1]}; uint16_t* stage_d[2] = {c->seq_stage_d[0], c->seq_stage_d[1]};
    cudaEvent_t ev_ready[2] = {c->seq_ev_ready[0], c->seq_ev_ready[1]}, ev_free[2] = {c->seq_ev_free[0], c->seq_ev_free[1]};
    fail(cudaMemsetAsync(d_rel, 0, sizeof(double) * 12 * n, c->stream));
    fail(cudaMemsetAsync(d_kind, 0, sizeof(int) * n, c->stream));
    fail(cudaMemsetAsync(d_reason, 0, sizeof(int) * n, c->stream));
    if (host_in) { fail(cudaEventRecord(c->ev_entry, c->stream)); fail(cudaStreamWaitEvent(c->copy_stream, c->ev_entry, 0)); }

    // frame t of every sequence: host -> staging buffer t & 1 on the copy stream
    bool staged_before[2] = {false, false};
    auto upload = [&](int t) {
        if (!host_in || t >= nframes || rc != DVO_OK) return;
        const int k = t & 1;
        if (staged_before[k]) fail(cudaStreamWaitEvent(c->copy_stream, ev_free[k], 0));
        fail(cudaMemcpy2DAsync(stage_g[k], P0, gray + (size_t)t * P0, (size_t)nframes * P0, P0, nseq, cudaMemcpyHostToDevice, c->copy_stream));
        fail(cudaMemcpy2DAsync(stage_d[k], P0 * 2, depth + (size_t)t * P0, (size_t)nframes * P0 * 2, P0 * 2, nseq, cudaMemcpyHostToDevice, c->copy_stream));
        fail(cudaEventRecord(ev_ready[k], c->copy_stream));
        staged_before[k] = true;
    };
    // staged (or device-resident) frame t -> level-0 regions of `frame` on the compute stream
    auto ingest = [&](int frame, int t) {
        if (rc != DVO_OK) return;
        uint8_t* dg = c->gray[frame] + lvl_at(c->geom, 0, 0); uint16_t* dd = c->depth[frame] + lvl_at(c->geom, 0, 0);
        if (frame == DVO_FRAME_NOW && c->now_valid[0]) {                                     // setRcvdFrameAsNowFrame keeps the outgoing now frame (:594-600)
            if (rc == DVO_OK) rc = launch_copy_bytes(c, c->prev_gray, dg, P0 * nseq);
            if (rc == DVO_OK) rc = launch_copy_bytes(c, c->prev_depth, dd, P0 * nseq * 2);
            for (int s2 = 0; s2 < nseq; ++s2) c->prev_valid[s2] = 1;
        }
        if (frame == DVO_FRAME_NOW) for (int s2 = 0; s2 < nseq; ++s2) c->now_valid[s2] = 1;
        if (host_in) {
            const int k = t & 1;
            fail(cudaStreamWaitEvent(c->stream, ev_ready[k], 0));
            if (rc == DVO_OK) rc = launch_copy_bytes(c, dg, stage_g[k], P0 * nseq);          // a kernel, not a D2D memcpy: the copy engines stay with the uploads
            if (rc == DVO_OK) rc = launch_copy_bytes(c, dd, stage_d[k], P0 * nseq * 2);
            fail(cudaEventRecord(ev_free[k], c->stream));
        } else {
            fail(cudaMemcpy2DAsync(dg, P0, gray + (size_t)t * P0, (size_t)nframes * P0, P0, nseq, cudaMemcpyDeviceToDevice, c->stream));
            fail(cudaMemcpy2DAsync(dd, P0 * 2, depth + (size_t)t * P0, (size_t)nframes * P0 * 2, P0 * 2, nseq, cudaMemcpyDeviceToDevice, c->stream));
        }
    };

    const int blk = (nseq + 127) / 128;
    const bool timing = c->timing;
    c->timing = false;                                                                       // per-stage timers synchronise; keep the pipeline asynchronous
    for (int s2 = 0; s2 < nseq; ++s2) { c->now_valid[s2] = 0; c->prev_valid[s2] = 0; }
    upload(0); upload(1);
    ingest(DVO_FRAME_REF, 0);
    if (rc == DVO_OK) rc = dvo_build_pyramids(c, 0, nseq, 1);
    if (rc == DVO_OK) rc = dvo_prepare(c, 0, nseq, 1);
    if (rc == DVO_OK) { seq_init_kernel<<<blk, 128, 0, c->stream>>>(nseq, nframes, c->pose0, c->seq_state, d_rel, d_kind, d_reason); c->launches++; fail(cudaGetLastError()); }
    int host_last_ref = 0;                                                                   // mirrors lastRef of every slot under the periodic rule alone
    for (int t = 1; t < nframes && rc == DVO_OK; ++t) {
        ingest(DVO_FRAME_NOW, t);                                                            // setRcvdFrameAsNowFrame
        upload(t + 1);                                                                       // next frame travels while this one is solved
        if (rc != DVO_OK) break;
        if ((rc = dvo_build_pyramids(c, 0, nseq, 2))) break;
        if ((rc = dvo_prepare(c, 0, nseq, 2))) break;
        bool switch_pass = t >= 2, first_solve = true;                                       // a previous now frame exists from t = 2 on
        if (!pol->use_quality_gates) {
            switch_pass = t >= 2 && pol->keyframe_every > 0 && (t - host_last_ref) == pol->keyframe_every && host_last_ref != t - 1;
            if (switch_pass) { host_last_ref = t - 1; first_solve = false; }
        }
        if (first_solve && (rc = dvo_run(c, 0, nseq, p))) break;                             // warm start: pose0 holds the previous frame's pose (:1931-1932)
        seq_gate_kernel<<<blk, 128, 0, c->stream>>>(nseq, nframes, t, *pol, finest, c->info, c->pose, c->pose0, c->seq_state, c->seq_mask, d_rel, d_kind, d_reason);
        c->launches++;
        if (fail(cudaGetLastError())) break;
        if (switch_pass) {
            c->active = c->seq_mask;                                                         // only the flagged slots do any work
            rc = dvo_promote_now_to_ref(c, 0, nseq);                                         // setPrevFrameAsRefFrame (:2196)
            if (rc == DVO_OK) rc = dvo_build_pyramids(c, 0, nseq, 1);
            if (rc == DVO_OK) rc = dvo_prepare(c, 0, nseq, 1);                               // computeDistTransfrmOfRef + preProcessRefFrame (:2197)
            if (rc == DVO_OK) rc = dvo_run(c, 0, nseq, p);                                   // re-run from identity (:2213-2220)
            c->active = nullptr;
            if (rc != DVO_OK) break;
            seq_finish_kernel<<<blk, 128, 0, c->stream>>>(nseq, nframes, t, c->seq_mask, c->pose, c->pose0, d_rel, d_kind, d_reason);
            c->launches++;
            if (fail(cudaGetLastError())) break;
        }
    }
    c->active = nullptr;
    c->timing = timing;
    if (rc == DVO_OK && global_poses) {
        if (c->seq_glob_n < n) {
            cudaFree(c->seq_glob); c->seq_glob = nullptr; c->seq_glob_n = 0;
            if (!fail(cudaMalloc((void**)&c->seq_glob, sizeof(double) * 19 * n))) c->seq_glob_n = n;
        }
        d_glob = c->seq_glob;
        if (rc == DVO_OK) {
            rc = launch_gop(c, nseq, nframes, d_kind, d_rel, d_glob);
            if (rc == DVO_OK) fail(cudaMemcpyAsync(global_poses, d_glob, sizeof(double) * 19 * n, cudaMemcpyDeviceToHost, c->stream));
        }
    }
    if (rc == DVO_OK) {
        fail(cudaMemcpyAsync(rel_poses, d_rel, sizeof(double) * 12 * n, cudaMemcpyDeviceToHost, c->stream));
        fail(cudaMemcpyAsync(kind, d_kind, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
        if (reason) fail(cudaMemcpyAsync(reason, d_reason, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    }
    cudaStreamSynchronize(c->copy_stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && rc == DVO_OK) { dvo_set_error("%s: %s", who, cudaGetErrorString(cudaGetLastError())); rc = DVO_ERR_CUDA; }
    return rc;
}
}  // namespace

extern "C" {

int dvo_run_sequences(dvo_ctx* c, int nseq, int nframes, const uint8_t* gray, const uint16_t* depth, const dvo_solver_params* p,
                      int keyframe_every, double* rel_poses, int* kind, double* global_poses) {
    dvo_keyframe_policy pol; memset(&pol, 0, sizeof(pol));
    pol.keyframe_every = keyframe_every; pol.use_quality_gates = 0; pol.laplacian_thresh = 3.0f; pol.visible_ratio_thresh = 0.8f; pol.min_reprojections = 50;
    return run_sequences_core(c, nseq, nframes, gray, depth, DVO_MEM_HOST, p, &pol, rel_poses, kind, nullptr, global_poses);
}

int dvo_run_sequences_gated(dvo_ctx* c, int nseq, int nframes, const uint8_t* gray, const uint16_t* depth, const dvo_solver_params* p,
                            const dvo_keyframe_policy* pol, double* rel_poses, int* kind, int* reason, double* global_poses) {
    return run_sequences_core(c, nseq, nframes, gray, depth, DVO_MEM_HOST, p, pol, rel_poses, kind, reason, global_poses);
}

int dvo_run_sequences_mem(dvo_ctx* c, int nseq, int nframes, const uint8_t* gray, const uint16_t* depth, int mem, const dvo_solver_params* p,
                          const dvo_keyframe_policy* pol, double* rel_poses, int* kind, int* reason, double* global_poses) {
    return run_sequences_core(c, nseq, nframes, gray, depth, mem, p, pol, rel_poses, kind, reason, global_poses);
}

int dvo_level_dims(dvo_ctx* c, int level, int* w, int* h) {
    if (!c || level < 0 || level >= c->geom.L) return DVO_ERR_ARG;
    if (w) *w = c->geom.w[level];
    if (h) *h = c->geom.h[level];
    return DVO_OK;
}

int dvo_get_level_buffer(dvo_ctx* c, int slot, int frame, int level, int which, void* dst, size_t bytes) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, slot, 1) || level < 0 || level >= c->geom.L || (frame != 0 && frame != 1) || !dst) { dvo_set_error("dvo_get_level_buffer: bad argument"); return DVO_ERR_ARG; }
    const size_t P = c->geom.P[level];
    const long long o = lvl_at(c->geom, level, slot);
    const void* src = nullptr; size_t es = 0;
    switch (which) {
        case DVO_BUF_GRAY: src = c->gray[frame] + o; es = 1; break;
        case DVO_BUF_DEPTH: if (!c->depth[frame]) { dvo_set_error("depth not stored for this frame"); return DVO_ERR_STATE; } src = c->depth[frame] + o; es = 2; break;
        case DVO_BUF_EDGE: src = c->edge[frame] + o; es = 1; break;
        case DVO_BUF_D2: if (frame != DVO_FRAME_NOW) { dvo_set_error("d2 exists for the now frame only"); return DVO_ERR_ARG; } src = c->d2 + o; es = 4; break;
        case DVO_BUF_DTN: case DVO_BUF_GX: case DVO_BUF_GY:
            if (frame != DVO_FRAME_NOW) { dvo_set_error("DT exists for the now frame only"); return DVO_ERR_ARG; } src = c->texel ? c->texel + o : nullptr; es = 16; break;
        default: dvo_set_error("dvo_get_level_buffer: unknown buffer %d", which); return DVO_ERR_ARG;
    }
    if (which == DVO_BUF_D2 && c->texel_mode && c->edt_band >= 0) {      // fused EDT + texel kernel: the 32-bit image is sparse, rebuild the dense one
        if (bytes < P * 4) { dvo_set_error("dvo_get_level_buffer: destination too small"); return DVO_ERR_ARG; }
        int32_t* d_tmp = nullptr;
        DVO_CUDA(cudaMalloc((void**)&d_tmp, P * 4));
        const int rc = launch_d2_from_texels(c, slot, level, d_tmp);
        cudaError_t ce = rc ? cudaSuccess : cudaMemcpyAsync(dst, d_tmp, P * 4, cudaMemcpyDeviceToHost, c->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
        cudaFree(d_tmp);
        if (rc) return rc;
        DVO_CUDA(ce);
        return DVO_OK;
    }
    if (which >= DVO_BUF_DTN) {
        if (bytes < P * 4) { dvo_set_error("dvo_get_level_buffer: destination too small"); return DVO_ERR_ARG; }
        std::vector<float> tmp(P * 4);
        float4* d_tmp = nullptr;
        if (!src) {      // packed-texel contexts keep no float images: evaluate them for this slot / level on demand
            DVO_CUDA(cudaMalloc((void**)&d_tmp, P * 16));
            const int rc = c->texel_mode ? launch_resolve_texels(c, slot, level, d_tmp) : launch_normgrad_into(c, slot, level, d_tmp);
            if (rc) { cudaFree(d_tmp); return rc; }
            src = d_tmp;
        }
        cudaError_t ce = cudaMemcpyAsync(tmp.data(), src, P * 16, cudaMemcpyDeviceToHost, c->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
        cudaFree(d_tmp);
        DVO_CUDA(ce);
        const int comp = which - DVO_BUF_DTN;
        float* d = (float*)dst;
        for (size_t i = 0; i < P; ++i) d[i] = tmp[4 * i + comp];
        return DVO_OK;
    }
    if (bytes < P * es) { dvo_set_error("dvo_get_level_buffer: destination too small"); return DVO_ERR_ARG; }
    DVO_CUDA(cudaMemcpyAsync(dst, src, P * es, cudaMemcpyDeviceToHost, c->stream));
    DVO_CUDA(cudaStreamSynchronize(c->stream));
    if (which == DVO_BUF_DEPTH && level == 0) {            // level-0 zero fix is applied at read-out (see preprocess.cu)
        uint16_t* d = (uint16_t*)dst;
        for (size_t i = 0; i < P; ++i) if (d[i] == 0) d[i] = 1;
    }
    return DVO_OK;
}

// The device keeps the point list in row-major pixel order; the reference enumerates column-major (xx outer, yy inner,
// src/SolveDVO.cpp:1237-1241).  perm[k] = device index of the k-th point in the reference's order.
static int reference_order(dvo_ctx* c, int slot, int level, int n, std::vector<int>& perm) {
    perm.resize(n);
    if (n == 0) return DVO_OK;
    std::vector<int> pix(n);
    DVO_CUDA(cudaMemcpyAsync(pix.data(), c->ptsPix + lvl_at(c->geom, level, slot), sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    DVO_CUDA(cudaStreamSynchronize(c->stream));
    const int w = c->geom.w[level], h = c->geom.h[level];
    std::vector<long long> key(n);
    for (int i = 0; i < n; ++i) { const int y = pix[i] / w, x = pix[i] - y * w; key[i] = ((long long)x * h + y) * (long long)n + i; }
    std::sort(key.begin(), key.end());
    for (int k = 0; k < n; ++k) perm[k] = (int)(key[k] % n);
    return DVO_OK;
}

int dvo_get_points(dvo_ctx* c, int slot, int level, float* X, float* Y, float* Z, int capacity, int* n) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, slot, 1) || level < 0 || level >= c->geom.L || !n) return DVO_ERR_ARG;
    int cnt = 0;
    DVO_CUDA(cudaMemcpyAsync(&cnt, c->npts + (size_t)slot * c->geom.L + level, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    DVO_CUDA(cudaStreamSynchronize(c->stream));
    *n = cnt;
    const int m = cnt < capacity ? cnt : capacity;
    if (m <= 0 || !(X || Y || Z)) return DVO_OK;
    std::vector<int> perm;
    int rc = reference_order(c, slot, level, cnt, perm);
    if (rc) return rc;
    const long long o = lvl_at(c->geom, level, slot);
    std::vector<float> tmp(cnt);
    float* dst[3] = {X, Y, Z}; const float* src[3] = {c->ptsX + o, c->ptsY + o, c->ptsZ + o};
    for (int a = 0; a < 3; ++a) {
        if (!dst[a]) continue;
        DVO_CUDA(cudaMemcpyAsync(tmp.data(), src[a], sizeof(float) * cnt, cudaMemcpyDeviceToHost, c->stream));
        DVO_CUDA(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < m; ++k) dst[a][k] = tmp[perm[k]];
    }
    return DVO_OK;
}

int dvo_eval_normal_equations(dvo_ctx* c, int slot, int level, const double* R9T3, int jacobian, int weight, int arithmetic,
                              float huber_k, double* H36, double* g6, double* sumsq, int* nvis, float* eps, float* w, float* u,
                              float* v, float* J) {
    dvo_solver_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.jacobian = jacobian; prm.weight = weight; prm.arithmetic = arithmetic; prm.huber_k = huber_k;
    return dvo_eval_normal_equations_ex(c, slot, level, R9T3, &prm, H36, g6, sumsq, nvis, eps, w, u, v, J);
}

int dvo_eval_normal_equations_ex(dvo_ctx* c, int slot, int level, const double* R9T3, const dvo_solver_params* prm,
                                 double* H36, double* g6, double* sumsq, int* nvis, float* eps, float* w, float* u, float* v, float* J) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!prm) return DVO_ERR_ARG;
    if (!range_ok(c, slot, 1) || level < 0 || level >= c->geom.L || !R9T3) return DVO_ERR_ARG;
    if (!c->haveK) { dvo_set_error("intrinsics not set"); return DVO_ERR_STATE; }
    int N = 0;
    DVO_CUDA(cudaMemcpyAsync(&N, c->npts + (size_t)slot * c->geom.L + level, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    DVO_CUDA(cudaStreamSynchronize(c->stream));
    double* d_pose = nullptr; double* d_out = nullptr; float* d_pp = nullptr;
    const bool pp = eps || w || u || v || J;
    const size_t n1 = (size_t)(N > 0 ? N : 1);
    {   // temporaries: released on every path
        cudaError_t e = cudaMalloc((void**)&d_pose, sizeof(double) * 12);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, sizeof(double) * 44);
        if (e == cudaSuccess && pp) e = cudaMalloc((void**)&d_pp, sizeof(float) * n1 * 10);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_pose, R9T3, sizeof(double) * 12, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) {
            dvo_set_error("dvo_eval_normal_equations: %s", cudaGetErrorString(e));
            cudaFree(d_pose); cudaFree(d_out); cudaFree(d_pp);
            return DVO_ERR_CUDA;
        }
    }
    float *de = nullptr, *dw = nullptr, *du = nullptr, *dv = nullptr, *dJ = nullptr;
    if (pp) { de = d_pp; dw = d_pp + n1; du = d_pp + 2 * n1; dv = d_pp + 3 * n1; dJ = d_pp + 4 * n1; }
    int rc = launch_eval(c, slot, level, d_pose, prm->jacobian, prm->weight, prm->arithmetic, prm->huber_k, prm->residual, d_out, de, dw, du, dv, dJ);
    if (rc == DVO_OK) {
        double out[44];
        std::vector<float> host(pp ? n1 * 10 : 0);
        cudaError_t e = cudaMemcpyAsync(out, d_out, sizeof(out), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && pp && N > 0) e = cudaMemcpyAsync(host.data(), d_pp, sizeof(float) * n1 * 10, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { dvo_set_error("dvo_eval_normal_equations: %s", cudaGetErrorString(e)); rc = DVO_ERR_CUDA; }
        else {
            if (g6) memcpy(g6, out, sizeof(double) * 6);
            if (H36) memcpy(H36, out + 6, sizeof(double) * 36);
            if (sumsq) *sumsq = out[42];
            if (nvis) *nvis = (int)out[43];
            if (pp && N > 0) {                       // per-point arrays are returned in the reference's (column-major) order
                std::vector<int> perm;
                rc = reference_order(c, slot, level, N, perm);
                if (rc == DVO_OK)
                    for (int k = 0; k < N; ++k) {
                        const size_t i = (size_t)perm[k];
                        if (eps) eps[k] = host[i];
                        if (w) w[k] = host[n1 + i];
                        if (u) u[k] = host[2 * n1 + i];
                        if (v) v[k] = host[3 * n1 + i];
                        if (J) for (int q = 0; q < 6; ++q) J[6 * (size_t)k + q] = host[4 * n1 + 6 * i + q];
                    }
            }
        }
    }
    cudaFree(d_pose); cudaFree(d_out); cudaFree(d_pp);
    return rc;
}

int dvo_get_trace(dvo_ctx* c, int slot, int level, double* trace) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, slot, 1) || level < 0 || level >= c->geom.L || !trace) return DVO_ERR_ARG;
    if (!c->trace) { dvo_set_error("dvo_get_trace: context created with trace_iters = 0"); return DVO_ERR_STATE; }
    const size_t n = (size_t)c->cfg.trace_iters * DVO_TRACE_DOUBLES;
    DVO_CUDA(cudaMemcpyAsync(trace, c->trace + ((size_t)slot * c->geom.L + level) * n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    DVO_CUDA(cudaStreamSynchronize(c->stream));
    return DVO_OK;
}

int dvo_get_energies(dvo_ctx* c, int slot, int level, float* energies, int capacity) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!range_ok(c, slot, 1) || level < 0 || level >= c->geom.L || !energies || capacity < 0) return DVO_ERR_ARG;
    const int n = capacity < DVO_ENERGY_ITERS ? capacity : DVO_ENERGY_ITERS;
    if (n == 0) return DVO_OK;
    DVO_CUDA(cudaMemcpyAsync(energies, c->energy + ((size_t)slot * c->geom.L + level) * DVO_ENERGY_ITERS, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    DVO_CUDA(cudaStreamSynchronize(c->stream));
    return DVO_OK;
}

int dvo_enable_timing(dvo_ctx* c, int on) {
    if (!c) return DVO_ERR_ARG;
    c->timing = on != 0;
    for (int i = 0; i < DVO_STAGE_COUNT; ++i) c->stage_ms[i] = 0.f;
    return DVO_OK;
}

int dvo_get_stage_ms(dvo_ctx* c, float* ms) {
    if (!c || !ms) return DVO_ERR_ARG;
    for (int i = 0; i < DVO_STAGE_COUNT; ++i) { ms[i] = c->stage_ms[i]; c->stage_ms[i] = 0.f; }
    return DVO_OK;
}

long long dvo_launch_count(dvo_ctx* c) { return c ? c->launches : 0; }

int dvo_gop_compose(dvo_ctx* c, int nseq, int nframes, const int* kind, const double* rel, double* out, int mem) {
    if (c) { const int jr = join_aux(c); if (jr) return jr; }
    if (!c || nseq < 1 || nframes < 1 || !kind || !rel || !out) return DVO_ERR_ARG;
    const size_t n = (size_t)nseq * nframes;
    if (mem == DVO_MEM_DEVICE) return launch_gop(c, nseq, nframes, kind, rel, out);
    int* d_kind = nullptr; double* d_rel = nullptr; double* d_out = nullptr;
    {
        cudaError_t e = cudaMalloc((void**)&d_kind, sizeof(int) * n);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_rel, sizeof(double) * 12 * n);
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, sizeof(double) * 19 * n);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_kind, kind, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_rel, rel, sizeof(double) * 12 * n, cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) {
            dvo_set_error("dvo_gop_compose: %s", cudaGetErrorString(e));
            cudaFree(d_kind); cudaFree(d_rel); cudaFree(d_out);
            return DVO_ERR_CUDA;
        }
    }
    int rc = launch_gop(c, nseq, nframes, d_kind, d_rel, d_out);
    if (rc == DVO_OK) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, sizeof(double) * 19 * n, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { dvo_set_error("dvo_gop_compose: %s", cudaGetErrorString(e)); rc = DVO_ERR_CUDA; }
    }
    cudaFree(d_kind); cudaFree(d_rel); cudaFree(d_out);
    return rc;
}

// ---- device-resident storage blobs (PyramidalStorageStruct keeps caller-provided level arrays on the device) ----
struct dvo_blob { void* d; size_t bytes; int device; };

int dvo_blob_create(size_t bytes, int device, dvo_blob** out) {
    if (!out || bytes == 0) { dvo_set_error("dvo_blob_create: bad argument"); return DVO_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); dvo_set_error("dvo_blob_create: no usable CUDA device (there is no CPU fallback)"); return DVO_ERR_CUDA; }
    if (device < 0 || device >= ndev) { dvo_set_error("dvo_blob_create: device %d out of range", device); return DVO_ERR_ARG; }
    DVO_CUDA(cudaSetDevice(device));
    dvo_blob* b = new (std::nothrow) dvo_blob();
    if (!b) return DVO_ERR_NOMEM;
    b->bytes = bytes; b->device = device; b->d = nullptr;
    if (cudaMalloc(&b->d, bytes) != cudaSuccess) { cudaGetLastError(); delete b; dvo_set_error("dvo_blob_create: cudaMalloc(%zu) failed", bytes); return DVO_ERR_NOMEM; }
    *out = b;
    return DVO_OK;
}
int dvo_blob_upload(dvo_blob* b, const void* host_src, size_t bytes) {
    if (!b || !host_src || bytes > b->bytes) { dvo_set_error("dvo_blob_upload: bad argument"); return DVO_ERR_ARG; }
    DVO_CUDA(cudaSetDevice(b->device));
    DVO_CUDA(cudaMemcpy(b->d, host_src, bytes, cudaMemcpyHostToDevice));
    return DVO_OK;
}
int dvo_blob_download(dvo_blob* b, void* host_dst, size_t bytes) {
    if (!b || !host_dst || bytes > b->bytes) { dvo_set_error("dvo_blob_download: bad argument"); return DVO_ERR_ARG; }
    DVO_CUDA(cudaSetDevice(b->device));
    DVO_CUDA(cudaMemcpy(host_dst, b->d, bytes, cudaMemcpyDeviceToHost));
    return DVO_OK;
}
void* dvo_blob_device_ptr(dvo_blob* b) { return b ? b->d : nullptr; }
size_t dvo_blob_bytes(dvo_blob* b) { return b ? b->bytes : 0; }
int dvo_blob_destroy(dvo_blob* b) {
    if (!b) return DVO_OK;
    cudaSetDevice(b->device);
    cudaFree(b->d);
    delete b;
    return DVO_OK;
}

}  // extern "C"
