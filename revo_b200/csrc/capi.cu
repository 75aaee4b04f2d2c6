// capi.cu -- implementation of the C ABI declared in include/revo_b200.h.
// Host-side orchestration only: memory layout in HBM, descriptor tables, launch sequencing.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <new>
#include <vector>

#include "internal.h"

using namespace revo;

namespace revo {

int cuda_fail(revo_ctx *ctx, cudaError_t e, const char *what)
{
    if (ctx) {
        char buf[512];
        snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
        ctx->last_error = buf;
    }
    (void)cudaGetLastError();
    return REVO_ERR_CUDA;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static bool is_device_ptr(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static int ensure_scratch(revo_ctx *ctx, size_t bytes)
{
    if (ctx->scratch_bytes >= bytes) return REVO_OK;
    if (ctx->scratch) {
        REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        REVO_CUDA(ctx, cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    bytes = align_up(bytes + bytes / 4, 1 << 20);
    REVO_CUDA(ctx, cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return REVO_OK;
}

static int ensure_pinned(revo_ctx *ctx, size_t bytes)
{
    if (ctx->pinned_bytes >= bytes) return REVO_OK;
    if (ctx->pinned) {
        REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        REVO_CUDA(ctx, cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
    }
    bytes = align_up(bytes * 2, 1 << 16);
    REVO_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
    return REVO_OK;
}

// Camera(fx,fy,cx,cy,w,h,scale) -- camerapyr.h:98-103 with scale = 1.0f/pow(2,lvl) (:141)
static void level_camera(const revo_camera &c0, int lvl, revo_camera *out)
{
    if (lvl == 0) { *out = c0; return; }
    const float scale = 1.0f / (float)pow(2.0, (double)lvl);
    out->fx = c0.fx * scale; out->fy = c0.fy * scale; out->cx = c0.cx * scale; out->cy = c0.cy * scale;
    out->width = (int32_t)((float)c0.width * scale);
    out->height = (int32_t)((float)c0.height * scale);
}

}  // namespace revo

// ---------------------------------------------------------------------------------------------------
// defaults
// ---------------------------------------------------------------------------------------------------
extern "C" {

void revo_pyr_config_default(revo_pyr_config *c)
{
    c->n_levels = 3; c->canny_threshold1 = 150; c->canny_threshold2 = 100;
    c->depth_min = 0.1f; c->depth_max = 5.2f; c->use_edge_hist = 1; c->n_percentage = 0.3f; c->patch0 = 20;
}

void revo_opt_config_default(revo_opt_config *c)
{
    const float ed[6] = {30, 20, 10, 5, 5, 5};
    c->lambda_success_fac = 0.5f; c->lambda_fail_fac = 2.0f;
    for (int l = 0; l < REVO_MAX_LEVELS; ++l) {
        c->lambda_initial[l] = 0.f; c->step_size_min[l] = 1e-16f; c->convergence_eps[l] = 0.999f;
        c->max_its_per_lvl[l] = 100; c->edge_distance_lvl[l] = ed[l];
    }
    c->huber_edge = 0.3f; c->use_edge_filter = 1; c->max_lm_tries = 0;
}

void revo_tracker_config_default(revo_tracker_config *c)
{
    c->check_init_values = 1; c->pyr_min_lvl = 2; c->pyr_max_lvl = 0;
    revo_opt_config_default(&c->opt);
}

const char *revo_strerror(int code)
{
    switch (code) {
        case REVO_OK: return "ok";
        case REVO_ERR_INVALID_ARG: return "invalid argument";
        case REVO_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
        case REVO_ERR_CUDA: return "CUDA error (see revo_last_error)";
        case REVO_ERR_NOT_KEYFRAME: return "optimization structure not built (makeKeyframe was not called)";
        case REVO_ERR_NOT_ORTHOGONAL: return "R is not a rotation matrix";
        case REVO_ERR_BAD_LEVEL: return "pyramid level out of range";
        case REVO_ERR_BUFFER_TOO_SMALL: return "destination buffer too small";
        case REVO_ERR_UNSUPPORTED: return "unsupported configuration";
        case REVO_ERR_COMM: return "multi-GPU exchange failed (setup error, or a peer rank did not answer in time)";
        default: return "unknown error";
    }
}

// ---------------------------------------------------------------------------------------------------
// pose helpers (host arithmetic)
// ---------------------------------------------------------------------------------------------------
int revo_quat_to_R9(const float *q, float *R)
{
    if (!q || !R) return REVO_ERR_INVALID_ARG;
    const double n2 = (double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2] + (double)q[3] * q[3];
    if (!(n2 > 0.0) || !std::isfinite(n2)) return REVO_ERR_INVALID_ARG;
    const double s = 1.0 / std::sqrt(n2);
    const double x = q[0] * s, y = q[1] * s, z = q[2] * s, w = q[3] * s;
    R[0] = (float)(1 - 2 * (y * y + z * z)); R[3] = (float)(2 * (x * y - z * w));     R[6] = (float)(2 * (x * z + y * w));
    R[1] = (float)(2 * (x * y + z * w));     R[4] = (float)(1 - 2 * (x * x + z * z)); R[7] = (float)(2 * (y * z - x * w));
    R[2] = (float)(2 * (x * z - y * w));     R[5] = (float)(2 * (y * z + x * w));     R[8] = (float)(1 - 2 * (x * x + y * y));
    return REVO_OK;
}

int revo_R9_to_quat(const float *Rf, float *q)
{
    if (!Rf || !q) return REVO_ERR_INVALID_ARG;
    double R[9];
    for (int i = 0; i < 9; ++i) R[i] = Rf[i];
    auto M = [&](int i, int j) { return R[j * 3 + i]; };
    double n2 = 0;   // ||R^T R - I||_F^2
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double d = -(i == j ? 1.0 : 0.0);
            for (int k = 0; k < 3; ++k) d += M(k, i) * M(k, j);
            n2 += d * d;
        }
    const double det = M(0, 0) * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) - M(0, 1) * (M(1, 0) * M(2, 2) - M(1, 2) * M(2, 0)) +
                       M(0, 2) * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
    if (!(std::sqrt(n2) < 1e-5) || !(det > 0)) return REVO_ERR_NOT_ORTHOGONAL;
    double o[4];
    double t = M(0, 0) + M(1, 1) + M(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        o[3] = 0.5 * t;
        t = 0.5 / t;
        o[0] = (M(2, 1) - M(1, 2)) * t;
        o[1] = (M(0, 2) - M(2, 0)) * t;
        o[2] = (M(1, 0) - M(0, 1)) * t;
    } else {
        int i = 0;
        if (M(1, 1) > M(0, 0)) i = 1;
        if (M(2, 2) > M(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
        o[i] = 0.5 * t;
        t = 0.5 / t;
        o[3] = (M(k, j) - M(j, k)) * t;
        o[j] = (M(j, i) + M(i, j)) * t;
        o[k] = (M(k, i) + M(i, k)) * t;
    }
    for (int c = 0; c < 4; ++c) q[c] = (float)o[c];
    return REVO_OK;
}

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
int revo_ctx_create(int device, revo_ctx **out)
{
    if (!out) return REVO_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        return REVO_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return REVO_ERR_INVALID_ARG;
    revo_ctx *ctx = new (std::nothrow) revo_ctx();
    if (!ctx) return REVO_ERR_INVALID_ARG;
    ctx->device = device;
    ctx->launches = 0;
    ctx->scratch = nullptr; ctx->scratch_bytes = 0; ctx->pinned = nullptr; ctx->pinned_bytes = 0; ctx->pinned_kf = nullptr; ctx->pinned_kf_bytes = 0; ctx->pinned_kf_busy = false;
    for (int i = 0; i < 2; ++i) { ctx->stage[i] = nullptr; ctx->stage_bytes[i] = 0; ctx->stage_used[i] = false; }
    ctx->stage_next = 0;
    ctx->track_ctas_per_pair = 0; ctx->track_threads = 0; ctx->track_max_clusters = 0;
    for (auto &v : ctx->ev_valid) v = false;
    ctx->split_rank = 0; ctx->split_world = 1; ctx->split_local = nullptr; ctx->split_seq = 0;
    for (auto &p : ctx->split_peers) p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        (void)cudaGetLastError();
        delete ctx;
        return REVO_ERR_CUDA;
    }
    // keep freed stream-ordered allocations cached in the pool (no trimming at synchronisation points)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    for (auto &st : ctx->lvl_stream) cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->depth_stream, cudaStreamNonBlocking);
    for (auto &e : ctx->lvl_gray) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &e : ctx->lvl_depth) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->lvl_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->lvl_canny0, cudaEventDisableTiming);
    for (auto &e : ctx->lvl_done) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &e : ctx->lvl_fill) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->pinned_kf_read, cudaEventDisableTiming);
    for (auto &e : ctx->stage_consumed) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    (void)cudaGetLastError();
    *out = ctx;
    return REVO_OK;
}

int revo_ctx_destroy(revo_ctx *ctx)
{
    if (!ctx) return REVO_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 16; ++i)
        if (ctx->split_peers[i] && ctx->split_peers[i] != ctx->split_local) cudaIpcCloseMemHandle(ctx->split_peers[i]);
    if (ctx->split_local) cudaFree(ctx->split_local);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pinned_kf) cudaFreeHost(ctx->pinned_kf);
    cudaEventDestroy(ctx->pinned_kf_read);
    for (int i = 0; i < 2; ++i) { if (ctx->stage[i]) cudaFree(ctx->stage[i]); cudaEventDestroy(ctx->stage_consumed[i]); }
    for (auto &e : ctx->ev) cudaEventDestroy(e);
    for (auto &st : ctx->lvl_stream) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    cudaStreamSynchronize(ctx->depth_stream);
    cudaStreamDestroy(ctx->depth_stream);
    for (auto &e : ctx->lvl_gray) cudaEventDestroy(e);
    for (auto &e : ctx->lvl_depth) cudaEventDestroy(e);
    cudaEventDestroy(ctx->lvl_fork);
    cudaEventDestroy(ctx->lvl_canny0);
    for (auto &e : ctx->lvl_done) cudaEventDestroy(e);
    for (auto &e : ctx->lvl_fill) cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    (void)cudaGetLastError();
    delete ctx;
    return REVO_OK;
}

int revo_ctx_synchronize(revo_ctx *ctx)
{
    if (!ctx) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

const char *revo_last_error(revo_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t revo_ctx_stream(revo_ctx *ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }
uint64_t revo_ctx_launch_count(revo_ctx *ctx) { return ctx ? ctx->launches : 0; }

int revo_ctx_last_timings(revo_ctx *ctx, float *pyramid_ms, float *keyframe_ms, float *track_kernel_ms)
{
    if (!ctx) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float *out[3] = {pyramid_ms, keyframe_ms, track_kernel_ms};
    for (int i = 0; i < 3; ++i) {
        if (!out[i]) continue;
        *out[i] = 0.f;
        if (ctx->ev_valid[i]) REVO_CUDA(ctx, cudaEventElapsedTime(out[i], ctx->ev[2 * i], ctx->ev[2 * i + 1]));
    }
    return REVO_OK;
}

int revo_ctx_last_upload_ms(revo_ctx *ctx, float *upload_ms)
{
    if (!ctx || !upload_ms) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    *upload_ms = 0.f;
    if (ctx->ev_valid[3]) REVO_CUDA(ctx, cudaEventElapsedTime(upload_ms, ctx->ev[6], ctx->ev[7]));
    return REVO_OK;
}

int revo_ctx_set_track_max_clusters(revo_ctx *ctx, int max_clusters)
{
    if (!ctx || max_clusters < 0) return REVO_ERR_INVALID_ARG;
    ctx->track_max_clusters = max_clusters;
    return REVO_OK;
}

int revo_ctx_set_track_shape(revo_ctx *ctx, int ctas_per_pair, int threads_per_cta)
{
    if (!ctx) return REVO_ERR_INVALID_ARG;
    if (ctas_per_pair != 0 && ctas_per_pair != 1 && ctas_per_pair != 2 && ctas_per_pair != 4 && ctas_per_pair != 8 &&
        ctas_per_pair != 16)
        return REVO_ERR_INVALID_ARG;
    if (threads_per_cta != 0 && threads_per_cta != 128 && threads_per_cta != 256 && threads_per_cta != 512 &&
        threads_per_cta != 1024)
        return REVO_ERR_INVALID_ARG;
    ctx->track_ctas_per_pair = ctas_per_pair;
    ctx->track_threads = threads_per_cta;
    return REVO_OK;
}

// Kept for ABI compatibility with round 1: the library has ONE tracking engine (one thread-block cluster per pair, track.cu);
// the task-queue and ping-pong engines measured slower at every batch size and live in scratch/experiments/ now.
int revo_ctx_set_track_engine(revo_ctx *ctx, int engine, int chunk_points)
{
    if (!ctx || chunk_points < 0) return REVO_ERR_INVALID_ARG;
    if (engine != 0 && engine != 1) return REVO_ERR_UNSUPPORTED;
    return REVO_OK;
}

// Grow the device's stream-ordered memory pool to at least `bytes` of cached, reusable memory now (one allocation +
// free on the context stream), so that later slab allocations of a steady-state stream never reach the driver.
int revo_ctx_reserve(revo_ctx *ctx, size_t bytes)
{
    if (!ctx) return REVO_ERR_INVALID_ARG;
    if (bytes == 0) return REVO_OK;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    void *p = nullptr;
    REVO_CUDA(ctx, cudaMallocAsync(&p, bytes, ctx->stream));
    REVO_CUDA(ctx, cudaFreeAsync(p, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

// ---------------------------------------------------------------------------------------------------
// pyramid construction
// ---------------------------------------------------------------------------------------------------
struct LevelGeom {
    int w, h, patch, hist_w, hist_h, cap, n_tiles;
    revo_camera cam;
};

static int level_geometry(const revo_pyr_config *cfg, const revo_camera *cam0, LevelGeom *g)
{
    if (cfg->n_levels < 1 || cfg->n_levels > REVO_MAX_LEVELS) return REVO_ERR_INVALID_ARG;
    if (cam0->width < 8 || cam0->height < 8) return REVO_ERR_INVALID_ARG;
    for (int l = 0; l < cfg->n_levels; ++l) {
        level_camera(*cam0, l, &g[l].cam);
        g[l].w = g[l].cam.width; g[l].h = g[l].cam.height;
        if (g[l].w < 4 || g[l].h < 4) return REVO_ERR_UNSUPPORTED;
        // the reference is only self-consistent for even sizes (pyrDown -> (n+1)/2, depth/Camera -> n/2:
        // imgpyramidrgbd.cpp:79-85, camerapyr.h:100); refuse the others instead of reading out of bounds
        if (l + 1 < cfg->n_levels && ((g[l].w & 1) || (g[l].h & 1))) return REVO_ERR_UNSUPPORTED;
        g[l].patch = std::max(1, cfg->patch0 >> l);
        g[l].hist_w = g[l].w / g[l].patch; g[l].hist_h = g[l].h / g[l].patch;
        g[l].cap = g[l].w * g[l].h / 2 + 1024;
        g[l].n_tiles = ((g[l].w + kTileW - 1) / kTileW) * ((g[l].h + kTileH - 1) / kTileH);
    }
    return REVO_OK;
}

// depth16 != nullptr: raw 16-bit depth (units of depth_scale metres), converted on the device (K0 k_depth_u16)
static int create_batch_impl(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, int n, const uint8_t *bgr,
                             int channels, const float *depth, const uint16_t *depth16, float depth_scale,
                             const double *timestamps, revo_pyr **pyr_out)
{
    if (!ctx || !cfg || !cam0 || !bgr || (!depth && !depth16) || !pyr_out || n < 1 || (channels != 3 && channels != 4))
        return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    LevelGeom g[REVO_MAX_LEVELS];
    int rc = level_geometry(cfg, cam0, g);
    if (rc) return rc;
    const int NL = cfg->n_levels;
    const int w0 = g[0].w, h0 = g[0].h;

    // ---- slab layout: [array][level][frame], every chunk 256-byte aligned ------------------------
    size_t off = 0;
    auto take = [&](size_t per_frame) { size_t o = off; off += align_up(per_frame, 256) * (size_t)n; return o; };
    size_t o_desc[REVO_MAX_LEVELS], o_gray[REVO_MAX_LEVELS], o_depth[REVO_MAX_LEVELS], o_edges[REVO_MAX_LEVELS],
        o_eorig[REVO_MAX_LEVELS], o_hist[REVO_MAX_LEVELS], o_pts[REVO_MAX_LEVELS], o_toff[REVO_MAX_LEVELS];
    for (int l = 0; l < NL; ++l) { o_desc[l] = off; off += align_up(sizeof(ImgLevel) * (size_t)n, 256); }
    const size_t o_counters = off; off += align_up(sizeof(int) * 2 * NL * (size_t)n, 256);   // n_pts, nz_patches
    for (int l = 0; l < NL; ++l) {
        const size_t px = (size_t)g[l].w * g[l].h;
        o_gray[l] = take(px); o_depth[l] = take(px * 4); o_edges[l] = take(px); o_eorig[l] = take(px);
        o_hist[l] = take((size_t)std::max(1, g[l].hist_w * g[l].hist_h));
        o_pts[l] = take((size_t)g[l].cap * 16);
        o_toff[l] = take(((size_t)g[l].n_tiles + 1) * 4);
    }
    const size_t o_labels = take((size_t)w0 * h0 * 4);
    const size_t o_flags = take((size_t)w0 * h0);
    const size_t total = off;

    // Host inputs are uploaded on the context's COPY stream (slab allocation, descriptor tables, bgr staging and depth),
    // the kernels run on the main stream behind an event: the H2D of the next batch overlaps the kernels of this one.
    const bool host_in = !is_device_ptr(bgr);
    cudaStream_t up = host_in ? ctx->copy_stream : ctx->stream;

    Slab *slab = new (std::nothrow) Slab();
    if (!slab) return REVO_ERR_INVALID_ARG;
    slab->n_frames = n; slab->live = n; slab->bytes = total; slab->mem = nullptr;
    slab->stream = ctx->stream; slab->ready = nullptr;
    cudaError_t e = cudaMallocAsync(&slab->mem, total, up);
    if (e != cudaSuccess) { delete slab; return cuda_fail(ctx, e, "cudaMallocAsync(slab)"); }
    uint8_t *base = (uint8_t *)slab->mem;
    auto chunk = [&](size_t o, size_t per_frame, int f) { return base + o + align_up(per_frame, 256) * (size_t)f; };

    size_t lbl_off[REVO_MAX_LEVELS];
    const size_t cnt_region = align_up((size_t)w0 * h0 / 64 + 256, 256);      // >= 4 (w/P)(h/P) bytes on every level (P = 20 >> l)
    {
        size_t o = 0;
        for (int l = 0; l < NL; ++l) { lbl_off[l] = o; o += align_up((size_t)g[l].w * g[l].h / 2 + 4096, 256); }
        bool fits = lbl_off[NL - 1] + (size_t)g[NL - 1].w * g[NL - 1].h * 4 <= (size_t)w0 * h0 * 4 && (size_t)NL * cnt_region <= (size_t)w0 * h0;
        for (int l = 0; l < NL; ++l) fits = fits && (size_t)std::max(1, g[l].hist_w * g[l].hist_h) * 4 <= cnt_region;
        if (!fits) return REVO_ERR_UNSUPPORTED;      // tiny images with many levels
    }
    std::vector<revo_pyr *> pyrs(n);
    std::vector<ImgLevel> host_desc((size_t)NL * n);
    for (int f = 0; f < n; ++f) {
        revo_pyr *p = new revo_pyr();
        p->slab = slab; p->index_in_slab = f; p->n_levels = NL; p->cfg = *cfg; p->cam0 = *cam0;
        p->timestamp = timestamps ? timestamps[f] : 0.0; p->kf_slab = nullptr; p->is_keyframe = false;
        for (int l = 0; l < NL; ++l) {
            ImgLevel &L = p->lv[l];
            const size_t px = (size_t)g[l].w * g[l].h;
            L.gray = chunk(o_gray[l], px, f);
            L.depth = (float *)chunk(o_depth[l], px * 4, f);
            L.edges = chunk(o_edges[l], px, f);
            // edgesOrigPyr differs from edgesPyr only where the edge fill-in may run (levels 1 and 2 with USE_EDGE_HIST,
            // imgpyramidrgbd.cpp:185-195): elsewhere both names are ONE plane
            L.edges_orig = (cfg->use_edge_hist && l >= 1 && l <= 2) ? chunk(o_eorig[l], px, f) : L.edges;
            L.hist = chunk(o_hist[l], (size_t)std::max(1, g[l].hist_w * g[l].hist_h), f);
            L.pts = (float4 *)chunk(o_pts[l], (size_t)g[l].cap * 16, f);
            L.n_pts = (int *)(base + o_counters) + ((size_t)f * NL + l) * 2;
            L.nz_patches = L.n_pts + 1;
            L.tile_off = (int *)chunk(o_toff[l], ((size_t)g[l].n_tiles + 1) * 4, f);
            // every level owns its part of the frame's scratch planes, so that the levels' chains can run concurrently: the
            // label plane (4 B per level-0 pixel) holds the Canny bit masks and the compaction's group counts (< w h / 2 bytes
            // per level) and, for keyframes, the EDT column distances (4 w h bytes, level by level); the flags plane the integer
            // patch counters of the histogram (w0 h0 / 64 bytes per level)
            L.labels = (int *)(chunk(o_labels, (size_t)w0 * h0 * 4, f) + lbl_off[l]);
            L.flags = chunk(o_flags, (size_t)w0 * h0, f) + (size_t)l * cnt_region;
            L.dt = nullptr; L.opt = nullptr;
            L.w = g[l].w; L.h = g[l].h; L.pts_cap = g[l].cap; L.patch = g[l].patch;
            L.hist_w = g[l].hist_w; L.hist_h = g[l].hist_h;
            L.fx = g[l].cam.fx; L.fy = g[l].cam.fy; L.cx = g[l].cam.cx; L.cy = g[l].cam.cy;
            host_desc[(size_t)l * n + f] = L;
        }
        pyrs[f] = p;
    }
    int stage_slot = -1;     // which of the context's two bgr staging buffers this call uploads into
    auto fail = [&](int code) {
        for (auto *p : pyrs) delete p;
        cudaStreamSynchronize(ctx->stream);   // error path only: nothing may still be using the slab
        cudaFreeAsync(slab->mem, up);
        delete slab;
        return code;
    };
    for (int l = 0; l < NL; ++l) {
        slab->d_desc[l] = (ImgLevel *)(base + o_desc[l]);
        e = cudaMemcpyAsync(slab->d_desc[l], &host_desc[(size_t)l * n], sizeof(ImgLevel) * (size_t)n, cudaMemcpyHostToDevice, up);
        if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memcpy(desc)"));
    }
    e = cudaMemsetAsync(base + o_counters, 0, sizeof(int) * 2 * NL * (size_t)n, up);
    if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memset(counters)"));

    // ---- inputs -------------------------------------------------------------------------------------
    const size_t bgr_frame = (size_t)w0 * h0 * channels;
    const uint8_t *d_bgr = bgr;
    if (host_in) {
        // persistent double buffer: no allocator dependency between this upload and the kernels of the previous batch
        stage_slot = ctx->stage_next;
        ctx->stage_next ^= 1;
        const size_t bgr_bytes = align_up(bgr_frame * (size_t)n, 256);
        const size_t need = bgr_bytes + (depth16 ? (size_t)w0 * h0 * 2 * (size_t)n : 0);
        if (ctx->stage_bytes[stage_slot] < need) {
            if (ctx->stage[stage_slot]) {
                cudaStreamSynchronize(ctx->stream);
                cudaStreamSynchronize(up);
                cudaFree(ctx->stage[stage_slot]);
                ctx->stage[stage_slot] = nullptr; ctx->stage_bytes[stage_slot] = 0; ctx->stage_used[stage_slot] = false;
            }
            e = cudaMalloc(&ctx->stage[stage_slot], need);
            if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "cudaMalloc(bgr staging)"));
            ctx->stage_bytes[stage_slot] = need;
        }
        if (ctx->stage_used[stage_slot]) cudaStreamWaitEvent(up, ctx->stage_consumed[stage_slot], 0);   // gray of two batches ago
        cudaEventRecord(ctx->ev[6], up);
        e = cudaMemcpyAsync(ctx->stage[stage_slot], bgr, bgr_frame * (size_t)n, cudaMemcpyHostToDevice, up);
        if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memcpy(bgr)"));
        d_bgr = (const uint8_t *)ctx->stage[stage_slot];
        if (depth16) {
            uint8_t *d16 = (uint8_t *)ctx->stage[stage_slot] + bgr_bytes;
            e = cudaMemcpyAsync(d16, depth16, (size_t)w0 * h0 * 2 * (size_t)n, cudaMemcpyHostToDevice, up);
            if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memcpy(depth16)"));
            depth16 = (const uint16_t *)d16;
        }
    }
    if (!depth16) {
        const size_t fb = (size_t)w0 * h0 * 4;
        e = cudaMemcpy2DAsync(base + o_depth[0], align_up(fb, 256), depth, fb, fb, (size_t)n, cudaMemcpyDefault, up);
        if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memcpy(depth)"));
    }
    if (host_in) {
        cudaEventRecord(ctx->ev[7], up);
        ctx->ev_valid[3] = true;
        cudaEvent_t uploaded;
        e = cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(uploaded, up);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, uploaded, 0);
        if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "upload event"));
        cudaEventDestroy(uploaded);   // released by the runtime once the wait has been satisfied
    }

    // ---- the pyramid (imgpyramidrgbd.cpp:43-96) -------------------------------------------------
    const double t1 = cfg->canny_threshold1, t2 = cfg->canny_threshold2;
    double lo = std::min(t1, t2), hi = std::max(t1, t2);
    lo = std::min(32767.0, lo); hi = std::min(32767.0, hi);
    if (lo > 0) lo *= lo;
    if (hi > 0) hi *= hi;
    const int low = (int)floor(lo), high = (int)floor(hi);

    // Launch structure: the level-0 Canny -> compaction chain (the longest) starts right behind the gray kernel; the gray
    // pyramid continues on the main stream and releases the chain of every level as its image appears; the depth pyramid runs
    // on its own stream (only the compaction of a level needs it).  Everything joins the main stream at the end.
    static const int serial_levels = getenv("REVO_PYR_SERIAL") ? atoi(getenv("REVO_PYR_SERIAL")) : 0;
    const bool fork = NL > 1 && !serial_levels;
    cudaStream_t main_stream = ctx->stream;
    cudaEventRecord(ctx->ev[0], main_stream);
    if (depth16) rc = launch_depth_u16(ctx, depth16, (size_t)w0 * h0, depth_scale, slab->d_desc[0], n, w0 * h0);
    if (fork) {
        cudaEventRecord(ctx->lvl_fork, main_stream);           // inputs (and the level-0 depth) are on the device
        ctx->stream = ctx->depth_stream;
        cudaStreamWaitEvent(ctx->stream, ctx->lvl_fork, 0);
        for (int l = 1; l < NL && !rc; ++l) {
            rc = launch_depth_half(ctx, slab->d_desc[l - 1], slab->d_desc[l], n, g[l].w, g[l].h, g[l - 1].w);
            cudaEventRecord(ctx->lvl_depth[l], ctx->stream);
        }
        ctx->stream = main_stream;
    }
    if (!rc) rc = launch_gray(ctx, d_bgr, (size_t)w0 * channels, channels, bgr_frame, slab->d_desc[0], n, w0, h0);
    if (stage_slot >= 0) {
        cudaEventRecord(ctx->stage_consumed[stage_slot], main_stream);     // (the depth stream read the staging buffer before lvl_fork)
        ctx->stage_used[stage_slot] = true;
    }
    for (int l = 0; l < NL && !rc; ++l) {
        if (l > 0) {
            if (fork) rc = launch_pyrdown(ctx, slab->d_desc[l - 1], slab->d_desc[l], n, g[l].w, g[l].h, g[l - 1].w, g[l - 1].h);
            else rc = launch_pyrdown_depth(ctx, slab->d_desc[l - 1], slab->d_desc[l], n, g[l].w, g[l].h, g[l - 1].w, g[l - 1].h);
        }
        if (fork) {
            cudaEventRecord(ctx->lvl_gray[l], main_stream);
            ctx->stream = ctx->lvl_stream[l];          // the launchers enqueue on ctx->stream
            cudaStreamWaitEvent(ctx->stream, ctx->lvl_gray[l], 0);
        }
        if (!rc) {
            alignas(64) unsigned char tmap[128];
            const bool tma = make_gray_tensor_map(tmap, base + o_gray[l], g[l].w, g[l].h, n, align_up((size_t)g[l].w * g[l].h, 256));
            rc = launch_canny(ctx, slab->d_desc[l], n, g[l].w, g[l].h, low, high, tma ? tmap : nullptr, g[l].patch,
                              base + o_flags + (size_t)l * cnt_region, align_up((size_t)w0 * h0, 256));
        }
        // fill-in is only defined for the reference's 3 patch sizes (levels 1,2); see SURVEY D5.  It reads the FINAL edge map
        // of the level above (after that level's own fill-in).
        const bool fill = cfg->use_edge_hist && l >= 1 && l <= 2;
        if (fork && l == 0) cudaEventRecord(ctx->lvl_canny0, ctx->stream);
        if (fork && fill) cudaStreamWaitEvent(ctx->stream, l == 1 ? ctx->lvl_canny0 : ctx->lvl_fill[l - 1], 0);
        if (!rc) rc = launch_hist_fill(ctx, slab->d_desc[l], l > 0 ? slab->d_desc[l - 1] : nullptr, n, g[l].w, g[l].h,
                                      g[l].patch, l > 0 ? g[l - 1].patch : g[l].patch, fill, cfg->n_percentage);
        if (fork && l >= 1) {
            cudaEventRecord(ctx->lvl_fill[l], ctx->stream);
            cudaStreamWaitEvent(ctx->stream, ctx->lvl_depth[l], 0);        // the compaction reads the level's depth image
        }
        if (!rc) rc = launch_compact(ctx, slab->d_desc[l], n, g[l].w, g[l].h, cfg->depth_min, cfg->depth_max);
        if (fork) cudaEventRecord(ctx->lvl_done[l], ctx->stream);
        ctx->stream = main_stream;
    }
    if (fork)
        for (int l = 0; l < NL; ++l) cudaStreamWaitEvent(main_stream, ctx->lvl_done[l], 0);
    if (rc) return fail(rc);
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->ev_valid[0] = true;
    if (cudaEventCreateWithFlags(&slab->ready, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(slab->ready, ctx->stream);
    for (int f = 0; f < n; ++f) pyr_out[f] = pyrs[f];
    return REVO_OK;
}

int revo_pyr_create_batch(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, int n, const uint8_t *bgr,
                          int channels, const float *depth, const double *timestamps, revo_pyr **pyr_out)
{
    if (!depth) return REVO_ERR_INVALID_ARG;
    return create_batch_impl(ctx, cfg, cam0, n, bgr, channels, depth, nullptr, 0.f, timestamps, pyr_out);
}

int revo_pyr_create_batch_u16(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, int n, const uint8_t *bgr,
                              int channels, const uint16_t *depth_raw, float depth_scale, const double *timestamps,
                              revo_pyr **pyr_out)
{
    if (!depth_raw) return REVO_ERR_INVALID_ARG;
    return create_batch_impl(ctx, cfg, cam0, n, bgr, channels, nullptr, depth_raw, depth_scale, timestamps, pyr_out);
}

int revo_pyr_create(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, const uint8_t *bgr, size_t bgr_stride,
                    int channels, const float *depth, size_t depth_stride, double timestamp, revo_pyr **pyr_out)
{
    if (!ctx || !cfg || !cam0 || !bgr || !depth || !pyr_out) return REVO_ERR_INVALID_ARG;
    const size_t tight_bgr = (size_t)cam0->width * channels, tight_d = (size_t)cam0->width * 4;
    if (bgr_stride == 0) bgr_stride = tight_bgr;
    if (depth_stride == 0) depth_stride = tight_d;
    if (bgr_stride == tight_bgr && depth_stride == tight_d)
        return revo_pyr_create_batch(ctx, cfg, cam0, 1, bgr, channels, depth, &timestamp, pyr_out);
    // strided inputs (cv::Mat ROI): repack rows on the host, then take the tight path
    if (is_device_ptr(bgr) || is_device_ptr(depth)) return REVO_ERR_UNSUPPORTED;
    std::vector<uint8_t> b(tight_bgr * cam0->height);
    std::vector<float> d((size_t)cam0->width * cam0->height);
    for (int y = 0; y < cam0->height; ++y) {
        memcpy(b.data() + tight_bgr * y, bgr + bgr_stride * y, tight_bgr);
        memcpy((uint8_t *)d.data() + tight_d * y, (const uint8_t *)depth + depth_stride * y, tight_d);
    }
    int rc = revo_pyr_create_batch(ctx, cfg, cam0, 1, b.data(), channels, d.data(), &timestamp, pyr_out);
    if (!rc) cudaStreamSynchronize(ctx->stream);   // the temporaries die here
    return rc;
}

// A pyramid built on another context's stream: order this context's stream after the build (no host sync).
// handles made by revo_pyr_copy_points_batch own one point list and nothing else
static bool points_only(const revo_pyr *p) { return p && !p->lv[0].gray; }

static void wait_for_build(revo_ctx *ctx, const revo_pyr *p)
{
    if (p && p->slab && p->slab->ready && p->slab->stream != ctx->stream) cudaStreamWaitEvent(ctx->stream, p->slab->ready, 0);
}

// One stream-ordered allocation for the keyframe structures (dt 4 B/px + tiled lookup texels 8 B/px, all levels) of
// every pyramid in `ps` that does not have them yet.
static int alloc_keyframe_mem(revo_ctx *ctx, revo_pyr *const *ps, int n)
{
    auto bytes_of = [](const revo_pyr *p) {
        size_t b = 0;
        for (int l = 0; l < p->n_levels; ++l) b += align_up(opt_bytes(p->lv[l].w, p->lv[l].h) + (size_t)p->lv[l].w * p->lv[l].h * 4, 256);
        return b;
    };
    size_t total = 0;
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (!ps[i]->kf_slab) { total += bytes_of(ps[i]); ++m; }
    if (!m) return REVO_OK;
    KfSlab *ks = new (std::nothrow) KfSlab();
    if (!ks) return REVO_ERR_INVALID_ARG;
    ks->mem = nullptr; ks->live = m;
    cudaError_t e = cudaMallocAsync(&ks->mem, total, ctx->stream);
    if (e != cudaSuccess) { delete ks; return cuda_fail(ctx, e, "cudaMallocAsync(keyframe)"); }
    uint8_t *mem = (uint8_t *)ks->mem;
    for (int i = 0; i < n; ++i) {
        revo_pyr *p = ps[i];
        if (p->kf_slab) continue;
        p->kf_slab = ks;
        for (int l = 0; l < p->n_levels; ++l) {
            const size_t px = (size_t)p->lv[l].w * p->lv[l].h, ob = opt_bytes(p->lv[l].w, p->lv[l].h);
            p->lv[l].opt = (uint2 *)mem;
            p->lv[l].dt = (float *)(mem + ob);
            mem += align_up(ob + px * 4, 256);
        }
    }
    return REVO_OK;
}

int revo_pyr_make_keyframe_batch(revo_ctx *ctx, int n, revo_pyr *const *pyrs)
{
    if (!ctx || !pyrs || n < 0) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<revo_pyr *> todo;
    for (int i = 0; i < n; ++i) {
        if (!pyrs[i]) return REVO_ERR_INVALID_ARG;
        if (points_only(pyrs[i])) return REVO_ERR_UNSUPPORTED;
        if (!pyrs[i]->is_keyframe && std::find(todo.begin(), todo.end(), pyrs[i]) == todo.end()) todo.push_back(pyrs[i]);
    }
    if (todo.empty()) return REVO_OK;
    const int m = (int)todo.size();
    const int NL = todo[0]->n_levels;
    for (auto *p : todo)
        if (p->n_levels != NL || p->lv[0].w != todo[0]->lv[0].w || p->lv[0].h != todo[0]->lv[0].h) return REVO_ERR_INVALID_ARG;
    for (size_t i = 0; i < todo.size(); ++i)
        if (i == 0 || todo[i]->slab != todo[i - 1]->slab) wait_for_build(ctx, todo[i]);
    {
        int rc = alloc_keyframe_mem(ctx, todo.data(), m);
        if (rc) return rc;
    }
    // temporary descriptor tables (with dt/opt set) in a stream-ordered allocation, filled from pinned mapped host memory by
    // a kernel (not by the copy engine, which may be busy for milliseconds with the frame uploads of the next batches)
    const size_t tab_bytes = sizeof(ImgLevel) * (size_t)NL * m;
    if (ctx->pinned_kf_busy) { REVO_CUDA(ctx, cudaEventSynchronize(ctx->pinned_kf_read)); ctx->pinned_kf_busy = false; }
    if (ctx->pinned_kf_bytes < tab_bytes + 16) {
        if (ctx->pinned_kf) REVO_CUDA(ctx, cudaFreeHost(ctx->pinned_kf));
        ctx->pinned_kf = nullptr;
        ctx->pinned_kf_bytes = align_up(2 * tab_bytes + 16, 1 << 16);
        REVO_CUDA(ctx, cudaMallocHost(&ctx->pinned_kf, ctx->pinned_kf_bytes));
    }
    ImgLevel *host = (ImgLevel *)ctx->pinned_kf;
    for (int l = 0; l < NL; ++l)
        for (int i = 0; i < m; ++i) host[(size_t)l * m + i] = todo[i]->lv[l];
    ImgLevel *d_tab = nullptr;
    REVO_CUDA(ctx, cudaMallocAsync((void **)&d_tab, tab_bytes + 16, ctx->stream));
    {
        int rc = launch_stage_in(ctx, host, d_tab, tab_bytes);
        if (rc) return rc;
        cudaEventRecord(ctx->pinned_kf_read, ctx->stream);
        ctx->pinned_kf_busy = true;
    }
    int rc = REVO_OK;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    for (int l = 0; l < NL && !rc; ++l) rc = launch_keyframe(ctx, d_tab + (size_t)l * m, m, todo[0]->lv[l].w, todo[0]->lv[l].h);
    cudaEventRecord(ctx->ev[3], ctx->stream);
    ctx->ev_valid[1] = true;
    cudaFreeAsync(d_tab, ctx->stream);
    if (rc) return rc;
    for (auto *p : todo) p->is_keyframe = true;
    return REVO_OK;
}

int revo_pyr_make_keyframe(revo_ctx *ctx, revo_pyr *pyr) { return revo_pyr_make_keyframe_batch(ctx, 1, &pyr); }

static void destroy_one(revo_ctx *ctx, revo_pyr *pyr)
{
    if (KfSlab *k = pyr->kf_slab) {
        if (--k->live == 0) {
            cudaFreeAsync(k->mem, ctx->stream);
            delete k;
        }
    }
    Slab *s = pyr->slab;
    if (s && --s->live == 0) {
        if (s->ready) cudaEventDestroy(s->ready);
        cudaFreeAsync(s->mem, ctx->stream);
        delete s;
    }
    delete pyr;
}

int revo_pyr_destroy(revo_ctx *ctx, revo_pyr *pyr)
{
    if (!pyr) return REVO_OK;
    if (!ctx) return REVO_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    destroy_one(ctx, pyr);
    (void)cudaGetLastError();
    return REVO_OK;
}

int revo_pyr_destroy_batch(revo_ctx *ctx, int n, revo_pyr *const *pyrs)
{
    if (!ctx || (n > 0 && !pyrs)) return REVO_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < n; ++i)
        if (pyrs[i]) destroy_one(ctx, pyrs[i]);
    (void)cudaGetLastError();
    return REVO_OK;
}

int revo_pyr_copy_points_batch(revo_ctx *ctx, int n, revo_pyr *const *pyrs, int lvl, revo_pyr **out)
{
    if (!ctx || n < 0 || (n > 0 && (!pyrs || !out))) return REVO_ERR_INVALID_ARG;
    if (n == 0) return REVO_OK;
    for (int i = 0; i < n; ++i) {
        if (!pyrs[i]) return REVO_ERR_INVALID_ARG;
        if (lvl < 0 || lvl >= pyrs[i]->n_levels || !pyrs[i]->lv[lvl].pts) return REVO_ERR_BAD_LEVEL;
    }
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    // The lists' lengths live on the device and are not read back here (no host synchronisation between two frames): a copy is
    // allocated at the source list's capacity (w h / 2 + 1024 points: 170 KB at level 2 of VGA) and filled to its length.
    for (int i = 0; i < n; ++i)
        if (i == 0 || pyrs[i]->slab != pyrs[i - 1]->slab) wait_for_build(ctx, pyrs[i]);
    {
        int rc = ensure_scratch(ctx, align_up(sizeof(PointListCopy) * (size_t)n, 256));
        if (rc) return rc;
    }
    std::vector<int> cnt((size_t)n);
    for (int i = 0; i < n; ++i) cnt[i] = pyrs[i]->lv[lvl].pts_cap;
    std::vector<size_t> off((size_t)n + 1);
    off[0] = align_up(8 * (size_t)n, 256);
    for (int i = 0; i < n; ++i) off[i + 1] = off[i] + align_up((size_t)std::max(cnt[i], 1) * 16, 256);
    Slab *slab = new (std::nothrow) Slab();
    if (!slab) return REVO_ERR_INVALID_ARG;
    memset(slab, 0, sizeof(*slab));
    slab->n_frames = n; slab->live = n; slab->bytes = off[n]; slab->stream = ctx->stream;
    cudaError_t e = cudaMallocAsync(&slab->mem, slab->bytes, ctx->stream);
    if (e != cudaSuccess) { delete slab; return cuda_fail(ctx, e, "cudaMallocAsync(point lists)"); }
    std::vector<PointListCopy> tab((size_t)n);
    for (int i = 0; i < n; ++i) {
        revo_pyr *p = new revo_pyr();
        memset(p->lv, 0, sizeof(p->lv));
        p->slab = slab; p->index_in_slab = i; p->n_levels = pyrs[i]->n_levels; p->cfg = pyrs[i]->cfg; p->cam0 = pyrs[i]->cam0;
        p->timestamp = pyrs[i]->timestamp; p->kf_slab = nullptr; p->is_keyframe = false;
        for (int l = 0; l < p->n_levels; ++l) {     // cameras and sizes of every level; arrays only at `lvl`
            const ImgLevel &S = pyrs[i]->lv[l];
            ImgLevel &L = p->lv[l];
            L.w = S.w; L.h = S.h; L.fx = S.fx; L.fy = S.fy; L.cx = S.cx; L.cy = S.cy; L.patch = S.patch; L.hist_w = S.hist_w; L.hist_h = S.hist_h;
        }
        ImgLevel &L = p->lv[lvl];
        L.pts = (float4 *)((uint8_t *)slab->mem + off[i]);
        L.n_pts = (int *)slab->mem + 2 * i;
        L.pts_cap = std::max(cnt[i], 1);
        tab[i] = PointListCopy{pyrs[i]->lv[lvl].pts, pyrs[i]->lv[lvl].n_pts, L.pts, L.n_pts, cnt[i]};
        out[i] = p;
    }
    PointListCopy *d_tab = (PointListCopy *)ctx->scratch;
    e = cudaMemcpyAsync(d_tab, tab.data(), sizeof(PointListCopy) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
    int rc = e == cudaSuccess ? launch_copy_point_lists(ctx, d_tab, n) : cuda_fail(ctx, e, "memcpy(point list table)");
    if (rc) {
        for (int i = 0; i < n; ++i) { delete out[i]; out[i] = nullptr; }
        cudaStreamSynchronize(ctx->stream);
        cudaFreeAsync(slab->mem, ctx->stream);
        delete slab;
        return rc;
    }
    return REVO_OK;
}

int revo_pyr_is_keyframe(const revo_pyr *pyr) { return pyr && pyr->is_keyframe; }
double revo_pyr_timestamp(const revo_pyr *pyr) { return pyr ? pyr->timestamp : 0.0; }

int revo_pyr_level_camera(const revo_pyr *pyr, int lvl, revo_camera *cam_out)
{
    if (!pyr || !cam_out) return REVO_ERR_INVALID_ARG;
    if (lvl < 0 || lvl >= pyr->n_levels) return REVO_ERR_BAD_LEVEL;
    const ImgLevel &L = pyr->lv[lvl];
    cam_out->fx = L.fx; cam_out->fy = L.fy; cam_out->cx = L.cx; cam_out->cy = L.cy; cam_out->width = L.w; cam_out->height = L.h;
    return REVO_OK;
}

int revo_pyr_num_edges(revo_ctx *ctx, const revo_pyr *pyr, int lvl, int *n_out)
{
    if (!ctx || !pyr || !n_out) return REVO_ERR_INVALID_ARG;
    if (lvl < 0 || lvl >= pyr->n_levels) return REVO_ERR_BAD_LEVEL;
    if (!pyr->lv[lvl].n_pts) return REVO_ERR_UNSUPPORTED;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    REVO_CUDA(ctx, cudaMemcpyAsync(n_out, pyr->lv[lvl].n_pts, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

int revo_pyr_download(revo_ctx *ctx, const revo_pyr *pyr, int lvl, int which, void *dst, size_t dst_bytes, size_t *bytes_out)
{
    if (!ctx || !pyr) return REVO_ERR_INVALID_ARG;
    if (lvl < 0 || lvl >= pyr->n_levels) return REVO_ERR_BAD_LEVEL;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    const ImgLevel &L = pyr->lv[lvl];
    const size_t px = (size_t)L.w * L.h;
    const void *src = nullptr;
    size_t bytes = 0;
    if (points_only(pyr) && !(which == REVO_ARRAY_EDGES3D_DEVICE_ORDER && L.pts)) return REVO_ERR_UNSUPPORTED;
    switch (which) {
        case REVO_ARRAY_GRAY: src = L.gray; bytes = px; break;
        case REVO_ARRAY_DEPTH: src = L.depth; bytes = px * 4; break;
        case REVO_ARRAY_EDGES: src = L.edges; bytes = px; break;
        case REVO_ARRAY_EDGES_ORIG: src = L.edges_orig; bytes = px; break;
        case REVO_ARRAY_HIST: src = L.hist; bytes = (size_t)L.hist_w * L.hist_h; break;
        case REVO_ARRAY_DT:
            if (!pyr->is_keyframe) return REVO_ERR_NOT_KEYFRAME;
            src = L.dt; bytes = px * 4; break;
        case REVO_ARRAY_OPTSTRUCT: {
            // the reference's float4 layout is materialised on demand from the distance transform
            if (!pyr->is_keyframe) return REVO_ERR_NOT_KEYFRAME;
            int rc = ensure_scratch(ctx, px * 16);
            if (rc) return rc;
            rc = launch_opt_struct_f4(ctx, L.dt, L.w, L.h, (float4 *)ctx->scratch);
            if (rc) return rc;
            src = ctx->scratch; bytes = px * 16; break;
        }
        case REVO_ARRAY_EDGES3D_DEVICE_ORDER: {
            int n = 0;
            int rc = revo_pyr_num_edges(ctx, pyr, lvl, &n);
            if (rc) return rc;
            src = L.pts; bytes = (size_t)n * 16; break;
        }
        case REVO_ARRAY_EDGES3D: {
            // reference order: column-major scan (imgpyramidrgbd.cpp:203-205)
            const size_t need = px * 16 + (size_t)(L.w + 4) * 4 + sizeof(ImgLevel) + 1024;
            int rc = ensure_scratch(ctx, need);
            if (rc) return rc;
            uint8_t *s = (uint8_t *)ctx->scratch;
            float4 *d_out = (float4 *)s;
            int *d_col = (int *)(s + align_up(px * 16, 256));
            int *d_n = d_col + L.w + 1;
            ImgLevel *d_desc = (ImgLevel *)(s + align_up(px * 16, 256) + align_up((size_t)(L.w + 4) * 4, 256));
            REVO_CUDA(ctx, cudaMemcpyAsync(d_desc, &L, sizeof(ImgLevel), cudaMemcpyHostToDevice, ctx->stream));
            rc = launch_edges3d_reference_order(ctx, d_desc, L.w, L.h, pyr->cfg.depth_min, pyr->cfg.depth_max, d_out, d_n, d_col);
            if (rc) return rc;
            int n = 0;
            REVO_CUDA(ctx, cudaMemcpyAsync(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            src = d_out; bytes = (size_t)n * 16; break;
        }
        default: return REVO_ERR_INVALID_ARG;
    }
    if (bytes_out) *bytes_out = bytes;
    if (!dst) return REVO_OK;   // size query
    if (dst_bytes < bytes) return REVO_ERR_BUFFER_TOO_SMALL;
    if (bytes) REVO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

int revo_pyr_colored_pcl(revo_ctx *ctx, const revo_pyr *pyr, int lvl, int dense, const uint8_t *bgr, int channels, float *out,
                         size_t capacity_points, int *n_out)
{
    if (!ctx || !pyr || !bgr || !n_out || (channels != 3 && channels != 4)) return REVO_ERR_INVALID_ARG;
    *n_out = 0;
    if (lvl < 0) return REVO_ERR_BAD_LEVEL;
    if (points_only(pyr)) return REVO_ERR_UNSUPPORTED;
    if (lvl > 2) return REVO_OK;      // the reference only has a colour image for levels 0..2 (imgpyramidrgbd.cpp:288-297)
    if (lvl >= pyr->n_levels) return REVO_ERR_BAD_LEVEL;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    const ImgLevel &L = pyr->lv[lvl];
    const int w0 = pyr->lv[0].w, h0 = pyr->lv[0].h;
    // scratch: colour pyramid (two ping-pong planes of the level-0 size) | column offsets | count | descriptor | cloud
    const size_t b_img = align_up((size_t)w0 * h0 * channels, 256), b_col = align_up((size_t)(L.w + 4) * 4, 256);
    const size_t b_out = align_up((size_t)L.w * L.h * 32, 256);
    int rc = ensure_scratch(ctx, 2 * b_img + b_col + 256 + 256 + b_out);
    if (rc) return rc;
    uint8_t *s = (uint8_t *)ctx->scratch;
    uint8_t *img[2] = {s, s + b_img};
    int *d_col = (int *)(s + 2 * b_img), *d_n = (int *)(s + 2 * b_img + b_col);
    ImgLevel *d_desc = (ImgLevel *)(s + 2 * b_img + b_col + 256);
    float *d_out = (float *)(s + 2 * b_img + b_col + 512);
    wait_for_build(ctx, pyr);
    REVO_CUDA(ctx, cudaMemcpyAsync(img[0], bgr, (size_t)w0 * h0 * channels, cudaMemcpyDefault, ctx->stream));
    int cur = 0, cw = w0, chh = h0;
    for (int l = 0; l < lvl && !rc; ++l) {
        rc = launch_pyrdown_color(ctx, img[cur], img[cur ^ 1], cw, chh, channels);
        cw = (cw + 1) / 2; chh = (chh + 1) / 2;
        cur ^= 1;
    }
    if (rc) return rc;
    if (cw != L.w || chh != L.h) return REVO_ERR_UNSUPPORTED;
    REVO_CUDA(ctx, cudaMemcpyAsync(d_desc, &L, sizeof(ImgLevel), cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_colored_pcl(ctx, d_desc, L.w, L.h, dense, pyr->cfg.depth_min, pyr->cfg.depth_max, img[cur], channels, out ? d_out : nullptr,
                            L.w * L.h, d_n, d_col);
    if (rc) return rc;
    int n = 0;
    REVO_CUDA(ctx, cudaMemcpyAsync(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    if (!out) return REVO_OK;
    if ((size_t)n > capacity_points) return REVO_ERR_BUFFER_TOO_SMALL;
    if (n) REVO_CUDA(ctx, cudaMemcpyAsync(out, d_out, (size_t)n * 32, cudaMemcpyDefault, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

int revo_pyr_upload_level(revo_ctx *ctx, revo_pyr *pyr, int lvl, const float *pts4, int n, const float *dt, const float *opt4)
{
    if (!ctx || !pyr) return REVO_ERR_INVALID_ARG;
    if (lvl < 0 || lvl >= pyr->n_levels) return REVO_ERR_BAD_LEVEL;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    ImgLevel &L = pyr->lv[lvl];
    const size_t px = (size_t)L.w * L.h;
    if (pts4) {
        if (n < 0 || n > L.pts_cap) return REVO_ERR_BUFFER_TOO_SMALL;
        if (n) REVO_CUDA(ctx, cudaMemcpyAsync(L.pts, pts4, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
        REVO_CUDA(ctx, cudaMemcpyAsync(L.n_pts, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (dt || opt4) {
        int rc = alloc_keyframe_mem(ctx, &pyr, 1);
        if (rc) return rc;
        if (dt) REVO_CUDA(ctx, cudaMemcpyAsync(L.dt, dt, px * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (opt4) {
            rc = ensure_scratch(ctx, px * 16);
            if (rc) return rc;
            REVO_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, opt4, px * 16, cudaMemcpyHostToDevice, ctx->stream));
            rc = launch_opt_pack_from_f4(ctx, (const float4 *)ctx->scratch, L.w, L.h, L.opt);
            if (rc) return rc;
            pyr->is_keyframe = true;
        }
    }
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

// ---------------------------------------------------------------------------------------------------
// tracking
// ---------------------------------------------------------------------------------------------------
static int fill_pair(const revo_pyr *ref, const revo_pyr *cur, int min_lvl, int max_lvl, const float *R9, const float *t3,
                     PairDesc *d)
{
    if (!ref || !cur) return REVO_ERR_INVALID_ARG;
    if (!ref->is_keyframe) return REVO_ERR_NOT_KEYFRAME;
    if (points_only(cur)) return REVO_ERR_UNSUPPORTED;
    if (min_lvl < max_lvl || max_lvl < 0 || min_lvl >= cur->n_levels || min_lvl >= ref->n_levels) return REVO_ERR_BAD_LEVEL;
    memset(d, 0, sizeof(*d));
    for (int l = max_lvl; l <= min_lvl; ++l) {
        const ImgLevel &c = cur->lv[l], &r = ref->lv[l];
        if (c.w != r.w || c.h != r.h) return REVO_ERR_INVALID_ARG;
        LevelIn &L = d->lvl[l];
        L.pts = c.pts; L.n_pts = c.n_pts; L.opt = r.opt;
        // calcErrorAndBuffers takes the camera from the reference frame (optimizer.cpp:80)
        L.fx = r.fx; L.fy = r.fy; L.cx = r.cx; L.cy = r.cy; L.w = r.w; L.h = r.h;
    }
    d->ref_dt_min = ref->lv[min_lvl].dt;
    memcpy(d->R, R9, sizeof(float) * 9);
    memcpy(d->t, t3, sizeof(float) * 3);
    return REVO_OK;
}

static int run_track(revo_ctx *ctx, TrackParams &prm, int n, revo_pyr *const *refs, revo_pyr *const *curs, const float *R9s,
                     const float *t3s, revo_track_result *results, double *records, revo_trace_entry *trace, int trace_cap,
                     int *trace_counts)
{
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    const int min_lvl = prm.mode == 0 ? prm.cfg.pyr_min_lvl : prm.level;
    const int max_lvl = prm.mode == 0 ? prm.cfg.pyr_max_lvl : prm.level;
    // descriptors are written straight into pinned, device-mapped host memory (see k_stage_in)
    {
        int rc = ensure_pinned(ctx, align_up(sizeof(PairDesc) * (size_t)n, 16));
        if (rc) return rc;
    }
    PairDesc *host = (PairDesc *)ctx->pinned;
    for (int i = 0; i < n; ++i) {
        int rc = fill_pair(refs[i], curs[i], min_lvl, max_lvl, R9s + 9 * (size_t)i, t3s + 3 * (size_t)i, &host[i]);
        if (rc) return rc;
        // one device-side wait per distinct batch (the pairs of a batch share two slabs)
        if (i == 0 || refs[i]->slab != refs[i - 1]->slab) wait_for_build(ctx, refs[i]);
        if (i == 0 || curs[i]->slab != curs[i - 1]->slab) wait_for_build(ctx, curs[i]);
    }
    if (!trace) trace_cap = 0;
    prm.trace_cap = trace_cap;
    prm.profile = getenv("REVO_TRACK_PROF") != nullptr;
    prm.speculate = !(getenv("REVO_TRACK_SPEC") && atoi(getenv("REVO_TRACK_SPEC")) == 0);   // A/B switch for profiling
    // device workspace: pairs | results | records | trace | trace counts
    const size_t b_pairs = align_up(sizeof(PairDesc) * (size_t)n, 256);
    const size_t b_res = align_up(sizeof(revo_track_result) * (size_t)n, 256);
    const size_t b_rec = align_up(sizeof(double) * 32 * (size_t)n, 256);
    const size_t b_tr = align_up(sizeof(revo_trace_entry) * (size_t)trace_cap * n, 256);
    const size_t b_tc = align_up(sizeof(int) * (size_t)n, 256);
    uint8_t *ws = nullptr;
    REVO_CUDA(ctx, cudaMallocAsync((void **)&ws, b_pairs + b_res + b_rec + b_tr + b_tc + 256, ctx->stream));
    PairDesc *d_pairs = (PairDesc *)ws;
    revo_track_result *d_res = (revo_track_result *)(ws + b_pairs);
    double *d_rec = (double *)(ws + b_pairs + b_res);
    revo_trace_entry *d_tr = trace_cap ? (revo_trace_entry *)(ws + b_pairs + b_res + b_rec) : nullptr;
    int *d_tc = (int *)(ws + b_pairs + b_res + b_rec + b_tr);
    int *d_wc = (int *)(ws + b_pairs + b_res + b_rec + b_tr + b_tc);
    int rc = REVO_OK;
    rc = launch_stage_in(ctx, host, d_pairs, sizeof(PairDesc) * (size_t)n);
    cudaError_t e = cudaMemsetAsync(d_res, 0, b_res + b_rec, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_wc, 0, 256, ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "track upload");
    cudaEventRecord(ctx->ev[4], ctx->stream);
    if (!rc) rc = launch_track(ctx, d_pairs, n, prm, d_res, d_rec, d_tr, d_tc, d_wc);
    cudaEventRecord(ctx->ev[5], ctx->stream);
    ctx->ev_valid[2] = true;
    if (!rc && results) {
        e = cudaMemcpyAsync(results, d_res, sizeof(revo_track_result) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "results download");
    }
    if (!rc && records) {
        e = cudaMemcpyAsync(records, d_rec, sizeof(double) * 32 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "records download");
    }
    if (!rc && trace && trace_cap) {
        e = cudaMemcpyAsync(trace, d_tr, sizeof(revo_trace_entry) * (size_t)trace_cap * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && trace_counts)
            e = cudaMemcpyAsync(trace_counts, d_tc, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "trace download");
    }
    unsigned long long prof[4] = {0, 0, 0, 0};
    const bool want_prof = getenv("REVO_TRACK_PROF") != nullptr;
    if (!rc && want_prof) cudaMemcpyAsync(prof, d_wc + 2, sizeof(prof), cudaMemcpyDeviceToHost, ctx->stream);
    cudaFreeAsync(ws, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess && !rc) rc = cuda_fail(ctx, e, "track kernel");
    if (!rc && want_prof && prof[3])
        fprintf(stderr, "[k_track prof] pairs %d evals %llu  cycles/eval: gather %.0f reduce %.0f serial+sync %.0f\n", n, prof[3],
                (double)prof[0] / prof[3], (double)prof[1] / prof[3], (double)prof[2] / prof[3]);
    return rc;
}

int revo_track_batch(revo_ctx *ctx, const revo_tracker_config *cfg, int n, revo_pyr *const *refs, revo_pyr *const *curs,
                     const float *R9s, const float *t3s, revo_track_result *results, revo_trace_entry *trace, int trace_cap,
                     int *trace_counts)
{
    if (!ctx || !cfg || !refs || !curs || !R9s || !t3s || !results || n < 0) return REVO_ERR_INVALID_ARG;
    if (n == 0) return REVO_OK;
    TrackParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.cfg = *cfg; prm.mode = 0; prm.split_world = 1;
    return run_track(ctx, prm, n, refs, curs, R9s, t3s, results, nullptr, trace, trace_cap, trace_counts);
}

int revo_track(revo_ctx *ctx, const revo_tracker_config *cfg, const revo_pyr *ref, const revo_pyr *cur, float *R9, float *t3,
               revo_track_result *result)
{
    if (!ctx || !cfg || !ref || !cur || !R9 || !t3) return REVO_ERR_INVALID_ARG;
    revo_track_result res;
    revo_pyr *r = const_cast<revo_pyr *>(ref), *c = const_cast<revo_pyr *>(cur);
    int rc = revo_track_batch(ctx, cfg, 1, &r, &c, R9, t3, &res, nullptr, 0, nullptr);
    if (rc) return rc;
    if (result) *result = res;
    if (res.rc) return res.rc;
    memcpy(R9, res.R, sizeof(float) * 9);
    memcpy(t3, res.t, sizeof(float) * 3);
    return REVO_OK;
}

int revo_track_level(revo_ctx *ctx, const revo_opt_config *cfg, const revo_pyr *ref, const revo_pyr *cur, int lvl, float *R9,
                     float *t3, revo_residual_info *res, float *err, int *n_evals)
{
    if (!ctx || !cfg || !ref || !cur || !R9 || !t3) return REVO_ERR_INVALID_ARG;
    TrackParams prm;
    memset(&prm, 0, sizeof(prm));
    revo_tracker_config_default(&prm.cfg);
    prm.cfg.opt = *cfg; prm.cfg.check_init_values = 0;
    prm.mode = 1; prm.level = lvl; prm.split_world = 1;
    revo_track_result out;
    revo_pyr *r = const_cast<revo_pyr *>(ref), *c = const_cast<revo_pyr *>(cur);
    int rc = run_track(ctx, prm, 1, &r, &c, R9, t3, &out, nullptr, nullptr, 0, nullptr);
    if (rc) return rc;
    if (out.rc) return out.rc;
    memcpy(R9, out.R, sizeof(float) * 9);
    memcpy(t3, out.t, sizeof(float) * 3);
    if (res) *res = out.res;
    if (err) *err = out.error;
    if (n_evals) *n_evals = out.n_evals[lvl];
    return REVO_OK;
}

int revo_eval(revo_ctx *ctx, const revo_opt_config *cfg, const revo_pyr *ref, const revo_pyr *cur, int lvl, const float *R9,
              const float *t3, double *record32)
{
    if (!ctx || !cfg || !ref || !cur || !R9 || !t3 || !record32) return REVO_ERR_INVALID_ARG;
    TrackParams prm;
    memset(&prm, 0, sizeof(prm));
    revo_tracker_config_default(&prm.cfg);
    prm.cfg.opt = *cfg; prm.cfg.check_init_values = 0;
    prm.mode = 2; prm.level = lvl; prm.split_world = 1;
    revo_pyr *r = const_cast<revo_pyr *>(ref), *c = const_cast<revo_pyr *>(cur);
    return run_track(ctx, prm, 1, &r, &c, R9, t3, nullptr, record32, nullptr, 0, nullptr);
}

// ---------------------------------------------------------------------------------------------------
// tracking-quality vote
// ---------------------------------------------------------------------------------------------------
static bool invert4(const double *m /* column-major */, double *inv)
{
    // Gauss-Jordan with partial pivoting on [m | I]
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { a[r][c] = m[c * 4 + r]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        for (int r = k + 1; r < 4; ++r)
            if (fabs(a[r][k]) > fabs(a[piv][k])) piv = r;
        if (fabs(a[piv][k]) < 1e-300) return false;
        if (piv != k)
            for (int c = 0; c < 8; ++c) std::swap(a[piv][c], a[k][c]);
        const double d = 1.0 / a[k][k];
        for (int c = 0; c < 8; ++c) a[k][c] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == k) continue;
            const double f = a[r][k];
            if (f != 0.0)
                for (int c = 0; c < 8; ++c) a[r][c] -= f * a[k][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) inv[c * 4 + r] = a[r][4 + c];
    return true;
}

// One vote's host half: transforms inv(estimatedPose) * pastWorldPose (tracker.cpp:147), pointers; false = bad argument.
static int quality_prepare(revo_ctx *ctx, const revo_pyr *cur, int hist_level, int n_past, revo_pyr *const *past, const float *past_world_poses16,
                           const float *estimated_pose16, int n_frames_voting, QualityArgs &a, revo_quality_result *out)
{
    memset(out, 0, sizeof(*out));
    memset(&a, 0, sizeof(a));
    out->status = REVO_TRACKER_STATE_OK;
    if (!cur || n_past < 0 || (n_past > 0 && (!past || !past_world_poses16)) || !estimated_pose16) return REVO_ERR_INVALID_ARG;
    if (hist_level < 0 || hist_level >= cur->n_levels) return REVO_ERR_BAD_LEVEL;
    int nf = n_past < n_frames_voting ? n_past : n_frames_voting;
    if (nf > 3) nf = 3;                        // histWeights has four entries (tracker.cpp:231-234)
    out->n_frames = nf > 0 ? nf : 0;
    const ImgLevel &L = cur->lv[hist_level];
    a.n_frames = out->n_frames; a.fx = L.fx; a.fy = L.fy; a.cx = L.cx; a.cy = L.cy; a.w = L.w; a.h = L.h;
    if (points_only(cur)) return REVO_ERR_UNSUPPORTED;
    a.depth = L.depth;
    // returnOrigEdges(histogramLevel): the Canny output before the fill-in (imgpyramidrgbd.h:69-77)
    a.edges = (cur->cfg.use_edge_hist && hist_level > 0) ? L.edges_orig : L.edges;
    if (nf <= 0) return REVO_OK;               // tracker.cpp:121: nothing to vote with
    double est[16], est_inv[16];
    for (int i = 0; i < 16; ++i) est[i] = estimated_pose16[i];
    if (!invert4(est, est_inv)) return REVO_ERR_INVALID_ARG;
    for (int f = 0; f < nf; ++f) {
        if (!past[f] || hist_level >= past[f]->n_levels || !past[f]->lv[hist_level].pts) return REVO_ERR_INVALID_ARG;
        const float *pw = past_world_poses16 + 16 * (size_t)f;
        double tr[16];
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) {
                double s = 0;
                for (int k = 0; k < 4; ++k) s += est_inv[k * 4 + r] * (double)pw[c * 4 + k];
                tr[c * 4 + r] = s;
            }
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) a.fr[f].R[c * 3 + r] = (float)tr[c * 4 + r];
        for (int r = 0; r < 3; ++r) a.fr[f].T[r] = (float)tr[12 + r];
        a.fr[f].pts = past[f]->lv[hist_level].pts;
        a.fr[f].n_pts = past[f]->lv[hist_level].n_pts;
        wait_for_build(ctx, past[f]);
    }
    wait_for_build(ctx, cur);
    return REVO_OK;
}

static void quality_finish(const int *c, revo_quality_result *out)
{
    static const float kHistWeights[4] = {0.f, 1.f, 1.25f, 1.5f};
    const int nf = out->n_frames;
    float measure = 0.f;
    for (int k = 0; k < 4; ++k) { out->histogram[k] = c[k]; out->overlaps[k] = c[4 + k]; }
    for (int k = 1; k <= nf; ++k) measure += (float)c[4 + k] * kHistWeights[k];     // tracker.cpp:176-181
    out->overlap_measure = measure;
    out->out_of_bounds = c[8];
    // tracker.cpp:183: histogram.size() = 1 + frames that took part
    out->status = (measure >= (float)c[4] || nf + 1 < 4) ? REVO_TRACKER_STATE_OK : REVO_TRACKER_STATE_NEW_KF;
}

int revo_track_quality_batch(revo_ctx *ctx, int n, revo_pyr *const *curs, int hist_level, const int *n_past, revo_pyr *const *past,
                             const float *past_world_poses16, const float *estimated_poses16, int n_frames_voting, revo_quality_result *out)
{
    if (!ctx || n < 0 || (n > 0 && (!curs || !n_past || !estimated_poses16 || !out))) return REVO_ERR_INVALID_ARG;
    if (n == 0) return REVO_OK;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<QualityArgs> args((size_t)n);
    int any = 0;
    for (int i = 0; i < n; ++i) {
        int rc = quality_prepare(ctx, curs[i], hist_level, n_past[i], past ? past + 3 * (size_t)i : nullptr,
                                 past_world_poses16 ? past_world_poses16 + 48 * (size_t)i : nullptr, estimated_poses16 + 16 * (size_t)i,
                                 n_frames_voting, args[i], &out[i]);
        if (rc) return rc;
        if (args[i].w != args[0].w || args[i].h != args[0].h) return REVO_ERR_INVALID_ARG;
        any |= out[i].n_frames > 0;
    }
    if (!any) return REVO_OK;
    const int w = args[0].w, h = args[0].h;
    const size_t words = ((size_t)w * h + 3) / 4, b_args = align_up(sizeof(QualityArgs) * (size_t)n, 256), b_cnt = align_up(64 * (size_t)n, 256);
    int rc = ensure_scratch(ctx, b_args + b_cnt + words * 4 * n);
    if (rc) return rc;
    QualityArgs *d_args = (QualityArgs *)ctx->scratch;
    int *d_counters = (int *)((uint8_t *)ctx->scratch + b_args);
    unsigned *d_mbits = (unsigned *)((uint8_t *)ctx->scratch + b_args + b_cnt);
    REVO_CUDA(ctx, cudaMemcpyAsync(d_args, args.data(), sizeof(QualityArgs) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_quality(ctx, d_args, n, w, h, curs[0]->cfg.depth_min, curs[0]->cfg.depth_max, d_mbits, d_counters);
    if (rc) return rc;
    std::vector<int> c(16 * (size_t)n);
    REVO_CUDA(ctx, cudaMemcpyAsync(c.data(), d_counters, 64 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; ++i)
        if (out[i].n_frames > 0) quality_finish(c.data() + 16 * (size_t)i, &out[i]);
    return REVO_OK;
}

int revo_track_quality(revo_ctx *ctx, const revo_pyr *cur, int hist_level, int n_past, revo_pyr *const *past,
                       const float *past_world_poses16, const float *estimated_pose16, int n_frames_voting, revo_quality_result *out)
{
    if (!ctx || !cur || !out) return REVO_ERR_INVALID_ARG;
    // one vote = a batch of one; the batch entry expects three past slots per vote
    revo_pyr *slots[3] = {nullptr, nullptr, nullptr};
    float poses[48] = {0};
    const int np = n_past < 0 ? n_past : (n_past < 3 ? n_past : 3);
    if (n_past > 0 && (!past || !past_world_poses16)) return REVO_ERR_INVALID_ARG;
    for (int f = 0; f < np; ++f) { slots[f] = past[f]; memcpy(poses + 16 * f, past_world_poses16 + 16 * f, 64); }
    revo_pyr *c = const_cast<revo_pyr *>(cur);
    const int npast = n_past;
    return revo_track_quality_batch(ctx, 1, &c, hist_level, &npast, slots, poses, estimated_pose16, n_frames_voting, out);
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU split of one pair (one process per GPU; mailboxes exchanged as CUDA IPC handles)
// ---------------------------------------------------------------------------------------------------
struct SplitBlob {
    cudaIpcMemHandle_t handle;   // 64 bytes
    int32_t rank, world;
    int32_t pid_lo, device;
    char pad[REVO_SPLIT_HANDLE_BYTES - 64 - 16];
};
static_assert(sizeof(SplitBlob) == REVO_SPLIT_HANDLE_BYTES, "blob size");

int revo_split_export(revo_ctx *ctx, int rank, int world, void *handle_out)
{
    if (!ctx || !handle_out || world < 1 || world > 16 || rank < 0 || rank >= world) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->split_local) {
        REVO_CUDA(ctx, cudaMalloc(&ctx->split_local, 16384));
        REVO_CUDA(ctx, cudaMemset(ctx->split_local, 0, 16384));
    }
    ctx->split_rank = rank; ctx->split_world = world;
    SplitBlob b;
    memset(&b, 0, sizeof(b));
    REVO_CUDA(ctx, cudaIpcGetMemHandle(&b.handle, ctx->split_local));
    b.rank = rank; b.world = world; b.device = ctx->device;
    memcpy(handle_out, &b, sizeof(b));
    return REVO_OK;
}

int revo_split_open(revo_ctx *ctx, const void *handles)
{
    if (!ctx || !handles || !ctx->split_local) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    const SplitBlob *b = (const SplitBlob *)handles;
    for (int r = 0; r < ctx->split_world; ++r) {
        if (b[r].rank != r || b[r].world != ctx->split_world) return REVO_ERR_COMM;
        if (r == ctx->split_rank) { ctx->split_peers[r] = ctx->split_local; continue; }
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, b[r].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cuda_fail(ctx, e, "cudaIpcOpenMemHandle"); return REVO_ERR_COMM; }
        ctx->split_peers[r] = p;
    }
    ctx->split_seq = 0;
    return REVO_OK;
}

int revo_track_split(revo_ctx *ctx, const revo_tracker_config *cfg, const revo_pyr *ref, const revo_pyr *cur, float *R9, float *t3,
                     revo_track_result *result)
{
    if (!ctx || !cfg || !ref || !cur || !R9 || !t3) return REVO_ERR_INVALID_ARG;
    if (ctx->split_world < 1 || !ctx->split_peers[ctx->split_rank]) return REVO_ERR_COMM;
    TrackParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.cfg = *cfg; prm.mode = 0;
    prm.split_rank = ctx->split_rank; prm.split_world = ctx->split_world;
    prm.split_seq0 = ctx->split_seq;
    ctx->split_seq += (1ull << 24);   // every launch owns a disjoint range of flag values
    for (int r = 0; r < 16; ++r) prm.split_peers[r] = ctx->split_peers[r];
    revo_track_result out;
    revo_pyr *r = const_cast<revo_pyr *>(ref), *c = const_cast<revo_pyr *>(cur);
    int rc = run_track(ctx, prm, 1, &r, &c, R9, t3, &out, nullptr, nullptr, 0, nullptr);
    if (rc) return rc;
    if (result) *result = out;
    if (out.rc) return out.rc;
    memcpy(R9, out.R, sizeof(float) * 9);
    memcpy(t3, out.t, sizeof(float) * 3);
    return REVO_OK;
}

}  // extern "C"
