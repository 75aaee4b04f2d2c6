// track_deep.cu -- EXPERIMENT, NOT PART OF THE LIBRARY: measured slower than track.cu on B200 (0.70-0.92 ms vs 0.62 ms on 128 VGA
// pairs, profiles/r1_track_engine_ab_shot.jsonl).  Kept for the record of what was tried.
//
// The cluster-per-pair tracking engine of track.cu with a DEEPER gather pipeline (engine 4, opt-in).
//
// Same algorithm, same work split, same reduction / exchange / LM step as k_track (see track.cu for the reference
// file:line map: system/tracker.cpp:265-353, system/optimizer.cpp:74-311, utils/LGSX.h:320-398); single GPU only (the
// multi-GPU edge split stays with track.cu).  The difference is the per-thread software pipeline of an evaluation:
// kDepth points in flight per thread instead of two.  k_track runs 16 warps per SM (128 registers per thread), and the
// gather of an evaluation is bound by memory latency x gathers in flight (one 256-bit gather per warp every ~76 cycles
// per SM against a latency of ~1000 cycles); registers are the only place to hold more gathers in flight, and a slot
// costs 8 registers of record + 3 (kSlim: the sub-pixel offsets are recomputed at retirement) or 5 of projection state.
// Points retire in list order, so every sum -- and with it every pose and every iteration count -- is bit-identical to
// k_track's: that is how the engine is validated (tests/test_gpu_track.py, engine fixture).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "internal.h"
#include "track_common.cuh"

namespace cg = cooperative_groups;

namespace revo {

// The 256-bit record gather of track_common.cuh (ldg_quad) with an optional cache hint (kHint):
//   0 plain, 1 L2::128B (a miss fills the whole 128-byte line = the records of the 3 pixels next to it in the row: the L2
//   line is allocated anyway, DRAM bandwidth is far from its limit in this kernel), 2 L2::64B, 3 L1::no_allocate,
//   4 L1::no_allocate + L2::128B.
#define REVO_LDG_QUAD(QUAL)                                                                                              \
    asm("ld.global.nc" QUAL ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                                   \
        : "=r"(r0.x), "=r"(r0.y), "=r"(r1.x), "=r"(r1.y), "=r"(r0.z), "=r"(r0.w), "=r"(r1.z), "=r"(r1.w)                 \
        : "l"(p))
template <int kHint>
__device__ __forceinline__ void ldg_quad_h(const uint4 *p, uint4 &r0, uint4 &r1)
{
    if (kHint == 1) REVO_LDG_QUAD(".L2::128B");
    else if (kHint == 2) REVO_LDG_QUAD(".L2::64B");
    else if (kHint == 3) REVO_LDG_QUAD(".L1::no_allocate");
    else if (kHint == 4) REVO_LDG_QUAD(".L1::no_allocate.L2::128B");
    else REVO_LDG_QUAD("");
}
#undef REVO_LDG_QUAD

// Dynamic shared memory: the thread-private cache of the level's 3-D points, float[3][pcap][kThreads] (x, y, z planes):
// thread t keeps the first `pcap` of ITS points of the current level there for all evaluations of the level, so an
// evaluation starts with shared-memory reads instead of an L2 round trip and re-reads no list bytes from L2 / HBM.
template <int kThreads, int kMinBlocks, int kDepth, bool kSlim, int kHint>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_track_deep(const PairDesc *__restrict__ pairs, int n_pairs, const TrackParams prm, revo_track_result *__restrict__ results,
        double *__restrict__ records, revo_trace_entry *__restrict__ trace, int *__restrict__ trace_counts,
        int *__restrict__ work_counter, int pcap)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int kWarps = kThreads / 32;

    extern __shared__ float s_pts[];
    float *const sx = s_pts + tid, *const sy = sx + (size_t)pcap * kThreads, *const sz = sy + (size_t)pcap * kThreads;

    __shared__ float warp_part[kWarps][32];
    __shared__ __align__(16) double cta_part[2][16][32];   // [parity][source rank]: partials pushed by the CTAs of the cluster
    __shared__ double rec[32];
    __shared__ __align__(8) uint64_t xbar[2];              // transaction barriers of the partial exchange (one per parity)
    __shared__ Ctrl ctrl;
    __shared__ LMState lm;

    const revo_opt_config &oc = prm.cfg.opt;
    const bool use_filter = oc.use_edge_filter != 0;
    const int n_members = C;
    const int member = crank;
    unsigned seq = 0;   // evaluation counter of this cluster (drives the double buffers)

    if (tid == 0) {
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (C > 1) cluster.sync(); else __syncthreads();

    // ---- reduction of a per-thread accumulator to `rec` (identical in every CTA of the cluster / every rank)
    auto reduce_record = [&](float (&acc)[32]) {
        const float mine = warp_transpose_reduce(acc, lane);
        warp_part[wid][lane] = mine;
        __syncthreads();
        const int par = seq & 1;
        // Every CTA pushes its 32-double partial into slot [its rank] of every CTA of the cluster (st.async over
        // distributed shared memory, 8 bytes per lane and destination) and waits on its OWN transaction barrier for
        // the C x 256 bytes of this evaluation: one-sided, no cluster barrier, no fence.  Two parities suffice: a CTA
        // can run at most one evaluation ahead of the slowest CTA of its cluster.
        if (wid == 0) {
            double s = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += (double)warp_part[w][lane];
            if (C == 1) {
                rec[lane] = s;
            } else {
                if (lane == 0) mbar_expect_tx(&xbar[par], (uint32_t)C * 256u);
                const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
                for (int r = 0; r < C; ++r) st_async_b64(&cta_part[par][crank][lane], (unsigned)r, bits, &xbar[par]);
                mbar_wait(&xbar[par], (seq >> 1) & 1u);
                double tot = 0;
                for (int r = 0; r < C; ++r) tot += cta_part[par][r][lane];   // rank order: deterministic
                rec[lane] = tot;
            }
        }
        seq++;
        __syncthreads();
    };

    long long prof_gather = 0, prof_reduce = 0, prof_serial = 0, prof_evals = 0;   // thread 0: cycles per phase
    int pair = cluster_id;
    while (pair < n_pairs) {
        const PairDesc &P = pairs[pair];
        const int min_lvl = prm.mode == 0 ? prm.cfg.pyr_min_lvl : prm.level;
        const int max_lvl = prm.mode == 0 ? prm.cfg.pyr_max_lvl : prm.level;
        int evals_lvl[REVO_MAX_LEVELS] = {0, 0, 0, 0, 0, 0};
        int used_identity = 0;
        int ntrace = 0;

        if (tid == 0) {
            for (int i = 0; i < 9; ++i) ctrl.R[i] = P.R[i];
            for (int i = 0; i < 3; ++i) ctrl.t[i] = P.t[i];
            ctrl.pair_skip = rotation_ok(P.R) ? 0 : 1;
            ctrl.level_done = 0;
        }
        __syncthreads();
        const bool skip = ctrl.pair_skip != 0;
        if (skip) {
            if (crank == 0 && tid == 0) {
                revo_track_result &o = results[pair];
                for (int i = 0; i < 9; ++i) o.R[i] = P.R[i];
                for (int i = 0; i < 3; ++i) o.t[i] = P.t[i];
                o.error = INFINITY;
                o.status = REVO_TRACKER_STATE_UNKNOWN;
                o.rc = REVO_ERR_NOT_ORTHOGONAL;
                o.res.good_pts_edges = o.res.bad_pts_edges = 0;
                o.res.sum_error_unweighted = o.res.sum_error_weighted = 0.f;
                for (int l = 0; l < REVO_MAX_LEVELS; ++l) { o.n_evals[l] = 0; o.n_pts[l] = 0; }
                o.used_identity_init = 0;
                if (trace_counts) trace_counts[pair] = 0;
            }
        } else {
            // ---- checkInitializationValues (tracker.cpp:265-283): cost at identity vs cost at (R,t), coarsest level
            if (prm.mode == 0 && prm.cfg.check_init_values) {
                const LevelIn L = P.lvl[min_lvl];
                const int n = *L.n_pts;
                const int lo = (int)((long long)n * member / n_members), hi = (int)((long long)n * (member + 1) / n_members);
                float acc[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                const float ed = oc.edge_distance_lvl[min_lvl];
                float R[9], t[3];
#pragma unroll
                for (int i = 0; i < 9; ++i) R[i] = ctrl.R[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) t[i] = ctrl.t[i];
                for (int i = lo + tid; i < hi; i += kThreads) {
                    const float4 p = __ldg(L.pts + i);
                    acc[0] += cost_point(p.x, p.y, p.z, L, P.ref_dt_min, ed, use_filter);
                    const float X = R[0] * p.x + R[3] * p.y + R[6] * p.z + t[0];
                    const float Y = R[1] * p.x + R[4] * p.y + R[7] * p.z + t[1];
                    const float Z = R[2] * p.x + R[5] * p.y + R[8] * p.z + t[2];
                    acc[1] += cost_point(X, Y, Z, L, P.ref_dt_min, ed, use_filter);
                }
                reduce_record(acc);
                if (tid == 0) {
                    if ((float)rec[0] < (float)rec[1]) {   // tracker.cpp:277
                        for (int i = 0; i < 9; ++i) ctrl.R[i] = (i % 4 == 0) ? 1.f : 0.f;
                        for (int i = 0; i < 3; ++i) ctrl.t[i] = 0.f;
                        ctrl.pair_skip = 2;   // marker: identity init used
                    }
                }
                __syncthreads();
                used_identity = ctrl.pair_skip == 2;
                __syncthreads();
            }

            if (tid == 0) {
                quat_from_R(ctrl.R, lm.q);
                for (int i = 0; i < 3; ++i) lm.t[i] = ctrl.t[i];
                lm.last_residual = INFINITY;
            }
            float last_good = 0.f, last_bad = 0.f, last_sw = 0.f, last_su = 0.f;

            for (int lvl = min_lvl; lvl >= max_lvl; --lvl) {
                const LevelIn Lin = P.lvl[lvl];
                const int n = *Lin.n_pts;
                // block-cyclic split of the list over the CTAs of the cluster (and the ranks of a GPU split): member m takes
                // the blocks m, m + M, m + 2M, ... of kThreads points -- balanced (the exchange waits for the slowest CTA)
                // and the cluster as a whole still sweeps the tile-major list front to back
                const int stride = n_members * kThreads;
                const int first_idx = member * kThreads + tid;
                const int n_iter = (n + stride - 1) / stride;          // uniform over the cluster
                const int n_cached = n_iter < pcap ? n_iter : pcap;
                const float4 *__restrict__ pts = Lin.pts;
                LevelConst L;
                L.fx = Lin.fx; L.fy = Lin.fy; L.cx = Lin.cx; L.cy = Lin.cy;
                L.umax = (float)(Lin.w - 2); L.vmax = (float)(Lin.h - 2); L.w = Lin.w; L.opt = Lin.opt;
                const float ed = oc.edge_distance_lvl[lvl];
                const float huber = oc.huber_edge;
                // this thread's points of the level -> its private columns of the shared-memory cache
                for (int k = 0; k < n_cached; ++k) {
                    const int i = first_idx + k * stride;
                    const float4 p = i < n ? __ldg(pts + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    sx[k * kThreads] = p.x; sy[k * kThreads] = p.y; sz[k * kThreads] = p.z;
                }
                auto fetch = [&](int k, bool &exists) -> float4 {
                    const int i = first_idx + k * stride;
                    exists = i < n;
                    if (k < n_cached) return make_float4(sx[k * kThreads], sy[k * kThreads], sz[k * kThreads], 1.f);
                    return __ldg(pts + (exists ? i : 0));
                };
                bool first = true;
                __syncthreads();
                while (true) {
                    float R[9], t[3];
#pragma unroll
                    for (int i = 0; i < 9; ++i) R[i] = ctrl.R[i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) t[i] = ctrl.t[i];
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                    const long long c_begin = prm.profile ? clock64() : 0;
                    // Software pipeline over kDepth register slots: slot j holds the projection state and the 256-bit
                    // record of one point; a slot is re-armed (next point projected, its gather issued) right after its
                    // point has been retired, so kDepth - 1 .. kDepth gathers are in flight per thread.  Points retire in
                    // list order, so the sums are bit-identical to the two-slot pipeline of track.cu.
                    if (n_iter > 0) {
                        float sa[kDepth], sb[kDepth], siz[kDepth], sdx[kDepth], sdy[kDepth];
                        bool sex[kDepth], sva[kDepth];
                        uint4 q0[kDepth], q1[kDepth];
                        auto arm = [&](int j, int k) {
                            bool ex;
                            const float4 p = fetch(k, ex);
                            const ProjB P = project_b(ex, p, L, R, t);
                            sa[j] = P.a; sb[j] = P.b; siz[j] = P.iz; sex[j] = P.exists; sva[j] = P.valid;
                            if (!kSlim) { sdx[j] = P.dx; sdy[j] = P.dy; }
                            ldg_quad_h<kHint>(P.bp, q0[j], q1[j]);
                        };
                        auto retire = [&](int j) {
                            ProjB P;
                            P.a = sa[j]; P.b = sb[j]; P.iz = siz[j]; P.exists = sex[j]; P.valid = sva[j]; P.bp = nullptr;
                            if (kSlim) {   // the sub-pixel offsets again from a, b (same expressions as project_b)
                                const float u = P.a * L.fx + L.cx;
                                const float v = P.b * L.fy + L.cy;
                                const int ix = P.valid ? (int)u : 0, iy = P.valid ? (int)v : 0;
                                P.dx = P.valid ? u - (float)ix : 0.f;
                                P.dy = P.valid ? v - (float)iy : 0.f;
                            } else {
                                P.dx = sdx[j]; P.dy = sdy[j];
                            }
                            finish_point_b(P, q0[j], q1[j], L, ed, use_filter, huber, acc);
                        };
#pragma unroll
                        for (int j = 0; j < kDepth; ++j)
                            if (j < n_iter) arm(j, j);
                        for (int k = 0; k < n_iter; k += kDepth) {
#pragma unroll
                            for (int j = 0; j < kDepth; ++j) {
                                if (k + j < n_iter) {
                                    retire(j);
                                    if (k + j + kDepth < n_iter) arm(j, k + j + kDepth);
                                }
                            }
                        }
                    }
                    const long long c_gather = prm.profile ? clock64() : 0;
                    reduce_record(acc);
                    const long long c_reduce = prm.profile ? clock64() : 0;
                    evals_lvl[lvl]++;
                    last_good = (float)rec[kRecGood]; last_bad = (float)rec[kRecBad];
                    last_sw = (float)rec[kRecSW]; last_su = (float)rec[kRecSU];

                    if (prm.mode == 2) {   // single evaluation: export the record
                        if (crank == 0 && tid < 32 && records) records[(size_t)pair * 32 + tid] = rec[tid];
                        break;
                    }

                    if (tid == 0) {
                        // Optimizer::trackFrames LM logic, optimizer.cpp:243-306 (track_common.cuh: lm_step)
                        revo_trace_entry te;
                        bool traced;
                        const bool done = lm_step(lm, rec, oc, lvl, first, ctrl.R, ctrl.t, &te, &traced);
                        if (traced) {
                            if (trace && crank == 0 && ntrace < prm.trace_cap) trace[(size_t)pair * prm.trace_cap + ntrace] = te;
                            ntrace++;
                        }
                        ctrl.level_done = done ? 1 : 0;
                    }
                    first = false;
                    __syncthreads();
                    if (prm.profile && tid == 0) {
                        const long long c_end = clock64();
                        prof_gather += c_gather - c_begin; prof_reduce += c_reduce - c_gather; prof_serial += c_end - c_reduce;
                        prof_evals++;
                    }
                    if (ctrl.level_done) break;
                }
                __syncthreads();
            }

            if (crank == 0 && tid == 0 && prm.mode != 2) {
                revo_track_result &o = results[pair];
                for (int i = 0; i < 9; ++i) o.R[i] = ctrl.R[i];
                for (int i = 0; i < 3; ++i) o.t[i] = ctrl.t[i];
                o.error = lm.last_residual;
                o.res.good_pts_edges = (int)last_good;
                o.res.bad_pts_edges = (int)last_bad;
                o.res.sum_error_weighted = last_sw;
                o.res.sum_error_unweighted = last_su;
                // tracker.cpp:351: good/bad < 4 -> NEW_KF (double division; bad == 0 -> inf -> OK)
                o.status = ((double)last_good / (double)last_bad < 4.0) ? REVO_TRACKER_STATE_NEW_KF : REVO_TRACKER_STATE_OK;
                o.rc = REVO_OK;
                for (int l = 0; l < REVO_MAX_LEVELS; ++l) {
                    o.n_evals[l] = evals_lvl[l];
                    o.n_pts[l] = (l >= max_lvl && l <= min_lvl) ? *P.lvl[l].n_pts : 0;
                }
                o.used_identity_init = used_identity;
                if (trace_counts) trace_counts[pair] = ntrace < prm.trace_cap ? ntrace : prm.trace_cap;
            }
        }
        // ---- next pair from the global work counter (cluster rank 0 fetches, everybody reads it over DSMEM)
        __syncthreads();
        if (crank == 0 && tid == 0) ctrl.next_pair = n_clusters + atomicAdd(work_counter, 1);
        if (C > 1) cluster.sync(); else __syncthreads();
        pair = *cluster.map_shared_rank(&ctrl.next_pair, 0);
        if (C > 1) cluster.sync(); else __syncthreads();
    }
    if (prm.profile && tid == 0 && crank == 0) {   // phase cycle counters behind the work counter (read back when REVO_TRACK_PROF is set)
        unsigned long long *prof = (unsigned long long *)(work_counter + 2);
        atomicAdd(prof + 0, (unsigned long long)prof_gather);
        atomicAdd(prof + 1, (unsigned long long)prof_reduce);
        atomicAdd(prof + 2, (unsigned long long)prof_serial);
        atomicAdd(prof + 3, (unsigned long long)prof_evals);
    }
    if (C > 1) cluster.sync();   // nobody may exit while a peer can still write into its shared memory
}

// ---- launcher -------------------------------------------------------------------
template <int kThreads, int kMinBlocks, int kDepth, bool kSlim, int kHint>
static int launch_deep_t(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, int ctas_per_pair,
                         revo_track_result *d_results, double *d_records, revo_trace_entry *d_trace, int *d_trace_counts,
                         int *d_work_counter)
{
    auto kern = k_track_deep<kThreads, kMinBlocks, kDepth, kSlim, kHint>;
    if (ctas_per_pair > 8) REVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const int env_pcap = getenv("REVO_TRACK_PCAP") ? atoi(getenv("REVO_TRACK_PCAP")) : -1;
    int pcap = env_pcap >= 0 ? env_pcap : (int)((112 * 1024 / kMinBlocks) / (12 * kThreads));
    if (pcap > 64) pcap = 64;
    const size_t dyn = (size_t)pcap * kThreads * 12;
    REVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas_per_pair;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = ctx->stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(ctas_per_pair);
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        (void)cudaGetLastError();
        max_clusters = ctx->prop.multiProcessorCount / ctas_per_pair;
        if (max_clusters < 1) max_clusters = 1;
    }
    const int env_maxc = getenv("REVO_TRACK_MAX_CLUSTERS") ? atoi(getenv("REVO_TRACK_MAX_CLUSTERS")) : 0;
    if (env_maxc > 0 && max_clusters > env_maxc) max_clusters = env_maxc;
    const int n_clusters = n_pairs < max_clusters ? n_pairs : max_clusters;
    cfg.gridDim = dim3(n_clusters * ctas_per_pair);
    REVO_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, d_pairs, n_pairs, prm, d_results, d_records, d_trace, d_trace_counts,
                                      d_work_counter, pcap));
    ctx->launches++;
    return REVO_OK;
}

// Shape from the context (revo_ctx_set_track_shape) / environment: REVO_DEEP_DEPTH (3..6; default 4),
// REVO_DEEP_MINBLOCKS (CTAs of 128 threads per SM: 4 -> 128 registers per thread, 3 -> 168; default 4), REVO_DEEP_SLIM
// (1: recompute the sub-pixel offsets at retirement; default 1), REVO_DEEP_HINT (cache hint of the gather, see ldg_quad_h;
// default 0; instantiated for the (4 blocks, depth 4) and (3 blocks, depth 5) shapes).
int launch_track_deep(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                      double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, int *d_work_counter)
{
    if (n_pairs <= 0) return REVO_OK;
    if (prm.split_world > 1) {
        ctx->last_error = "track engine 4 (deep pipeline) does not support the multi-GPU edge split";
        return REVO_ERR_INVALID_ARG;
    }
    const int C = ctx->track_ctas_per_pair > 0 ? ctx->track_ctas_per_pair : 8;
    const int depth = getenv("REVO_DEEP_DEPTH") ? atoi(getenv("REVO_DEEP_DEPTH")) : 4;
    const int minb = getenv("REVO_DEEP_MINBLOCKS") ? atoi(getenv("REVO_DEEP_MINBLOCKS")) : 4;
    const bool slim = getenv("REVO_DEEP_SLIM") ? atoi(getenv("REVO_DEEP_SLIM")) != 0 : true;
    const int hint = getenv("REVO_DEEP_HINT") ? atoi(getenv("REVO_DEEP_HINT")) : 0;
    const int T = ctx->track_threads > 0 ? ctx->track_threads : 128;
#define REVO_DEEP_ARGS ctx, d_pairs, n_pairs, prm, C, d_results, d_records, d_trace, d_trace_counts, d_work_counter
#define REVO_DEEP_CASE(TT, MB, DD)                                                                \
    if (T == TT && minb == MB && depth == DD && hint == 0)                                        \
        return slim ? launch_deep_t<TT, MB, DD, true, 0>(REVO_DEEP_ARGS) : launch_deep_t<TT, MB, DD, false, 0>(REVO_DEEP_ARGS);
#define REVO_DEEP_HINT(TT, MB, DD, HH)                                                            \
    if (T == TT && minb == MB && depth == DD && hint == HH && slim) return launch_deep_t<TT, MB, DD, true, HH>(REVO_DEEP_ARGS);
    REVO_DEEP_CASE(128, 4, 3)
    REVO_DEEP_CASE(128, 4, 4)
    REVO_DEEP_CASE(128, 3, 4)
    REVO_DEEP_CASE(128, 3, 5)
    REVO_DEEP_CASE(128, 3, 6)
    REVO_DEEP_CASE(256, 2, 4)
    REVO_DEEP_HINT(128, 4, 4, 1)
    REVO_DEEP_HINT(128, 4, 4, 2)
    REVO_DEEP_HINT(128, 4, 4, 3)
    REVO_DEEP_HINT(128, 4, 4, 4)
    REVO_DEEP_HINT(128, 3, 5, 1)
    REVO_DEEP_HINT(128, 3, 5, 2)
    REVO_DEEP_HINT(128, 3, 5, 3)
    REVO_DEEP_HINT(128, 3, 5, 4)
#undef REVO_DEEP_HINT
#undef REVO_DEEP_CASE
#undef REVO_DEEP_ARGS
    ctx->last_error = "track engine 4: unsupported (threads, min blocks, depth, hint) combination";
    return REVO_ERR_INVALID_ARG;
}

}  // namespace revo
