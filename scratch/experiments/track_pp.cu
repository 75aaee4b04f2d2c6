// track_pp.cu -- K9/K10, engine "ping-pong": the cluster engine of track.cu with the serial part of an evaluation taken
// off the gathering warps.
//
// Replaces (reference file:line, fabianschenk/REVO) -- the same functions as track.cu:
//   TrackerNew::trackFrames / checkInitializationValues / evalCostFunction   system/tracker.cpp:294-353, 265-283, 357-393
//   Optimizer::trackFrames (LM loop)                                          system/optimizer.cpp:235-311
//   Optimizer::calcErrorAndBuffers + getInterpolatedElement43                 system/optimizer.cpp:74-191, optimizer.h:173-185
//   Optimizer::calculateWarpUpdate + LGS6::update/finish                      system/optimizer.cpp:192-234, utils/LGSX.h:320-326,392-398
//   Eigen LDLT 6x6 solve, Sophus::SE3f exp / product                          system/optimizer.cpp:258-266
//
// In track.cu every evaluation ends with ~5 k cycles in which the CTAs of the pair only wait: cross-warp reduction,
// cluster exchange, the 6x6 solve + SE3 update on one thread, pose broadcast.  Here a cluster works on TWO pairs
// ("slots") at once and its CTAs are warp-specialised:
//   * 4 GATHER warps alternate between the slots: wait for the slot's job (pose, level) -> fused PASS A + PASS B over
//     this thread's points of that slot (branch-free, software-pipelined 256-bit gathers, shared-memory point cache per
//     slot) -> transposing shuffle reduction -> per-warp row in shared memory -> arrive on the slot's barrier -> other slot;
//   * 1 SOLVER warp does everything serial for both slots: sums the warp rows, exchanges the 32-double partial with the
//     other CTAs of the cluster (st.async + transaction barrier, as in track.cu), runs the LM state machine (lm_step)
//     redundantly per CTA, handles level switches / the next pair from the work counter, and publishes the slot's next job.
// While the solver warp works on slot A the gather warps are busy with slot B, so the serial latency is hidden whenever
// an evaluation's gather takes at least as long as the solve.  Hand-over between the warp roles is by mbarriers in
// shared memory (no block-wide barrier inside the loop).  No tensor cores (no dense contraction on this path).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "internal.h"
#include "track_common.cuh"

namespace cg = cooperative_groups;

namespace revo {

constexpr int kPPGatherWarps = 4;
constexpr int kPPGatherThreads = 32 * kPPGatherWarps;
constexpr int kPPThreads = kPPGatherThreads + 32;
constexpr int kJobEval = 0, kJobCost = 1, kJobExit = 2;

struct PPJob {                  // published by the solver warp, read by the gather warps
    float R[9], t[3];
    LevelIn L;                  // level of the pair this evaluation runs on
    const float *ref_dt_min;    // for the init-cost job
    int n;                      // points of the level
    int kind, lvl, new_level;
};

struct PPSlot {                 // solver-warp state of the pair a slot is working on
    LMState lm;
    float R[9], t[3];           // pose under evaluation / accepted pose between levels
    float last_good, last_bad, last_sw, last_su;
    int evals_lvl[REVO_MAX_LEVELS];
    int pair, lvl, first, phase, ntrace, used_identity, done;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int kMinBlocks>
__global__ void __launch_bounds__(kPPThreads, kMinBlocks)
k_track_pp(const PairDesc *__restrict__ pairs, int n_pairs, const TrackParams prm, revo_track_result *__restrict__ results,
           double *__restrict__ records, revo_trace_entry *__restrict__ trace, int *__restrict__ trace_counts,
           int *__restrict__ work_counter, int pcap)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool is_solver = wid == kPPGatherWarps;

    extern __shared__ float s_pts[];                           // [slot][3][pcap][kPPGatherThreads]
    __shared__ float warp_part[2][kPPGatherWarps][32];
    __shared__ __align__(16) double cta_part[2][2][16][32];    // [slot][parity][source rank]
    __shared__ double rec_s[2][32];
    __shared__ PPJob job[2];
    __shared__ PPSlot slot[2];
    __shared__ __align__(8) uint64_t job_ready[2], part_ready[2], xbar[2][2], np_bar[2];
    __shared__ __align__(8) unsigned long long next_pair_s[2];

    const revo_opt_config &oc = prm.cfg.opt;
    const bool use_filter = oc.use_edge_filter != 0;
    const int min_lvl = prm.mode == 0 ? prm.cfg.pyr_min_lvl : prm.level;
    const int max_lvl = prm.mode == 0 ? prm.cfg.pyr_max_lvl : prm.level;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&job_ready[s], 1);
            mbar_init(&part_ready[s], kPPGatherWarps);
            mbar_init(&xbar[s][0], 1);
            mbar_init(&xbar[s][1], 1);
            mbar_init(&np_bar[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (C > 1) cluster.sync(); else __syncthreads();

    if (!is_solver) {
        // =========================================== gather warps ===========================================
        const int gtid = tid;
        unsigned ph_job[2] = {0u, 0u};
        bool slot_done[2] = {false, false};
        for (int s = 0;; s ^= 1) {
            if (slot_done[s]) {
                if (slot_done[s ^ 1]) break;
                continue;
            }
            mbar_wait(&job_ready[s], ph_job[s]);
            ph_job[s] ^= 1u;
            const PPJob &J = job[s];
            const int kind = J.kind;
            if (kind == kJobExit) {
                slot_done[s] = true;
                if (slot_done[s ^ 1]) break;
                continue;
            }
            float R[9], t[3];
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = J.R[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) t[i] = J.t[i];
            const LevelIn Lin = J.L;
            const int n = J.n, lvl = J.lvl;
            float acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.f;
            const float ed = oc.edge_distance_lvl[lvl];
            if (kind == kJobCost) {
                // checkInitializationValues (tracker.cpp:265-283): cost at identity vs cost at (R,t), coarsest level
                const float *dtm = J.ref_dt_min;
                const int lo = (int)((long long)n * crank / C), hi = (int)((long long)n * (crank + 1) / C);
                for (int i = lo + gtid; i < hi; i += kPPGatherThreads) {
                    const float4 p = __ldg(Lin.pts + i);
                    acc[0] += cost_point(p.x, p.y, p.z, Lin, dtm, ed, use_filter);
                    const float X = R[0] * p.x + R[3] * p.y + R[6] * p.z + t[0];
                    const float Y = R[1] * p.x + R[4] * p.y + R[7] * p.z + t[1];
                    const float Z = R[2] * p.x + R[5] * p.y + R[8] * p.z + t[2];
                    acc[1] += cost_point(X, Y, Z, Lin, dtm, ed, use_filter);
                }
            } else {
                float *const sx = s_pts + (size_t)s * 3 * pcap * kPPGatherThreads + gtid;
                float *const sy = sx + (size_t)pcap * kPPGatherThreads, *const sz = sy + (size_t)pcap * kPPGatherThreads;
                const int stride = C * kPPGatherThreads;
                const int first_idx = crank * kPPGatherThreads + gtid;
                const int n_iter = (n + stride - 1) / stride;          // uniform over the cluster
                const int n_cached = n_iter < pcap ? n_iter : pcap;
                const float4 *__restrict__ pts = Lin.pts;
                if (J.new_level) {   // this thread's points of the level -> its private columns of the slot's cache
                    for (int k = 0; k < n_cached; ++k) {
                        const int i = first_idx + k * stride;
                        const float4 p = i < n ? __ldg(pts + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                        sx[k * kPPGatherThreads] = p.x; sy[k * kPPGatherThreads] = p.y; sz[k * kPPGatherThreads] = p.z;
                    }
                }
                LevelConst L;
                L.fx = Lin.fx; L.fy = Lin.fy; L.cx = Lin.cx; L.cy = Lin.cy;
                L.umax = (float)(Lin.w - 2); L.vmax = (float)(Lin.h - 2); L.w = Lin.w; L.opt = Lin.opt;
                const float huber = oc.huber_edge;
                auto fetch = [&](int k, bool &exists) -> float4 {
                    const int i = first_idx + k * stride;
                    exists = i < n;
                    if (k < n_cached) return make_float4(sx[k * kPPGatherThreads], sy[k * kPPGatherThreads], sz[k * kPPGatherThreads], 1.f);
                    return __ldg(pts + (exists ? i : 0));
                };
                if (n_iter > 0) {
                    bool eA, eB;
                    float4 p = fetch(0, eA);
                    ProjB A = project_b(eA, p, L, R, t), B;
                    uint4 a0, a1, b0, b1;
                    ldg_quad(A.bp, a0, a1);
                    int k = 0;
                    while (true) {
                        const bool hasB = k + 1 < n_iter;
                        if (hasB) {
                            p = fetch(k + 1, eB);
                            B = project_b(eB, p, L, R, t);
                            ldg_quad(B.bp, b0, b1);
                        }
                        finish_point_b(A, a0, a1, L, ed, use_filter, huber, acc);
                        if (!hasB) break;
                        const bool hasA = k + 2 < n_iter;
                        if (hasA) {
                            p = fetch(k + 2, eA);
                            A = project_b(eA, p, L, R, t);
                            ldg_quad(A.bp, a0, a1);
                        }
                        finish_point_b(B, b0, b1, L, ed, use_filter, huber, acc);
                        if (!hasA) break;
                        k += 2;
                    }
                }
            }
            const float mine = warp_transpose_reduce(acc, lane);
            warp_part[s][wid][lane] = mine;
            __syncwarp();
            if (lane == 0) mbar_arrive(&part_ready[s]);
        }
    } else {
        // =========================================== solver warp ============================================
        unsigned ph_part[2] = {0u, 0u}, ph_np[2] = {0u, 0u};
        unsigned seq[2] = {0u, 0u};                    // evaluations of the slot so far (parity of the exchange buffers)
        bool slot_done[2] = {false, false};

        // lane 0: publish the job of slot s from its state (the gather warps are waiting on job_ready[s])
        auto publish = [&](int s, int kind, int new_level) {
            PPSlot &S = slot[s];
            PPJob &J = job[s];
            J.kind = kind;
            if (kind != kJobExit) {
                const PairDesc &P = pairs[S.pair];
                for (int i = 0; i < 9; ++i) J.R[i] = S.R[i];
                for (int i = 0; i < 3; ++i) J.t[i] = S.t[i];
                J.L = P.lvl[S.lvl];
                J.ref_dt_min = P.ref_dt_min;
                J.n = *J.L.n_pts;
                J.lvl = S.lvl;
                J.new_level = new_level;
            }
            mbar_arrive(&job_ready[s]);
        };
        // lane 0 (every CTA identically): start level `lvl` of the slot's pair from the accepted pose in S.R / S.t
        auto begin_level = [&](int s, int lvl) {
            PPSlot &S = slot[s];
            S.lvl = lvl; S.first = 1; S.phase = 1;
            publish(s, kJobEval, 1);
        };
        auto write_result = [&](int s, bool skipped) {   // lane 0 of cluster rank 0
            PPSlot &S = slot[s];
            const PairDesc &P = pairs[S.pair];
            if (prm.mode == 2 || crank != 0) return;
            revo_track_result &o = results[S.pair];
            if (skipped) {
                for (int i = 0; i < 9; ++i) o.R[i] = P.R[i];
                for (int i = 0; i < 3; ++i) o.t[i] = P.t[i];
                o.error = INFINITY;
                o.status = REVO_TRACKER_STATE_UNKNOWN;
                o.rc = REVO_ERR_NOT_ORTHOGONAL;
                o.res.good_pts_edges = o.res.bad_pts_edges = 0;
                o.res.sum_error_unweighted = o.res.sum_error_weighted = 0.f;
                for (int l = 0; l < REVO_MAX_LEVELS; ++l) { o.n_evals[l] = 0; o.n_pts[l] = 0; }
                o.used_identity_init = 0;
                if (trace_counts) trace_counts[S.pair] = 0;
                return;
            }
            for (int i = 0; i < 9; ++i) o.R[i] = S.R[i];
            for (int i = 0; i < 3; ++i) o.t[i] = S.t[i];
            o.error = S.lm.last_residual;
            o.res.good_pts_edges = (int)S.last_good;
            o.res.bad_pts_edges = (int)S.last_bad;
            o.res.sum_error_weighted = S.last_sw;
            o.res.sum_error_unweighted = S.last_su;
            // tracker.cpp:351: good/bad < 4 -> NEW_KF (double division; bad == 0 -> inf -> OK)
            o.status = ((double)S.last_good / (double)S.last_bad < 4.0) ? REVO_TRACKER_STATE_NEW_KF : REVO_TRACKER_STATE_OK;
            o.rc = REVO_OK;
            for (int l = 0; l < REVO_MAX_LEVELS; ++l) {
                o.n_evals[l] = S.evals_lvl[l];
                o.n_pts[l] = (l >= max_lvl && l <= min_lvl) ? *P.lvl[l].n_pts : 0;
            }
            o.used_identity_init = S.used_identity;
            if (trace_counts) trace_counts[S.pair] = S.ntrace < prm.trace_cap ? S.ntrace : prm.trace_cap;
        };
        // Warp-collective: give slot s the pair `pair` (or the next ones from the work counter while pairs are rejected) and
        // publish its first job; marks the slot done and publishes the exit job when the pairs are exhausted.
        auto start_slot = [&](int s, int pair) {
            while (true) {
                int st = 0;     // 0 = exhausted, 1 = started, 2 = rejected (fetch another)
                if (lane == 0) {
                    PPSlot &S = slot[s];
                    S.pair = pair;
                    if (pair >= n_pairs) {
                        publish(s, kJobExit, 0);
                        st = 0;
                    } else {
                        const PairDesc &P = pairs[pair];
                        for (int i = 0; i < 9; ++i) S.R[i] = P.R[i];
                        for (int i = 0; i < 3; ++i) S.t[i] = P.t[i];
                        S.last_good = S.last_bad = S.last_sw = S.last_su = 0.f;
                        for (int l = 0; l < REVO_MAX_LEVELS; ++l) S.evals_lvl[l] = 0;
                        S.ntrace = 0; S.used_identity = 0; S.first = 1;
                        S.lm.last_residual = INFINITY;
                        if (!rotation_ok(P.R)) {
                            write_result(s, true);
                            st = 2;
                        } else {
                            st = 1;
                            if (prm.mode == 0 && prm.cfg.check_init_values) {
                                S.phase = 0; S.lvl = min_lvl;
                                publish(s, kJobCost, 0);
                            } else {
                                quat_from_R(S.R, S.lm.q);
                                for (int i = 0; i < 3; ++i) S.lm.t[i] = S.t[i];
                                begin_level(s, min_lvl);
                            }
                        }
                    }
                }
                st = __shfl_sync(kFull, st, 0);
                if (st == 1) return;
                if (st == 0) { slot_done[s] = true; return; }
                // rejected pair: the next one from the work counter (cluster rank 0 fetches, pushes it to every CTA)
                if (lane == 0) {
                    mbar_expect_tx(&np_bar[s], 8u);
                    if (crank == 0) {
                        const unsigned long long np = (unsigned long long)(2 * n_clusters + atomicAdd(work_counter, 1));
                        for (int r = 0; r < C; ++r) st_async_b64(&next_pair_s[s], (unsigned)r, np, &np_bar[s]);
                    }
                }
                mbar_wait(&np_bar[s], ph_np[s]);
                ph_np[s] ^= 1u;
                pair = (int)next_pair_s[s];
            }
        };
        auto next_pair_for = [&](int s) -> int {      // warp-collective
            if (lane == 0) {
                mbar_expect_tx(&np_bar[s], 8u);
                if (crank == 0) {
                    const unsigned long long np = (unsigned long long)(2 * n_clusters + atomicAdd(work_counter, 1));
                    for (int r = 0; r < C; ++r) st_async_b64(&next_pair_s[s], (unsigned)r, np, &np_bar[s]);
                }
            }
            mbar_wait(&np_bar[s], ph_np[s]);
            ph_np[s] ^= 1u;
            return (int)next_pair_s[s];
        };

        start_slot(0, 2 * cluster_id);
        start_slot(1, 2 * cluster_id + 1);

        for (int s = 0;; s ^= 1) {
            if (slot_done[s]) {
                if (slot_done[s ^ 1]) break;
                continue;
            }
            // ---- the gather warps have delivered the evaluation of slot s
            mbar_wait(&part_ready[s], ph_part[s]);
            ph_part[s] ^= 1u;
            const int par = seq[s] & 1;
            double v = 0;
#pragma unroll
            for (int w = 0; w < kPPGatherWarps; ++w) v += (double)warp_part[s][w][lane];
            if (C > 1) {
                if (lane == 0) mbar_expect_tx(&xbar[s][par], (uint32_t)C * 256u);
                const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
                for (int r = 0; r < C; ++r) st_async_b64(&cta_part[s][par][crank][lane], (unsigned)r, bits, &xbar[s][par]);
                mbar_wait(&xbar[s][par], (seq[s] >> 1) & 1u);
                v = 0;
                for (int r = 0; r < C; ++r) v += cta_part[s][par][r][lane];   // rank order: deterministic
            }
            seq[s]++;
            rec_s[s][lane] = v;
            __syncwarp();
            // ---- state machine of the slot's pair (lane 0; identical in every CTA of the cluster)
            int act = 0;     // 0 = job published, 1 = pair finished
            if (lane == 0) {
                PPSlot &S = slot[s];
                const double *rec = rec_s[s];
                const int lvl = S.lvl;
                if (S.phase == 0) {
                    if ((float)rec[0] < (float)rec[1]) {   // tracker.cpp:277
                        for (int k = 0; k < 9; ++k) S.R[k] = (k % 4 == 0) ? 1.f : 0.f;
                        for (int k = 0; k < 3; ++k) S.t[k] = 0.f;
                        S.used_identity = 1;
                    }
                    quat_from_R(S.R, S.lm.q);
                    for (int k = 0; k < 3; ++k) S.lm.t[k] = S.t[k];
                    begin_level(s, min_lvl);
                } else {
                    S.evals_lvl[lvl]++;
                    S.last_good = (float)rec[kRecGood]; S.last_bad = (float)rec[kRecBad];
                    S.last_sw = (float)rec[kRecSW]; S.last_su = (float)rec[kRecSU];
                    if (prm.mode == 2) {   // single evaluation: export the record
                        if (crank == 0 && records)
                            for (int k = 0; k < 32; ++k) records[(size_t)S.pair * 32 + k] = rec[k];
                        act = 1;
                    } else {
                        revo_trace_entry te;
                        bool traced;
                        const bool done = lm_step(S.lm, rec, oc, lvl, S.first != 0, S.R, S.t, &te, &traced);
                        S.first = 0;
                        if (traced) {
                            if (trace && crank == 0 && S.ntrace < prm.trace_cap) trace[(size_t)S.pair * prm.trace_cap + S.ntrace] = te;
                            S.ntrace++;
                        }
                        if (!done) publish(s, kJobEval, 0);
                        else if (lvl > max_lvl) begin_level(s, lvl - 1);
                        else { write_result(s, false); act = 1; }
                    }
                }
            }
            act = __shfl_sync(kFull, act, 0);
            if (act == 1) start_slot(s, next_pair_for(s));
        }
    }
    __syncthreads();
    if (C > 1) cluster.sync();   // nobody may exit while a peer can still write into its shared memory
}

// ---- launcher -------------------------------------------------------------------------------------------------
int launch_track_pp(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                    double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, int *d_work_counter)
{
    if (n_pairs <= 0) return REVO_OK;
    constexpr int kMinBlocks = 3;
    auto kern = k_track_pp<kMinBlocks>;
    int C = ctx->track_ctas_per_pair > 0 ? ctx->track_ctas_per_pair : 8;
    if (C > 8) REVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const int env_pcap = getenv("REVO_TRACK_PCAP") ? atoi(getenv("REVO_TRACK_PCAP")) : -1;
    int pcap = env_pcap >= 0 ? env_pcap : 12;
    if (pcap > 48) pcap = 48;
    const size_t dyn = (size_t)2 * 3 * pcap * kPPGatherThreads * sizeof(float);
    REVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(kPPThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = ctx->stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.gridDim = dim3(C);
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        (void)cudaGetLastError();
        max_clusters = ctx->prop.multiProcessorCount / C;
        if (max_clusters < 1) max_clusters = 1;
    }
    const int env_maxc = getenv("REVO_TRACK_MAX_CLUSTERS") ? atoi(getenv("REVO_TRACK_MAX_CLUSTERS")) : 0;
    if (env_maxc > 0 && max_clusters > env_maxc) max_clusters = env_maxc;
    const int want = (n_pairs + 1) / 2;     // two pairs per cluster
    const int n_clusters = want < max_clusters ? want : max_clusters;
    cfg.gridDim = dim3(n_clusters * C);
    REVO_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, d_pairs, n_pairs, prm, d_results, d_records, d_trace, d_trace_counts,
                                      d_work_counter, pcap));
    ctx->launches++;
    return REVO_OK;
}

}  // namespace revo
