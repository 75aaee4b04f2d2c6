// track.cu -- K9/K10: the coarse-to-fine Gauss-Newton / Levenberg-Marquardt edge alignment as ONE
// persistent kernel.
//
// Replaces (reference file:line, fabianschenk/REVO):
//   TrackerNew::trackFrames / checkInitializationValues / evalCostFunction   system/tracker.cpp:294-353, 265-283, 357-393
//   Optimizer::trackFrames (LM loop)                                          system/optimizer.cpp:235-311
//   Optimizer::calcErrorAndBuffers (PASS A) + getInterpolatedElement43       system/optimizer.cpp:74-191, optimizer.h:173-185
//   Optimizer::calculateWarpUpdate (PASS B) + LGS6::update/finish             system/optimizer.cpp:192-234, utils/LGSX.h:320-326,392-398
//   Eigen LDLT 6x6 solve, Sophus::SE3f exp / product                          system/optimizer.cpp:258-266
//
// Design (B200): a frame pair is owned by one thread-block CLUSTER (1..16 CTAs); clusters pull pairs from a global work
// counter (persistent kernel).  PASS A and PASS B are fused: every evaluation at a pose warps each 3-D edge point, fetches
// the 2x2 texels around its projection from the tiled lookup structure (four 64-bit gathers), forms the residual, Huber weight and 1x6
// Jacobian and accumulates the 21+6 normal-equation terms + statistics in registers -- the 7 SoA buffers of the reference
// never exist.  A thread keeps its points of the level in a private column of a shared-memory cache and visits only points
// that exist; two points are in flight per thread (software pipeline) to cover the point -> record latency.  The 32-value
// record is reduced with a transposing warp-shuffle tree, across warps through shared memory, across the CTAs of the cluster
// by a one-sided st.async exchange over distributed shared memory (transaction barrier, no cluster barrier), and -- when
// one pair is split over several GPUs -- across GPUs through peer-mapped mailboxes over NVLink inside the same kernel.
// Every CTA then runs the identical LM step (6x6 LDL^T, SE3 exp / product, accept test) redundantly: bitwise-equal inputs,
// so no broadcast is needed, and all levels and all LM iterations of a pair run without a host round trip.
// Two thirds of the LM tries of this optimizer are REJECTED (every iteration restarts at lambda = 0), and the pose that
// follows a rejection depends only on what is known when the try starts (accepted normal equations, next lambda): lane 0
// of warp 1 computes it while warp 0 waits for the record exchange, so after a rejection the next evaluation starts
// without the solve on the critical path (lm_step in track_common.cuh picks it up).  The LM step itself runs in float32
// like the reference's Eigen / Sophus types (a dependent chain: the float LDL^T + solve is ~0.9 k cycles, and a dependent FP64
// FMA costs twice the latency of an FP32 one on this chip, profiles/r2_lm_probe.txt).
// No tensor cores: there is no dense contraction here.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "internal.h"
#include "track_common.cuh"

namespace cg = cooperative_groups;

namespace revo {

// ---- mailbox for the multi-GPU split ----------------------------------------
struct Mailbox {
    double data[2][16][32];
    unsigned long long flag[2][16];
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// x, y, z of one cached point: three 32-bit shared loads / stores at immediate offsets from one running address
template <int kThreads>
__device__ __forceinline__ void lds3(uint32_t addr, float &x, float &y, float &z)
{
    asm volatile("ld.shared.f32 %0, [%3];\n\tld.shared.f32 %1, [%3+%4];\n\tld.shared.f32 %2, [%3+%5];"
                 : "=f"(x), "=f"(y), "=f"(z)
                 : "r"(addr), "n"(kThreads * 4), "n"(kThreads * 8));
}
template <int kThreads>
__device__ __forceinline__ void sts3(uint32_t addr, float x, float y, float z)
{
    asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+%4], %2;\n\tst.shared.f32 [%0+%5], %3;"
                 :: "r"(addr), "f"(x), "f"(y), "f"(z), "n"(kThreads * 4), "n"(kThreads * 8) : "memory");
}

__device__ __forceinline__ int opt_tiles_per_row_dev(int w) { return (w + 3) >> 2; }

// keeps a per-level constant in a register (the compiler otherwise re-derives it from the constant bank for every point)
__device__ __forceinline__ float pin(float x)
{
    asm volatile("" : "+f"(x));
    return x;
}

// ---- the kernel ---------------------------------------------------------------
// Dynamic shared memory: the thread-private cache of the level's 3-D points, float[pcap][3][kThreads]: thread t keeps the
// first `pcap` of ITS points of the current level there for all evaluations of the level, so an evaluation starts with
// shared-memory reads instead of an L2 round trip and re-reads no list bytes from L2 / HBM.
template <int kThreads, int kMinBlocks, int kHint>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_track(const PairDesc *__restrict__ pairs, int n_pairs, const TrackParams prm, revo_track_result *__restrict__ results,
        double *__restrict__ records, revo_trace_entry *__restrict__ trace, int *__restrict__ trace_counts,
        int *__restrict__ work_counter, int pcap)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int kWarps = kThreads / 32;

    extern __shared__ float s_pts[];   // [pcap][3][kThreads]: x, y, z of cached point k of thread t at (k * 3 + c) * kThreads + t
    const uint32_t s_base = smem_u32(s_pts) + 4u * (uint32_t)tid;
    constexpr uint32_t kPtStride = 3u * kThreads * 4u;

    __shared__ float warp_part[kWarps][32];
    __shared__ __align__(16) double cta_part[2][16][32];   // [parity][source rank]: partials pushed by the CTAs of the cluster
    __shared__ double total[2][32];                        // split mode: CTA 0 publishes the cross-GPU total here
    __shared__ double rec[2][32];                          // [lm.acc]: record of the accepted pose, [lm.acc ^ 1]: the latest one
    __shared__ lmreal recs[2][32];                         // the same records divided by their good count, as the LM step reads them
    __shared__ __align__(8) uint64_t xbar[2];              // transaction barriers of the partial exchange (one per parity)
    __shared__ Ctrl ctrl;
    __shared__ LMState lm;
    __shared__ Trial trial[3];                             // poses: being evaluated / speculative successor / fresh proposal
    __shared__ SpecIn specin[2];                           // input of the speculating thread, by evaluation parity

    const revo_opt_config &oc = prm.cfg.opt;
    const bool use_filter = oc.use_edge_filter != 0;
    const int world = prm.split_world > 1 ? prm.split_world : 1;
    const int n_members = world * C;
    const int member = (world > 1 ? prm.split_rank : 0) * C + crank;
    const bool speculate = prm.speculate != 0 && kWarps >= 2;
    const unsigned long long policy = kHint == 2 ? l2_policy_evict_last() : 0ull;
    unsigned seq = 0;   // evaluation counter of this cluster (drives the double buffers)

    if (tid == 0) {
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (C > 1) cluster.sync(); else __syncthreads();

    // The successor of the try being evaluated in case it is rejected (lane 0 of warp 1, while warp 0 exchanges the record).
    auto speculate_successor = [&](int cur) {
        const SpecIn s = specin[seq & 1];
        if (s.active) lm_propose(recs[s.acc], lm.q[s.pacc], lm.t[s.pacc], s.lambda, trial[cur == 2 ? 0 : cur + 1]);
    };

    // warp 0: lane l publishes value l of the record, in double and scaled for the LM step (n = value kRecGood)
    auto publish = [&](double tot, int wb) {
        rec[wb][lane] = tot;
        const int nlo = __shfl_sync(kFull, __double2loint(tot), kRecGood), nhi = __shfl_sync(kFull, __double2hiint(tot), kRecGood);
        recs[wb][lane] = lm_scaled(tot, __hiloint2double(nhi, nlo));
    };

    // ---- reduction of a per-thread accumulator to rec[wb] (identical in every CTA of the cluster / every rank).  On return
    // the record is visible to WARP 0 only (it goes straight on to the LM step); the other warps meet it at the block-wide
    // barrier behind that step.  spec_now: lane 0 of warp 1 computes the reject-successor of trial[cur] meanwhile.
    auto reduce_record = [&](float (&acc)[32], bool spec_now, int cur, int wb) {
        const float mine = warp_transpose_reduce(acc, lane);
        warp_part[wid][lane] = mine;
        __syncthreads();
        const int par = seq & 1;
        if (world == 1) {
            // Every CTA pushes its 32-double partial into slot [its rank] of every CTA of the cluster (st.async over
            // distributed shared memory, 8 bytes per lane and destination) and waits on its OWN transaction barrier for
            // the C x 256 bytes of this evaluation: one-sided, no cluster barrier, no fence.  Two parities suffice: a CTA
            // can run at most one evaluation ahead of the slowest CTA of its cluster.
            if (wid == 0) {
                double s = 0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) s += (double)warp_part[w][lane];
                double tot = s;
                if (C > 1) {
                    if (lane == 0) mbar_expect_tx(&xbar[par], (uint32_t)C * 256u);
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
                    for (int r = 0; r < C; ++r) st_async_b64(&cta_part[par][crank][lane], (unsigned)r, bits, &xbar[par]);
                    mbar_wait(&xbar[par], (seq >> 1) & 1u);
                    tot = 0;
                    for (int r = 0; r < C; ++r) tot += cta_part[par][r][lane];   // rank order: deterministic
                }
                publish(tot, wb);
            } else if (spec_now && tid == 32) {
                speculate_successor(cur);
            }
        } else {
            if (tid < 32) {
                double s = 0;
#pragma unroll
                for (int w = 0; w < kWarps; ++w) s += (double)warp_part[w][tid];
                cta_part[par][0][tid] = s;
            }
            cluster.sync();
            // cross-GPU exchange: CTA 0 of each rank pushes the rank partial into every rank's mailbox,
            // then waits for all `world` partials of this evaluation and sums them in rank order.
            const unsigned long long fl = prm.split_seq0 + (unsigned long long)seq + 1ull;
            if (crank == 0 && tid < 32) {
                double s = 0;
                for (int r = 0; r < C; ++r) s += *cluster.map_shared_rank(&cta_part[par][0][tid], r);
                for (int g = 0; g < world; ++g) {
                    Mailbox *mb = (Mailbox *)prm.split_peers[g];
                    mb->data[par][prm.split_rank][tid] = s;
                }
                __threadfence_system();
                __syncwarp();
                if (tid < world) st_release_sys(&((Mailbox *)prm.split_peers[tid])->flag[par][prm.split_rank], fl);
                Mailbox *mine_mb = (Mailbox *)prm.split_peers[prm.split_rank];
                bool timed_out = false;
                if (tid < world) {
                    // watchdog (~5 s): a peer that never launches (or died) must not hang this GPU.  The failure is
                    // published in slot 31 of the record (otherwise always 0): every CTA of this rank sees it, abandons
                    // the pair and the result carries REVO_ERR_COMM.
                    const long long t0 = clock64();
                    while (ld_acquire_sys(&mine_mb->flag[par][tid]) < fl) {
                        if (clock64() - t0 > 10000000000ll) { timed_out = true; break; }
                    }
                }
                const bool any_timeout = __any_sync(kFull, timed_out);
                double tot = 0;
                for (int g = 0; g < world; ++g) tot += ((volatile double *)mine_mb->data[par][g])[tid];
                if (tid == 31) tot = any_timeout ? 1.0 : 0.0;
                total[par][tid] = tot;
            } else if (spec_now && tid == 32) {
                speculate_successor(cur);      // overlaps the NVLink round trip
            }
            cluster.sync();
            if (tid < 32) publish(*cluster.map_shared_rank(&total[par][tid], 0), wb);
        }
        seq++;
        if (wid == 0) __syncwarp();
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
        bool comm_failed = false;   // split mode: an exchange timed out (uniform over the CTAs of this rank)

        if (tid == 0) {
            for (int i = 0; i < 9; ++i) trial[0].R[i] = P.R[i];
            for (int i = 0; i < 3; ++i) trial[0].t[i] = P.t[i];
            ctrl.cur = 0;
            ctrl.pair_skip = rotation_ok(P.R) ? 0 : 1;
            ctrl.level_done = 0;
            lm.acc = 0; lm.pacc = 0;
            specin[0].active = 0; specin[1].active = 0;
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
                for (int i = 0; i < 9; ++i) R[i] = trial[0].R[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) t[i] = trial[0].t[i];
                for (int i = lo + tid; i < hi; i += kThreads) {
                    const float4 p = __ldg(L.pts + i);
                    acc[0] += cost_point(p.x, p.y, p.z, L, P.ref_dt_min, ed, use_filter);
                    const float X = R[0] * p.x + R[3] * p.y + R[6] * p.z + t[0];
                    const float Y = R[1] * p.x + R[4] * p.y + R[7] * p.z + t[1];
                    const float Z = R[2] * p.x + R[5] * p.y + R[8] * p.z + t[2];
                    acc[1] += cost_point(X, Y, Z, L, P.ref_dt_min, ed, use_filter);
                }
                reduce_record(acc, false, 0, 1);
                __syncthreads();
                const double *r = rec[1];
                if (world > 1 && r[31] != 0.0) comm_failed = true;
                const bool take_identity = (float)r[0] < (float)r[1];   // tracker.cpp:277 (uniform: every thread reads the same record)
                __syncthreads();
                if (tid == 0 && take_identity) {
                    for (int i = 0; i < 9; ++i) trial[0].R[i] = (i % 4 == 0) ? 1.f : 0.f;
                    for (int i = 0; i < 3; ++i) trial[0].t[i] = 0.f;
                }
                used_identity = take_identity ? 1 : 0;
                __syncthreads();
            }

            if (tid == 0) {
                quat_from_R<lmreal>(trial[0].R, lm.q[0]);
                for (int i = 0; i < 3; ++i) lm.t[0][i] = (lmreal)trial[0].t[i];
                lm.last_residual = INFINITY;
            }
            float last_good = 0.f, last_bad = 0.f, last_sw = 0.f, last_su = 0.f;   // of the last evaluation (thread 0 only)

            for (int lvl = min_lvl; lvl >= max_lvl && !comm_failed; --lvl) {
                const LevelIn Lin = P.lvl[lvl];
                const int n = *Lin.n_pts;
                // block-cyclic split of the list over the CTAs of the cluster (and the ranks of a GPU split): member m takes
                // the blocks m, m + M, m + 2M, ... of kThreads points -- balanced (the exchange waits for the slowest CTA)
                // and the cluster as a whole still sweeps the tile-major list front to back
                const int stride = n_members * kThreads;
                const int first_idx = member * kThreads + tid;
                // this thread's points: first_idx, first_idx + stride, ... (the count differs by at most one over the cluster)
                const int my_iter = first_idx < n ? (n - first_idx + stride - 1) / stride : 0;
                const int my_cached = my_iter < pcap ? my_iter : pcap;
                const float4 *__restrict__ pts = Lin.pts;
                LevelConst L;
                L.fx = Lin.fx; L.fy = Lin.fy; L.cx = Lin.cx; L.cy = Lin.cy;
                L.umax = pin((float)(Lin.w - 2)); L.vmax = pin((float)(Lin.h - 2)); L.tw16 = (unsigned)opt_tiles_per_row_dev(Lin.w) << 4; L.opt = Lin.opt;
                const float ed_eff = pin(use_filter ? oc.edge_distance_lvl[lvl] : INFINITY);
                const float huber = pin(oc.huber_edge);
                const float kqfx = pin(Lin.fx * (1.0f / 32764.0f)), kqfy = pin(Lin.fy * (1.0f / 32764.0f));
                // this thread's points of the level -> its private slots of the shared-memory cache
                for (int k = 0; k < my_cached; ++k) {
                    const float4 p = ldg_point<kHint>(pts + first_idx + (size_t)k * stride);
                    sts3<kThreads>(s_base + (uint32_t)k * kPtStride, p.x, p.y, p.z);
                }
                bool first = true;
                __syncthreads();
                while (true) {
                    const int cur = ctrl.cur;
                    const int wb = lm.acc ^ 1;     // record buffer this evaluation writes (lm.acc: the accepted pose's record)
                    float R[9], t[3];
#pragma unroll
                    for (int i = 0; i < 9; ++i) R[i] = trial[cur].R[i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) t[i] = trial[cur].t[i];
                    float acc[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                    const long long c_begin = prm.profile ? clock64() : 0;
                    // Two pipelined segments (cached points, then the uncached tail of a long level), each with two
                    // register sets (A/B): while point k is being finished the 256-bit gather of point k+1 is in flight.
                    auto segment = [&](auto from_smem, int k0, int k1) {
                        constexpr bool kS = decltype(from_smem)::value;
                        if (k0 >= k1) return;
                        uint32_t sp = s_base + (uint32_t)k0 * kPtStride;
                        const float4 *gp = pts + first_idx + (size_t)k0 * stride;
                        auto arm = [&](ProjB &Q, uint2 (&q)[4]) {
                            float x, y, z;
                            if constexpr (kS) {
                                lds3<kThreads>(sp, x, y, z);
                                sp += kPtStride;
                            } else {
                                const float4 p = ldg_point<kHint>(gp);
                                gp += stride;
                                x = p.x; y = p.y; z = p.z;
                            }
                            Q = project_b(x, y, z, L, R, t);
                            q[0] = ldg_texel<kHint>(L.opt + Q.i00, policy); q[1] = ldg_texel<kHint>(L.opt + Q.i10, policy);
                            q[2] = ldg_texel<kHint>(L.opt + Q.i01, policy); q[3] = ldg_texel<kHint>(L.opt + Q.i11, policy);
                        };
                        ProjB A, B;
                        uint2 qa[4], qb[4];
                        arm(A, qa);
                        int left = k1 - k0 - 1;   // points of the segment not yet armed
                        while (true) {
                            if (left > 0) arm(B, qb);
                            finish_point_b(A, qa[0], qa[1], qa[2], qa[3], kqfx, kqfy, ed_eff, huber, acc);
                            if (left <= 0) break;
                            if (left > 1) arm(A, qa);
                            finish_point_b(B, qb[0], qb[1], qb[2], qb[3], kqfx, kqfy, ed_eff, huber, acc);
                            if (left <= 1) break;
                            left -= 2;
                        }
                    };
                    segment(std::true_type{}, 0, my_cached);
                    segment(std::false_type{}, my_cached, my_iter);
                    acc[kRecBad] = (float)my_iter - acc[kRecGood];   // every visited point exists
                    const long long c_gather = prm.profile ? clock64() : 0;
                    const bool spec_now = speculate && !first && prm.mode != 2;
                    reduce_record(acc, spec_now, cur, wb);
                    const long long c_reduce = prm.profile ? clock64() : 0;
                    evals_lvl[lvl]++;
                    if (prm.mode == 2) {   // single evaluation: export the record (warp 0 wrote it)
                        if (crank == 0 && tid < 32 && records) records[(size_t)pair * 32 + tid] = rec[wb][tid];
                        break;
                    }
                    if (tid == 0) {
                        const double *r = rec[wb];
                        last_good = (float)r[kRecGood]; last_bad = (float)r[kRecBad];
                        last_sw = (float)r[kRecSW]; last_su = (float)r[kRecSU];
                        // Optimizer::trackFrames LM logic, optimizer.cpp:243-306 (track_common.cuh: lm_step).  reduce_record
                        // advanced seq: the evaluation just finished used specin[(seq - 1) & 1], the next reads specin[seq & 1]
                        revo_trace_entry te;
                        bool traced;
                        int next = cur;
                        LMOrder order;
                        const bool done = lm_step(lm, trial, next, rec, spec_now, specin[(seq - 1) & 1], specin[seq & 1], order, oc, lvl,
                                                  first, &te, &traced);
                        // a fresh proposal is needed (first evaluation, accepted try, or no speculation)
                        if (order.propose) lm_propose(recs[order.acc], lm.q[order.pacc], lm.t[order.pacc], order.lambda, trial[order.slot]);
                        if (traced) {
                            if (trace && crank == 0 && ntrace < prm.trace_cap) trace[(size_t)pair * prm.trace_cap + ntrace] = te;
                            ntrace++;
                        }
                        ctrl.cur = next;
                        ctrl.level_done = done ? 1 : 0;
                    }
                    first = false;
                    __syncthreads();
                    if (world > 1 && rec[wb][31] != 0.0) { comm_failed = true; break; }   // an exchange timed out (uniform)
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
                const Trial &fin = trial[ctrl.cur];
                for (int i = 0; i < 9; ++i) o.R[i] = fin.R[i];
                for (int i = 0; i < 3; ++i) o.t[i] = fin.t[i];
                o.error = lm.last_residual;
                o.res.good_pts_edges = (int)last_good;
                o.res.bad_pts_edges = (int)last_bad;
                o.res.sum_error_weighted = last_sw;
                o.res.sum_error_unweighted = last_su;
                // tracker.cpp:351: good/bad < 4 -> NEW_KF (double division; bad == 0 -> inf -> OK)
                o.status = ((double)last_good / (double)last_bad < 4.0) ? REVO_TRACKER_STATE_NEW_KF : REVO_TRACKER_STATE_OK;
                o.rc = comm_failed ? REVO_ERR_COMM : REVO_OK;
                if (comm_failed) o.status = REVO_TRACKER_STATE_UNKNOWN;
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
    if (C > 1 || world > 1) cluster.sync();   // nobody may exit while a peer can still write into its shared memory
}

// Descriptor upload without the copy engine: the host writes the pair descriptors into pinned, device-mapped memory and
// this kernel pulls them into the workspace.  A cudaMemcpyAsync would queue behind the multi-hundred-megabyte frame
// uploads of the next batches on the same H2D engine and stall the tracker for milliseconds.
__global__ void __launch_bounds__(256) k_stage_in(const uint4 *__restrict__ src_mapped_host, uint4 *__restrict__ dst, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src_mapped_host[i];
}

int launch_stage_in(revo_ctx *ctx, const void *src_mapped_host, void *dst, size_t bytes)
{
    const size_t n16 = (bytes + 15) / 16;
    const int blocks = (int)((n16 + 255) / 256 < 64 ? (n16 + 255) / 256 : 64);
    k_stage_in<<<blocks < 1 ? 1 : blocks, 256, 0, ctx->stream>>>((const uint4 *)src_mapped_host, (uint4 *)dst, n16);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "k_stage_in launch");
    ctx->launches++;
    return REVO_OK;
}

// ---- launcher -------------------------------------------------------------------
template <int kThreads, int kMinBlocks, int kHint>
static int launch_track_t(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, int ctas_per_pair,
                          revo_track_result *d_results, double *d_records, revo_trace_entry *d_trace, int *d_trace_counts,
                          int *d_work_counter)
{
    auto kern = k_track<kThreads, kMinBlocks, kHint>;
    if (ctas_per_pair > 8) REVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // points per thread cached in shared memory: 144 KB of the SM's 228 KB over the resident CTAs (24 points per thread at
    // 4 x 128 threads: a VGA level 0 has ~20-25; the gathers bypass L1, so little L1 is needed)
    const int env_pcap = getenv("REVO_TRACK_PCAP") ? atoi(getenv("REVO_TRACK_PCAP")) : -1;
    int pcap = env_pcap >= 0 ? env_pcap : (int)((144 * 1024 / kMinBlocks) / (12 * kThreads));
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
    // persistent: as many clusters as can be co-resident, never more than there are pairs
    cfg.gridDim = dim3(ctas_per_pair);
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        (void)cudaGetLastError();
        max_clusters = ctx->prop.multiProcessorCount / ctas_per_pair;
        if (max_clusters < 1) max_clusters = 1;
    }
    // optional cap on the number of pairs in flight (their lookup structures should stay L2-resident)
    const int env_maxc = getenv("REVO_TRACK_MAX_CLUSTERS") ? atoi(getenv("REVO_TRACK_MAX_CLUSTERS")) : ctx->track_max_clusters;
    if (env_maxc > 0 && max_clusters > env_maxc) max_clusters = env_maxc;
    const int n_clusters = n_pairs < max_clusters ? n_pairs : max_clusters;
    cfg.gridDim = dim3(n_clusters * ctas_per_pair);
    REVO_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, d_pairs, n_pairs, prm, d_results, d_records, d_trace, d_trace_counts,
                                      d_work_counter, pcap));
    ctx->launches++;
    return REVO_OK;
}

int launch_track(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                 double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, int *d_work_counter)
{
    if (n_pairs <= 0) return REVO_OK;
    // Default shape (measured on B200, scratch/track_bench.py): clusters of 8 CTAs; 128-thread CTAs (4 per SM, so that an
    // SM interleaves four different pairs) once more than one wave of 256-thread clusters would be needed.
    const int slots256 = 2 * ctx->prop.multiProcessorCount;
    int C = ctx->track_ctas_per_pair > 0 ? ctx->track_ctas_per_pair : 8;
    const int T = ctx->track_threads > 0 ? ctx->track_threads : ((long long)n_pairs * C > slots256 ? 128 : 256);
    // A/B switch for profiling only: REVO_TRACK_HINT = 0 plain gathers, 1 L1::no_allocate, 2 (default) L2 eviction priorities
    const int hint = getenv("REVO_TRACK_HINT") ? atoi(getenv("REVO_TRACK_HINT")) : 2;
#define REVO_TRACK_ARGS ctx, d_pairs, n_pairs, prm, C, d_results, d_records, d_trace, d_trace_counts, d_work_counter
    switch (T) {
        case 128:
            return hint == 0 ? launch_track_t<128, 4, 0>(REVO_TRACK_ARGS)
                 : hint == 2 ? launch_track_t<128, 4, 2>(REVO_TRACK_ARGS) : launch_track_t<128, 4, 1>(REVO_TRACK_ARGS);
        case 512: return launch_track_t<512, 1, 2>(REVO_TRACK_ARGS);
        case 1024: return launch_track_t<1024, 1, 2>(REVO_TRACK_ARGS);
        default:
            return hint == 0 ? launch_track_t<256, 2, 0>(REVO_TRACK_ARGS)
                 : hint == 2 ? launch_track_t<256, 2, 2>(REVO_TRACK_ARGS) : launch_track_t<256, 2, 1>(REVO_TRACK_ARGS);
    }
#undef REVO_TRACK_ARGS
}

}  // namespace revo
