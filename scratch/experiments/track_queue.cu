// track_queue.cu -- K9/K10, engine "queue": the coarse-to-fine Gauss-Newton / Levenberg-Marquardt edge alignment of a
// BATCH of frame pairs as one persistent kernel driven by a chip-wide task queue.
//
// Replaces (reference file:line, fabianschenk/REVO) -- the same functions as track.cu:
//   TrackerNew::trackFrames / checkInitializationValues / evalCostFunction   system/tracker.cpp:294-353, 265-283, 357-393
//   Optimizer::trackFrames (LM loop)                                          system/optimizer.cpp:235-311
//   Optimizer::calcErrorAndBuffers + getInterpolatedElement43                 system/optimizer.cpp:74-191, optimizer.h:173-185
//   Optimizer::calculateWarpUpdate + LGS6::update/finish                      system/optimizer.cpp:192-234, utils/LGSX.h:320-326,392-398
//   Eigen LDLT 6x6 solve, Sophus::SE3f exp / product                          system/optimizer.cpp:258-266
//
// Why a second engine: with one cluster per pair (track.cu) a launch lasts as long as its slowest pair (pairs need
// 25..60 evaluations), every evaluation pays reduction + solve latency with the pair's CTAs idle, and the coarse
// levels have too few points to fill the CTAs they own.  Here NO CTA owns a pair:
//   * an evaluation of a pair at a pose is cut into K chunks of its 3-D edge list; each chunk is a TASK in a bounded
//     multi-producer / multi-consumer ring in global memory (ticket counters, per-slot sequence words);
//   * every CTA of the persistent grid loops: pop a task -> fused PASS A + PASS B over the chunk (branch-free,
//     software-pipelined gathers, 21+6+4 sums in registers) -> transposing warp-shuffle reduction -> one 32-double
//     partial in the pair's slot table -> atomic arrival counter;
//   * the CTA whose arrival completes the evaluation ("last arriver") sums the K partials in chunk order
//     (deterministic, independent of which CTA ran which chunk), runs the LM state machine (6x6 LDL^T, SE3 update,
//     accept / reject, level switch) for that pair and pushes the tasks of the pair's NEXT evaluation.
// No CTA ever waits for another pair's progress, all SMs stay busy until the work runs out, and the tail of a launch
// is one pair's critical path with its chunks spread over the whole chip.  No tensor cores (no dense contraction).
//
// Memory ordering: payload stores -> __threadfence() -> flag/counter (producer); poll with volatile loads -> match ->
// __threadfence() -> payload loads through L2 (__ldcg) (consumer).  Pair state is only touched by the pair's current
// last arriver; successive last arrivers are ordered through the queue publication and the arrival counter.
#include <stdlib.h>

#include "internal.h"
#include "track_common.cuh"

namespace revo {

constexpr int kQMaxChunks = 64;
constexpr int kTaskEval = 0, kTaskCost = 1, kTaskExit = 2;

struct QTask {             // one ring slot, 128 bytes
    unsigned seq;          // 2*lap: free for the producer of lap `lap`; 2*lap+1: holds the task of that lap
    int kind;
    int pair, lvl;
    int begin, end;        // point range of the level's list
    int chunk, n_chunks;
    float R[9], t[3];      // pose of the evaluation
    int pad[12];
};
static_assert(sizeof(QTask) == 128, "slot size");

struct QPair {             // mutable per-pair state (global memory; touched by the pair's current last arriver only)
    LMState lm;
    float R[9], t[3];      // pose under evaluation / accepted pose between levels
    float last_good, last_bad, last_sw, last_su;
    int evals_lvl[REVO_MAX_LEVELS];
    int lvl, first, phase, n_chunks, n_pts, ntrace, used_identity;
    int done;              // arrival counter of the evaluation in flight
};
static_assert(sizeof(QPair) % 8 == 0, "QPair is staged through shared memory in 8-byte words");

struct QCtl {
    unsigned head, tail;   // consumer / producer ticket counters
    int pairs_done;
    int abort;             // watchdog: a spin loop gave up
    unsigned long long t_start;                 // globaltimer of the first CTA to start (profile)
    unsigned long long stat[8];                 // profile: cycles in pop / gather / partial / last-arriver, #tasks, #last arrivals
    unsigned long long pair_finish_ns[1];       // profile: [n_pairs] finish time of every pair relative to t_start (flexible tail)
};

// ---- small PTX helpers ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(unsigned *p, unsigned v)
{
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr long long kWatchdogCycles = 6000000000ll;   // ~3 s at 1.9 GHz: far beyond any legitimate wait

struct QArgs {
    const PairDesc *pairs;
    int n_pairs;
    revo_track_result *results;
    double *records;
    revo_trace_entry *trace;
    int *trace_counts;
    QCtl *ctl;
    QTask *ring;
    unsigned cap_mask, cap_log2;
    QPair *qp;
    double *partials;      // [n_pairs][kQMaxChunks][32]
    int smin, kmax;
};

// Warp-collective: publish K tasks of one evaluation (or `K` exit tasks).  Pose from shared memory.
__device__ __forceinline__ void push_tasks(const QArgs &a, int lane, int kind, int pair, int lvl, int n_pts, int K, const float *R,
                                           const float *t)
{
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(&a.ctl->tail, (unsigned)K);
    base = __shfl_sync(kFull, base, 0);
    const int S = K > 0 ? (n_pts + K - 1) / K : 0;
    for (int c = lane; c < K; c += 32) {
        const unsigned ticket = base + (unsigned)c;
        QTask *slot = a.ring + (ticket & a.cap_mask);
        const unsigned lap2 = (ticket >> a.cap_log2) * 2u;
        const long long t0 = clock64();
        while (ld_volatile_u32(&slot->seq) != lap2) {   // the consumer of the previous lap has not released it yet
            __nanosleep(64);
            if (clock64() - t0 > kWatchdogCycles) { atomicExch(&a.ctl->abort, 1); break; }
        }
        slot->kind = kind; slot->pair = pair; slot->lvl = lvl;
        const int b = kind == kTaskExit ? 0 : min(n_pts, c * S);
        const int e = kind == kTaskExit ? 0 : min(n_pts, b + S);
        slot->begin = b; slot->end = e; slot->chunk = c; slot->n_chunks = K;
        if (kind != kTaskExit) {
#pragma unroll
            for (int i = 0; i < 9; ++i) slot->R[i] = R[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) slot->t[i] = t[i];
        }
        __threadfence();
        st_volatile_u32(&slot->seq, lap2 + 1u);
    }
    __syncwarp();
}

__device__ __forceinline__ int chunks_for(const QArgs &a, int n_pts)
{
    int K = (n_pts + a.smin - 1) / a.smin;
    K = K < 1 ? 1 : K;
    return K > a.kmax ? a.kmax : K;
}

template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks) k_track_queue(const QArgs a, const TrackParams prm)
{
    constexpr int kWarps = kThreads / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    __shared__ float warp_part[kWarps][32];
    __shared__ __align__(16) QTask task;
    __shared__ __align__(16) double rec[32];
    __shared__ __align__(16) QPair sp;            // staged copy of the pair state (last arriver only)
    __shared__ int s_action[4];                   // lane 0 -> warp: {what, lvl, n_pts, K}

    const revo_opt_config &oc = prm.cfg.opt;
    const bool use_filter = oc.use_edge_filter != 0;
    const int min_lvl = prm.mode == 0 ? prm.cfg.pyr_min_lvl : prm.level;
    const int max_lvl = prm.mode == 0 ? prm.cfg.pyr_max_lvl : prm.level;
    constexpr int kWords = (int)(sizeof(QPair) / 8);

    // actions computed by lane 0 of the acting warp
    constexpr int kActNone = 0, kActPush = 1, kActFinished = 2;

    // ---- pair finished: result record (lane 0) ---------------------------------------------------------------
    auto write_result = [&](int pair, const PairDesc &P, bool skipped) {
        if (prm.mode == 2) return;
        revo_track_result &o = a.results[pair];
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
            if (a.trace_counts) a.trace_counts[pair] = 0;
            return;
        }
        for (int i = 0; i < 9; ++i) o.R[i] = sp.R[i];
        for (int i = 0; i < 3; ++i) o.t[i] = sp.t[i];
        o.error = sp.lm.last_residual;
        o.res.good_pts_edges = (int)sp.last_good;
        o.res.bad_pts_edges = (int)sp.last_bad;
        o.res.sum_error_weighted = sp.last_sw;
        o.res.sum_error_unweighted = sp.last_su;
        // tracker.cpp:351: good/bad < 4 -> NEW_KF (double division; bad == 0 -> inf -> OK)
        o.status = ((double)sp.last_good / (double)sp.last_bad < 4.0) ? REVO_TRACKER_STATE_NEW_KF : REVO_TRACKER_STATE_OK;
        o.rc = REVO_OK;
        for (int l = 0; l < REVO_MAX_LEVELS; ++l) {
            o.n_evals[l] = sp.evals_lvl[l];
            o.n_pts[l] = (l >= max_lvl && l <= min_lvl) ? *P.lvl[l].n_pts : 0;
        }
        o.used_identity_init = sp.used_identity;
        if (a.trace_counts) a.trace_counts[pair] = sp.ntrace < prm.trace_cap ? sp.ntrace : prm.trace_cap;
    };

    // lane 0: start level `lvl` from the accepted pose in sp.R/sp.t (optimizer.cpp:241-249 happen on the first record)
    auto begin_level = [&](const PairDesc &P, int lvl) {
        sp.lvl = lvl; sp.first = 1; sp.phase = 1;
        sp.n_pts = *P.lvl[lvl].n_pts;
        sp.n_chunks = chunks_for(a, sp.n_pts);
        sp.done = 0;
        s_action[0] = kActPush; s_action[1] = kTaskEval;
    };

    // Warp-collective tail of every state change: store the staged pair state, then publish / finish.
    auto commit = [&](int pair) {
        __syncwarp();
        const int what = s_action[0];
        if (what == kActPush) {
            unsigned long long *g = (unsigned long long *)(a.qp + pair);
            const unsigned long long *s = (const unsigned long long *)&sp;
            for (int i = lane; i < kWords; i += 32) g[i] = s[i];
            __threadfence();
            __syncwarp();
            push_tasks(a, lane, s_action[1], pair, sp.lvl, sp.n_pts, sp.n_chunks, sp.R, sp.t);
        } else if (what == kActFinished) {
            int last = 0;
            if (lane == 0 && prm.profile) a.ctl->pair_finish_ns[pair] = globaltimer_ns() - *(volatile unsigned long long *)&a.ctl->t_start;
            if (lane == 0) last = atomicAdd(&a.ctl->pairs_done, 1) == a.n_pairs - 1;
            last = __shfl_sync(kFull, last, 0);
            if (last) {   // everything is done: one exit task per CTA of the grid
                int left = (int)gridDim.x;
                while (left > 0) {
                    const int k = left < 1024 ? left : 1024;
                    push_tasks(a, lane, kTaskExit, 0, 0, 0, k, sp.R, sp.t);
                    left -= k;
                }
            }
        }
        __syncwarp();
    };

    long long pc_pop = 0, pc_gather = 0, pc_part = 0, pc_la = 0, pc_nla = 0;
    if (prm.profile && tid == 0) atomicCAS(&a.ctl->t_start, 0ull, globaltimer_ns());
    // ---- phase 0: every CTA starts a share of the pairs (warp 0) ------------------------------------------------
    if (wid == 0) {
        for (int pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
            const PairDesc &P = a.pairs[pair];
            if (lane == 0) {
                s_action[0] = kActNone;
                for (int i = 0; i < 9; ++i) sp.R[i] = P.R[i];
                for (int i = 0; i < 3; ++i) sp.t[i] = P.t[i];
                sp.last_good = sp.last_bad = sp.last_sw = sp.last_su = 0.f;
                for (int l = 0; l < REVO_MAX_LEVELS; ++l) sp.evals_lvl[l] = 0;
                sp.ntrace = 0; sp.used_identity = 0; sp.done = 0; sp.first = 1;
                sp.lm.last_residual = INFINITY;
                if (!rotation_ok(P.R)) {
                    write_result(pair, P, true);
                    s_action[0] = kActFinished;
                } else if (prm.mode == 0 && prm.cfg.check_init_values) {
                    // checkInitializationValues (tracker.cpp:265-283): cost at identity vs cost at (R,t), coarsest level
                    sp.phase = 0; sp.lvl = min_lvl;
                    sp.n_pts = *P.lvl[min_lvl].n_pts;
                    sp.n_chunks = chunks_for(a, sp.n_pts);
                    s_action[0] = kActPush; s_action[1] = kTaskCost;
                } else {
                    quat_from_R(sp.R, sp.lm.q);
                    for (int i = 0; i < 3; ++i) sp.lm.t[i] = sp.t[i];
                    begin_level(P, min_lvl);
                }
            }
            commit(pair);
        }
    }
    __syncthreads();

    // ---- phase 1: consume tasks until the exit task arrives --------------------------------------------------
    unsigned long long n_tasks = 0;
    while (true) {
        const long long c0 = prm.profile ? clock64() : 0;
        if (tid == 0) {
            const unsigned ticket = atomicAdd(&a.ctl->head, 1u);
            QTask *slot = a.ring + (ticket & a.cap_mask);
            const unsigned lap2 = (ticket >> a.cap_log2) * 2u;
            const long long t0 = clock64();
            bool ok = true;
            while (ld_volatile_u32(&slot->seq) != lap2 + 1u) {
                __nanosleep(100);
                if (clock64() - t0 > kWatchdogCycles || *(volatile int *)&a.ctl->abort) { atomicExch(&a.ctl->abort, 1); ok = false; break; }
            }
            __threadfence();
            if (ok) {
                const int4 *src = (const int4 *)slot;
                int4 *dst = (int4 *)&task;
#pragma unroll
                for (int i = 0; i < 5; ++i) dst[i] = __ldcg(src + i);   // 80 bytes: header + pose
                __threadfence();
                st_volatile_u32(&slot->seq, lap2 + 2u);                  // free for the next lap
            } else {
                task.kind = kTaskExit;
            }
        }
        __syncthreads();                                                 // B1: task visible
        const long long c1 = prm.profile ? clock64() : 0;
        const int kind = task.kind;
        if (kind == kTaskExit) break;
        const int pair = task.pair, lvl = task.lvl;
        const int begin = task.begin, end = task.end, chunk = task.chunk, n_chunks = task.n_chunks;
        float R[9], t[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) R[i] = task.R[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) t[i] = task.t[i];
        const PairDesc &P = a.pairs[pair];
        const LevelIn Lin = P.lvl[lvl];
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        const float ed = oc.edge_distance_lvl[lvl];

        if (kind == kTaskCost) {
            const float *dtm = P.ref_dt_min;
            for (int i = begin + tid; i < end; i += kThreads) {
                const float4 p = __ldg(Lin.pts + i);
                acc[0] += cost_point(p.x, p.y, p.z, Lin, dtm, ed, use_filter);
                const float X = R[0] * p.x + R[3] * p.y + R[6] * p.z + t[0];
                const float Y = R[1] * p.x + R[4] * p.y + R[7] * p.z + t[1];
                const float Z = R[2] * p.x + R[5] * p.y + R[8] * p.z + t[2];
                acc[1] += cost_point(X, Y, Z, Lin, dtm, ed, use_filter);
            }
        } else {
            LevelConst L;
            L.fx = Lin.fx; L.fy = Lin.fy; L.cx = Lin.cx; L.cy = Lin.cy;
            L.umax = (float)(Lin.w - 2); L.vmax = (float)(Lin.h - 2); L.w = Lin.w; L.opt = Lin.opt;
            const float huber = oc.huber_edge;
            const float4 *__restrict__ pts = Lin.pts;
            const int safe = begin < end ? begin : 0;
            // Software pipeline, two register sets (A/B): while point k is finished, the two texel-pair gathers of
            // point k+1 and the list entry of point k+2 are in flight.
            int i = begin + tid;
            bool eA = i < end;
            float4 pA = __ldg(pts + (eA ? i : safe));
            i += kThreads;
            bool eB = i < end;
            float4 pB = __ldg(pts + (eB ? i : safe));
            ProjB A = project_b(eA, pA, L, R, t);
            uint4 a0, a1;
            ldg_quad(A.bp, a0, a1);
            while (true) {
                i += kThreads;
                const bool eC = i < end;
                const float4 pC = __ldg(pts + (eC ? i : safe));
                const ProjB B = project_b(eB, pB, L, R, t);
                uint4 b0, b1;
                ldg_quad(B.bp, b0, b1);
                finish_point_b(A, a0, a1, L, ed, use_filter, huber, acc);
                if (!eB) break;
                i += kThreads;
                const bool eD = i < end;
                const float4 pD = __ldg(pts + (eD ? i : safe));
                A = project_b(eC, pC, L, R, t);
                ldg_quad(A.bp, a0, a1);
                finish_point_b(B, b0, b1, L, ed, use_filter, huber, acc);
                if (!eC) break;
                eB = eD;
                pB = pD;
            }
        }

        // ---- chunk partial: transposing shuffle tree -> shared -> double -> the pair's slot table
        const long long c2 = prm.profile ? clock64() : 0;
        const float mine = warp_transpose_reduce(acc, lane);
        warp_part[wid][lane] = mine;
        __syncthreads();                                                 // B2
        if (wid == 0) {
            double s = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += (double)warp_part[w][lane];
            double *slotp = a.partials + ((size_t)pair * kQMaxChunks + chunk) * 32;
            __stcg(slotp + lane, s);
            __threadfence();
            __syncwarp();
            int last = 0;
            if (lane == 0) last = atomicAdd(&a.qp[pair].done, 1) == n_chunks - 1;
            last = __shfl_sync(kFull, last, 0);
            const long long c3 = prm.profile ? clock64() : 0;
            if (prm.profile) { pc_pop += c1 - c0; pc_gather += c2 - c1; pc_part += c3 - c2; }
            if (last) {
                // ---- last arriver: the evaluation of `pair` is complete
                __threadfence();
                const double *pp = a.partials + (size_t)pair * kQMaxChunks * 32 + lane;
                double tot = 0;
                for (int c = 0; c < n_chunks; ++c) tot += __ldcg(pp + (size_t)c * 32);
                rec[lane] = tot;
                {
                    const unsigned long long *g = (const unsigned long long *)(a.qp + pair);
                    unsigned long long *s8 = (unsigned long long *)&sp;
                    for (int k = lane; k < kWords; k += 32) s8[k] = __ldcg(g + k);
                }
                __syncwarp();
                if (lane == 0) {
                    s_action[0] = kActNone;
                    if (sp.phase == 0) {
                        if ((float)rec[0] < (float)rec[1]) {   // tracker.cpp:277
                            for (int k = 0; k < 9; ++k) sp.R[k] = (k % 4 == 0) ? 1.f : 0.f;
                            for (int k = 0; k < 3; ++k) sp.t[k] = 0.f;
                            sp.used_identity = 1;
                        }
                        quat_from_R(sp.R, sp.lm.q);
                        for (int k = 0; k < 3; ++k) sp.lm.t[k] = sp.t[k];
                        begin_level(P, min_lvl);
                    } else {
                        sp.evals_lvl[lvl]++;
                        sp.last_good = (float)rec[kRecGood]; sp.last_bad = (float)rec[kRecBad];
                        sp.last_sw = (float)rec[kRecSW]; sp.last_su = (float)rec[kRecSU];
                        if (prm.mode == 2) {   // single evaluation: export the record
                            if (a.records)
                                for (int k = 0; k < 32; ++k) a.records[(size_t)pair * 32 + k] = rec[k];
                            s_action[0] = kActFinished;
                        } else {
                            revo_trace_entry te;
                            bool traced;
                            const bool done = lm_step(sp.lm, rec, oc, lvl, sp.first != 0, sp.R, sp.t, &te, &traced);
                            sp.first = 0;
                            if (traced) {
                                if (a.trace && sp.ntrace < prm.trace_cap) a.trace[(size_t)pair * prm.trace_cap + sp.ntrace] = te;
                                sp.ntrace++;
                            }
                            if (!done) {
                                sp.done = 0;
                                s_action[0] = kActPush; s_action[1] = kTaskEval;
                            } else if (lvl > max_lvl) {
                                begin_level(P, lvl - 1);
                            } else {
                                write_result(pair, P, false);
                                s_action[0] = kActFinished;
                            }
                        }
                    }
                }
                commit(pair);
                if (prm.profile) { pc_la += clock64() - c3; pc_nla++; }
            }
        }
        ++n_tasks;
    }
    if (prm.profile && tid == 0) {
        atomicAdd(&a.ctl->stat[0], (unsigned long long)pc_pop);
        atomicAdd(&a.ctl->stat[1], (unsigned long long)pc_gather);
        atomicAdd(&a.ctl->stat[2], (unsigned long long)pc_part);
        atomicAdd(&a.ctl->stat[3], (unsigned long long)pc_la);
        atomicAdd(&a.ctl->stat[4], n_tasks);
        atomicAdd(&a.ctl->stat[5], (unsigned long long)pc_nla);
        atomicAdd(&a.ctl->stat[6], 1ull);
    }
}

// ---- launcher -------------------------------------------------------------------------------------------------
size_t track_queue_workspace_bytes(int n_pairs, int grid_cap, unsigned *cap_out)
{
    unsigned need = (unsigned)n_pairs * kQMaxChunks + (unsigned)grid_cap + 64u;
    unsigned cap = 1024;
    while (cap < need) cap <<= 1;
    if (cap_out) *cap_out = cap;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    return al(sizeof(QCtl) + 8 * (size_t)n_pairs) + al(sizeof(QTask) * (size_t)cap) + al(sizeof(QPair) * (size_t)n_pairs) +
           al(sizeof(double) * 32 * kQMaxChunks * (size_t)n_pairs);
}

template <int kThreads, int kMinBlocks>
static int launch_q(revo_ctx *ctx, QArgs &a, const TrackParams &prm, int oversub_x4)
{
    auto kern = k_track_queue<kThreads, kMinBlocks>;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
    }
    const int full = per_sm * ctx->prop.multiProcessorCount;
    // chunks per evaluation: enough tasks in flight to keep `oversub` x the resident CTAs fed when all pairs are active
    int kmax = (int)(((long long)full * oversub_x4 / 4 + a.n_pairs - 1) / a.n_pairs);
    kmax = kmax < 1 ? 1 : (kmax > kQMaxChunks ? kQMaxChunks : kmax);
    a.kmax = kmax;
    long long want = (long long)a.n_pairs * kmax;
    int grid = (int)(want < full ? want : full);
    if (grid < 1) grid = 1;
    kern<<<grid, kThreads, 0, ctx->stream>>>(a, prm);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "k_track_queue launch");
    ctx->launches++;
    return REVO_OK;
}

int launch_track_queue(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                       double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, void *d_ws, size_t ws_bytes)
{
    if (n_pairs <= 0) return REVO_OK;
    unsigned cap = 0;
    const int grid_cap = 8 * ctx->prop.multiProcessorCount;
    const size_t need = track_queue_workspace_bytes(n_pairs, grid_cap, &cap);
    if (ws_bytes < need) return REVO_ERR_INVALID_ARG;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    uint8_t *p = (uint8_t *)d_ws;
    QArgs a;
    a.pairs = d_pairs; a.n_pairs = n_pairs; a.results = d_results; a.records = d_records; a.trace = d_trace;
    a.trace_counts = d_trace_counts;
    a.ctl = (QCtl *)p; p += al(sizeof(QCtl) + 8 * (size_t)n_pairs);
    a.ring = (QTask *)p; p += al(sizeof(QTask) * (size_t)cap);
    a.qp = (QPair *)p; p += al(sizeof(QPair) * (size_t)n_pairs);
    a.partials = (double *)p;
    a.cap_mask = cap - 1;
    unsigned lg = 0;
    while ((1u << lg) < cap) ++lg;
    a.cap_log2 = lg;
    // ring sequence words and the control block start at zero (= every slot free for lap 0)
    REVO_CUDA(ctx, cudaMemsetAsync(d_ws, 0, al(sizeof(QCtl) + 8 * (size_t)n_pairs) + al(sizeof(QTask) * (size_t)cap), ctx->stream));
    const int env_smin = getenv("REVO_Q_SMIN") ? atoi(getenv("REVO_Q_SMIN")) : 0;
    const int env_over = getenv("REVO_Q_OVERSUB_X4") ? atoi(getenv("REVO_Q_OVERSUB_X4")) : 0;
    const int env_thr = getenv("REVO_Q_THREADS") ? atoi(getenv("REVO_Q_THREADS")) : 0;
    const int T = ctx->track_threads > 0 ? ctx->track_threads : (env_thr > 0 ? env_thr : 128);
    a.smin = ctx->track_chunk_points > 0 ? ctx->track_chunk_points : (env_smin > 0 ? env_smin : 4 * T);
    const int over = env_over > 0 ? env_over : 8;   // 2.0 x
    switch (T) {
        case 256: return launch_q<256, 2>(ctx, a, prm, over);
        case 512: return launch_q<512, 1>(ctx, a, prm, over);
        default: return launch_q<128, 4>(ctx, a, prm, over);
    }
}

}  // namespace revo
