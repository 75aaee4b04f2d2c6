// track_lean.cu -- EXPERIMENT, NOT PART OF THE LIBRARY and NOT YET RUN ON HARDWARE (written after the GPU budget of round 1
// was spent; the evidence so far is the static SASS instruction count, see scratch/experiments/README.md).
//
// The cluster-per-pair tracking engine of track.cu with a leaner gather loop.  ncu of k_track says a warp spends 45 % of
// its time on instruction issue + fixed-latency dependencies, and the SASS of its loop holds ~172 instructions per point,
// of which ~25 only re-derive addresses of the shared-memory point cache (generic -> shared window, thread id, pcap from
// the constant bank, predicated smem/global variants of every load) and a few handle points that do not exist.  Here:
//   * a thread only visits the points it really has (the per-thread trip count differs by at most one over the cluster),
//     so the `exists` flag and its selects disappear;
//   * the cached points are read through a running 32-bit shared-memory address with immediate offsets ([k][3][T]
//     layout: one add per point), the uncached tail of a long level through a running global pointer, in two separate
//     pipelined segments instead of one loop with both variants predicated;
//   * the rigid transform is three FMA chains (9 instead of 12 instructions);
//   * optional ld.global.nc.L1::no_allocate on the gather (kHint 3; +5 % in the deep-pipeline experiment).
// Same work split, reduction, exchange and LM step as k_track (reference file:line map there); single GPU only.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

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


// x, y, z of one cached point: three 32-bit shared loads at immediate offsets from one running address
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

// project_b of track_common.cuh for a point that exists, with the rigid transform as three FMA chains
__device__ __forceinline__ ProjB project_l(float x, float y, float z, const LevelConst &L, const float *__restrict__ R,
                                           const float *__restrict__ t)
{
    ProjB o;
    const float Wx = fmaf(R[6], z, fmaf(R[3], y, fmaf(R[0], x, t[0])));
    const float Wy = fmaf(R[7], z, fmaf(R[4], y, fmaf(R[1], x, t[1])));
    const float Wz = fmaf(R[8], z, fmaf(R[5], y, fmaf(R[2], x, t[2])));
    const float iz = rcp_approx(Wz);
    const float a = Wx * iz, b = Wy * iz;
    const float u = a * L.fx + L.cx;
    const float v = b * L.fy + L.cy;
    const bool inb = (u > 1.f && v > 1.f && u < L.umax && v < L.vmax);   // NaN-safe (optimizer.cpp:100)
    o.exists = true;
    o.valid = inb;
    const int ix = inb ? (int)u : 0, iy = inb ? (int)v : 0;
    o.dx = inb ? u - (float)ix : 0.f;
    o.dy = inb ? v - (float)iy : 0.f;
    o.a = inb ? a : 0.f;
    o.b = inb ? b : 0.f;
    o.iz = inb ? iz : 0.f;
    o.bp = L.opt + 2u * (unsigned)(iy * L.w + ix);
    return o;
}

// finish_point_b of track_common.cuh with the per-level constants pre-combined and pinned to registers by the caller:
// kqfx = fx / 32764, kqfy = fy / 32764 (gradient scale), ed_eff = edge filter distance or +inf when the filter is off.
// The "bad" counter is not kept here: every visited point exists, so bad = visited - good (set by the caller).
__device__ __forceinline__ void finish_point_l(const ProjB &P, const uint4 r0, const uint4 r1, float kqfx, float kqfy, float ed_eff,
                                               float huber, float (&acc)[32])
{
    // getInterpolatedElement43, optimizer.h:173-185
    const float dxdy = P.dx * P.dy;
    const float w11 = dxdy, w01 = P.dy - dxdy, w10 = P.dx - dxdy, w00 = 1.f - P.dx - P.dy + dxdy;
    float gx00, gy00, gx10, gy10, gx01, gy01, gx11, gy11;
    unpack_grad(r0.z, gx00, gy00); unpack_grad(r0.w, gx10, gy10);
    unpack_grad(r1.z, gx01, gy01); unpack_grad(r1.w, gx11, gy11);
    const float gx = (w11 * gx11 + w01 * gx01 + w10 * gx10 + w00 * gx00) * kqfx;   // optimizer.cpp:119
    const float gy = (w11 * gy11 + w01 * gy01 + w10 * gy10 + w00 * gy00) * kqfy;   // optimizer.cpp:120
    const float r = w11 * __uint_as_float(r1.y) + w01 * __uint_as_float(r1.x) + w10 * __uint_as_float(r0.y) + w00 * __uint_as_float(r0.x);
    const bool pass = P.valid && !(r > ed_eff);                                     // optimizer.cpp:100,112
    const float hub = huber * rcp_approx(fmaxf(r, huber));                          // optimizer.h:159: r <= huber ? 1 : huber / r
    const float wr = pass ? ((r <= huber) ? 1.f : hub) : 0.f;
    const float rs = pass ? r : 0.f;
    acc[kRecGood] += pass ? 1.f : 0.f;
    // calculateWarpUpdate, optimizer.cpp:204-228, factored through a = x/z, b = y/z, s = a gx + b gy
    const float z = P.iz, a = P.a, b = P.b;
    const float s = a * gx + b * gy;
    float J[6];
    J[0] = z * gx;
    J[1] = z * gy;
    J[2] = -(s * z);
    J[3] = -(b * s + gy);
    J[4] = a * s + gx;
    J[5] = a * gy - b * gx;
    // LGS6::update, LGSX.h:392-398 (upper triangle only; A is symmetric)
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float wi = wr * J[i];
#pragma unroll
        for (int j = i; j < 6; ++j) acc[k++] += wi * J[j];
    }
    const float rw = rs * wr;
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[kRecB + i] += rw * J[i];
    acc[kRecSW] += rw * rs;     // optimizer.cpp:131
    acc[kRecSU] += rs * rs;
}

// ---- packed accumulation (kPack): fma.rn.f32x2 (FFMA2 on sm_100a) ---------------------------------------------------------
// The 21 + 6 normal-equation sums as 12 register pairs + 3 scalars: with P0 = (J0,J1), P1 = (J2,J3), P2 = (J4,J5) and
// WPk = w * Pk, the diagonal 2x2 blocks are WPk * Pk (+ one scalar cross term each) and every off-diagonal block is
// WPa * Pb and WPa * swap(Pb): 15 packed + 3 scalar FMAs instead of 27 scalar ones, at the price of two swapped pairs.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b)
{
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

struct PackedAcc {
    float2 d0, d1, d2;                       // (A00,A11) (A22,A33) (A44,A55)
    float2 e01a, e01b, e02a, e02b, e12a, e12b;   // (A02,A13) (A03,A12) (A04,A15) (A05,A14) (A24,A35) (A25,A34)
    float2 b0, b1, b2;                       // (b0,b1) (b2,b3) (b4,b5)
    float a01, a23, a45, sw, su, good;
    __device__ __forceinline__ void clear()
    {
        const float2 z = make_float2(0.f, 0.f);
        d0 = d1 = d2 = e01a = e01b = e02a = e02b = e12a = e12b = b0 = b1 = b2 = z;
        a01 = a23 = a45 = sw = su = good = 0.f;
    }
    // -> record order of track_common.cuh (LGS6 upper triangle row by row, then b, sums, counts)
    __device__ __forceinline__ void unpack(float (&acc)[32], float visited) const
    {
        acc[0] = d0.x;  acc[1] = a01;    acc[2] = e01a.x;  acc[3] = e01b.x;  acc[4] = e02a.x;  acc[5] = e02b.x;
        acc[6] = d0.y;  acc[7] = e01b.y; acc[8] = e01a.y;  acc[9] = e02b.y;  acc[10] = e02a.y;
        acc[11] = d1.x; acc[12] = a23;   acc[13] = e12a.x; acc[14] = e12b.x;
        acc[15] = d1.y; acc[16] = e12b.y; acc[17] = e12a.y;
        acc[18] = d2.x; acc[19] = a45;
        acc[20] = d2.y;
        acc[21] = b0.x; acc[22] = b0.y; acc[23] = b1.x; acc[24] = b1.y; acc[25] = b2.x; acc[26] = b2.y;
        acc[kRecSW] = sw; acc[kRecSU] = su; acc[kRecGood] = good; acc[kRecBad] = visited - good; acc[31] = 0.f;
    }
};

__device__ __forceinline__ void finish_point_p(const ProjB &P, const uint4 r0, const uint4 r1, float kqfx, float kqfy, float ed_eff,
                                               float huber, PackedAcc &S)
{
    const float dxdy = P.dx * P.dy;
    const float w11 = dxdy, w01 = P.dy - dxdy, w10 = P.dx - dxdy, w00 = 1.f - P.dx - P.dy + dxdy;
    float gx00, gy00, gx10, gy10, gx01, gy01, gx11, gy11;
    unpack_grad(r0.z, gx00, gy00); unpack_grad(r0.w, gx10, gy10);
    unpack_grad(r1.z, gx01, gy01); unpack_grad(r1.w, gx11, gy11);
    const float gx = (w11 * gx11 + w01 * gx01 + w10 * gx10 + w00 * gx00) * kqfx;
    const float gy = (w11 * gy11 + w01 * gy01 + w10 * gy10 + w00 * gy00) * kqfy;
    const float r = w11 * __uint_as_float(r1.y) + w01 * __uint_as_float(r1.x) + w10 * __uint_as_float(r0.y) + w00 * __uint_as_float(r0.x);
    const bool pass = P.valid && !(r > ed_eff);
    const float hub = huber * rcp_approx(fmaxf(r, huber));
    const float wr = pass ? ((r <= huber) ? 1.f : hub) : 0.f;
    const float rs = pass ? r : 0.f;
    S.good += pass ? 1.f : 0.f;
    const float z = P.iz, a = P.a, b = P.b;
    const float s = a * gx + b * gy;
    const float J0 = z * gx, J1 = z * gy, J2 = -(s * z), J3 = -(b * s + gy), J4 = a * s + gx, J5 = a * gy - b * gx;
    const float2 P0 = make_float2(J0, J1), P1 = make_float2(J2, J3), P2 = make_float2(J4, J5);
    const float2 P1s = make_float2(J3, J2), P2s = make_float2(J5, J4);
    const float2 W = make_float2(wr, wr);
    const float2 WP0 = fmul2(W, P0), WP1 = fmul2(W, P1), WP2 = fmul2(W, P2);
    S.d0 = ffma2(WP0, P0, S.d0);  S.d1 = ffma2(WP1, P1, S.d1);  S.d2 = ffma2(WP2, P2, S.d2);
    S.a01 = fmaf(WP0.x, J1, S.a01); S.a23 = fmaf(WP1.x, J3, S.a23); S.a45 = fmaf(WP2.x, J5, S.a45);
    S.e01a = ffma2(WP0, P1, S.e01a); S.e01b = ffma2(WP0, P1s, S.e01b);
    S.e02a = ffma2(WP0, P2, S.e02a); S.e02b = ffma2(WP0, P2s, S.e02b);
    S.e12a = ffma2(WP1, P2, S.e12a); S.e12b = ffma2(WP1, P2s, S.e12b);
    const float rw = rs * wr;
    const float2 RW = make_float2(rw, rw);
    S.b0 = ffma2(RW, P0, S.b0); S.b1 = ffma2(RW, P1, S.b1); S.b2 = ffma2(RW, P2, S.b2);
    S.sw = fmaf(rw, rs, S.sw);
    S.su = fmaf(rs, rs, S.su);
}

// keeps a per-level constant in a register (the compiler otherwise re-derives it from the constant bank for every point)
__device__ __forceinline__ float pin(float x)
{
    asm volatile("" : "+f"(x));
    return x;
}

// Dynamic shared memory: the thread-private cache of the level's 3-D points, float[3][pcap][kThreads] (x, y, z planes):
// thread t keeps the first `pcap` of ITS points of the current level there for all evaluations of the level, so an
// evaluation starts with shared-memory reads instead of an L2 round trip and re-reads no list bytes from L2 / HBM.
template <int kThreads, int kMinBlocks, int kHint, bool kPack>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_track_lean(const PairDesc *__restrict__ pairs, int n_pairs, const TrackParams prm, revo_track_result *__restrict__ results,
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
                // this thread's points: first_idx, first_idx + stride, ... (the count differs by at most one over the cluster)
                const int my_iter = first_idx < n ? (n - first_idx + stride - 1) / stride : 0;
                const int my_cached = my_iter < pcap ? my_iter : pcap;
                const float4 *__restrict__ pts = Lin.pts;
                LevelConst L;
                L.fx = Lin.fx; L.fy = Lin.fy; L.cx = Lin.cx; L.cy = Lin.cy;
                L.umax = pin((float)(Lin.w - 2)); L.vmax = pin((float)(Lin.h - 2)); L.w = Lin.w; L.opt = Lin.opt;
                const float ed_eff = pin(use_filter ? oc.edge_distance_lvl[lvl] : INFINITY);
                const float huber = pin(oc.huber_edge);
                const float kqfx = pin(Lin.fx * (1.0f / 32764.0f)), kqfy = pin(Lin.fy * (1.0f / 32764.0f));
                // this thread's points of the level -> its private slots of the shared-memory cache
                for (int k = 0; k < my_cached; ++k) {
                    const float4 p = __ldg(pts + first_idx + (size_t)k * stride);
                    sts3<kThreads>(s_base + (uint32_t)k * kPtStride, p.x, p.y, p.z);
                }
                bool first = true;
                __syncthreads();
                while (true) {
                    float R[9], t[3];
#pragma unroll
                    for (int i = 0; i < 9; ++i) R[i] = ctrl.R[i];
#pragma unroll
                    for (int i = 0; i < 3; ++i) t[i] = ctrl.t[i];
                    float acc[32];
                    PackedAcc S;
                    if (kPack) {
                        S.clear();
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                    }
                    const long long c_begin = prm.profile ? clock64() : 0;
                    // Two pipelined segments (cached points, then the uncached tail of a long level), each with two
                    // register sets (A/B): while point k is being finished the 256-bit gather of point k+1 is in flight.
                    auto segment = [&](auto from_smem, int k0, int k1) {
                        constexpr bool kS = decltype(from_smem)::value;
                        if (k0 >= k1) return;
                        uint32_t sp = s_base + (uint32_t)k0 * kPtStride;
                        const float4 *gp = pts + first_idx + (size_t)k0 * stride;
                        auto arm = [&](ProjB &P, uint4 &q0, uint4 &q1) {
                            float x, y, z;
                            if constexpr (kS) {
                                lds3<kThreads>(sp, x, y, z);
                                sp += kPtStride;
                            } else {
                                const float4 p = __ldg(gp);
                                gp += stride;
                                x = p.x; y = p.y; z = p.z;
                            }
                            P = project_l(x, y, z, L, R, t);
                            ldg_quad_h<kHint>(P.bp, q0, q1);
                        };
                        ProjB A, B;
                        uint4 a0, a1, b0, b1;
                        arm(A, a0, a1);
                        int left = k1 - k0 - 1;   // points of the segment not yet armed
                        while (true) {
                            if (left > 0) arm(B, b0, b1);
                            if (kPack) finish_point_p(A, a0, a1, kqfx, kqfy, ed_eff, huber, S); else finish_point_l(A, a0, a1, kqfx, kqfy, ed_eff, huber, acc);
                            if (left <= 0) break;
                            if (left > 1) arm(A, a0, a1);
                            if (kPack) finish_point_p(B, b0, b1, kqfx, kqfy, ed_eff, huber, S); else finish_point_l(B, b0, b1, kqfx, kqfy, ed_eff, huber, acc);
                            if (left <= 1) break;
                            left -= 2;
                        }
                    };
                    segment(std::true_type{}, 0, my_cached);
                    segment(std::false_type{}, my_cached, my_iter);
                    if (kPack) S.unpack(acc, (float)my_iter);
                    else acc[kRecBad] = (float)my_iter - acc[kRecGood];   // every visited point exists
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
template <int kThreads, int kMinBlocks, int kHint, bool kPack>
static int launch_lean_t(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, int ctas_per_pair,
                         revo_track_result *d_results, double *d_records, revo_trace_entry *d_trace, int *d_trace_counts,
                         int *d_work_counter)
{
    auto kern = k_track_lean<kThreads, kMinBlocks, kHint, kPack>;
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

// Shape like launch_track (track.cu); REVO_LEAN_HINT: 0 plain gather, 3 ld.global.nc.L1::no_allocate (default 0);
// REVO_LEAN_PACK=1: packed fma.rn.f32x2 accumulation (finish_point_p).
int launch_track_lean(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                      double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, int *d_work_counter)
{
    if (n_pairs <= 0) return REVO_OK;
    if (prm.split_world > 1) {
        ctx->last_error = "the lean tracking engine does not support the multi-GPU edge split";
        return REVO_ERR_INVALID_ARG;
    }
    const int slots256 = 2 * ctx->prop.multiProcessorCount;
    const int C = ctx->track_ctas_per_pair > 0 ? ctx->track_ctas_per_pair : 8;
    const int T = ctx->track_threads > 0 ? ctx->track_threads : ((long long)n_pairs * C > slots256 ? 128 : 256);
    const int hint = getenv("REVO_LEAN_HINT") ? atoi(getenv("REVO_LEAN_HINT")) : 0;
#define REVO_LEAN_ARGS ctx, d_pairs, n_pairs, prm, C, d_results, d_records, d_trace, d_trace_counts, d_work_counter
    const bool pack = getenv("REVO_LEAN_PACK") ? atoi(getenv("REVO_LEAN_PACK")) != 0 : false;
#define REVO_LEAN_SHAPES(HH, PP)                                                  \
    switch (T) {                                                                  \
        case 128: return launch_lean_t<128, 4, HH, PP>(REVO_LEAN_ARGS);           \
        case 512: return launch_lean_t<512, 1, HH, PP>(REVO_LEAN_ARGS);           \
        default: return launch_lean_t<256, 2, HH, PP>(REVO_LEAN_ARGS);            \
    }
    if (hint == 3 && pack) { REVO_LEAN_SHAPES(3, true) }
    if (hint == 3) { REVO_LEAN_SHAPES(3, false) }
    if (pack) { REVO_LEAN_SHAPES(0, true) }
    REVO_LEAN_SHAPES(0, false)
#undef REVO_LEAN_SHAPES
#undef REVO_LEAN_ARGS
}

}  // namespace revo
