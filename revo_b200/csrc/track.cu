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
// Design (B200): a frame pair is owned by one thread-block CLUSTER (1..16 CTAs); clusters pull pairs from a
// global work counter (persistent kernel).  PASS A and PASS B are fused: every evaluation at a pose warps each
// 3-D edge point, fetches the 4 {gx,gy,dt} texels, forms the residual, Huber weight and 1x6 Jacobian and
// accumulates the 21+6 normal-equation terms + 4 statistics in registers -- the 7 SoA buffers of the reference
// never exist.  Two points are in flight per thread (their 8 texel gathers are issued back to back) to cover the
// dependent pts -> texel latency.  The 32-value record is reduced with a transposing warp-shuffle tree, across
// warps through shared memory, across the CTAs of the cluster through distributed shared memory (one cluster
// barrier per evaluation), and -- when one pair is split over several GPUs -- across GPUs through peer-mapped
// mailboxes over NVLink inside the same kernel.  Every CTA then runs the identical 6x6 LDL^T solve, SE3 update
// and accept/reject test redundantly (bitwise-equal inputs, so no broadcast is needed): all levels and all LM
// iterations of a pair run without a host round trip.  No tensor cores: there is no dense contraction here.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "internal.h"

namespace cg = cooperative_groups;

namespace revo {

constexpr unsigned kFull = 0xffffffffu;

// ---- record layout ---------------------------------------------------------
// [0..20] sum w v_i v_j (i<=j, LGS6 slot order), [21..26] sum w r v_i, [27] sum w r^2, [28] sum r^2,
// [29] good, [30] bad, [31] unused.
constexpr int kRecA = 0, kRecB = 21, kRecSW = 27, kRecSU = 28, kRecGood = 29, kRecBad = 30;

struct Ctrl {
    // written by thread 0 of every CTA (identically), read by all threads
    float R[9];
    float t[3];
    int level_done;
    int pair_skip;
    int next_pair;
};

struct LMState {
    double q[4], t[3];    // accepted pose (Sophus SE3: unit quaternion xyzw + translation)
    double qn[4], tn[3];  // trial pose
    double A[21], b[6], n;
    double inc[6];
    float lastErr, last_residual, lambda;
    int iteration, incTry, tries;
};

// ---- small double-precision SE3 / solver helpers (thread 0 only) --------------
__device__ __forceinline__ void quat_to_R(const double *q, double *R /* col-major */)
{
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[3] = txy - twz;       R[6] = txz + twy;
    R[1] = txy + twz;       R[4] = 1 - (txx + tzz); R[7] = tyz - twx;
    R[2] = txz - twy;       R[5] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen quaternion-from-matrix (Shepperd), as SO3(Matrix3) does (so3.hpp:419). R col-major float.
__device__ void quat_from_R(const float *Rf, double *q)
{
    double R[9];
    for (int i = 0; i < 9; ++i) R[i] = Rf[i];
#define RMAT(i, j) R[(j) * 3 + (i)]
    double t = RMAT(0, 0) + RMAT(1, 1) + RMAT(2, 2);
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (RMAT(2, 1) - RMAT(1, 2)) * t;
        q[1] = (RMAT(0, 2) - RMAT(2, 0)) * t;
        q[2] = (RMAT(1, 0) - RMAT(0, 1)) * t;
    } else {
        int i = 0;
        if (RMAT(1, 1) > RMAT(0, 0)) i = 1;
        if (RMAT(2, 2) > RMAT(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(RMAT(i, i) - RMAT(j, j) - RMAT(k, k) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (RMAT(k, j) - RMAT(j, k)) * t;
        q[j] = (RMAT(j, i) + RMAT(i, j)) * t;
        q[k] = (RMAT(k, i) + RMAT(i, k)) * t;
    }
#undef RMAT
}

// ||R R^T - I||_F < 1e-5 and det > 0: the Sophus ENSUREs of so3.hpp:419-424 (float epsilon, common.hpp:152).
__device__ bool rotation_ok(const float *Rf)
{
    double n2 = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)Rf[k * 3 + i] * (double)Rf[k * 3 + j];
            s -= (i == j) ? 1.0 : 0.0;
            n2 += s * s;
        }
    const double det = (double)Rf[0] * ((double)Rf[4] * Rf[8] - (double)Rf[7] * Rf[5]) -
                       (double)Rf[3] * ((double)Rf[1] * Rf[8] - (double)Rf[7] * Rf[2]) +
                       (double)Rf[6] * ((double)Rf[1] * Rf[5] - (double)Rf[4] * Rf[2]);
    return (sqrt(n2) < 1e-5) && (det > 0);
}

// Sophus::SE3::exp (se3.hpp:723-748, so3.hpp:531-564) in double; one sincos: sin t = 2 s c, 1 - cos t = 2 s^2.
__device__ __forceinline__ void se3_exp(const double *xi, double *q, double *t)
{
    const double ox = xi[3], oy = xi[4], oz = xi[5];
    const double theta_sq = ox * ox + oy * oy + oz * oz;
    const double theta = sqrt(theta_sq);
    double imag, re, c1, c2;
    if (theta < 1e-5) {   // Sophus::Constants<float>::epsilon()
        const double t4 = theta_sq * theta_sq;
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * t4;
        re = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * t4;
        // V = R(q) there (se3.hpp:735-737) = I + 2 re imag Om + 2 imag^2 Om^2
        c1 = 2.0 * re * imag;
        c2 = 2.0 * imag * imag;
    } else {
        double s, c;
        sincos(0.5 * theta, &s, &c);
        const double inv_t = __drcp_rn(theta), inv_t2 = inv_t * inv_t;
        imag = s * inv_t;
        re = c;
        c1 = 2.0 * s * s * inv_t2;                       // (1 - cos t) / t^2
        c2 = (theta - 2.0 * s * c) * inv_t2 * inv_t;     // (t - sin t) / t^3
    }
    q[0] = imag * ox; q[1] = imag * oy; q[2] = imag * oz; q[3] = re;
    // V = I + c1 Om + c2 Om^2 ; Om = hat(omega), Om^2 = omega omega^T - |omega|^2 I
    const double v00 = 1 + c2 * (ox * ox - theta_sq), v01 = -c1 * oz + c2 * ox * oy, v02 = c1 * oy + c2 * ox * oz;
    const double v10 = c1 * oz + c2 * ox * oy, v11 = 1 + c2 * (oy * oy - theta_sq), v12 = -c1 * ox + c2 * oy * oz;
    const double v20 = -c1 * oy + c2 * ox * oz, v21 = c1 * ox + c2 * oy * oz, v22 = 1 + c2 * (oz * oz - theta_sq);
    t[0] = v00 * xi[0] + v01 * xi[1] + v02 * xi[2];
    t[1] = v10 * xi[0] + v11 * xi[1] + v12 * xi[2];
    t[2] = v20 * xi[0] + v21 * xi[1] + v22 * xi[2];
}

// (qa,ta) * (qb,tb) with Sophus' renormalisation (se3.hpp:317-321, so3.hpp:335-352)
__device__ __forceinline__ void se3_mul(const double *qa, const double *ta, const double *qb, const double *tb, double *q, double *t)
{
    double ux = qa[1] * tb[2] - qa[2] * tb[1], uy = qa[2] * tb[0] - qa[0] * tb[2], uz = qa[0] * tb[1] - qa[1] * tb[0];
    ux += ux; uy += uy; uz += uz;
    const double cx = qa[1] * uz - qa[2] * uy, cy = qa[2] * ux - qa[0] * uz, cz = qa[0] * uy - qa[1] * ux;
    t[0] = ta[0] + (tb[0] + qa[3] * ux + cx);
    t[1] = ta[1] + (tb[1] + qa[3] * uy + cy);
    t[2] = ta[2] + (tb[2] + qa[3] * uz + cz);
    const double ax = qa[0], ay = qa[1], az = qa[2], aw = qa[3], bx = qb[0], by = qb[1], bz = qb[2], bw = qb[3];
    double w = aw * bw - ax * bx - ay * by - az * bz;
    double x = aw * bx + ax * bw + ay * bz - az * by;
    double y = aw * by + ay * bw + az * bx - ax * bz;
    double z = aw * bz + az * bw + ax * by - ay * bx;
    const double sn = x * x + y * y + z * z + w * w;
    if (sn != 1.0) {
        const double s = 2.0 * __drcp_rn(1.0 + sn);
        x *= s; y *= s; z *= s; w *= s;
    }
    q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

// Solve (A/n with diag * lam1) x = b/n for the symmetric positive (semi-)definite 6x6 normal equations
// (system/optimizer.cpp:258-262, "A.ldlt().solve(b)").  LDL^T in double, fully unrolled so that everything
// stays in registers; no pivoting (the matrix is a damped sum of outer products; Eigen's diagonal pivoting
// only changes rounding, which double precision makes irrelevant at the float tolerance of this path).
// Non-positive / non-finite pivots are treated like Eigen's pseudo-inverse of D: that component becomes 0.
__device__ __forceinline__ void solve6(const double *Au /* 21 upper slots */, const double *b, double inv_n, double lam1, double *x)
{
    double a[6][6];
    {
        int s = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) a[j][i] = Au[s++] * inv_n;   // lower triangle
    }
    double y[6], invd[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { a[i][i] *= lam1; y[i] = b[i] * inv_n; }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double dk = a[k][k];
        const double id = (dk > 0.0 && dk < 1e300) ? __drcp_rn(dk) : 0.0;
        invd[k] = id;
#pragma unroll
        for (int j = k + 1; j < 6; ++j) {
            const double ljk = a[j][k] * id;
#pragma unroll
            for (int i = j; i < 6; ++i) a[i][j] -= a[i][k] * ljk;
        }
#pragma unroll
        for (int i = k + 1; i < 6; ++i) a[i][k] *= id;   // L
    }
#pragma unroll
    for (int i = 1; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < i; ++j) y[i] -= a[i][j] * y[j];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] *= invd[i];
#pragma unroll
    for (int i = 4; i >= 0; --i)
#pragma unroll
        for (int j = i + 1; j < 6; ++j) y[i] -= a[j][i] * y[j];
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = y[i];
}

// ---- per-point work: PASS A + PASS B fused ---------------------------------------
struct Proj {
    float Wx, Wy, iz, dx, dy;
    const uint4 *bp;
    int state;   // 0 = no point, 1 = in bounds (texels wanted), 2 = out of bounds
};

// optimizer.cpp:93-100: warp, project, bounds test
__device__ __forceinline__ Proj project(bool exists, const float4 p, const LevelIn &L, const float *__restrict__ R,
                                        const float *__restrict__ t)
{
    Proj o;
    o.Wx = R[0] * p.x + R[3] * p.y + R[6] * p.z + t[0];
    o.Wy = R[1] * p.x + R[4] * p.y + R[7] * p.z + t[1];
    const float Wz = R[2] * p.x + R[5] * p.y + R[8] * p.z + t[2];
    // the reference divides (Wx/Wz*fx+cx); one correctly rounded reciprocal is shared by u, v and the Jacobian
    // (differs from the quotient by <= 1 ulp, far inside the float tolerance of this path)
    o.iz = __frcp_rn(Wz);
    const float u = o.Wx * o.iz * L.fx + L.cx;
    const float v = o.Wy * o.iz * L.fy + L.cy;
    const bool inb = (u > 1.f && v > 1.f && u < (float)(L.w - 2) && v < (float)(L.h - 2));   // NaN-safe (:100)
    const int ix = inb ? (int)u : 0, iy = inb ? (int)v : 0;
    o.dx = u - (float)ix;
    o.dy = v - (float)iy;
    o.bp = L.opt + (unsigned)(iy * L.w + ix);
    o.state = exists ? (inb ? 1 : 2) : 0;
    return o;
}

// snorm16 pair -> floats (scale folded in by the caller)
__device__ __forceinline__ void unpack_grad(uint32_t g, float &gx, float &gy)
{
    gx = (float)(short)(g & 0xffffu);
    gy = (float)((int)g >> 16);
}

// r0 = pair record of row iy (texels (ix,iy),(ix+1,iy)), r1 = pair record of row iy+1
__device__ __forceinline__ void finish_point(const Proj &P, const uint4 r0, const uint4 r1, const LevelIn &L, float edge_dist,
                                             bool use_filter, float huber, float (&acc)[32])
{
    if (P.state == 0) return;
    if (P.state == 2) { acc[kRecBad] += 1.f; return; }
    // getInterpolatedElement43, optimizer.h:173-185
    const float dxdy = P.dx * P.dy;
    const float w11 = dxdy, w01 = P.dy - dxdy, w10 = P.dx - dxdy, w00 = 1.f - P.dx - P.dy + dxdy;
    float gx00, gy00, gx10, gy10, gx01, gy01, gx11, gy11;
    unpack_grad(r0.z, gx00, gy00); unpack_grad(r0.w, gx10, gy10);
    unpack_grad(r1.z, gx01, gy01); unpack_grad(r1.w, gx11, gy11);
    constexpr float kq = 1.0f / 32764.0f;
    const float gxi = (w11 * gx11 + w01 * gx01 + w10 * gx10 + w00 * gx00) * kq;
    const float gyi = (w11 * gy11 + w01 * gy01 + w10 * gy10 + w00 * gy00) * kq;
    const float r = w11 * __uint_as_float(r1.y) + w01 * __uint_as_float(r1.x) + w10 * __uint_as_float(r0.y) + w00 * __uint_as_float(r0.x);
    if (use_filter && r > edge_dist) {                                             // optimizer.cpp:112
        acc[kRecBad] += 1.f;
        return;
    }
    const float wr = (r <= huber) ? 1.f : __fdividef(huber, r);                    // optimizer.h:159
    const float gx = L.fx * gxi, gy = L.fy * gyi;                                  // optimizer.cpp:119-120
    // calculateWarpUpdate, optimizer.cpp:204-228
    // Same six entries, factored through a = x/z, b = y/z, t = a gx + b gy (12 flops instead of ~30):
    //   v2 = -(a gx + b gy)/z, v3 = -(b t + gy), v4 = a t + gx, v5 = a gy - b gx.
    const float z = P.iz;
    const float a = P.Wx * z, b = P.Wy * z;
    const float t = a * gx + b * gy;
    float J[6];
    J[0] = z * gx;
    J[1] = z * gy;
    J[2] = -(t * z);
    J[3] = -(b * t + gy);
    J[4] = a * t + gx;
    J[5] = a * gy - b * gx;
    // LGS6::update, LGSX.h:392-398 (upper triangle only; A is symmetric)
    int s = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float wi = wr * J[i];
#pragma unroll
        for (int j = i; j < 6; ++j) acc[s++] += wi * J[j];
    }
    const float rw = r * wr;
#pragma unroll
    for (int i = 0; i < 6; ++i) acc[kRecB + i] += rw * J[i];
    acc[kRecSW] += rw * r;     // optimizer.cpp:131
    acc[kRecSU] += r * r;
    acc[kRecGood] += 1.f;
}

// evalCostFunction (tracker.cpp:357-393) for one pose
__device__ __forceinline__ float cost_point(float X, float Y, float Z, const LevelIn &L, const float *__restrict__ dt, float edge_dist,
                                            bool use_filter)
{
    const float nx = L.fx * X / Z + L.cx;    // tracker.cpp:378-379
    const float ny = L.fy * Y / Z + L.cy;
    if (nx >= 0.f && nx < (float)L.w && ny >= 0.f && ny < (float)L.h) {
        const float r = __ldg(dt + (size_t)floorf(ny) * L.w + (size_t)floorf(nx));
        if (use_filter && r > edge_dist) return 0.f;
        return r;
    }
    return 0.f;
}

// After the call lane L holds the warp total of v[L].
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane)
{
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool hi = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = hi ? v[i] : v[i + half];
            const float keep = hi ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, half);
        }
    }
    return v[0];
}

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


// ---- the kernel ---------------------------------------------------------------
template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_track(const PairDesc *__restrict__ pairs, int n_pairs, const TrackParams prm, revo_track_result *__restrict__ results,
        double *__restrict__ records, revo_trace_entry *__restrict__ trace, int *__restrict__ trace_counts,
        int *__restrict__ work_counter)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int kWarps = kThreads / 32;

    __shared__ float warp_part[kWarps][32];
    __shared__ double cta_part[2][32];   // read by the other CTAs of the cluster (DSMEM)
    __shared__ double total[2][32];      // split mode: CTA 0 publishes the cross-GPU total here
    __shared__ double rec[32];
    __shared__ Ctrl ctrl;
    __shared__ LMState lm;

    const revo_opt_config &oc = prm.cfg.opt;
    const bool use_filter = oc.use_edge_filter != 0;
    const int world = prm.split_world > 1 ? prm.split_world : 1;
    const int n_members = world * C;
    const int member = (world > 1 ? prm.split_rank : 0) * C + crank;
    unsigned seq = 0;   // evaluation counter of this cluster (drives the double buffers)

    // ---- reduction of a per-thread accumulator to `rec` (identical in every CTA of the cluster / every rank)
    auto reduce_record = [&](float (&acc)[32]) {
        const float mine = warp_transpose_reduce(acc, lane);
        warp_part[wid][lane] = mine;
        __syncthreads();
        const int par = seq & 1;
        if (tid < 32) {
            double s = 0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += (double)warp_part[w][tid];
            cta_part[par][tid] = s;
        }
        if (C > 1 || world > 1) cluster.sync(); else __syncthreads();
        if (world == 1) {
            if (tid < 32) {
                double s = cta_part[par][tid];
                if (C > 1) {
                    s = 0;
                    for (int r = 0; r < C; ++r) s += *cluster.map_shared_rank(&cta_part[par][tid], r);
                }
                rec[tid] = s;
            }
        } else {
            // cross-GPU exchange: CTA 0 of each rank pushes the rank partial into every rank's mailbox,
            // then waits for all `world` partials of this evaluation and sums them in rank order.
            const unsigned long long fl = prm.split_seq0 + (unsigned long long)seq + 1ull;
            if (crank == 0 && tid < 32) {
                double s = 0;
                for (int r = 0; r < C; ++r) s += *cluster.map_shared_rank(&cta_part[par][tid], r);
                for (int g = 0; g < world; ++g) {
                    Mailbox *mb = (Mailbox *)prm.split_peers[g];
                    mb->data[par][prm.split_rank][tid] = s;
                }
                __threadfence_system();
                __syncwarp();
                if (tid < world) st_release_sys(&((Mailbox *)prm.split_peers[tid])->flag[par][prm.split_rank], fl);
                Mailbox *mine_mb = (Mailbox *)prm.split_peers[prm.split_rank];
                if (tid < world) {
                    while (ld_acquire_sys(&mine_mb->flag[par][tid]) < fl) { }
                }
                __syncwarp();
                double tot = 0;
                for (int g = 0; g < world; ++g) tot += ((volatile double *)mine_mb->data[par][g])[tid];
                total[par][tid] = tot;
            }
            cluster.sync();
            if (tid < 32) rec[tid] = *cluster.map_shared_rank(&total[par][tid], 0);
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
                const LevelIn L = P.lvl[lvl];
                const int n = *L.n_pts;
                // block-cyclic split of the list over the CTAs of the cluster (and the ranks of a GPU split): member m takes
                // the blocks m, m + M, m + 2M, ... of kThreads points -- balanced (the cluster barrier waits for the slowest
                // CTA) and the cluster as a whole still sweeps the tile-major list front to back
                const int lo = 0, hi = n;
                const int stride = n_members * kThreads;
                const float ed = oc.edge_distance_lvl[lvl];
                const float huber = oc.huber_edge;
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
                    // Software pipeline, two register sets (A/B): while point k is being finished, the four texel
                    // gathers of point k+1 and the list entry of point k+2 are already in flight.
                    {
                        const float4 *__restrict__ pts = L.pts;
                        int i = member * kThreads + tid;
                        bool eA = i < hi;
                        float4 pA = __ldg(pts + (eA ? i : lo));
                        i += stride;
                        bool eB = i < hi;
                        float4 pB = __ldg(pts + (eB ? i : lo));
                        Proj A = project(eA, pA, L, R, t);
                        uint4 a0 = __ldg(A.bp), a1 = __ldg(A.bp + L.w);
                        while (true) {
                            i += stride;
                            const bool eC = i < hi;
                            const float4 pC = __ldg(pts + (eC ? i : lo));
                            const Proj B = project(eB, pB, L, R, t);
                            const uint4 b0 = __ldg(B.bp), b1 = __ldg(B.bp + L.w);
                            finish_point(A, a0, a1, L, ed, use_filter, huber, acc);
                            if (!eB) break;
                            i += stride;
                            const bool eD = i < hi;
                            const float4 pD = __ldg(pts + (eD ? i : lo));
                            A = project(eC, pC, L, R, t);
                            a0 = __ldg(A.bp); a1 = __ldg(A.bp + L.w);
                            finish_point(B, b0, b1, L, ed, use_filter, huber, acc);
                            if (!eC) break;
                            eB = eD;
                            pB = pD;
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
                        // ---------------- Optimizer::trackFrames LM logic, optimizer.cpp:243-306 ----------------
                        const float err = (float)(rec[kRecSW] / rec[kRecGood]);    // :190
                        bool propose = false, done = false;
                        if (first) {
                            lm.lastErr = err;
                            lm.last_residual = err;
                            lm.lambda = oc.lambda_initial[lvl];
                            lm.iteration = 0; lm.incTry = 0; lm.tries = 0;
                            for (int i = 0; i < 21; ++i) lm.A[i] = rec[kRecA + i];
                            for (int i = 0; i < 6; ++i) lm.b[i] = rec[kRecB + i];
                            lm.n = rec[kRecGood];
                            propose = true;
                        } else {
                            const bool accepted = err < lm.lastErr;                // :273
                            if (trace && crank == 0 && ntrace < prm.trace_cap) {
                                revo_trace_entry &e = trace[(size_t)pair * prm.trace_cap + ntrace];
                                e.error = err; e.lambda = lm.lambda; e.accepted = accepted ? 1 : 0;
                                e.good = (int)rec[kRecGood]; e.bad = (int)rec[kRecBad]; e.level = lvl;
                            }
                            ntrace++;
                            if (accepted) {
                                for (int i = 0; i < 4; ++i) lm.q[i] = lm.qn[i];
                                for (int i = 0; i < 3; ++i) lm.t[i] = lm.tn[i];
                                for (int i = 0; i < 21; ++i) lm.A[i] = rec[kRecA + i];
                                for (int i = 0; i < 6; ++i) lm.b[i] = rec[kRecB + i];
                                lm.n = rec[kRecGood];
                                if (err / lm.lastErr > oc.convergence_eps[lvl]) lm.iteration = oc.max_its_per_lvl[lvl];   // :279-283
                                lm.last_residual = lm.lastErr = err;
                                if (lm.lambda <= 0.2f) lm.lambda = 0.f; else lm.lambda *= oc.lambda_success_fac;          // :286-289
                                lm.iteration++;     // for-loop increment after the break (:291)
                                lm.incTry = 0;
                                propose = true;
                            } else {
                                double dot = 0;
                                for (int i = 0; i < 6; ++i) dot += lm.inc[i] * lm.inc[i];
                                if (!((float)dot > oc.step_size_min[lvl])) {                                               // :294
                                    done = true;
                                } else {
                                    if (lm.lambda == 0.f) lm.lambda = 0.2f;                                                // :300-303
                                    else lm.lambda *= powf(oc.lambda_fail_fac, (float)lm.incTry);
                                    propose = true;
                                }
                            }
                        }
                        if (propose && !done) {
                            if (lm.iteration >= oc.max_its_per_lvl[lvl]) done = true;
                            else if (oc.max_lm_tries > 0 && lm.tries >= oc.max_lm_tries) done = true;
                        }
                        if (propose && !done) {
                            // solve (A/n with diag *(1+lambda)) inc = (sum w r v)/n     :258-262
                            solve6(lm.A, lm.b, __drcp_rn(lm.n), (double)(1.f + lm.lambda), lm.inc);
                            lm.incTry++; lm.tries++;
                            double qe[4], te[3];
                            se3_exp(lm.inc, qe, te);
                            se3_mul(qe, te, lm.q, lm.t, lm.qn, lm.tn);              // :266 exp(inc) * referenceToFrame
                            double Rn[9];
                            quat_to_R(lm.qn, Rn);
                            for (int i = 0; i < 9; ++i) ctrl.R[i] = (float)Rn[i];
                            for (int i = 0; i < 3; ++i) ctrl.t[i] = (float)lm.tn[i];
                        }
                        if (done) {
                            // next level (or the result) starts from the accepted pose      :308-309
                            double Ra[9];
                            quat_to_R(lm.q, Ra);
                            for (int i = 0; i < 9; ++i) ctrl.R[i] = (float)Ra[i];
                            for (int i = 0; i < 3; ++i) ctrl.t[i] = (float)lm.t[i];
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
    if (C > 1 || world > 1) cluster.sync();   // nobody may exit while a peer can still read its shared memory
}

// ---- launcher -------------------------------------------------------------------
template <int kThreads, int kMinBlocks>
static int launch_track_t(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, int ctas_per_pair,
                          revo_track_result *d_results, double *d_records, revo_trace_entry *d_trace, int *d_trace_counts,
                          int *d_work_counter)
{
    auto kern = k_track<kThreads, kMinBlocks>;
    if (ctas_per_pair > 8) REVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas_per_pair;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
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
    const int n_clusters = n_pairs < max_clusters ? n_pairs : max_clusters;
    cfg.gridDim = dim3(n_clusters * ctas_per_pair);
    REVO_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, d_pairs, n_pairs, prm, d_results, d_records, d_trace, d_trace_counts,
                                      d_work_counter));
    ctx->launches++;
    return REVO_OK;
}

int launch_track(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                 double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, int *d_work_counter)
{
    if (n_pairs <= 0) return REVO_OK;
    const int T = ctx->track_threads > 0 ? ctx->track_threads : 256;
    // register budget: "dense" = 85 registers/thread (3 CTAs of 256 or 6 of 128 per SM) instead of 128
    static const int dense_mode = getenv("REVO_TRACK_DENSE") ? atoi(getenv("REVO_TRACK_DENSE")) : 0;
    const bool dense = dense_mode == 1;
    int C = ctx->track_ctas_per_pair;
    if (C <= 0) {
        // automatic: fill the CTA slots of the chip (SMs x resident CTAs of this shape); a pair alone gets a
        // full portable cluster
        const int per_sm = (T <= 128 ? 4 : (T <= 256 ? 2 : 1)) + (dense && T <= 256 ? (T <= 128 ? 2 : 1) : 0) +
                           (dense_mode == 2 && T <= 128 ? 1 : 0);
        const int slots = ctx->prop.multiProcessorCount * per_sm;
        C = 1;
        while (C < 8 && n_pairs * (C * 2) <= slots) C *= 2;
    }
#define REVO_TRACK_ARGS ctx, d_pairs, n_pairs, prm, C, d_results, d_records, d_trace, d_trace_counts, d_work_counter
    switch (T) {
        case 128:
            if (dense_mode == 2) return launch_track_t<128, 5>(REVO_TRACK_ARGS);   // 102 registers, 5 CTAs / SM
            return dense ? launch_track_t<128, 6>(REVO_TRACK_ARGS) : launch_track_t<128, 4>(REVO_TRACK_ARGS);
        case 512: return launch_track_t<512, 1>(REVO_TRACK_ARGS);
        case 1024: return launch_track_t<1024, 1>(REVO_TRACK_ARGS);
        default: return dense ? launch_track_t<256, 3>(REVO_TRACK_ARGS) : launch_track_t<256, 2>(REVO_TRACK_ARGS);
    }
#undef REVO_TRACK_ARGS
}

}  // namespace revo
