// track_common.cuh -- device helpers shared by the two tracking engines (track.cu: one cluster per pair;
// track_queue.cu: chip-wide task queue): record layout, SE3 / 6x6 solver in double, the fused PASS A + PASS B
// per-point work and the transposing warp reduction.
//
// Reference (fabianschenk/REVO): system/optimizer.cpp:74-311, system/optimizer.h:156-185, utils/LGSX.h:196-398,
// thirdparty/Sophus/sophus/se3.hpp:317-321,723-748, so3.hpp:335-352,419-424,531-564, system/tracker.cpp:357-393.
#pragma once
#include <math.h>

#include "internal.h"

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
__device__ inline void quat_from_R(const float *Rf, double *q)
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
__device__ inline bool rotation_ok(const float *Rf)
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

// Sophus::SE3::exp (se3.hpp:723-748, so3.hpp:531-564) in double.  The four coefficients sin(t/2)/t, cos(t/2),
// (1 - cos t)/t^2 and (t - sin t)/t^3 are even functions of t; for the increments of a tracker (t < 0.5 rad, in practice
// < 0.05) they are evaluated as power series in t^2 (8 terms: truncation < 1e-17 relative) as four independent Horner
// chains: no sqrt, sincos or division on the serial critical path of an evaluation.  Larger angles take the closed form.
__device__ __forceinline__ void se3_exp(const double *xi, double *q, double *t)
{
    const double ox = xi[3], oy = xi[4], oz = xi[5];
    const double s = ox * ox + oy * oy + oz * oz;   // theta^2
    double imag, re, c1, c2;
    if (s < 1e-10) {   // theta < Sophus::Constants<float>::epsilon() = 1e-5
        const double t4 = s * s;
        imag = 0.5 - (1.0 / 48.0) * s + (1.0 / 3840.0) * t4;
        re = 1.0 - (1.0 / 8.0) * s + (1.0 / 384.0) * t4;
        // V = R(q) there (se3.hpp:735-737) = I + 2 re imag Om + 2 imag^2 Om^2
        c1 = 2.0 * re * imag;
        c2 = 2.0 * imag * imag;
    } else if (s < 0.25) {
        // coefficients: 1/(2^(2k+1) (2k+1)!), 1/(4^k (2k)!), 1/(2k+2)!, 1/(2k+3)!  with alternating sign
        imag = 1.0 / 42849873690624000.0;
        re = 1.0 / 1428329123020800.0;
        c1 = 1.0 / 20922789888000.0;
        c2 = 1.0 / 355687428096000.0;
        imag = imag * -s + 1.0 / 51011754393600.0;     re = re * -s + 1.0 / 1961990553600.0;
        c1 = c1 * -s + 1.0 / 87178291200.0;             c2 = c2 * -s + 1.0 / 1307674368000.0;
        imag = imag * -s + 1.0 / 81749606400.0;         re = re * -s + 1.0 / 3715891200.0;
        c1 = c1 * -s + 1.0 / 479001600.0;               c2 = c2 * -s + 1.0 / 6227020800.0;
        imag = imag * -s + 1.0 / 185794560.0;           re = re * -s + 1.0 / 10321920.0;
        c1 = c1 * -s + 1.0 / 3628800.0;                 c2 = c2 * -s + 1.0 / 39916800.0;
        imag = imag * -s + 1.0 / 645120.0;              re = re * -s + 1.0 / 46080.0;
        c1 = c1 * -s + 1.0 / 40320.0;                   c2 = c2 * -s + 1.0 / 362880.0;
        imag = imag * -s + 1.0 / 3840.0;                re = re * -s + 1.0 / 384.0;
        c1 = c1 * -s + 1.0 / 720.0;                     c2 = c2 * -s + 1.0 / 5040.0;
        imag = imag * -s + 1.0 / 48.0;                  re = re * -s + 1.0 / 8.0;
        c1 = c1 * -s + 1.0 / 24.0;                      c2 = c2 * -s + 1.0 / 120.0;
        imag = imag * -s + 0.5;                         re = re * -s + 1.0;
        c1 = c1 * -s + 0.5;                             c2 = c2 * -s + 1.0 / 6.0;
    } else {
        const double theta = sqrt(s);
        double sn, cs;
        sincos(0.5 * theta, &sn, &cs);
        const double inv_t = __drcp_rn(theta), inv_t2 = inv_t * inv_t;
        imag = sn * inv_t;
        re = cs;
        c1 = 2.0 * sn * sn * inv_t2;                        // (1 - cos t) / t^2
        c2 = (theta - 2.0 * sn * cs) * inv_t2 * inv_t;      // (t - sin t) / t^3
    }
    q[0] = imag * ox; q[1] = imag * oy; q[2] = imag * oz; q[3] = re;
    // V = I + c1 Om + c2 Om^2 ; Om = hat(omega), Om^2 = omega omega^T - |omega|^2 I
    const double v00 = 1 + c2 * (ox * ox - s), v01 = -c1 * oz + c2 * ox * oy, v02 = c1 * oy + c2 * ox * oz;
    const double v10 = c1 * oz + c2 * ox * oy, v11 = 1 + c2 * (oy * oy - s), v12 = -c1 * ox + c2 * oy * oz;
    const double v20 = -c1 * oy + c2 * ox * oz, v21 = c1 * ox + c2 * oy * oz, v22 = 1 + c2 * (oz * oz - s);
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

// ---- one step of the Levenberg-Marquardt state machine (thread-serial) ------------------------------------------
// Optimizer::trackFrames, system/optimizer.cpp:243-306, restated as "consume the record of the evaluation that just
// finished, decide, and name the next pose to evaluate".  `first`: the record was taken at the level's start pose
// (optimizer.cpp:246-249); otherwise at the trial pose (lm.qn, lm.tn).  Returns true when the level is finished;
// R_out/t_out then hold the accepted pose (:308-309), else the next trial pose exp(inc) * referenceToFrame (:266).
// *traced is set when an LM try was judged (te, if not null, receives it).
__device__ __forceinline__ bool lm_step(LMState &lm, const double *rec, const revo_opt_config &oc, int lvl, bool first,
                                        float *R_out, float *t_out, revo_trace_entry *te, bool *traced)
{
    const float err = (float)(rec[kRecSW] / rec[kRecGood]);    // :190
    bool propose = false, done = false;
    *traced = false;
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
        *traced = true;
        if (te) {
            te->error = err; te->lambda = lm.lambda; te->accepted = accepted ? 1 : 0;
            te->good = (int)rec[kRecGood]; te->bad = (int)rec[kRecBad]; te->level = lvl;
        }
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
                else {                                                                                 // pow(fail_fac, incTry)
                    float pw = 1.f;
                    for (int k = 0; k < lm.incTry; ++k) pw *= oc.lambda_fail_fac;
                    lm.lambda *= pw;
                }
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
        double qe[4], te3[3];
        se3_exp(lm.inc, qe, te3);
        se3_mul(qe, te3, lm.q, lm.t, lm.qn, lm.tn);              // :266 exp(inc) * referenceToFrame
        double Rn[9];
        quat_to_R(lm.qn, Rn);
        for (int i = 0; i < 9; ++i) R_out[i] = (float)Rn[i];
        for (int i = 0; i < 3; ++i) t_out[i] = (float)lm.tn[i];
    }
    if (done) {
        // next level (or the result) starts from the accepted pose      :308-309
        double Ra[9];
        quat_to_R(lm.q, Ra);
        for (int i = 0; i < 9; ++i) R_out[i] = (float)Ra[i];
        for (int i = 0; i < 3; ++i) t_out[i] = (float)lm.t[i];
    }
    return done;
}

// ---- per-point work: PASS A + PASS B fused ---------------------------------------
// One 256-bit load (LDG.E.ENL2.256 on sm_100a) of the 32-byte QUAD record of pixel (ix,iy): the four distance-transform
// values and the four packed gradients the bilinear fetch of optimizer.h:173-185 needs.  Returned as the two row
// records r0 = {dt(x,y), dt(x+1,y), g(x,y), g(x+1,y)}, r1 = the same for row y+1.  One gather and one address per point
// instead of two (or four texel fetches); on its own this measured neutral -- the gather phase is bound neither by L1
// wavefronts nor by per-thread memory parallelism (profiles/r1_k_track_v6_hotspots.txt) -- but it is the cheapest fetch.
__device__ __forceinline__ void ldg_quad(const uint4 *p, uint4 &r0, uint4 &r1)
{
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0.x), "=r"(r0.y), "=r"(r1.x), "=r"(r1.y), "=r"(r0.z), "=r"(r0.w), "=r"(r1.z), "=r"(r1.w)
                 : "l"(p));
}

// snorm16 pair -> floats (scale folded in by the caller)
__device__ __forceinline__ void unpack_grad(uint32_t g, float &gx, float &gy)
{
    gx = (float)(short)(g & 0xffffu);
    gy = (float)((int)g >> 16);
}

// ---- branch-free per-point work (all engines) -------------------------------------------------------------------
// optimizer.cpp:93-131 + calculateWarpUpdate (:204-228) + LGS6::update (LGSX.h:392-398).  A point that does not exist,
// projects out of bounds or fails the edge filter runs through the same straight-line code with weight 0 (its texel fetch is redirected to texel 0 and
// its projection is zeroed so that no inf/NaN can reach the sums).  Straight-line code lets the compiler interleave
// the arithmetic of one point with the address computation and gathers of the next, and no lane ever waits for a
// divergent neighbour.  The two divisions are single MUFU.RCP (<= 1 ulp, far inside the float tolerance of the path).
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct ProjB {
    float a, b, iz, dx, dy;   // a = Wx/Wz, b = Wy/Wz (0 when invalid)
    const uint4 *bp;
    bool exists, valid;
};

struct LevelConst {           // per-level constants of an evaluation, kept in registers
    float fx, fy, cx, cy, umax, vmax;
    int w;
    const uint4 *opt;
};

__device__ __forceinline__ ProjB project_b(bool exists, const float4 p, const LevelConst &L, const float *__restrict__ R,
                                           const float *__restrict__ t)
{
    ProjB o;
    const float Wx = R[0] * p.x + R[3] * p.y + R[6] * p.z + t[0];
    const float Wy = R[1] * p.x + R[4] * p.y + R[7] * p.z + t[1];
    const float Wz = R[2] * p.x + R[5] * p.y + R[8] * p.z + t[2];
    const float iz = rcp_approx(Wz);
    const float a = Wx * iz, b = Wy * iz;
    const float u = a * L.fx + L.cx;
    const float v = b * L.fy + L.cy;
    const bool inb = (u > 1.f && v > 1.f && u < L.umax && v < L.vmax);   // NaN-safe (optimizer.cpp:100)
    o.exists = exists;
    o.valid = exists && inb;
    const int ix = o.valid ? (int)u : 0, iy = o.valid ? (int)v : 0;
    o.dx = o.valid ? u - (float)ix : 0.f;
    o.dy = o.valid ? v - (float)iy : 0.f;
    o.a = o.valid ? a : 0.f;
    o.b = o.valid ? b : 0.f;
    o.iz = o.valid ? iz : 0.f;
    o.bp = L.opt + 2u * (unsigned)(iy * L.w + ix);
    return o;
}

__device__ __forceinline__ void finish_point_b(const ProjB &P, const uint4 r0, const uint4 r1, const LevelConst &L, float edge_dist,
                                               bool use_filter, float huber, float (&acc)[32])
{
    // getInterpolatedElement43, optimizer.h:173-185
    const float dxdy = P.dx * P.dy;
    const float w11 = dxdy, w01 = P.dy - dxdy, w10 = P.dx - dxdy, w00 = 1.f - P.dx - P.dy + dxdy;
    float gx00, gy00, gx10, gy10, gx01, gy01, gx11, gy11;
    unpack_grad(r0.z, gx00, gy00); unpack_grad(r0.w, gx10, gy10);
    unpack_grad(r1.z, gx01, gy01); unpack_grad(r1.w, gx11, gy11);
    constexpr float kq = 1.0f / 32764.0f;
    const float gx = (w11 * gx11 + w01 * gx01 + w10 * gx10 + w00 * gx00) * (kq * L.fx);   // optimizer.cpp:119
    const float gy = (w11 * gy11 + w01 * gy01 + w10 * gy10 + w00 * gy00) * (kq * L.fy);   // optimizer.cpp:120
    const float r = w11 * __uint_as_float(r1.y) + w01 * __uint_as_float(r1.x) + w10 * __uint_as_float(r0.y) + w00 * __uint_as_float(r0.x);
    const bool pass = P.valid && !(use_filter && r > edge_dist);                   // optimizer.cpp:100,112
    const float hub = huber * rcp_approx(fmaxf(r, huber));                          // optimizer.h:159: r <= huber ? 1 : huber / r
    const float wr = pass ? ((r <= huber) ? 1.f : hub) : 0.f;
    const float rs = pass ? r : 0.f;
    acc[kRecGood] += pass ? 1.f : 0.f;
    acc[kRecBad] += (P.exists && !pass) ? 1.f : 0.f;
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


// ---- mbarrier / st.async PTX (cluster exchange without a cluster-wide fence) ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// 8 bytes into the shared memory of CTA `dst_rank` of this cluster (same offset as `local_ptr`), completing 8 bytes of
// the transaction count of that CTA's mbarrier (same offset as `local_bar`): STAS.64 on sm_100a.
__device__ __forceinline__ void st_async_b64(void *local_ptr, unsigned dst_rank, unsigned long long v, uint64_t *local_bar)
{
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(dst_rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(dst_rank));
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(ra), "l"(v), "r"(rb) : "memory");
}

}  // namespace revo
