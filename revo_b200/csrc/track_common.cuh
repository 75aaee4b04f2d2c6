// track_common.cuh -- device helpers of the tracking kernel (track.cu): record layout, SE3 / 6x6 solver in double, the
// Levenberg-Marquardt state machine, the fused PASS A + PASS B per-point work and the transposing warp reduction.
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

// arithmetic type of the LM step (see "SE3 / solver helpers" below)
#ifdef REVO_LM_DOUBLE
typedef double lmreal;
#else
typedef float lmreal;
#endif

struct Ctrl {
    // written by thread 0 of every CTA (identically), read by all threads
    int cur;          // index of the Trial whose pose (R, t) the next evaluation uses
    int level_done;
    int pair_skip;
    int next_pair;
};

// A pose to evaluate: exp(inc) * accepted pose for one value of lambda (or a level's start pose: only R, t are set).
struct Trial {
    lmreal qn[4], tn[3];   // Sophus SE3: unit quaternion xyzw + translation
    lmreal inc[6];         // the increment it was made from (step-size test of a rejected try, optimizer.cpp:294)
    float R[9], t[3];      // the same pose as the evaluation reads it (column-major R)
    float lambda;          // the lambda the normal equations were solved with
    int pad;
};

struct LMState {
    lmreal q[2][4], t[2][3];   // accepted pose, double-buffered: [pacc] is current (a reader of the other half is never disturbed)
    float lastErr, last_residual, lambda;
    int iteration, incTry, tries;
    int acc;                   // index of the record buffer that holds the normal equations of the accepted pose
    int pacc;
};

// What the speculating thread may read during ONE evaluation (written by thread 0 at the end of the previous LM step, double-
// buffered by evaluation parity): where the accepted normal equations / pose are and the lambda the NEXT try would use if
// the try being evaluated is rejected.
struct SpecIn {
    int acc, pacc, active;
    float lambda;
};

// What lm_step asks for: "solve with `lambda` on the record rec[acc] from the accepted pose [pacc] into trial[slot]".
struct LMOrder {
    int propose, slot, acc, pacc;
    float lambda;
};

// ---- SE3 / solver helpers of the LM step ------------------------------------------------------------------------------
// Templates over the arithmetic type.  The library instantiates them with lmreal = float: that is the arithmetic of the
// reference (Eigen::Matrix<float,6,6>::ldlt(), Sophus::SE3f); the record sums feeding the step stay in double.  The step is
// a chain of ~500 dependent operations on ONE thread: a dependent DFMA has ~9 cycles of latency on B200 against ~4.7 for an
// FFMA (scratch/probes/lm_probe.cu), which made the double-precision step ~3.7 k cycles per evaluation on the critical
// path of every pair.  (Spreading the 6x6 solve over the lanes of a warp was tried and is slower than one thread: every
// elimination step then hangs on two ~25-cycle shuffles.)  REVO_LM_DOUBLE switches the step back to double for A/B.

__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
template <typename T> __device__ __forceinline__ T lm_rcp(T x);
template <> __device__ __forceinline__ float lm_rcp<float>(float x)      // MUFU.RCP + one Newton step (<= 1 ulp), no subroutine call
{
    const float r = rcp_approx(x);
    return fmaf(r, fmaf(-x, r, 1.f), r);
}
template <> __device__ __forceinline__ double lm_rcp<double>(double x) { return __drcp_rn(x); }
template <typename T> __device__ __forceinline__ T lm_fma(T a, T b, T c);
template <> __device__ __forceinline__ float lm_fma<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double lm_fma<double>(double a, double b, double c) { return fma(a, b, c); }

template <typename T>
__device__ __forceinline__ void quat_to_R(const T *q, T *R /* col-major */)
{
    const T x = q[0], y = q[1], z = q[2], w = q[3];
    const T tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const T twx = tx * w, twy = ty * w, twz = tz * w;
    const T txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[3] = txy - twz;       R[6] = txz + twy;
    R[1] = txy + twz;       R[4] = 1 - (txx + tzz); R[7] = tyz - twx;
    R[2] = txz - twy;       R[5] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen quaternion-from-matrix (Shepperd), as SO3(Matrix3) does (so3.hpp:419). R col-major float; once per pair, in double.
template <typename T>
__device__ inline void quat_from_R(const float *Rf, T *qo)
{
    double R[9], q[4];
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
    for (int i = 0; i < 4; ++i) qo[i] = (T)q[i];
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

// Sophus::SE3::exp (se3.hpp:723-748, so3.hpp:531-564).  The four coefficients sin(t/2)/t, cos(t/2), (1 - cos t)/t^2 and
// (t - sin t)/t^3 are even functions of t; for the increments of a tracker (t < 0.5 rad, in practice < 0.05) they are
// evaluated as power series in t^2 as four independent Horner chains: no sqrt,
// sincos or division on the serial critical path of an evaluation, and -- unlike the closed forms evaluated in float, as
// the reference does -- no cancellation in (1 - cos t) and (t - sin t).  Larger angles take the closed form.
template <typename T>
__device__ __forceinline__ void se3_exp(const T *xi, T *q, T *t)
{
    const T ox = xi[3], oy = xi[4], oz = xi[5];
    const T s = ox * ox + oy * oy + oz * oz;   // theta^2
    T imag, re, c1, c2;
    if (s < (T)1e-10) {   // theta < Sophus::Constants<float>::epsilon() = 1e-5
        const T t4 = s * s;
        imag = (T)0.5 - (T)(1.0 / 48.0) * s + (T)(1.0 / 3840.0) * t4;
        re = (T)1.0 - (T)(1.0 / 8.0) * s + (T)(1.0 / 384.0) * t4;
        // V = R(q) there (se3.hpp:735-737) = I + 2 re imag Om + 2 imag^2 Om^2
        c1 = (T)2.0 * re * imag;
        c2 = (T)2.0 * imag * imag;
    } else if (s < (T)0.25) {
        // coefficient k of (-s)^k: 1/(2^(2k+1) (2k+1)!), 1/(4^k (2k)!), 1/(2k+2)!, 1/(2k+3)!.  8 terms leave < 1e-17 (double),
        // 5 terms < 4e-10 (float: below half an ulp of every coefficient function on this range)
        constexpr int N = sizeof(T) == 4 ? 5 : 8;
        constexpr double ci[8] = {0.5, 1.0 / 48.0, 1.0 / 3840.0, 1.0 / 645120.0, 1.0 / 185794560.0, 1.0 / 81749606400.0,
                                  1.0 / 51011754393600.0, 1.0 / 42849873690624000.0};
        constexpr double cr[8] = {1.0, 1.0 / 8.0, 1.0 / 384.0, 1.0 / 46080.0, 1.0 / 10321920.0, 1.0 / 3715891200.0,
                                  1.0 / 1961990553600.0, 1.0 / 1428329123020800.0};
        constexpr double c1c[8] = {0.5, 1.0 / 24.0, 1.0 / 720.0, 1.0 / 40320.0, 1.0 / 3628800.0, 1.0 / 479001600.0,
                                   1.0 / 87178291200.0, 1.0 / 20922789888000.0};
        constexpr double c2c[8] = {1.0 / 6.0, 1.0 / 120.0, 1.0 / 5040.0, 1.0 / 362880.0, 1.0 / 39916800.0, 1.0 / 6227020800.0,
                                   1.0 / 1307674368000.0, 1.0 / 355687428096000.0};
        imag = (T)ci[N - 1]; re = (T)cr[N - 1]; c1 = (T)c1c[N - 1]; c2 = (T)c2c[N - 1];
#pragma unroll
        for (int k = N - 2; k >= 0; --k) {      // four independent Horner chains
            imag = imag * -s + (T)ci[k];
            re = re * -s + (T)cr[k];
            c1 = c1 * -s + (T)c1c[k];
            c2 = c2 * -s + (T)c2c[k];
        }
    } else {
        const double sd = (double)s, theta = sqrt(sd);      // rare (a step of more than 0.5 rad): closed form in double
        double sn, cs;
        sincos(0.5 * theta, &sn, &cs);
        const double inv_t = 1.0 / theta, inv_t2 = inv_t * inv_t;
        imag = (T)(sn * inv_t);
        re = (T)cs;
        c1 = (T)(2.0 * sn * sn * inv_t2);                        // (1 - cos t) / t^2
        c2 = (T)((theta - 2.0 * sn * cs) * inv_t2 * inv_t);      // (t - sin t) / t^3
    }
    q[0] = imag * ox; q[1] = imag * oy; q[2] = imag * oz; q[3] = re;
    // V = I + c1 Om + c2 Om^2 ; Om = hat(omega), Om^2 = omega omega^T - |omega|^2 I
    const T v00 = 1 + c2 * (ox * ox - s), v01 = -c1 * oz + c2 * ox * oy, v02 = c1 * oy + c2 * ox * oz;
    const T v10 = c1 * oz + c2 * ox * oy, v11 = 1 + c2 * (oy * oy - s), v12 = -c1 * ox + c2 * oy * oz;
    const T v20 = -c1 * oy + c2 * ox * oz, v21 = c1 * ox + c2 * oy * oz, v22 = 1 + c2 * (oz * oz - s);
    t[0] = v00 * xi[0] + v01 * xi[1] + v02 * xi[2];
    t[1] = v10 * xi[0] + v11 * xi[1] + v12 * xi[2];
    t[2] = v20 * xi[0] + v21 * xi[1] + v22 * xi[2];
}

// (qa,ta) * (qb,tb) with Sophus' renormalisation (se3.hpp:317-321, so3.hpp:335-352)
template <typename T>
__device__ __forceinline__ void se3_mul(const T *qa, const T *ta, const T *qb, const T *tb, T *q, T *t)
{
    T ux = qa[1] * tb[2] - qa[2] * tb[1], uy = qa[2] * tb[0] - qa[0] * tb[2], uz = qa[0] * tb[1] - qa[1] * tb[0];
    ux += ux; uy += uy; uz += uz;
    const T cx = qa[1] * uz - qa[2] * uy, cy = qa[2] * ux - qa[0] * uz, cz = qa[0] * uy - qa[1] * ux;
    t[0] = ta[0] + (tb[0] + qa[3] * ux + cx);
    t[1] = ta[1] + (tb[1] + qa[3] * uy + cy);
    t[2] = ta[2] + (tb[2] + qa[3] * uz + cz);
    const T ax = qa[0], ay = qa[1], az = qa[2], aw = qa[3], bx = qb[0], by = qb[1], bz = qb[2], bw = qb[3];
    T w = aw * bw - ax * bx - ay * by - az * bz;
    T x = aw * bx + ax * bw + ay * bz - az * by;
    T y = aw * by + ay * bw + az * bx - ax * bz;
    T z = aw * bz + az * bw + ax * by - ay * bx;
    const T sn = x * x + y * y + z * z + w * w;
    if (sn != (T)1.0) {
        const T s = (T)2.0 * lm_rcp<T>((T)1.0 + sn);
        x *= s; y *= s; z *= s; w *= s;
    }
    q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

// Solve (A with diag * lam1) x = b for the symmetric positive (semi-)definite 6x6 normal equations
// (system/optimizer.cpp:258-262, "A.ldlt().solve(b)"; the reference divides A and b by the number of constraints first,
// LGSX.h:320-326; lm_scaled below does the same when the record leaves double precision, one value per lane of the warp that
// reduced it, so that the float range is never an issue and the serial step starts from ready-made floats).
// LDL^T without pivoting, fully unrolled so that everything stays in registers (the matrix is a damped sum of outer
// products; Eigen's diagonal pivoting only changes rounding).  Non-positive / non-finite pivots are treated like Eigen's
// pseudo-inverse of D: that component becomes 0.
template <typename T>
__device__ __forceinline__ void solve6(const T *Au /* 21 upper slots of A / n */, const T *b /* (sum w r v) / n */, T lam1, T *x)
{
    T a[6][6];     // lower triangle: a[i][j], i >= j; after step k column k holds L(:,k) below the diagonal
    {
        int s = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) a[j][i] = Au[s++];
    }
    T y[6], invd[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { a[i][i] *= lam1; y[i] = b[i]; }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const T dk = a[k][k];
        const T id = (dk > (T)0.0 && dk < (T)1e30) ? lm_rcp<T>(dk) : (T)0.0;
        invd[k] = id;
#pragma unroll
        for (int j = k + 1; j < 6; ++j) {
            const T ljk = a[j][k] * id;     // L(j,k)
#pragma unroll
            for (int i = j; i < 6; ++i) a[i][j] = lm_fma<T>(-ljk, a[i][k], a[i][j]);
            y[j] = lm_fma<T>(-ljk, y[k], y[j]);        // forward substitution L y = b rides along
            a[j][k] = ljk;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] *= invd[i];
#pragma unroll
    for (int j = 5; j >= 1; --j)                 // L^T x = D^-1 y, column by column: x_j is final when column j is used
#pragma unroll
        for (int i = 0; i < j; ++i) y[i] = lm_fma<T>(-a[j][i], y[j], y[i]);
#pragma unroll
    for (int i = 0; i < 6; ++i) x[i] = y[i];
}

// ---- the Levenberg-Marquardt state machine ---------------------------------------------------------------------------
// Optimizer::trackFrames, system/optimizer.cpp:243-306, restated as "consume the record of the evaluation that just
// finished, decide, and name the next pose to evaluate".

// The proposal for one lambda: solve (A with diag * (1 + lambda)) inc = sum w r v, pose = exp(inc) * accepted pose
// (optimizer.cpp:258-266).  (q, t): the accepted pose.
__device__ __forceinline__ void lm_pose_from_inc(const lmreal *inc, const lmreal *q, const lmreal *t, lmreal *qn, lmreal *tn, float *R,
                                                 float *tf)
{
    lmreal qe[4], te3[3];
    se3_exp<lmreal>(inc, qe, te3);
    se3_mul<lmreal>(qe, te3, q, t, qn, tn);              // :266 exp(inc) * referenceToFrame
    lmreal Rn[9];
    quat_to_R<lmreal>(qn, Rn);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = (float)Rn[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) tf[i] = (float)tn[i];
}
// One value of the record as the LM step uses it: A / n and (sum w r v) / n in lmreal (LGS6::finish, LGSX.h:320-326); n: the
// number of good points (slot kRecGood of the same record).
__device__ __forceinline__ lmreal lm_scaled(double v, double n)
{
    return (lmreal)v * lm_rcp<lmreal>((lmreal)n);
}

// recs: the scaled record (lm_scaled) of the accepted pose
__device__ __forceinline__ void lm_propose(const lmreal *recs, const lmreal *q, const lmreal *t, float lambda, Trial &o)
{
    solve6<lmreal>(recs + kRecA, recs + kRecB, (lmreal)(1.f + lambda), o.inc);
    lm_pose_from_inc(o.inc, q, t, o.qn, o.tn, o.R, o.t);
    o.lambda = lambda;
}
// lambda of the next try if the try made with (lambda, incTry) is rejected (optimizer.cpp:300-303; incTry counts the tries of
// the current iteration including that one)
__device__ __forceinline__ float lm_reject_lambda(float lambda, int incTry, float fail_fac)
{
    if (lambda == 0.f) return 0.2f;
    float pw = 1.f;                                  // std::pow(lambdaFailFac, incTry)
    for (int k = 0; k < incTry; ++k) pw *= fail_fac;
    return lambda * pw;
}

// One step.  rec[2][32]: the evaluation that just finished wrote rec[lm.acc ^ 1]; rec[lm.acc] holds the normal equations of
// the accepted pose.  trial[3]: trial[cur] is the pose that was evaluated; `first`: it was the level's start pose
// (optimizer.cpp:246-249).  spec: the reject-successor of trial[cur] may already sit in trial[(cur + 1) % 3], computed by
// another thread while the record was being exchanged (see k_track) -- it is used if it was made for exactly the lambda
// this step arrives at (spec_ran: that thread did run during this evaluation, with input `spec`).  On return cur names the pose to evaluate next or, when the level is finished (true is returned),
// the accepted pose (:308-309); if order.propose is set the caller still has to compute that pose (lm_propose / lm_propose_warp
// with the fields of `order`).  spec_next: what the speculating thread may use during the next evaluation.
// *traced is set when an LM try was judged (te, if not null, receives it).
__device__ __forceinline__ bool lm_step(LMState &lm, Trial *trial, int &cur, const double (*rec)[32], bool spec_ran,
                                        const SpecIn &spec, SpecIn &spec_next, LMOrder &order, const revo_opt_config &oc, int lvl,
                                        bool first, revo_trace_entry *te, bool *traced)
{
    const double *r = rec[lm.acc ^ 1];
    const float err = (float)(r[kRecSW] / r[kRecGood]);    // :190
    bool propose = false, done = false, rejected = false;
    *traced = false;
    if (first) {
        lm.lastErr = err;
        lm.last_residual = err;
        lm.lambda = oc.lambda_initial[lvl];
        lm.iteration = 0; lm.incTry = 0; lm.tries = 0;
        lm.acc ^= 1;
        propose = true;
    } else {
        const bool accepted = err < lm.lastErr;                // :273
        *traced = true;
        if (te) {
            te->error = err; te->lambda = lm.lambda; te->accepted = accepted ? 1 : 0;
            te->good = (int)r[kRecGood]; te->bad = (int)r[kRecBad]; te->level = lvl;
        }
        if (accepted) {
            const int np = lm.pacc ^ 1;
            for (int i = 0; i < 4; ++i) lm.q[np][i] = trial[cur].qn[i];
            for (int i = 0; i < 3; ++i) lm.t[np][i] = trial[cur].tn[i];
            lm.pacc = np;
            lm.acc ^= 1;                                       // the record just taken becomes the accepted one
            if (err / lm.lastErr > oc.convergence_eps[lvl]) lm.iteration = oc.max_its_per_lvl[lvl];   // :279-283
            lm.last_residual = lm.lastErr = err;
            if (lm.lambda <= 0.2f) lm.lambda = 0.f; else lm.lambda *= oc.lambda_success_fac;          // :286-289
            lm.iteration++;     // for-loop increment after the break (:291)
            lm.incTry = 0;
            propose = true;
        } else {
            lmreal dot = 0;
            for (int i = 0; i < 6; ++i) dot += trial[cur].inc[i] * trial[cur].inc[i];
            if (!((float)dot > oc.step_size_min[lvl])) {                                               // :294
                done = true;
            } else {
                lm.lambda = lm_reject_lambda(lm.lambda, lm.incTry, oc.lambda_fail_fac);                // :300-303
                propose = true;
                rejected = true;
            }
        }
    }
    if (propose && !done) {
        if (lm.iteration >= oc.max_its_per_lvl[lvl]) done = true;
        else if (oc.max_lm_tries > 0 && lm.tries >= oc.max_lm_tries) done = true;
    }
    const int spec_slot = cur == 2 ? 0 : cur + 1, fresh_slot = cur == 0 ? 2 : cur - 1;
    order.propose = 0;
    if (propose && !done) {
        lm.incTry++; lm.tries++;
        if (rejected && spec_ran && spec.active && spec.lambda == lm.lambda && spec.acc == lm.acc && spec.pacc == lm.pacc) {
            cur = spec_slot;                                   // already computed, bit for bit what the order below would give
        } else {
            order.propose = 1; order.slot = fresh_slot; order.acc = lm.acc; order.pacc = lm.pacc; order.lambda = lm.lambda;
            cur = fresh_slot;
        }
        spec_next.acc = lm.acc; spec_next.pacc = lm.pacc;
        spec_next.lambda = lm_reject_lambda(lm.lambda, lm.incTry, oc.lambda_fail_fac);
        spec_next.active = 1;
    }
    if (done) {
        // next level (or the result) starts from the accepted pose      :308-309
        lmreal Ra[9];
        quat_to_R<lmreal>(lm.q[lm.pacc], Ra);
        Trial &o = trial[fresh_slot];
        for (int i = 0; i < 9; ++i) o.R[i] = (float)Ra[i];
        for (int i = 0; i < 3; ++i) o.t[i] = (float)lm.t[lm.pacc][i];
        cur = fresh_slot;
        spec_next.active = 0;
    }
    return done;
}

// ---- per-point work: PASS A + PASS B fused ---------------------------------------
// The four texels around a projected point (optimizer.h:173-185), each one 64-bit load from the tiled lookup structure
// (internal.h: opt_texel_index): .x = dt (float32 bits), .y = snorm16 gx | gy.
// kHint 0: plain; 1: ld.global.nc.L1::no_allocate; 2: additionally L2::evict_last through a cache policy, i.e. the keyframe
// texels of the pairs in flight outlive the streamed data (point lists, other kernels' traffic) in L2.
template <int kHint>
__device__ __forceinline__ uint2 ldg_texel(const uint2 *p, unsigned long long policy)
{
    uint2 v;
    if (kHint == 2)
        asm("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(policy));
    else if (kHint == 1)
        asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else
        asm("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last()
{
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// A 3-D edge point of the current frame: read once per level (then cached in shared memory), so it should neither stay in
// L1 nor displace the texels in L2 (kHint 2: L2 evict_first through a cache policy).
template <int kHint>
__device__ __forceinline__ float4 ldg_point(const float4 *p)
{
    if (kHint == 2) {
        float4 v;
        unsigned long long pol;
        asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
            : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
        return v;
    }
    return __ldg(p);
}

// snorm16 pair -> floats (scale folded in by the caller)
__device__ __forceinline__ void unpack_grad(uint32_t g, float &gx, float &gy)
{
    gx = (float)(short)(g & 0xffffu);
    gy = (float)((int)g >> 16);
}

// ---- per-point work -----------------------------------------------------------------------------------------------
// optimizer.cpp:93-131 + calculateWarpUpdate (:204-228) + LGS6::update (LGSX.h:392-398).  A thread only visits points that
// exist; a point that projects out of bounds or fails the edge filter runs through the same straight-line code with weight
// 0 (its texel fetch is redirected to texel 0 and its projection is zeroed so that no inf/NaN can reach the sums), so the
// compiler can interleave the arithmetic of one point with the address computation and gather of the next and no lane waits
// for a divergent neighbour.  The two divisions are single MUFU.RCP (<= 1 ulp, far inside the float tolerance of the path).

struct ProjB {
    float a, b, iz, dx, dy;   // a = Wx/Wz, b = Wy/Wz (0 when the point projects out of bounds)
    unsigned i00, i10, i01, i11;   // texel indices (opt_texel_index) of (ix,iy), (ix+1,iy), (ix,iy+1), (ix+1,iy+1)
    bool valid;
};

struct LevelConst {           // per-level constants of an evaluation, kept in registers
    float fx, fy, cx, cy, umax, vmax;
    unsigned tw16;            // tiles per row x 16 texels
    const uint2 *opt;
};

// optimizer.cpp:93-100: rigid transform (three FMA chains), projection, NaN-safe bounds test, texel addresses
__device__ __forceinline__ ProjB project_b(float x, float y, float z, const LevelConst &L, const float *__restrict__ R,
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
    o.valid = inb;
    const int ix = inb ? (int)u : 0, iy = inb ? (int)v : 0;
    o.dx = inb ? u - (float)ix : 0.f;
    o.dy = inb ? v - (float)iy : 0.f;
    o.a = inb ? a : 0.f;
    o.b = inb ? b : 0.f;
    o.iz = inb ? iz : 0.f;
    // x part: (tile column) * 16 + (x & 3); y part: (tile row) * tw16 + (y & 3) * 4
    const unsigned ux = (unsigned)ix, uy = (unsigned)iy;
    const unsigned cx0 = ((ux & ~3u) << 2) | (ux & 3u), cx1 = (((ux + 1u) & ~3u) << 2) | ((ux + 1u) & 3u);
    const unsigned cy0 = (uy >> 2) * L.tw16 + ((uy & 3u) << 2), cy1 = ((uy + 1u) >> 2) * L.tw16 + (((uy + 1u) & 3u) << 2);
    o.i00 = cy0 + cx0; o.i10 = cy0 + cx1; o.i01 = cy1 + cx0; o.i11 = cy1 + cx1;
    return o;
}

// optimizer.cpp:101-131 + calculateWarpUpdate (:204-228) + LGS6::update (LGSX.h:392-398) for a point that exists.
// t00 .. t11: the texels of (ix,iy), (ix+1,iy), (ix,iy+1), (ix+1,iy+1).  kqfx = fx / 32764, kqfy = fy / 32764 (gradient
// scale), ed_eff = edge filter distance or +inf when the filter is off.  The "bad" counter is not kept here: every visited
// point exists, so bad = visited - good (set by the caller).
__device__ __forceinline__ void finish_point_b(const ProjB &P, const uint2 t00, const uint2 t10, const uint2 t01, const uint2 t11, float kqfx,
                                               float kqfy, float ed_eff, float huber, float (&acc)[32])
{
    // getInterpolatedElement43, optimizer.h:173-185
    const float dxdy = P.dx * P.dy;
    const float w11 = dxdy, w01 = P.dy - dxdy, w10 = P.dx - dxdy, w00 = 1.f - P.dx - P.dy + dxdy;
    float gx00, gy00, gx10, gy10, gx01, gy01, gx11, gy11;
    unpack_grad(t00.y, gx00, gy00); unpack_grad(t10.y, gx10, gy10);
    unpack_grad(t01.y, gx01, gy01); unpack_grad(t11.y, gx11, gy11);
    const float gx = (w11 * gx11 + w01 * gx01 + w10 * gx10 + w00 * gx00) * kqfx;   // optimizer.cpp:119
    const float gy = (w11 * gy11 + w01 * gy01 + w10 * gy10 + w00 * gy00) * kqfy;   // optimizer.cpp:120
    const float r = w11 * __uint_as_float(t11.x) + w01 * __uint_as_float(t01.x) + w10 * __uint_as_float(t10.x) + w00 * __uint_as_float(t00.x);
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
