/*
 * revo_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the edge-based RGB-D tracking hot path of
 * fabianschenk/REVO, written from the reference's behaviour, function by
 * function.  It exists only so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg can check or time the CUDA
 * product against it.  Nothing under revo_b200/ may include, link or call it.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout).  Third-party arithmetic that is NOT in the
 * reference tree is restated from the published algorithm:
 *   - OpenCV imgproc (reference pins "OpenCV 3", CMakeLists.txt:46): cvtColor,
 *     pyrDown, Canny(aperture 3, L2gradient), distanceTransform(L2, PRECISE).
 *     Parity target is python cv2 4.13 in this image; tests/test_oracle.py
 *     pins each restatement bit-exactly against cv2.
 *   - Eigen >=3.3 (CMakeLists.txt:128, README.md:20): fixed-size products,
 *     LDLT<Matrix6f> (pivoted), Quaternion<->Matrix3 conversions.
 *   - Sophus SE3/SO3 is vendored in the reference tree
 *     (thirdparty/Sophus/sophus/{se3,so3}.hpp) and pinned by its own KATs
 *     (thirdparty/Sophus/test/core/test_se3.cpp:137-146), see tests/golden.
 *
 * Build twice (see oracle/Makefile): REAL=float  -> "reference-as-is"
 * (sequential float32 sums, exactly the reference's arithmetic order) and
 * REAL=double -> "truth".  Inputs (point lists, lookup structure) are float32
 * in both, as in the reference.
 *
 * PARITY STATUS: REVO ships no tests and its build (CMake, Eigen, OpenCV C++, Boost) cannot run in this image, but
 * its hot-path SOURCES compile from where they lie: `make -C oracle ref` builds imgpyramidrgbd.cpp, optimizer.cpp,
 * tracker.cpp, LGSX.h (+ Logging.cpp) verbatim against the API shims of oracle/shim/ into oracle/_ref/librevo_ref.so.
 * tests/test_oracle_ref.py pins this restatement to that library -- every pyramid array, Optimizer::trackFrames
 * (evaluation count, LM trace, pose, residual info), the evaluation record, TrackerNew::trackFrames, evalCostFunction, the
 * quality vote -- BIT FOR BIT in float32, and tests/golden/ref_golden.npz (tests/golden/make_ref_golden.py) carries the
 * same outputs to machines without /root/reference.  What the shims themselves supply (not reference code) is pinned
 * separately: (a) SE3 exp / product by Sophus' KATs and sympy, (b) the four OpenCV kernels against cv2 4.13 outputs,
 * (c) LDLT (Eigen's published algorithm, restated) against numpy.linalg.solve.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifndef REAL
#define REAL float
#endif
typedef REAL real;

#if defined(__GNUC__)
#define ORC_API __attribute__((visibility("default")))
#else
#define ORC_API
#endif

/* ------------------------------------------------------------------------ */
/* helpers                                                                   */
/* ------------------------------------------------------------------------ */
static inline real r_sqrt(real x) { return (real)sqrt((double)x); }
static inline real r_sin(real x) { return sizeof(real) == 4 ? (real)sinf((float)x) : (real)sin((double)x); }
static inline real r_cos(real x) { return sizeof(real) == 4 ? (real)cosf((float)x) : (real)cos((double)x); }
static inline real r_abs(real x) { return x < 0 ? -x : x; }

ORC_API int orc_sizeof_real(void) { return (int)sizeof(real); }

/* ------------------------------------------------------------------------ */
/* Sophus SE3 / SO3 (thirdparty/Sophus/sophus/se3.hpp, so3.hpp) + the Eigen  */
/* quaternion conversions they call.  Quaternions are stored (x,y,z,w) like   */
/* Eigen::Quaternion::coeffs().  Rotation matrices are COLUMN-major 9 reals   */
/* (Eigen::Matrix3f::data()).                                                 */
/* ------------------------------------------------------------------------ */
#define RM(R, i, j) ((R)[(j) * 3 + (i)])

/* Eigen/src/Geometry/Quaternion.h quaternionbase_assign_impl<Other,3,3>
 * (Shepperd's method) -- used by SO3(Matrix3) at so3.hpp:419. */
static void quat_from_R(const real *R, real *q)
{
    real t = RM(R, 0, 0) + RM(R, 1, 1) + RM(R, 2, 2);
    if (t > (real)0) {
        t = r_sqrt(t + (real)1.0);
        q[3] = (real)0.5 * t;
        t = (real)0.5 / t;
        q[0] = (RM(R, 2, 1) - RM(R, 1, 2)) * t;
        q[1] = (RM(R, 0, 2) - RM(R, 2, 0)) * t;
        q[2] = (RM(R, 1, 0) - RM(R, 0, 1)) * t;
    } else {
        int i = 0;
        if (RM(R, 1, 1) > RM(R, 0, 0)) i = 1;
        if (RM(R, 2, 2) > RM(R, i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = r_sqrt(RM(R, i, i) - RM(R, j, j) - RM(R, k, k) + (real)1.0);
        q[i] = (real)0.5 * t;
        t = (real)0.5 / t;
        q[3] = (RM(R, k, j) - RM(R, j, k)) * t;
        q[j] = (RM(R, j, i) + RM(R, i, j)) * t;
        q[k] = (RM(R, k, i) + RM(R, i, k)) * t;
    }
}

/* Eigen QuaternionBase::toRotationMatrix -- so3.hpp:280-282 (matrix()). */
ORC_API void orc_quat_to_R(const real *q, real *R)
{
    const real x = q[0], y = q[1], z = q[2], w = q[3];
    const real tx = (real)2 * x, ty = (real)2 * y, tz = (real)2 * z;
    const real twx = tx * w, twy = ty * w, twz = tz * w;
    const real txx = tx * x, txy = ty * x, txz = tz * x;
    const real tyy = ty * y, tyz = tz * y, tzz = tz * z;
    RM(R, 0, 0) = (real)1 - (tyy + tzz);
    RM(R, 0, 1) = txy - twz;
    RM(R, 0, 2) = txz + twy;
    RM(R, 1, 0) = txy + twz;
    RM(R, 1, 1) = (real)1 - (txx + tzz);
    RM(R, 1, 2) = tyz - twx;
    RM(R, 2, 0) = txz - twy;
    RM(R, 2, 1) = tyz + twx;
    RM(R, 2, 2) = (real)1 - (txx + tyy);
}

/* Sophus::Constants<Scalar>::epsilon -- common.hpp:143,152 */
static inline real sophus_eps(void) { return sizeof(real) == 4 ? (real)1e-5 : (real)1e-10; }

/* SE3(Matrix3 R, Point t) -- se3.hpp:438-440 -> SO3(R) so3.hpp:419-424.
 * Returns 0 ok, 1 if R is not orthogonal (the reference abort()s there:
 * rotation_matrix.hpp:20-28 isOrthogonal = ||R R^T - I||_F < eps), 2 if
 * det(R) <= 0. */
ORC_API int orc_se3_from_Rt(const real *R, const real *t, real *q, real *t_out)
{
    real n2 = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            real s = 0;
            for (int k = 0; k < 3; ++k) s += RM(R, i, k) * RM(R, j, k);
            s -= (i == j) ? (real)1 : (real)0;
            n2 += s * s;
        }
    int rc = 0;
    if (!(r_sqrt(n2) < sophus_eps())) rc = 1;
    real det = RM(R, 0, 0) * (RM(R, 1, 1) * RM(R, 2, 2) - RM(R, 1, 2) * RM(R, 2, 1)) -
               RM(R, 0, 1) * (RM(R, 1, 0) * RM(R, 2, 2) - RM(R, 1, 2) * RM(R, 2, 0)) +
               RM(R, 0, 2) * (RM(R, 1, 0) * RM(R, 2, 1) - RM(R, 1, 1) * RM(R, 2, 0));
    if (!(det > (real)0) && rc == 0) rc = 2;
    quat_from_R(R, q);
    t_out[0] = t[0]; t_out[1] = t[1]; t_out[2] = t[2];
    return rc;
}

/* Eigen quaternion product (a*b), coefficient formula of Quaternion.h
 * internal::quat_product (scalar path). */
static void quat_mul(const real *a, const real *b, real *o)
{
    const real ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const real bx = b[0], by = b[1], bz = b[2], bw = b[3];
    real w = aw * bw - ax * bx - ay * by - az * bz;
    real x = aw * bx + ax * bw + ay * bz - az * by;
    real y = aw * by + ay * bw + az * bx - ax * bz;
    real z = aw * bz + az * bw + ax * by - ay * bx;
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

/* Eigen QuaternionBase::_transformVector: v + w*uv + q.vec x uv, uv = 2 (q.vec x v)
 * -- called by SO3::operator*(Point) so3.hpp:318-320. */
static void quat_rotate(const real *q, const real *v, real *o)
{
    real ux = q[1] * v[2] - q[2] * v[1];
    real uy = q[2] * v[0] - q[0] * v[2];
    real uz = q[0] * v[1] - q[1] * v[0];
    ux += ux; uy += uy; uz += uz;
    real cx = q[1] * uz - q[2] * uy;
    real cy = q[2] * ux - q[0] * uz;
    real cz = q[0] * uy - q[1] * ux;
    o[0] = v[0] + q[3] * ux + cx;
    o[1] = v[1] + q[3] * uy + cy;
    o[2] = v[2] + q[3] * uz + cz;
}

/* SE3::operator*  se3.hpp:285-289 + operator*= :317-321 + SO3::operator*=
 * so3.hpp:335-352 (quaternion renormalisation q *= 2/(1+|q|^2) if |q|^2 != 1). */
ORC_API void orc_se3_mul(const real *qa, const real *ta, const real *qb, const real *tb, real *q, real *t)
{
    real rt[3], qq[4];
    quat_rotate(qa, tb, rt);
    real tx = ta[0] + rt[0], ty = ta[1] + rt[1], tz = ta[2] + rt[2];
    quat_mul(qa, qb, qq);
    real sn = qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3];
    if (sn != (real)1.0) {
        real s = (real)2.0 / ((real)1.0 + sn);
        qq[0] *= s; qq[1] *= s; qq[2] *= s; qq[3] *= s;
    }
    q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3];
    t[0] = tx; t[1] = ty; t[2] = tz;
}

/* SE3::exp  se3.hpp:723-748, SO3::expAndTheta so3.hpp:531-564.
 * xi = (upsilon[3], omega[3]). */
ORC_API void orc_se3_exp(const real *xi, real *q, real *t)
{
    const real ox = xi[3], oy = xi[4], oz = xi[5];
    real theta_sq = ox * ox + oy * oy + oz * oz;
    real theta = r_sqrt(theta_sq);
    real half_theta = (real)0.5 * theta;
    real imag, re;
    if (theta < sophus_eps()) {
        real theta_po4 = theta_sq * theta_sq;
        imag = (real)0.5 - (real)(1.0 / 48.0) * theta_sq + (real)(1.0 / 3840.0) * theta_po4;
        re = (real)1 - (real)(1.0 / 8.0) * theta_sq + (real)(1.0 / 384.0) * theta_po4;
    } else {
        real s = r_sin(half_theta);
        imag = s / theta;
        re = r_cos(half_theta);
    }
    q[0] = imag * ox; q[1] = imag * oy; q[2] = imag * oz; q[3] = re;

    /* Omega = hat(omega), Omega_sq = Omega*Omega */
    real Om[9] = {0}, Om2[9], V[9];
    RM(Om, 0, 1) = -oz; RM(Om, 0, 2) = oy;
    RM(Om, 1, 0) = oz;  RM(Om, 1, 2) = -ox;
    RM(Om, 2, 0) = -oy; RM(Om, 2, 1) = ox;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            real s = 0;
            for (int k = 0; k < 3; ++k) s += RM(Om, i, k) * RM(Om, k, j);
            RM(Om2, i, j) = s;
        }
    if (theta < sophus_eps()) {
        orc_quat_to_R(q, V);
    } else {
        real c1 = ((real)1 - r_cos(theta)) / theta_sq;
        real c2 = (theta - r_sin(theta)) / (theta_sq * theta);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                RM(V, i, j) = ((i == j) ? (real)1 : (real)0) + c1 * RM(Om, i, j) + c2 * RM(Om2, i, j);
    }
    for (int i = 0; i < 3; ++i)
        t[i] = RM(V, i, 0) * xi[0] + RM(V, i, 1) * xi[1] + RM(V, i, 2) * xi[2];
}

/* ------------------------------------------------------------------------ */
/* Eigen::LDLT<Matrix<float,6,6>>::compute + solve  (Eigen/src/Cholesky/     */
/* LDLT.h, 3.3.x; lower, in place, diagonal pivoting) -- called at            */
/* system/optimizer.cpp:262  "inc = A.ldlt().solve(b)".                       */
/* A is column-major 6x6, only the lower triangle is read.                    */
/* ------------------------------------------------------------------------ */
ORC_API void orc_ldlt_solve6(const real *Ain, const real *b, real *x)
{
    enum { N = 6 };
    real M[N * N];
    int tr[N];
#define AM(i, j) M[(j) * N + (i)]
    memcpy(M, Ain, sizeof(M));
    for (int k = 0; k < N; ++k) {
        /* biggest |diagonal| in the remaining block */
        int p = k;
        real big = r_abs(AM(k, k));
        for (int i = k + 1; i < N; ++i)
            if (r_abs(AM(i, i)) > big) { big = r_abs(AM(i, i)); p = i; }
        tr[k] = p;
        if (p != k) {
            /* symmetric swap of rows/cols k and p in the lower triangle */
            for (int j = 0; j < k; ++j) { real tmp = AM(k, j); AM(k, j) = AM(p, j); AM(p, j) = tmp; }
            for (int i = p + 1; i < N; ++i) { real tmp = AM(i, k); AM(i, k) = AM(i, p); AM(i, p) = tmp; }
            for (int i = k + 1; i < p; ++i) { real tmp = AM(i, k); AM(i, k) = AM(p, i); AM(p, i) = tmp; }
            { real tmp = AM(k, k); AM(k, k) = AM(p, p); AM(p, p) = tmp; }
        }
        /* A10 = row k cols 0..k-1 ; A20 = rows k+1.. cols 0..k-1 ; A21 = rows k+1.., col k */
        if (k > 0) {
            real temp[N];
            for (int j = 0; j < k; ++j) temp[j] = AM(j, j) * AM(k, j);
            real s = 0;
            for (int j = 0; j < k; ++j) s += AM(k, j) * temp[j];
            AM(k, k) -= s;
            for (int i = k + 1; i < N; ++i) {
                real s2 = 0;
                for (int j = 0; j < k; ++j) s2 += AM(i, j) * temp[j];
                AM(i, k) -= s2;
            }
        }
        real piv = AM(k, k);
        /* Eigen: if pivot is not (exactly) zero, divide the column */
        if (r_abs(piv) > (real)0)
            for (int i = k + 1; i < N; ++i) AM(i, k) /= piv;
    }
    /* solve: dst = P b ; L^-1 ; D^-1 (pseudo) ; L^-T ; P^T */
    real y[N];
    for (int i = 0; i < N; ++i) y[i] = b[i];
    for (int k = 0; k < N; ++k)
        if (tr[k] != k) { real tmp = y[k]; y[k] = y[tr[k]]; y[tr[k]] = tmp; }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < i; ++j) y[i] -= AM(i, j) * y[j];
    {
        const real tol = (real)1 / (sizeof(real) == 4 ? (real)FLT_MAX : (real)DBL_MAX);
        for (int i = 0; i < N; ++i) {
            if (r_abs(AM(i, i)) > tol) y[i] /= AM(i, i);
            else y[i] = 0;
        }
    }
    for (int i = N - 1; i >= 0; --i)
        for (int j = i + 1; j < N; ++j) y[i] -= AM(j, i) * y[j];
    for (int k = N - 1; k >= 0; --k)
        if (tr[k] != k) { real tmp = y[k]; y[k] = y[tr[k]]; y[tr[k]] = tmp; }
    for (int i = 0; i < N; ++i) x[i] = y[i];
#undef AM
}

/* ------------------------------------------------------------------------ */
/* Optimizer  (system/optimizer.{h,cpp}, utils/LGSX.h)                        */
/* ------------------------------------------------------------------------ */
typedef struct {
    /* OptimizerSettings, system/optimizer.h:46-111 (fields the hot path reads) */
    float lambda_success_fac;     /* 0.5  :53 */
    float lambda_fail_fac;        /* 2.0  :54 */
    float lambda_initial[6];      /* 0    :63 */
    float step_size_min[6];       /* 1e-16:55 */
    float convergence_eps[6];     /* 0.999:65 */
    int   max_its_per_lvl[6];     /* 100  :56 */
    float edge_distance_lvl[6];   /* {30,20,10,5,5,5} :59 */
    float huber_edge;             /* 0.3  :75 */
    int   use_edge_filter;        /* false in ctor :80; TrackerSettings sets true (tracker.h:46) */
} orc_opt_cfg;

typedef struct { float fx, fy, cx, cy; int w, h; } orc_cam;

typedef struct {
    /* Optimizer::ResidualInfo  optimizer.h:117-139 */
    int good, bad;
    real sum_w, sum_unw;
} orc_resinfo;

typedef struct {
    /* 7 SoA buffers of optimizer.h:145-151 */
    real *x, *y, *z, *dx, *dy, *res, *wgt;
} orc_buffers;

ORC_API void orc_opt_cfg_default(orc_opt_cfg *c)
{
    c->lambda_success_fac = 0.5f;
    c->lambda_fail_fac = 2.0f;
    const float ed[6] = {30, 20, 10, 5, 5, 5};
    for (int l = 0; l < 6; ++l) {
        c->lambda_initial[l] = 0.f;
        c->step_size_min[l] = 1e-16f;
        c->convergence_eps[l] = 0.999f;
        c->max_its_per_lvl[l] = 100;
        c->edge_distance_lvl[l] = ed[l];
    }
    c->huber_edge = 0.3f;
    c->use_edge_filter = 1;
}

/* Optimizer::calcErrorAndBuffers  system/optimizer.cpp:74-191 (PASS A) with
 * getInterpolatedElement43 optimizer.h:173-185 and getWeightOfEvoR :156-160. */
static real calc_error_and_buffers(const float *pts4, int n, const float *opt4, const orc_cam *cam,
                                   const real *R, const real *T, const orc_opt_cfg *cfg, int lvl,
                                   orc_resinfo *ri, orc_buffers *buf)
{
    ri->good = ri->bad = 0;
    ri->sum_w = ri->sum_unw = 0;
    const int w = cam->w, h = cam->h;
    const real fx = cam->fx, fy = cam->fy, cx = cam->cx, cy = cam->cy;
    const real edge_dist = cfg->edge_distance_lvl[lvl];
    const real huber = cfg->huber_edge;
    for (int c = 0; c < n; ++c) {
        const real px = pts4[4 * c + 0], py = pts4[4 * c + 1], pz = pts4[4 * c + 2];
        /* :93  Wxp = R * p + T (Eigen coefficient-wise 3x3 product, then + T) */
        const real Wx = (RM(R, 0, 0) * px + RM(R, 0, 1) * py + RM(R, 0, 2) * pz) + T[0];
        const real Wy = (RM(R, 1, 0) * px + RM(R, 1, 1) * py + RM(R, 1, 2) * pz) + T[1];
        const real Wz = (RM(R, 2, 0) * px + RM(R, 2, 1) * py + RM(R, 2, 2) * pz) + T[2];
        const real u = Wx / Wz * fx + cx;   /* :94 */
        const real v = Wy / Wz * fy + cy;   /* :95 */
        if (!(u > 1 && v > 1 && u < w - 2 && v < h - 2)) { ri->bad++; continue; }  /* :100 */
        /* getInterpolatedElement43 */
        const int ix = (int)u, iy = (int)v;
        const real dx = u - ix, dy = v - iy, dxdy = dx * dy;
        const float *bp = opt4 + 4 * ((size_t)ix + (size_t)iy * w);
        const float *b01 = bp + 4 * w, *b11 = bp + 4 + 4 * w, *b10 = bp + 4;
        real interp[3];
        for (int k = 0; k < 3; ++k)
            interp[k] = dxdy * (real)b11[k] + (dy - dxdy) * (real)b01[k] + (dx - dxdy) * (real)b10[k] +
                        ((real)1 - dx - dy + dxdy) * (real)bp[k];
        const real residual = interp[2];
        if ((residual > edge_dist) && cfg->use_edge_filter) { ri->bad++; continue; }   /* :112 */
        const real w_r = (residual <= huber) ? (real)1 : huber / residual;             /* optimizer.h:159 */
        if (buf) {
            const int e = ri->good;
            buf->x[e] = Wx; buf->y[e] = Wy; buf->z[e] = Wz;
            buf->dx[e] = fx * interp[0];
            buf->dy[e] = fy * interp[1];
            buf->res[e] = residual;
            buf->wgt[e] = w_r;
        }
        const real res_2 = residual * residual;
        ri->sum_w += (w_r * res_2);       /* :131 sequential accumulation */
        ri->sum_unw += res_2;
        ri->good++;
    }
    return ri->sum_w / (real)ri->good;    /* :190 (NaN when good == 0, as the reference) */
}

/* Optimizer::calculateWarpUpdate system/optimizer.cpp:192-234 (PASS B) with
 * LGS6::initialize/update/finish utils/LGSX.h:196-204,392-398,320-326.
 * A is column-major 6x6 (full, symmetric), b 6, err 1. */
static void calculate_warp_update(const orc_buffers *buf, int good, real *A, real *b, real *err)
{
    for (int i = 0; i < 36; ++i) A[i] = 0;
    for (int i = 0; i < 6; ++i) b[i] = 0;
    real error = 0;
    size_t n = 0;
    for (int i = 0; i < good; ++i) {
        const real px = buf->x[i], py = buf->y[i], pz = buf->z[i];
        const real r = buf->res[i], gx = buf->dx[i], gy = buf->dy[i], wgt = buf->wgt[i];
        real v[6];
        const real z = (real)1.0 / pz;
        const real z_sqr = (real)1.0 / (pz * pz);
        v[0] = z * gx + 0;
        v[1] = 0 + z * gy;
        v[2] = (-px * z_sqr) * gx + (-py * z_sqr) * gy;
        /* :220-223  the literal 1.0 promotes these two to double */
        v[3] = (real)((double)((-px * py * z_sqr) * gx) + (-(1.0 + (double)(py * py * z_sqr))) * (double)gy);
        v[4] = (real)((1.0 + (double)(px * px * z_sqr)) * (double)gx + (double)((px * py * z_sqr) * gy));
        v[5] = (-py * z) * gx + (px * z) * gy;
        /* LGS6::update: A += J J^T w ; b -= J (r w) ; error += r r w ; n++ */
        for (int c = 0; c < 6; ++c)
            for (int rr = 0; rr < 6; ++rr) A[c * 6 + rr] += (v[rr] * v[c]) * wgt;
        const real rw = r * wgt;
        for (int k = 0; k < 6; ++k) b[k] -= v[k] * rw;
        error += r * r * wgt;
        n++;
    }
    const real fn = (real)n;
    for (int i = 0; i < 36; ++i) A[i] /= fn;
    for (int i = 0; i < 6; ++i) b[i] /= fn;
    *err = error / fn;
}

/* One fused evaluation at a pose (the unit the CUDA kernel computes):
 * rec[0..20] = upper triangle of sum(w v v^T) in LGS6 slot order
 * (0,0..5),(1,1..5),...,(5,5) (utils/LGSX.h:212-314), rec[21..26] = sum(w r v)
 * (= -b*n), rec[27] = sum(w r^2), rec[28] = sum(r^2), rec[29] = good,
 * rec[30] = bad, rec[31] = 0.  UN-normalised sums, always in double (the
 * record is the parity unit, compared with tolerance). */
ORC_API void orc_eval_record(const float *pts4, int n, const float *opt4, const orc_cam *cam,
                             const real *R, const real *T, const orc_opt_cfg *cfg, int lvl, double *rec)
{
    orc_resinfo ri;
    orc_buffers buf;
    real *mem = (real *)malloc(sizeof(real) * 7 * (size_t)(n > 0 ? n : 1));
    buf.x = mem; buf.y = mem + n; buf.z = mem + 2 * (size_t)n; buf.dx = mem + 3 * (size_t)n;
    buf.dy = mem + 4 * (size_t)n; buf.res = mem + 5 * (size_t)n; buf.wgt = mem + 6 * (size_t)n;
    calc_error_and_buffers(pts4, n, opt4, cam, R, T, cfg, lvl, &ri, &buf);
    real A[36], b[6], err;
    calculate_warp_update(&buf, ri.good, A, b, &err);
    const double fn = (double)ri.good;
    int s = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) rec[s++] = (double)A[j * 6 + i] * fn;
    for (int k = 0; k < 6; ++k) rec[21 + k] = -(double)b[k] * fn;
    rec[27] = (double)ri.sum_w;
    rec[28] = (double)ri.sum_unw;
    rec[29] = (double)ri.good;
    rec[30] = (double)ri.bad;
    rec[31] = 0;
    if (ri.good == 0) for (int i = 0; i < 27; ++i) rec[i] = 0;
    free(mem);
}

typedef struct {
    /* one LM try, logged for trace parity */
    float error;
    float lambda;     /* lambda used for this solve */
    int accepted;
    int good, bad;
} orc_trace_entry;

/* Optimizer::trackFrames  system/optimizer.cpp:235-311.
 * R (col-major 9), T in/out.  Returns last accepted residual (last_residual).
 * trace (optional, capacity trace_cap) receives one entry per PASS A after the
 * first; *n_evals counts PASS A calls (incl. the first).
 * max_evals <= 0: run to the reference's own termination. >0: hard cap on the
 * number of LM tries (fixed-iteration test mode). rc: 0 ok, 1/2 = SE3 ctor
 * precondition violated (reference abort()s). */
ORC_API real orc_track_level(const float *pts4, int n, const float *opt4, const orc_cam *cam,
                             real *R, real *T, const orc_opt_cfg *cfg, int lvl, orc_resinfo *ri,
                             orc_trace_entry *trace, int trace_cap, int *n_evals, int max_tries, int *rc_out)
{
    real q[4], t[3];
    int rc = orc_se3_from_Rt(R, T, q, t);                 /* :241 */
    if (rc_out) *rc_out = rc;
    orc_buffers buf;
    real *mem = (real *)malloc(sizeof(real) * 7 * (size_t)(n > 0 ? n : 1));
    buf.x = mem; buf.y = mem + n; buf.z = mem + 2 * (size_t)n; buf.dx = mem + 3 * (size_t)n;
    buf.dy = mem + 4 * (size_t)n; buf.res = mem + 5 * (size_t)n; buf.wgt = mem + 6 * (size_t)n;
    int evals = 0, tries = 0, ntrace = 0;
    real lastErr = calc_error_and_buffers(pts4, n, opt4, cam, R, T, cfg, lvl, ri, &buf);   /* :243 */
    evals++;
    real last_residual = lastErr;
    real lambda = cfg->lambda_initial[lvl];
    int stop = 0;
    for (int iteration = 0; iteration < cfg->max_its_per_lvl[lvl] && !stop; iteration++) {
        real A[36], b[6], err;
        calculate_warp_update(&buf, ri->good, A, b, &err);     /* :252 */
        int incTry = 0;
        while (1) {
            if (max_tries > 0 && tries >= max_tries) { stop = 1; break; }
            real Al[36], nb[6], inc[6];
            for (int i = 0; i < 6; ++i) nb[i] = -b[i];          /* :258 */
            memcpy(Al, A, sizeof(Al));
            for (int i = 0; i < 6; ++i) Al[i * 6 + i] *= 1 + lambda;   /* :261 */
            orc_ldlt_solve6(Al, nb, inc);
            incTry++; tries++;
            real qe[4], te[3], qn[4], tn[3], Rn[9];
            orc_se3_exp(inc, qe, te);
            orc_se3_mul(qe, te, q, t, qn, tn);                  /* :266 */
            orc_quat_to_R(qn, Rn);
            real error = calc_error_and_buffers(pts4, n, opt4, cam, Rn, tn, cfg, lvl, ri, &buf);   /* :268 */
            evals++;
            int accepted = (error < lastErr);
            if (trace && ntrace < trace_cap) {
                trace[ntrace].error = (float)error; trace[ntrace].lambda = (float)lambda;
                trace[ntrace].accepted = accepted; trace[ntrace].good = ri->good; trace[ntrace].bad = ri->bad;
                ntrace++;
            }
            if (accepted) {                                     /* :273 */
                memcpy(q, qn, sizeof(qn)); memcpy(t, tn, sizeof(tn));
                if (error / lastErr > cfg->convergence_eps[lvl]) iteration = cfg->max_its_per_lvl[lvl];
                last_residual = lastErr = error;
                if (lambda <= (real)0.2f) lambda = 0;
                else lambda *= cfg->lambda_success_fac;
                break;
            } else {
                real dot = 0;
                for (int i = 0; i < 6; ++i) dot += inc[i] * inc[i];
                if (!(dot > cfg->step_size_min[lvl])) {         /* :294 */
                    iteration = cfg->max_its_per_lvl[lvl];
                    break;
                }
                if (lambda == 0) lambda = (real)0.2f;
                else lambda *= (real)pow((double)cfg->lambda_fail_fac, (double)incTry);   /* :303 std::pow */
            }
        }
    }
    orc_quat_to_R(q, R);                                        /* :308 */
    T[0] = t[0]; T[1] = t[1]; T[2] = t[2];
    if (n_evals) *n_evals = evals;
    free(mem);
    return last_residual;
}

/* TrackerNew::evalCostFunction  system/tracker.cpp:357-393.  dt is the float
 * distance transform of the reference (key) frame at minLvl. */
ORC_API real orc_eval_cost_function(const float *pts4, int n, const float *dt, const orc_cam *cam,
                                    const real *R, const real *T, const orc_opt_cfg *cfg, int lvl)
{
    real total = 0;
    const real fx = cam->fx, fy = cam->fy, cx = cam->cx, cy = cam->cy;
    for (int ir = 0; ir < n; ++ir) {
        const real px = pts4[4 * ir + 0], py = pts4[4 * ir + 1], pz = pts4[4 * ir + 2];
        real nx = (RM(R, 0, 0) * px + RM(R, 0, 1) * py + RM(R, 0, 2) * pz) + T[0];
        real ny = (RM(R, 1, 0) * px + RM(R, 1, 1) * py + RM(R, 1, 2) * pz) + T[1];
        real nz = (RM(R, 2, 0) * px + RM(R, 2, 1) * py + RM(R, 2, 2) * pz) + T[2];
        nx = fx * nx / nz + cx;     /* :378 */
        ny = fy * ny / nz + cy;
        if (nx >= 0 && nx < cam->w && ny >= 0 && ny < cam->h) {
            const real residual = dt[(size_t)floor((double)ny) * cam->w + (size_t)floor((double)nx)];
            if (residual > cfg->edge_distance_lvl[lvl] && cfg->use_edge_filter) continue;   /* :383 */
            total += residual;
        }
    }
    return total;
}

/* One pyramid level as the tracker sees it. */
typedef struct {
    const float *cur_pts4; int cur_n;     /* current frame: return3DEdges(lvl)   */
    const float *ref_opt4;                /* key frame: returnOptimizationStructure(lvl) */
    const float *ref_dt;                  /* key frame: returnDistTransform(lvl) */
    orc_cam cam;
} orc_level;

/* TrackerNew::trackFrames  system/tracker.cpp:294-353 with
 * checkInitializationValues :265-283.  levels[] is indexed by pyramid level.
 * status: 0 OK, 2 NEW_KF (tracker.h:61-66 enum order). evals_per_lvl (size 6, optional). */
ORC_API int orc_track_frames(const orc_level *levels, int min_lvl, int max_lvl, int check_init,
                             const orc_opt_cfg *cfg, real *R, real *T, real *error_out,
                             orc_resinfo *ri_out, int *evals_per_lvl, int *rc_out)
{
    if (check_init) {
        const orc_level *L = &levels[min_lvl];
        real I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z[3] = {0, 0, 0};
        real costEye = orc_eval_cost_function(L->cur_pts4, L->cur_n, L->ref_dt, &L->cam, I, Z, cfg, min_lvl);
        real costInit = orc_eval_cost_function(L->cur_pts4, L->cur_n, L->ref_dt, &L->cam, R, T, cfg, min_lvl);
        if (costEye < costInit) {     /* :277 */
            memcpy(R, I, sizeof(I)); memcpy(T, Z, sizeof(Z));
        }
    }
    real error = (real)INFINITY;
    orc_resinfo ri; memset(&ri, 0, sizeof(ri));
    int rc_all = 0;
    for (int lvl = min_lvl; lvl >= max_lvl; --lvl) {
        const orc_level *L = &levels[lvl];
        int ne = 0, rc = 0;
        error = orc_track_level(L->cur_pts4, L->cur_n, L->ref_opt4, &L->cam, R, T, cfg, lvl, &ri, 0, 0, &ne, 0, &rc);
        if (evals_per_lvl) evals_per_lvl[lvl] = ne;
        if (rc && !rc_all) rc_all = rc;
    }
    if (error_out) *error_out = error;
    if (ri_out) *ri_out = ri;
    if (rc_out) *rc_out = rc_all;
    /* :351  good/bad < 4 -> NEW_KF (double division; bad==0 -> inf -> OK) */
    if ((double)ri.good / (double)ri.bad < 4) return 2;
    return 0;
}

/* orc_track_frames with the LM trace of every level (parity tests: "same trace" = same accept / reject sequence, not only the same
 * number of evaluations).  trace receives the entries of the levels min_lvl .. max_lvl back to back (capacity trace_cap in
 * total), trace_per_lvl[lvl] the number of entries of that level. */
ORC_API int orc_track_frames_traced(const orc_level *levels, int min_lvl, int max_lvl, int check_init,
                                    const orc_opt_cfg *cfg, real *R, real *T, real *error_out,
                                    orc_resinfo *ri_out, int *evals_per_lvl, int *rc_out,
                                    orc_trace_entry *trace, int trace_cap, int *trace_per_lvl)
{
    if (check_init) {
        const orc_level *L = &levels[min_lvl];
        real I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z[3] = {0, 0, 0};
        real costEye = orc_eval_cost_function(L->cur_pts4, L->cur_n, L->ref_dt, &L->cam, I, Z, cfg, min_lvl);
        real costInit = orc_eval_cost_function(L->cur_pts4, L->cur_n, L->ref_dt, &L->cam, R, T, cfg, min_lvl);
        if (costEye < costInit) {     /* tracker.cpp:277 */
            memcpy(R, I, sizeof(I)); memcpy(T, Z, sizeof(Z));
        }
    }
    real error = (real)INFINITY;
    orc_resinfo ri; memset(&ri, 0, sizeof(ri));
    int rc_all = 0, used = 0;
    for (int lvl = min_lvl; lvl >= max_lvl; --lvl) {
        const orc_level *L = &levels[lvl];
        int ne = 0, rc = 0;
        error = orc_track_level(L->cur_pts4, L->cur_n, L->ref_opt4, &L->cam, R, T, cfg, lvl, &ri, trace ? trace + used : 0,
                                trace ? trace_cap - used : 0, &ne, 0, &rc);
        int nt = ne - 1;
        if (nt > trace_cap - used) nt = trace_cap - used;
        if (nt < 0) nt = 0;
        if (trace_per_lvl) trace_per_lvl[lvl] = nt;
        used += nt;
        if (evals_per_lvl) evals_per_lvl[lvl] = ne;
        if (rc && !rc_all) rc_all = rc;
    }
    if (error_out) *error_out = error;
    if (ri_out) *ri_out = ri;
    if (rc_out) *rc_out = rc_all;
    if ((double)ri.good / (double)ri.bad < 4) return 2;   /* tracker.cpp:351 */
    return 0;
}

/* Thread count of the OpenMP batch harness (bench.py pins it explicitly: torch.distributed.run exports OMP_NUM_THREADS=1). */
#ifdef _OPENMP
#include <omp.h>
#endif
ORC_API int orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* Batch of independent frame pairs over OpenMP threads: the CPU-baseline
 * harness ("all host threads the reference path can use": the reference
 * tracker itself is single-threaded per pair). levels is n_pairs*6 entries. */
ORC_API void orc_track_frames_batch(const orc_level *levels, int n_pairs, int min_lvl, int max_lvl, int check_init,
                                    const orc_opt_cfg *cfg, real *R9s, real *T3s, real *errors, int *status,
                                    int *evals /* n_pairs*6 */)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < n_pairs; ++p) {
        int rc;
        status[p] = orc_track_frames(levels + (size_t)p * 6, min_lvl, max_lvl, check_init, cfg, R9s + 9 * (size_t)p,
                                     T3s + 3 * (size_t)p, errors + p, 0, evals ? evals + 6 * (size_t)p : 0, &rc);
    }
}

/* ------------------------------------------------------------------------ */
/* Image pyramid: the reference's hand-written loops                         */
/* (datastructures/imgpyramidrgbd.{h,cpp})                                    */
/* ------------------------------------------------------------------------ */

/* FilterSubsampleWithHoles  imgpyramidrgbd.h:218-249 */
ORC_API void orc_subsample_depth_holes(const float *in, int w_in, int h_in, float *out)
{
    const int w = w_in / 2, h = h_in / 2;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            float acc = 0.f, ngood = 0.f, p;
            p = in[(2 * x + 0) + (size_t)(2 * y + 0) * w_in]; if (p > 0.0f) { acc += p; ngood++; }
            p = in[(2 * x + 1) + (size_t)(2 * y + 0) * w_in]; if (p > 0.0f) { acc += p; ngood++; }
            p = in[(2 * x + 0) + (size_t)(2 * y + 1) * w_in]; if (p > 0.0f) { acc += p; ngood++; }
            p = in[(2 * x + 1) + (size_t)(2 * y + 1) * w_in]; if (p > 0.0f) { acc += p; ngood++; }
            if (ngood > 0) acc /= ngood;
            out[x + (size_t)y * w] = acc;
        }
}

/* generateDistHistogram  imgpyramidrgbd.cpp:146-172.  hist is (h/P)x(w/P) u8
 * (u8 ++ wraps, as cv::Mat_<uchar>).  Returns nonzero-patch fraction.
 * Pixels whose patch index falls outside the (h/P)x(w/P) grid (sizes not
 * divisible by P) are ignored here; the reference would write out of bounds. */
ORC_API float orc_dist_histogram(const uint8_t *edges, int w, int h, int P, uint8_t *hist)
{
    const int hw = w / P, hh = h / P;
    memset(hist, 0, (size_t)hw * hh);
    for (int yy = 0; yy < h; ++yy)
        for (int xx = 0; xx < w; ++xx)
            if (edges[(size_t)yy * w + xx] > 0) {
                const int py = yy / P, px = xx / P;
                if (py < hh && px < hw) hist[(size_t)py * hw + px]++;
            }
    int nz = 0;
    for (int i = 0; i < hw * hh; ++i) nz += hist[i] != 0;
    return (float)nz / (float)(hw * hh);
}

/* fillInEdges  imgpyramidrgbd.cpp:111-145: top = edges of level lvl-1 (after
 * ITS fill-in), hist = this level's histogram (patch P), P_low = patch size
 * of level lvl-1. edges_mod (this level) is updated in place. */
ORC_API void orc_fill_in_edges(const uint8_t *top, int w_top, int h_top, const uint8_t *hist, int hist_w, int hist_h,
                               int P, int P_low, uint8_t *edges_mod, int w, int h)
{
    const int P2 = P * P;
    for (int yy = 0; yy < h_top; ++yy)
        for (int xx = 0; xx < w_top; ++xx) {
            if ((yy % 2 == 1) && (xx % 2 == 1)) {
                const int py = yy / P_low, px = xx / P_low;
                if (py >= hist_h || px >= hist_w) continue;
                if (hist[(size_t)py * hist_w + px] < P2 * 0.05) {       /* int < double compare */
                    if (top[(size_t)yy * w_top + xx] > 0) {
                        const int oy = yy / 2, ox = xx / 2;
                        if (oy < h && ox < w) edges_mod[(size_t)oy * w + ox] = 255;
                    }
                }
            }
        }
}

/* 3-D edge list of addLevelEdge  imgpyramidrgbd.cpp:199-226: COLUMN-major
 * scan (xx outer, yy inner). out4 has capacity w*h float4. Returns count. */
ORC_API int orc_edges3d(const uint8_t *edges, const float *depth, const orc_cam *cam, float dmin, float dmax, float *out4)
{
    int n = 0;
    for (int xx = 0; xx < cam->w; ++xx)
        for (int yy = 0; yy < cam->h; ++yy) {
            const float Z = depth[(size_t)yy * cam->w + xx];
            if (isfinite(Z) && Z > dmin && Z < dmax) {
                if (edges[(size_t)yy * cam->w + xx] > 0) {
                    const float X = Z * (xx - cam->cx) / cam->fx;
                    const float Y = Z * (yy - cam->cy) / cam->fy;
                    out4[4 * (size_t)n + 0] = X; out4[4 * (size_t)n + 1] = Y;
                    out4[4 * (size_t)n + 2] = Z; out4[4 * (size_t)n + 3] = 1.0f;
                    n++;
                }
            }
        }
    return n;
}

/* buildOptimizationStructure  imgpyramidrgbd.cpp:255-276.  The reference
 * leaves rows 0 and h-1 and every .w uninitialised (malloc); the oracle (and
 * the CUDA product) define them as 0.  Row-wrapped neighbours at x=0 / x=w-1
 * are reproduced (those texels are never read by the optimizer: u in (1,w-2)). */
ORC_API void orc_build_opt_structure(const float *dt, int w, int h, float *opt4)
{
    memset(opt4, 0, sizeof(float) * 4 * (size_t)w * h);
    for (size_t i = (size_t)w; i < (size_t)w * (h - 1); ++i) {
        opt4[4 * i + 0] = 0.5f * (dt[i - 1] - dt[i + 1]);
        opt4[4 * i + 1] = 0.5f * (dt[i - w] - dt[i + w]);
        opt4[4 * i + 2] = dt[i];
    }
}

/* ------------------------------------------------------------------------ */
/* OpenCV kernels the reference calls (restated; pinned against cv2 4.13)    */
/* ------------------------------------------------------------------------ */

/* cv::cvtColor(BGR(A)->GRAY) 8U  (imgpyramidrgbd.cpp:53): 15-bit fixed point,
 * Y = (B*3735 + G*19235 + R*9798 + 16384) >> 15  (OpenCV 4.x color_yuv/rgb2gray). */
ORC_API void orc_gray_bgr(const uint8_t *bgr, int w, int h, size_t stride, int ch, uint8_t *gray)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t *row = bgr + (size_t)y * stride;
        for (int x = 0; x < w; ++x) {
            const int B = row[ch * x + 0], G = row[ch * x + 1], Rr = row[ch * x + 2];
            gray[(size_t)y * w + x] = (uint8_t)((B * 3735 + G * 19235 + Rr * 9798 + 16384) >> 15);
        }
    }
}

static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

/* cv::pyrDown 8U (imgpyramidrgbd.cpp:82): separable [1 4 6 4 1], REFLECT_101,
 * dst = ((h+1)/2, (w+1)/2), out = (sum + 128) >> 8. */
ORC_API void orc_pyrdown_u8(const uint8_t *src, int w, int h, uint8_t *dst)
{
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    static const int k[5] = {1, 4, 6, 4, 1};
    int *rowbuf = (int *)malloc(sizeof(int) * (size_t)dw * 5);
    for (int y = 0; y < dh; ++y) {
        for (int r = 0; r < 5; ++r) {
            const int sy = reflect101(2 * y + r - 2, h);
            const uint8_t *s = src + (size_t)sy * w;
            for (int x = 0; x < dw; ++x) {
                int acc = 0;
                for (int c = 0; c < 5; ++c) acc += k[c] * s[reflect101(2 * x + c - 2, w)];
                rowbuf[r * dw + x] = acc;
            }
        }
        for (int x = 0; x < dw; ++x) {
            int acc = 0;
            for (int r = 0; r < 5; ++r) acc += k[r] * rowbuf[r * dw + x];
            dst[(size_t)y * dw + x] = (uint8_t)((acc + 128) >> 8);
        }
    }
    free(rowbuf);
}

/* cv::Canny(gray, edges, t1, t2, 3, true)  (imgpyramidrgbd.cpp:184):
 * 3x3 Sobel (REPLICATE), mag = dx^2+dy^2 (int), NMS with the TG22 fixed-point
 * sector test and asymmetric >/>= comparisons, hysteresis = every 8-connected
 * component of candidates that contains a strong pixel. */
ORC_API void orc_canny(const uint8_t *gray, int w, int h, double t1, double t2, uint8_t *out)
{
    double lo_t = t1 < t2 ? t1 : t2, hi_t = t1 < t2 ? t2 : t1;
    if (lo_t > 32767.0) lo_t = 32767.0;
    if (hi_t > 32767.0) hi_t = 32767.0;
    if (lo_t > 0) lo_t *= lo_t;
    if (hi_t > 0) hi_t *= hi_t;
    const int low = (int)floor(lo_t), high = (int)floor(hi_t);
    const size_t npx = (size_t)w * h;
    short *dx = (short *)malloc(sizeof(short) * npx), *dy = (short *)malloc(sizeof(short) * npx);
    /* mag padded by one zero pixel on every side */
    const int mw = w + 2;
    int *mag = (int *)calloc((size_t)mw * (h + 2), sizeof(int));
    for (int y = 0; y < h; ++y) {
        const uint8_t *r0 = gray + (size_t)(y > 0 ? y - 1 : 0) * w;
        const uint8_t *r1 = gray + (size_t)y * w;
        const uint8_t *r2 = gray + (size_t)(y < h - 1 ? y + 1 : h - 1) * w;
        for (int x = 0; x < w; ++x) {
            const int xm = x > 0 ? x - 1 : 0, xp = x < w - 1 ? x + 1 : w - 1;
            const int gx = (r0[xp] - r0[xm]) + 2 * (r1[xp] - r1[xm]) + (r2[xp] - r2[xm]);
            const int gy = (r2[xm] - r0[xm]) + 2 * (r2[x] - r0[x]) + (r2[xp] - r0[xp]);
            dx[(size_t)y * w + x] = (short)gx;
            dy[(size_t)y * w + x] = (short)gy;
            mag[(size_t)(y + 1) * mw + (x + 1)] = gx * gx + gy * gy;
        }
    }
    /* map: 0 = not an edge, 1 = weak candidate, 2 = strong */
    uint8_t *map = (uint8_t *)calloc(npx, 1);
    int *stack = (int *)malloc(sizeof(int) * npx);
    int sp = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const int *mc = mag + (size_t)(y + 1) * mw + (x + 1);
            const int m = mc[0];
            if (!(m > low)) continue;
            const int xs = dx[(size_t)y * w + x], ys = dy[(size_t)y * w + x];
            const int ax = abs(xs), ay = abs(ys) << 15;
            const int tg22x = ax * 13573;
            int cand = 0;
            if (ay < tg22x) {
                cand = (m > mc[-1] && m >= mc[1]);
            } else {
                const int tg67x = tg22x + (ax << 16);
                if (ay > tg67x) cand = (m > mc[-mw] && m >= mc[mw]);
                else {
                    const int s = ((xs ^ ys) < 0) ? -1 : 1;
                    cand = (m > mc[-mw - s] && m > mc[mw + s]);
                }
            }
            if (cand) {
                if (m > high) { map[(size_t)y * w + x] = 2; stack[sp++] = y * w + x; }
                else map[(size_t)y * w + x] = 1;
            }
        }
    while (sp > 0) {
        const int p = stack[--sp];
        const int y = p / w, x = p % w;
        for (int dyy = -1; dyy <= 1; ++dyy)
            for (int dxx = -1; dxx <= 1; ++dxx) {
                const int yy = y + dyy, xx = x + dxx;
                if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
                if (map[(size_t)yy * w + xx] == 1) { map[(size_t)yy * w + xx] = 2; stack[sp++] = yy * w + xx; }
            }
    }
    for (size_t i = 0; i < npx; ++i) out[i] = map[i] == 2 ? 255 : 0;
    free(dx); free(dy); free(mag); free(map); free(stack);
}

/* cv::distanceTransform(255-edges, CV_DIST_L2, CV_DIST_MASK_PRECISE)
 * (imgpyramidrgbd.cpp:241): exact Euclidean DT, out = sqrtf((float)d2) with d2
 * the integer squared distance to the nearest edge pixel (edges > 0).
 * Meijster's two-scan algorithm in exact integer arithmetic.  An image with
 * no edge pixel yields 65536.0f, the constant cv2 4.13's own trueDistTrans returns
 * (its IPP branch, used only for images < 2^14 px, returns 2^64 instead -- see oracle.py). */
ORC_API void orc_edt_l2(const uint8_t *edges, int w, int h, float *dt)
{
    const int INF = w + h + 1;   /* > any real 1-D distance */
    int *g = (int *)malloc(sizeof(int) * (size_t)w * h);
    for (int x = 0; x < w; ++x) {
        g[x] = edges[x] > 0 ? 0 : INF;
        for (int y = 1; y < h; ++y)
            g[(size_t)y * w + x] = edges[(size_t)y * w + x] > 0 ? 0 : (g[(size_t)(y - 1) * w + x] >= INF ? INF : g[(size_t)(y - 1) * w + x] + 1);
        for (int y = h - 2; y >= 0; --y)
            if (g[(size_t)(y + 1) * w + x] < g[(size_t)y * w + x] && g[(size_t)(y + 1) * w + x] + 1 < g[(size_t)y * w + x])
                g[(size_t)y * w + x] = g[(size_t)(y + 1) * w + x] + 1;
    }
    int *s = (int *)malloc(sizeof(int) * w), *t = (int *)malloc(sizeof(int) * w);
    for (int y = 0; y < h; ++y) {
        const int *gr = g + (size_t)y * w;
        int q = -1;
        /* lower envelope over columns with a finite g */
        for (int u = 0; u < w; ++u) {
            if (gr[u] >= INF) continue;
            const long long gu2 = (long long)gr[u] * gr[u];
            while (q >= 0) {
                const long long i = s[q];
                const long long fi = ((long long)t[q] - i) * ((long long)t[q] - i) + (long long)gr[i] * gr[i];
                const long long fu = ((long long)t[q] - u) * ((long long)t[q] - u) + gu2;
                if (fi > fu) q--; else break;
            }
            if (q < 0) { q = 0; s[0] = u; t[0] = 0; }
            else {
                const long long i = s[q];
                /* Sep(i,u) = floor((u^2 - i^2 + g(u)^2 - g(i)^2) / (2(u-i))) */
                const long long num = (long long)u * u - i * i + gu2 - (long long)gr[i] * gr[i];
                const long long den = 2 * ((long long)u - i);
                long long sep = num >= 0 ? num / den : -((-num + den - 1) / den);
                const long long wq = 1 + sep;
                if (wq < w) { q++; s[q] = u; t[q] = (int)wq; }
            }
        }
        if (q < 0) {
            for (int x = 0; x < w; ++x) dt[(size_t)y * w + x] = 65536.0f;
            continue;
        }
        for (int x = w - 1; x >= 0; --x) {
            const long long i = s[q];
            const long long d2 = ((long long)x - i) * ((long long)x - i) + (long long)gr[i] * gr[i];
            dt[(size_t)y * w + x] = sqrtf((float)d2);
            if (x == t[q]) q--;
        }
    }
    free(g); free(s); free(t);
}
