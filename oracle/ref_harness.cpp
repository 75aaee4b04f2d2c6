// ref_harness.cpp -- TEST INFRASTRUCTURE (oracle/): a C ABI around the REFERENCE'S OWN hot-path classes, compiled from the
// sources where they lie under /root/reference (never copied into this repository):
//     datastructures/imgpyramidrgbd.{h,cpp}, datastructures/camerapyr.h      ImgPyramidRGBD, CameraPyr, ImgPyramidSettings
//     system/optimizer.{h,cpp}, utils/LGSX.h                                  Optimizer, LGS6
//     system/tracker.{h,cpp}                                                  TrackerNew
//     utils/Logging.{h,cpp}, utils/timer.h
// against the API shims of oracle/shim/ (Eigen, OpenCV and Sophus are not in this image; the four OpenCV kernels and the SE3
// arithmetic are served by oracle/revo_oracle.c, which is pinned separately against cv2 4.13 and the reference's sympy
// Sophus).  Built by `make -C oracle ref` into oracle/_ref/librevo_ref.so when /root/reference is present; used by
// tests/test_oracle_ref.py to pin the restatement of oracle/revo_oracle.c + oracle/oracle.py (the checker of the CUDA path)
// to what the reference's code really computes, and to generate the golden vectors of tests/golden/ref_*.npz.
// Private members are reached with the usual test-only `#define private public`.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>

#define private public
#define protected public
#include "system/tracker.h"
#undef private
#undef protected

namespace {

// the reference logs through std::cout at "info" level from inside the loops (and resets its own threshold,
// optimizer.cpp:78,135): swallow it
std::string g_last_log;      // what the reference printed during the last call (the LM trace is read from it)
struct Quiet {
    std::streambuf *old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) { LOG_THRESHOLD(i3d::nothing); LOG_FORMAT(false, false, false, false); }
    ~Quiet() { std::cout.rdbuf(old); g_last_log = sink.str(); }
};

struct RefSettings {
    ImgPyramidSettings pyr;
    std::shared_ptr<CameraPyr> cams;
    RefSettings(int w, int h, float fx, float fy, float cx, float cy, int n_levels, int canny1, int canny2, float dmin, float dmax,
                int use_hist, float n_percentage)
        : pyr(std::string(""))
    {
        pyr.PYR_MAX_LVL = 0;
        pyr.PYR_MIN_LVL = n_levels - 1;
        pyr.width = (size_t)w;
        pyr.height = (size_t)h;
        pyr.K = Eigen::Matrix3f::Identity();
        pyr.K(0, 0) = fx; pyr.K(1, 1) = fy; pyr.K(0, 2) = cx; pyr.K(1, 2) = cy;
        pyr.cannyThreshold1 = canny1;
        pyr.cannyThreshold2 = canny2;
        pyr.DEPTH_MIN = dmin;
        pyr.DEPTH_MAX = dmax;
        pyr.USE_EDGE_HIST = use_hist != 0;
        pyr.nPercentage = n_percentage;
        pyr.DO_UNDISTORT = false;
        pyr.DO_GAUSSIAN_SMOOTHING_BEFORE_CANNY = false;
        cams = std::make_shared<CameraPyr>(pyr);
    }
};

struct RefPyr {
    std::shared_ptr<RefSettings> st;
    std::shared_ptr<ImgPyramidRGBD> p;
};

}  // namespace

extern "C" {

#define REF_API __attribute__((visibility("default")))

// the reference's own log output of the last call (optimizer.cpp:270: one "goodPts: .. error = .." line per LM try)
REF_API long ref_last_log(char *dst, long cap)
{
    const long n = (long)g_last_log.size();
    if (dst && cap > 0) {
        const long m = n < cap - 1 ? n : cap - 1;
        std::memcpy(dst, g_last_log.data(), (size_t)m);
        dst[m] = 0;
    }
    return n;
}

// ---- ImgPyramidRGBD ------------------------------------------------------------------------------------------------
REF_API void *ref_pyr_create(int w, int h, float fx, float fy, float cx, float cy, int n_levels, int canny1, int canny2, float dmin,
                             float dmax, int use_hist, float n_percentage, const uint8_t *bgr, int channels, const float *depth,
                             double timestamp)
{
    Quiet q;
    RefPyr *r = new RefPyr();
    r->st = std::make_shared<RefSettings>(w, h, fx, fy, cx, cy, n_levels, canny1, canny2, dmin, dmax, use_hist, n_percentage);
    cv::Mat rgb(h, w, channels == 4 ? CV_8UC4 : CV_8UC3), d(h, w, CV_32FC1);
    std::memcpy(rgb.data, bgr, (size_t)w * h * channels);
    std::memcpy(d.data, depth, (size_t)w * h * 4);
    r->p = std::make_shared<ImgPyramidRGBD>(r->st->pyr, r->st->cams, rgb, d, timestamp);      // imgpyramidrgbd.cpp:43-96
    return r;
}

REF_API void ref_pyr_make_keyframe(void *h)
{
    Quiet q;
    ((RefPyr *)h)->p->makeKeyframe();                                                         // imgpyramidrgbd.cpp:231-252
}

REF_API void ref_pyr_destroy(void *h) { delete (RefPyr *)h; }

REF_API int ref_pyr_level_size(void *h, int lvl, int *w, int *hh, float *cam4)
{
    const Camera &c = ((RefPyr *)h)->p->cameraPyr->at(lvl);
    *w = (int)c.width; *hh = (int)c.height;
    cam4[0] = c.fx; cam4[1] = c.fy; cam4[2] = c.cx; cam4[3] = c.cy;
    return 0;
}

REF_API int ref_pyr_num_edges(void *h, int lvl) { return (int)((RefPyr *)h)->p->return3DEdges((uint)lvl).cols(); }

// which: 0 gray, 1 depth, 2 edges (after fill-in), 3 edgesOrig, 4 hist, 5 edges3D (4 x N, column-major = N x 4 rows), 6 dt,
// 7 optimizationStructure (h*w float4; rows 0 and h-1 and .w are uninitialised malloc memory in the reference: zeroed here)
REF_API long ref_pyr_get(void *h, int lvl, int which, void *dst, long dst_bytes)
{
    ImgPyramidRGBD &p = *((RefPyr *)h)->p;
    const cv::Mat *m = nullptr;
    switch (which) {
        case 0: m = &p.grayPyr.at(lvl); break;
        case 1: m = &p.depthPyr.at(lvl); break;
        case 2: m = &p.edgesPyr.at(lvl); break;
        case 3: m = &p.edgesOrigPyr.at(lvl); break;
        case 4: m = &p.histPyr.at(lvl); break;
        case 6: m = &p.dtPyr.at(lvl); break;
        default: break;
    }
    if (m) {
        const long n = (long)(m->total() * m->elemSize());
        if (dst && dst_bytes >= n) std::memcpy(dst, m->data, n);
        return n;
    }
    if (which == 5) {
        const Eigen::MatrixXf &e = p.return3DEdges((uint)lvl);
        const long n = (long)e.cols() * 16;
        if (dst && dst_bytes >= n) std::memcpy(dst, e.data(), n);
        return n;
    }
    if (which == 7) {
        const Camera &c = p.cameraPyr->at(lvl);
        const long n = (long)c.width * c.height * 16;
        if (dst && dst_bytes >= n) {
            const float *src = (const float *)p.returnOptimizationStructure((uint)lvl);
            float *o = (float *)dst;
            std::memcpy(o, src, n);
            const size_t w = c.width, hh = c.height;
            for (size_t i = 0; i < w * hh; ++i) {
                o[4 * i + 3] = 0.f;
                if (i < w || i >= w * (hh - 1)) o[4 * i] = o[4 * i + 1] = o[4 * i + 2] = 0.f;
            }
        }
        return n;
    }
    return -1;
}

// generateColoredPcl (imgpyramidrgbd.cpp:279-327): returns the number of columns; dst receives 8 x N column-major
REF_API long ref_pyr_colored_pcl(void *h, int lvl, int dense, float *dst, long dst_floats)
{
    Quiet q;
    Eigen::MatrixXf clr;
    ((RefPyr *)h)->p->generateColoredPcl((uint)lvl, clr, dense != 0);
    const long n = (long)clr.cols();
    if (dst && dst_floats >= n * 8) std::memcpy(dst, clr.data(), (size_t)n * 8 * 4);
    return n;
}

// ---- Optimizer -------------------------------------------------------------------------------------------------------
struct ref_opt_cfg {      // the fields of OptimizerSettings the loop reads (optimizer.h:87-111), same layout as oracle.OptCfg
    float lambda_success_fac, lambda_fail_fac;
    float lambda_initial[6], step_size_min[6], convergence_eps[6];
    int max_its_per_lvl[6];
    float edge_distance_lvl[6];
    float huber_edge;
    int use_edge_filter;
};

static OptimizerSettings to_settings(const ref_opt_cfg *c)
{
    OptimizerSettings s;
    s.lambdaSuccessFac = c->lambda_success_fac;
    s.lambdaFailFac = c->lambda_fail_fac;
    for (int l = 0; l < 6; ++l) {
        s.lambdaInitial[l] = c->lambda_initial[l];
        s.stepSizeMin[l] = c->step_size_min[l];
        s.convergenceEps[l] = c->convergence_eps[l];
        s.maxItsPerLvl[l] = c->max_its_per_lvl[l];
        s.edgeDistanceLvl[l] = c->edge_distance_lvl[l];
    }
    s.huber_edge = c->huber_edge;
    s.USE_EDGE_FILTER = c->use_edge_filter != 0;
    s.maxImgSize = cv::Size2i(1920, 1080);      // the 7 SoA buffers (optimizer.cpp:48-60); the default 640x480 is the reference's
    return s;
}

// Optimizer::trackFrames (optimizer.cpp:235-311) on level lvl.  R9 column-major (Eigen::Matrix3f::data()), in/out.
REF_API float ref_opt_track_level(void *ref, void *cur, const ref_opt_cfg *cfg, int lvl, float *R9, float *t3, int *good, int *bad,
                                  float *sum_w, float *sum_unw)
{
    Quiet q;
    Optimizer opt(to_settings(cfg));
    Eigen::Matrix3f R;
    Eigen::Vector3f T;
    std::memcpy(R.data(), R9, 36);
    std::memcpy(T.data(), t3, 12);
    Optimizer::ResidualInfo ri;
    const float err = opt.trackFrames(((RefPyr *)ref)->p, ((RefPyr *)cur)->p, R, T, lvl, ri);
    std::memcpy(R9, R.data(), 36);
    std::memcpy(t3, T.data(), 12);
    *good = ri.goodPtsEdges; *bad = ri.badPtsEdges; *sum_w = ri.sumErrorWeighted; *sum_unw = ri.sumErrorUnweighted;
    return err;
}

// One evaluation: calcErrorAndBuffers (PASS A, optimizer.cpp:74-191) + calculateWarpUpdate (PASS B, :192-234) with
// LGS6::initialize / update / finish (LGSX.h:196-204,392-398,320-326).  A36: ls.A (6x6, as finished: divided by n),
// b6: ls.b (NEGATIVE sum, divided by n), returns the mean weighted error.
REF_API float ref_opt_eval(void *ref, void *cur, const ref_opt_cfg *cfg, int lvl, const float *R9, const float *t3, int *good, int *bad,
                           float *sum_w, float *sum_unw, float *A36, float *b6, float *ls_error)
{
    Quiet q;
    Optimizer opt(to_settings(cfg));
    Eigen::Matrix3f R;
    Eigen::Vector3f T;
    std::memcpy(R.data(), R9, 36);
    std::memcpy(T.data(), t3, 12);
    Optimizer::ResidualInfo ri;
    const float err = opt.calcErrorAndBuffers(((RefPyr *)ref)->p, ((RefPyr *)cur)->p, R, T, ri, (uint)lvl, true);
    lsd_slam::LGS6 ls;
    opt.calculateWarpUpdate(ls, ri.goodPtsEdges);
    *good = ri.goodPtsEdges; *bad = ri.badPtsEdges; *sum_w = ri.sumErrorWeighted; *sum_unw = ri.sumErrorUnweighted;
    std::memcpy(A36, ls.A.data(), 36 * 4);
    std::memcpy(b6, ls.b.data(), 6 * 4);
    *ls_error = ls.error;
    return err;
}

// ---- TrackerNew ----------------------------------------------------------------------------------------------------
struct RefTracker {
    std::shared_ptr<RefSettings> st;
    std::unique_ptr<TrackerNew> trk;
};

REF_API void *ref_tracker_create(void *any_pyr, const ref_opt_cfg *cfg, int check_init_values, int check_tracking_results,
                                 int n_frames_voting)
{
    Quiet q;
    RefTracker *t = new RefTracker();
    t->st = ((RefPyr *)any_pyr)->st;
    TrackerSettings ts{std::string("")};
    ts.CHECK_INIT_VALUES = check_init_values != 0;
    ts.CHECK_TRACKING_RESULTS = check_tracking_results != 0;
    ts.nFramesHistogramVoting = n_frames_voting;
    ts.optimizerSettings = to_settings(cfg);
    t->trk.reset(new TrackerNew(ts, t->st->pyr));
    return t;
}
REF_API void ref_tracker_destroy(void *h) { delete (RefTracker *)h; }

// TrackerNew::trackFrames (tracker.cpp:294-353): returns TrackerStatus
REF_API int ref_tracker_track(void *h, void *ref, void *cur, float *R9, float *t3, float *error)
{
    Quiet q;
    Eigen::Matrix3f R;
    Eigen::Vector3f T;
    std::memcpy(R.data(), R9, 36);
    std::memcpy(T.data(), t3, 12);
    float err = 0.f;
    const int st = (int)((RefTracker *)h)->trk->trackFrames(R, T, err, ((RefPyr *)ref)->p, ((RefPyr *)cur)->p);
    std::memcpy(R9, R.data(), 36);
    std::memcpy(t3, T.data(), 12);
    *error = err;
    return st;
}
// evalCostFunction (tracker.cpp:357-393)
REF_API float ref_tracker_eval_cost(void *h, void *ref, void *cur, const float *R9, const float *t3, int min_lvl)
{
    Quiet q;
    Eigen::Matrix3f R;
    Eigen::Vector3f T;
    std::memcpy(R.data(), R9, 36);
    std::memcpy(T.data(), t3, 12);
    return ((RefTracker *)h)->trk->evalCostFunction(R, T, (uint)min_lvl, ((RefPyr *)cur)->p, ((RefPyr *)ref)->p);
}
// addOldPclAndPose (tracker.cpp:209-224) with return3DEdges(histogramLevel) of `pyr`, as REVO::start does (system.cpp:173,259)
REF_API void ref_tracker_add_old(void *h, void *pyr, const float *world_pose16, double ts)
{
    Quiet q;
    RefTracker *t = (RefTracker *)h;
    Eigen::Matrix4f P;
    std::memcpy(P.data(), world_pose16, 64);
    t->trk->addOldPclAndPose(((RefPyr *)pyr)->p->return3DEdges((uint)t->trk->histogramLevel), P, ts);
}
REF_API void ref_tracker_clear_past(void *h) { ((RefTracker *)h)->trk->clearUpPastLists(); }
REF_API int ref_tracker_num_past(void *h) { return (int)((RefTracker *)h)->trk->mPastPcl.size(); }
// assessTrackingQuality (tracker.cpp:118-201): returns TrackerStatus
REF_API int ref_tracker_assess(void *h, void *cur, const float *estimated_pose16)
{
    Quiet q;
    LOG_THRESHOLD(i3d::info);      // the histogram / overlap counts are read from the reference's own log lines
    Eigen::Matrix4f P;
    std::memcpy(P.data(), estimated_pose16, 64);
    ((RefPyr *)cur)->p->frameId = 0;
    return (int)((RefTracker *)h)->trk->assessTrackingQuality(P, ((RefPyr *)cur)->p);
}

// ---- LGS6 on its own (utils/LGSX.h:196-204,392-398,320-326) ---------------------------------------------------------
REF_API void ref_lgs6(const float *J6, const float *res, const float *w, int n, float *A36, float *b6, float *error, int finish)
{
    lsd_slam::LGS6 ls;
    ls.initialize((size_t)n);
    for (int i = 0; i < n; ++i) {
        lsd_slam::Vector6 v;
        for (int k = 0; k < 6; ++k) v[k] = J6[6 * i + k];
        ls.update(v, res[i], w[i]);
    }
    if (finish) ls.finish();
    std::memcpy(A36, ls.A.data(), 36 * 4);
    std::memcpy(b6, ls.b.data(), 6 * 4);
    *error = ls.error;
}

// Optimizer::getInterpolatedElement43 (optimizer.h:173-185)
REF_API void ref_interp43(const float *opt4, int width, float x, float y, float *out3)
{
    OptimizerSettings s;
    Optimizer opt(s);
    const Eigen::Vector3f v = opt.getInterpolatedElement43((const Eigen::Vector4f *)opt4, x, y, width);
    out3[0] = v[0]; out3[1] = v[1]; out3[2] = v[2];
}

}  // extern "C"
