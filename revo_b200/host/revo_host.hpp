// revo_host.hpp -- C++ host-side mirror of the reference's classes for the hot path, on top of the C ABI
// (include/revo_b200.h).  Header only.  Same class names, method names, argument meaning and defaults as
//   datastructures/camerapyr.h      (ImgPyramidSettings, Camera, CameraPyr)
//   datastructures/imgpyramidrgbd.h (ImgPyramidRGBD)
//   system/optimizer.h              (OptimizerSettings, Optimizer, Optimizer::ResidualInfo)
//   system/tracker.h                (TrackerSettings, TrackerNew)
// of fabianschenk/REVO, so that system/system.cpp compiles against it with the three includes swapped.
//
// Differences that are deliberate:
//  * where the reference exit(0)s / assert()s / lets Sophus abort(), these classes throw revo::Error carrying the
//    C-ABI status code (REVO_ERR_NOT_KEYFRAME, REVO_ERR_BAD_LEVEL, REVO_ERR_NOT_ORTHOGONAL, ...);
//  * images live in HBM: the cv::Mat / Eigen accessors return host COPIES (revo_pyr_download);
//  * Eigen / OpenCV are optional: without them the POD types revo::Mat3f / revo::Vec3f / revo::Image are used
//    (this image has neither Eigen nor OpenCV C++ headers).  Define REVO_HOST_WITH_EIGEN / REVO_HOST_WITH_SOPHUS / REVO_HOST_WITH_OPENCV
//    to get the reference's exact signatures (Eigen::Matrix3f&, cv::Mat) as overloads.
#pragma once

#include <cmath>
#include <cstring>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/revo_b200.h"

#ifdef REVO_HOST_WITH_EIGEN
#include <Eigen/Core>
#endif
#ifdef REVO_HOST_WITH_SOPHUS   // implies REVO_HOST_WITH_EIGEN
#include <sophus/se3.hpp>
#endif
#ifdef REVO_HOST_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

namespace revo {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &what) : std::runtime_error("revo_b200: " + what), code(c) {}
};

// Column-major 3x3, exactly Eigen::Matrix3f's storage.
struct Mat3f {
    float m[9];
    static Mat3f Identity() { Mat3f r{}; r.m[0] = r.m[4] = r.m[8] = 1.f; return r; }
    float &operator()(int i, int j) { return m[j * 3 + i]; }
    float operator()(int i, int j) const { return m[j * 3 + i]; }
    float *data() { return m; }
    const float *data() const { return m; }
};
struct Vec3f {
    float v[3];
    static Vec3f Zero() { return Vec3f{{0.f, 0.f, 0.f}}; }
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float *data() { return v; }
    const float *data() const { return v; }
};
// Pose as Sophus::SE3f stores it (se3.hpp: unit quaternion in Eigen coefficient order x, y, z, w + translation).  The
// conversions run in the C ABI (revo_quat_to_R9 / revo_R9_to_quat); a matrix that is not a rotation throws
// revo::Error(REVO_ERR_NOT_ORTHOGONAL) where Sophus' constructor would abort() (so3.hpp:419-424).
struct SE3f {
    float q[4];
    float t[3];
    static SE3f Identity() { return SE3f{{0.f, 0.f, 0.f, 1.f}, {0.f, 0.f, 0.f}}; }
};
template <typename T>
struct Image {   // minimal stand-in for cv::Mat_<T> (row-major, tight)
    int rows = 0, cols = 0, channels = 1;
    std::vector<T> data;
    T &at(int y, int x, int c = 0) { return data[((size_t)y * cols + x) * channels + c]; }
    const T &at(int y, int x, int c = 0) const { return data[((size_t)y * cols + x) * channels + c]; }
};

class Context {
public:
    explicit Context(int device = 0) {
        int rc = revo_ctx_create(device, &h_);
        if (rc) throw Error(rc, revo_strerror(rc));
    }
    ~Context() { revo_ctx_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    revo_ctx *handle() const { return h_; }
    // pre-size the device memory pool for the steady state of a stream (revo_ctx_reserve)
    void reserve(size_t bytes) const { check(revo_ctx_reserve(h_, bytes)); }
    // 0 = automatic, 1 = one thread-block cluster per pair, 2 = chip-wide task queue (revo_ctx_set_track_engine)
    void setTrackEngine(int engine, int chunk_points = 0) const { check(revo_ctx_set_track_engine(h_, engine, chunk_points)); }
    void synchronize() const { check(revo_ctx_synchronize(h_)); }
    void check(int rc) const {
        if (!rc) return;
        std::string msg = revo_strerror(rc);
        if (rc == REVO_ERR_CUDA) msg += std::string(": ") + revo_last_error(h_);
        throw Error(rc, msg);
    }
private:
    revo_ctx *h_ = nullptr;
};

}  // namespace revo

// ------------------------------------------------------------------------------------------------------------
// datastructures/camerapyr.h
// ------------------------------------------------------------------------------------------------------------
class ImgPyramidSettings {   // camerapyr.h:27-89 (YAML parsing is the caller's business; same fields, same defaults)
public:
    int PYR_MIN_LVL = 2, PYR_MAX_LVL = 0;
    float DEPTH_MIN = 0.1f, DEPTH_MAX = 5.2f;
    float fx = 560.f, fy = 560.f, cx = 320.f, cy = 240.f;   // K
    size_t width = 640, height = 480;
    int cannyThreshold1 = 150, cannyThreshold2 = 100;
    bool DO_UNDISTORT = false;      // dead in the reference (distCoeff never populated)
    bool USE_EDGE_HIST = true;
    float nPercentage = 0.3f;
    inline int nLevels() const { return PYR_MIN_LVL - PYR_MAX_LVL + 1; }   // camerapyr.h:68-71

    revo_pyr_config c_config() const {
        revo_pyr_config c;
        revo_pyr_config_default(&c);
        c.n_levels = nLevels(); c.canny_threshold1 = cannyThreshold1; c.canny_threshold2 = cannyThreshold2;
        c.depth_min = DEPTH_MIN; c.depth_max = DEPTH_MAX; c.use_edge_hist = USE_EDGE_HIST; c.n_percentage = nPercentage;
        return c;
    }
    revo_camera c_camera() const { return revo_camera{fx, fy, cx, cy, (int32_t)width, (int32_t)height}; }
};

class Camera {   // camerapyr.h:90-111
public:
    Camera(float fx, float fy, float cx, float cy, size_t width, size_t height)
        : fx(fx), fy(fy), cx(cx), cy(cy), width(width), height(height), area(width * height) {}
    Camera(float fx, float fy, float cx, float cy, size_t width, size_t height, float scale)
        : fx(fx * scale), fy(fy * scale), cx(cx * scale), cy(cy * scale), width(static_cast<size_t>(width * scale)),
          height(static_cast<size_t>(height * scale)), area(this->height * this->width) {}
    float fx, fy, cx, cy;
    size_t width, height, area;
};

class CameraPyr {   // camerapyr.h:113-193 (the unused mPclTemplate is not built)
public:
    explicit CameraPyr(const ImgPyramidSettings &s) {
        const int nLevels = s.nLevels();
        if (nLevels <= 0) return;
        camPyr.push_back(Camera(s.fx, s.fy, s.cx, s.cy, s.width, s.height));
        for (int lvl = 1; lvl <= nLevels; ++lvl) {
            const float scale = 1.0f / (float)std::pow(2, lvl);
            camPyr.push_back(Camera(s.fx, s.fy, s.cx, s.cy, s.width, s.height, scale));
        }
    }
    inline int size() const { return (int)camPyr.size(); }
    const Camera &at(int lvl) const { return camPyr.at(lvl); }
    std::vector<Camera> camPyr;
};

// ------------------------------------------------------------------------------------------------------------
// datastructures/imgpyramidrgbd.h
// ------------------------------------------------------------------------------------------------------------
class ImgPyramidRGBD {
public:
    // ImgPyramidRGBD(settings, cameraPyr, fullResRgb, fullResDepth, timestamp)  imgpyramidrgbd.h:39-41
    // rgb: 8UC3/8UC4 BGR(A), stride in bytes (0 = tight); depth: 32FC1 metres.
    ImgPyramidRGBD(const std::shared_ptr<revo::Context> &ctx, const ImgPyramidSettings &settings,
                   const std::shared_ptr<CameraPyr> &cameraPyr, const uint8_t *rgb, size_t rgb_stride, int channels,
                   const float *depth, size_t depth_stride, double timestamp)
        : cameraPyr(cameraPyr), frameId(0), ctx_(ctx), mSettings(settings) {
        const revo_pyr_config cfg = settings.c_config();
        const revo_camera cam = settings.c_camera();
        ctx_->check(revo_pyr_create(ctx_->handle(), &cfg, &cam, rgb, rgb_stride, channels, depth, depth_stride, timestamp, &h_));
        ctx_->check(revo_ctx_synchronize(ctx_->handle()));
        std::memset(T_w_f, 0, sizeof(T_w_f));
        T_w_f[0] = T_w_f[5] = T_w_f[10] = T_w_f[15] = 1.f;
    }
#ifdef REVO_HOST_WITH_OPENCV
    ImgPyramidRGBD(const std::shared_ptr<revo::Context> &ctx, const ImgPyramidSettings &settings,
                   const std::shared_ptr<CameraPyr> &cameraPyr, const cv::Mat &fullResRgb, const cv::Mat &fullResDepth,
                   const double timestamp)
        : ImgPyramidRGBD(ctx, settings, cameraPyr, fullResRgb.data, fullResRgb.step, fullResRgb.channels(),
                         (const float *)fullResDepth.data, fullResDepth.step, timestamp) {}
#endif
    ~ImgPyramidRGBD() { if (h_) revo_pyr_destroy(ctx_->handle(), h_); }
    ImgPyramidRGBD(const ImgPyramidRGBD &) = delete;
    ImgPyramidRGBD &operator=(const ImgPyramidRGBD &) = delete;

    std::shared_ptr<CameraPyr> cameraPyr;
    int frameId;

    void makeKeyframe() { ctx_->check(revo_pyr_make_keyframe(ctx_->handle(), h_)); }   // imgpyramidrgbd.cpp:231-252

    // ---- return methods, imgpyramidrgbd.h:45-117 (host copies) ----
    revo::Mat3f returnK(unsigned lvl) const {
        revo_camera c = cam(lvl);
        revo::Mat3f K = revo::Mat3f::Identity();
        K(0, 0) = c.fx; K(1, 1) = c.fy; K(0, 2) = c.cx; K(1, 2) = c.cy;
        return K;
    }
    revo::Image<float> returnDistTransform(unsigned lvl) const { return fetch<float>(lvl, REVO_ARRAY_DT, 1); }
    revo::Image<uint8_t> returnEdges(unsigned lvl) const { return fetch<uint8_t>(lvl, REVO_ARRAY_EDGES, 1); }
    revo::Image<uint8_t> returnOrigEdges(unsigned lvl) const {
        if (mSettings.USE_EDGE_HIST && int(lvl) > mSettings.PYR_MAX_LVL) return fetch<uint8_t>(lvl, REVO_ARRAY_EDGES_ORIG, 1);
        return returnEdges(lvl);
    }
    revo::Image<float> returnDepth(unsigned lvl) const { return fetch<float>(lvl, REVO_ARRAY_DEPTH, 1); }
    revo::Image<uint8_t> returnGray(unsigned lvl) const { return fetch<uint8_t>(lvl, REVO_ARRAY_GRAY, 1); }
    // 4 x N, column-major (= N consecutive float4), in the reference's column-major scan order
    std::vector<float> return3DEdges(unsigned lvl) const { return fetch_vec<float>(lvl, REVO_ARRAY_EDGES3D); }
    // w*h float4 {gx, gy, dt, 0}; throws REVO_ERR_NOT_KEYFRAME where the reference exit(0)s (imgpyramidrgbd.h:113-117)
    std::vector<float> returnOptimizationStructure(unsigned lvl) const { return fetch_vec<float>(lvl, REVO_ARRAY_OPTSTRUCT); }
    inline double returnTimestamp() const { return revo_pyr_timestamp(h_); }
    inline unsigned returnMaxLvl() const { return mSettings.PYR_MAX_LVL; }
    inline unsigned returnMinLvl() const { return mSettings.PYR_MIN_LVL; }
    int return3DEdgesCount(unsigned lvl) const {
        int n = 0;
        ctx_->check(revo_pyr_num_edges(ctx_->handle(), h_, (int)lvl, &n));
        return n;
    }
    // generateColoredPcl(lvl, clrPcl, densePcl), imgpyramidrgbd.cpp:279-327: 8 x N column-major, compacted on the device.  The
    // reference reads its own clone of the colour image (rgbFullSize); here the caller hands the same image in again.
    std::vector<float> generateColoredPcl(unsigned lvl, const uint8_t *rgb, int channels, bool densePcl = false) const {
        int n = 0;
        ctx_->check(revo_pyr_colored_pcl(ctx_->handle(), h_, (int)lvl, densePcl, rgb, channels, nullptr, 0, &n));
        std::vector<float> pcl((size_t)n * 8);
        if (n) ctx_->check(revo_pyr_colored_pcl(ctx_->handle(), h_, (int)lvl, densePcl, rgb, channels, pcl.data(), (size_t)n, &n));
        return pcl;
    }
    // pose bookkeeping of the keyframe (row-major 4x4 here; the reference keeps Eigen::Matrix4f)
    void setTwf(const float T[16]) { std::memcpy(T_w_f, T, sizeof(T_w_f)); }
    const float *getTransKFtoWorld() const { return T_w_f; }
    void prepareKfForStorage() {}   // effectively a no-op in the reference too (imgpyramidrgbd.h:158-160)
    inline bool isPointOkDepth(const float d) const { return std::isfinite(d) && d > mSettings.DEPTH_MIN && d < mSettings.DEPTH_MAX; }

    revo_pyr *handle() const { return h_; }
    const std::shared_ptr<revo::Context> &context() const { return ctx_; }

private:
    revo_camera cam(unsigned lvl) const {
        revo_camera c;
        ctx_->check(revo_pyr_level_camera(h_, (int)lvl, &c));
        return c;
    }
    template <typename T>
    std::vector<T> fetch_vec(unsigned lvl, int which) const {
        size_t bytes = 0;
        ctx_->check(revo_pyr_download(ctx_->handle(), h_, (int)lvl, which, nullptr, 0, &bytes));
        std::vector<T> v(bytes / sizeof(T));
        if (bytes) ctx_->check(revo_pyr_download(ctx_->handle(), h_, (int)lvl, which, v.data(), bytes, nullptr));
        return v;
    }
    template <typename T>
    revo::Image<T> fetch(unsigned lvl, int which, int channels) const {
        revo::Image<T> im;
        revo_camera c = cam(lvl);
        im.rows = c.height; im.cols = c.width; im.channels = channels;
        im.data = fetch_vec<T>(lvl, which);
        return im;
    }
    std::shared_ptr<revo::Context> ctx_;
    ImgPyramidSettings mSettings;
    revo_pyr *h_ = nullptr;
    float T_w_f[16];
};

// ------------------------------------------------------------------------------------------------------------
// system/optimizer.h
// ------------------------------------------------------------------------------------------------------------
#define PYRAMID_LEVELS 6
class OptimizerSettings {   // optimizer.h:42-112 (fields the hot path reads, same defaults)
public:
    OptimizerSettings() {
        lambdaSuccessFac = 0.5f; lambdaFailFac = 2.0f;
        const int edgeDistance[6] = {30, 20, 10, 5, 5, 5};
        for (int l = 0; l < PYRAMID_LEVELS; ++l) {
            lambdaInitial[l] = 0; stepSizeMin[l] = 1e-16f; convergenceEps[l] = 0.999f; maxItsPerLvl[l] = 100;
            edgeDistanceLvl[l] = (float)edgeDistance[l];
        }
        maxIncTry = 10; huber_edge = 0.3f; USE_EDGE_FILTER = false; nPyrLvl = 3;
    }
    float lambdaSuccessFac, lambdaFailFac;
    float lambdaInitial[PYRAMID_LEVELS], stepSizeMin[PYRAMID_LEVELS], convergenceEps[PYRAMID_LEVELS];
    int maxItsPerLvl[PYRAMID_LEVELS];
    float edgeDistanceLvl[PYRAMID_LEVELS];
    int maxIncTry;
    float huber_edge;
    bool USE_EDGE_FILTER;
    int nPyrLvl;

    revo_opt_config c_config() const {
        revo_opt_config c;
        revo_opt_config_default(&c);
        c.lambda_success_fac = lambdaSuccessFac; c.lambda_fail_fac = lambdaFailFac;
        for (int l = 0; l < PYRAMID_LEVELS; ++l) {
            c.lambda_initial[l] = lambdaInitial[l]; c.step_size_min[l] = stepSizeMin[l]; c.convergence_eps[l] = convergenceEps[l];
            c.max_its_per_lvl[l] = maxItsPerLvl[l]; c.edge_distance_lvl[l] = edgeDistanceLvl[l];
        }
        c.huber_edge = huber_edge; c.use_edge_filter = USE_EDGE_FILTER;
        return c;
    }
};

class Optimizer {
public:
    class ResidualInfo {   // optimizer.h:117-139
    public:
        ResidualInfo() { clearAll(); }
        int goodPtsEdges, badPtsEdges, badOutOfBounds;
        float sumErrorUnweighted, sumErrorWeighted, sumSignedRes;
        void clearCountings() { goodPtsEdges = badPtsEdges = 0; }
        void clearErrors() { sumErrorUnweighted = sumErrorWeighted = sumSignedRes = 0.0f; }
        void clearAll() { clearCountings(); clearErrors(); badOutOfBounds = 0; }
    };
    Optimizer(const std::shared_ptr<revo::Context> &ctx, const OptimizerSettings &settings) : ctx_(ctx), mSettings(settings) {}

    // float trackFrames(ref, cur, R&, T&, lvl, resInfo&)   optimizer.h:168-169 / optimizer.cpp:235-311
    float trackFrames(const std::shared_ptr<ImgPyramidRGBD> &refFrame, const std::shared_ptr<ImgPyramidRGBD> &currFrame,
                      revo::Mat3f &R, revo::Vec3f &T, int lvl, ResidualInfo &resInfo) {
        const revo_opt_config cfg = mSettings.c_config();
        revo_residual_info ri;
        float err = 0.f;
        int n_evals = 0;
        ctx_->check(revo_track_level(ctx_->handle(), &cfg, refFrame->handle(), currFrame->handle(), lvl, R.data(), T.data(), &ri, &err,
                                     &n_evals));
        resInfo.goodPtsEdges = ri.good_pts_edges; resInfo.badPtsEdges = ri.bad_pts_edges;
        resInfo.sumErrorUnweighted = ri.sum_error_unweighted; resInfo.sumErrorWeighted = ri.sum_error_weighted;
        lastEvaluations = n_evals;
        return err;
    }
#ifdef REVO_HOST_WITH_EIGEN
    float trackFrames(const std::shared_ptr<ImgPyramidRGBD> &refFrame, const std::shared_ptr<ImgPyramidRGBD> &currFrame,
                      Eigen::Matrix3f &R, Eigen::Vector3f &T, int lvl, ResidualInfo &resInfo) {
        revo::Mat3f r; revo::Vec3f t;
        std::memcpy(r.m, R.data(), sizeof(r.m)); std::memcpy(t.v, T.data(), sizeof(t.v));
        const float e = trackFrames(refFrame, currFrame, r, t, lvl, resInfo);
        std::memcpy(R.data(), r.m, sizeof(r.m)); std::memcpy(T.data(), t.v, sizeof(t.v));
        return e;
    }
#endif
    int lastEvaluations = 0;
private:
    std::shared_ptr<revo::Context> ctx_;
    OptimizerSettings mSettings;
};

// ------------------------------------------------------------------------------------------------------------
// system/tracker.h
// ------------------------------------------------------------------------------------------------------------
class TrackerSettings {   // tracker.h:31-50
public:
    TrackerSettings() { optimizerSettings.USE_EDGE_FILTER = true; }   // tracker.h:46 default
    bool CHECK_TRACKING_RESULTS = true;
    bool CHECK_INIT_VALUES = true;
    OptimizerSettings optimizerSettings;
    int nFramesHistogramVoting = 3;
};

class TrackerNew {
public:
    enum TrackerStatus { TRACKER_STATE_OK, TRACKER_STATE_LOST, TRACKER_STATE_NEW_KF, TRACKER_STATE_UNKNOWN };   // tracker.h:60-65
    int histogramLevel = 2;
    TrackerNew(const std::shared_ptr<revo::Context> &ctx, const TrackerSettings &config, const ImgPyramidSettings &pyrConfig)
        : ctx_(ctx), mSettings(config), mPyrConfig(pyrConfig) {}

    // TrackerStatus trackFrames(R&, T&, error&, refFrame, currFrame)   tracker.h:69-70 / tracker.cpp:294-353
    TrackerStatus trackFrames(revo::Mat3f &R, revo::Vec3f &T, float &error, const std::shared_ptr<ImgPyramidRGBD> &refFrame,
                              const std::shared_ptr<ImgPyramidRGBD> &currFrame) {
        revo_tracker_config cfg;
        revo_tracker_config_default(&cfg);
        cfg.check_init_values = mSettings.CHECK_INIT_VALUES;
        cfg.pyr_min_lvl = mPyrConfig.PYR_MIN_LVL; cfg.pyr_max_lvl = mPyrConfig.PYR_MAX_LVL;
        cfg.opt = mSettings.optimizerSettings.c_config();
        ctx_->check(revo_track(ctx_->handle(), &cfg, refFrame->handle(), currFrame->handle(), R.data(), T.data(), &lastResult));
        error = lastResult.error;
        return (TrackerStatus)lastResult.status;
    }
#ifdef REVO_HOST_WITH_EIGEN
    TrackerStatus trackFrames(Eigen::Matrix3f &R, Eigen::Vector3f &T, float &error, const std::shared_ptr<ImgPyramidRGBD> &refFrame,
                              const std::shared_ptr<ImgPyramidRGBD> &currFrame) {
        revo::Mat3f r; revo::Vec3f t;
        std::memcpy(r.m, R.data(), sizeof(r.m)); std::memcpy(t.v, T.data(), sizeof(t.v));
        const TrackerStatus s = trackFrames(r, t, error, refFrame, currFrame);
        std::memcpy(R.data(), r.m, sizeof(r.m)); std::memcpy(T.data(), t.v, sizeof(t.v));
        return s;
    }
#endif
    // the same call with the pose in Sophus::SE3f form (the reference converts R, T <-> SE3f inside Optimizer::trackFrames,
    // optimizer.cpp:240,308-309)
    TrackerStatus trackFrames(revo::SE3f &T_ref_cur, float &error, const std::shared_ptr<ImgPyramidRGBD> &refFrame,
                              const std::shared_ptr<ImgPyramidRGBD> &currFrame) {
        revo::Mat3f r; revo::Vec3f t;
        ctx_->check(revo_quat_to_R9(T_ref_cur.q, r.data()));
        std::memcpy(t.v, T_ref_cur.t, sizeof(t.v));
        const TrackerStatus s = trackFrames(r, t, error, refFrame, currFrame);
        ctx_->check(revo_R9_to_quat(r.data(), T_ref_cur.q));
        std::memcpy(T_ref_cur.t, t.v, sizeof(t.v));
        return s;
    }
#ifdef REVO_HOST_WITH_SOPHUS
    TrackerStatus trackFrames(Sophus::SE3f &T_ref_cur, float &error, const std::shared_ptr<ImgPyramidRGBD> &refFrame,
                              const std::shared_ptr<ImgPyramidRGBD> &currFrame) {
        revo::SE3f p;
        const Eigen::Quaternionf uq = T_ref_cur.unit_quaternion();
        p.q[0] = uq.x(); p.q[1] = uq.y(); p.q[2] = uq.z(); p.q[3] = uq.w();
        for (int i = 0; i < 3; ++i) p.t[i] = T_ref_cur.translation()[i];
        const TrackerStatus s = trackFrames(p, error, refFrame, currFrame);
        T_ref_cur = Sophus::SE3f(Eigen::Quaternionf(p.q[3], p.q[0], p.q[1], p.q[2]), Eigen::Vector3f(p.t[0], p.t[1], p.t[2]));
        return s;
    }
#endif
    // void addOldPclAndPose(pcl, worldPose, timeStamp)   tracker.h:77 / tracker.cpp:209-224.  Like the reference, which stores
    // return3DEdges(histogramLevel) BY VALUE, the tracker keeps a device copy of that one list (revo_pyr_copy_points_batch), not the
    // frame: the pyramid (and the batch slab it may live in) can be released.  worldPose: column-major 4x4 (Eigen::Matrix4f::data()).
    void addOldPclAndPose(const std::shared_ptr<ImgPyramidRGBD> &frame, const float *worldPose16, double timeStamp) {
        revo_pyr *src = frame->handle(), *copy = nullptr;
        ctx_->check(revo_pyr_copy_points_batch(ctx_->handle(), 1, &src, histogramLevel, &copy));
        Past p;
        std::shared_ptr<revo::Context> ctx = ctx_;
        p.list = std::shared_ptr<revo_pyr>(copy, [ctx](revo_pyr *h) { revo_pyr_destroy(ctx->handle(), h); });
        p.ts = timeStamp;
        std::memcpy(p.pose, worldPose16, sizeof(p.pose));
        mPast.push_back(p);
        // The reference's lists grow until the next keyframe, but only the first nFramesHistogramVoting entries ever vote
        // (tracker.cpp:138) and clearUpPastLists keeps the last ones: what lies in between can never be read again.
        const size_t nv = (size_t)mSettings.nFramesHistogramVoting;
        while (mPast.size() > 2 * nv) mPast.erase(mPast.begin() + (long)nv);
    }
    void clearUpPastLists() {   // tracker.cpp:249-257
        while ((int)mPast.size() > mSettings.nFramesHistogramVoting) mPast.pop_front();
    }
    // TrackerStatus assessTrackingQuality(estimatedPose, currFrame)   tracker.h:74 / tracker.cpp:118-201
    TrackerStatus assessTrackingQuality(const float *estimatedPose16, const std::shared_ptr<ImgPyramidRGBD> &currFrame) {
        if (mPast.empty() || !mSettings.CHECK_TRACKING_RESULTS) return TRACKER_STATE_OK;
        std::vector<revo_pyr *> hs;
        std::vector<float> poses;
        for (const Past &p : mPast) {
            hs.push_back(p.list.get());
            poses.insert(poses.end(), p.pose, p.pose + 16);
        }
        ctx_->check(revo_track_quality(ctx_->handle(), currFrame->handle(), histogramLevel, (int)hs.size(), hs.data(), poses.data(),
                                       estimatedPose16, mSettings.nFramesHistogramVoting, &lastQuality));
        return (TrackerStatus)lastQuality.status;
    }
#ifdef REVO_HOST_WITH_EIGEN
    void addOldPclAndPose(const std::shared_ptr<ImgPyramidRGBD> &frame, const Eigen::Matrix4f &worldPose, double timeStamp) {
        addOldPclAndPose(frame, worldPose.data(), timeStamp);
    }
    TrackerStatus assessTrackingQuality(const Eigen::Matrix4f &estimatedPose, const std::shared_ptr<ImgPyramidRGBD> &currFrame) {
        return assessTrackingQuality(estimatedPose.data(), currFrame);
    }
#endif
    revo_track_result lastResult{};
    revo_quality_result lastQuality{};
private:
    struct Past {
        std::shared_ptr<revo_pyr> list;   // owns a copy of the level-histogramLevel 3-D edge list only
        float pose[16];
        double ts;
    };
    std::deque<Past> mPast;
    std::shared_ptr<revo::Context> ctx_;
    const TrackerSettings mSettings;
    const ImgPyramidSettings mPyrConfig;
};
