/*
 * revo_b200.h -- C ABI of the B200-native edge-based RGB-D tracking hot path.
 *
 * Drop-in boundary for the path  ImgPyramidRGBD -> TrackerNew -> Optimizer  of
 * fabianschenk/REVO.  Plain C: opaque handles, POD structs, pointers + sizes,
 * int status codes (0 = ok).  Nothing here exits, aborts or throws (the
 * reference's conventions are exit(0)/assert/Sophus abort(); each entry point
 * documents the status code that replaces them).  No torch/Eigen/OpenCV types.
 *
 * Citations are file:line in the reference checkout (fabianschenk/REVO @ eb949c0).
 *
 * Conventions
 *  - R is a COLUMN-major 3x3 float matrix (9 floats), i.e. exactly
 *    Eigen::Matrix3f::data(); t is 3 floats (Eigen::Vector3f::data()).
 *    (R,t) maps current-frame points into the reference (key) frame:
 *    p_ref = R p_cur + t  (system/system.cpp:191-192).
 *  - Image pointers may be host or device pointers (detected per call).
 *  - Every function is thread-safe for distinct contexts; one context
 *    serialises its calls on its own CUDA stream.
 *  - There is NO CPU fallback: without a CUDA device revo_ctx_create returns
 *    REVO_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef REVO_B200_H
#define REVO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define REVO_API __attribute__((visibility("default")))
#else
#define REVO_API
#endif

#define REVO_MAX_LEVELS 6 /* PYRAMID_LEVELS, system/optimizer.h:36 */

/* ---- status codes ------------------------------------------------------ */
enum {
    REVO_OK = 0,
    REVO_ERR_INVALID_ARG = 1,
    REVO_ERR_NO_DEVICE = 2,      /* no CUDA device / driver: there is no CPU path */
    REVO_ERR_CUDA = 3,           /* a CUDA runtime call failed; see revo_last_error() */
    REVO_ERR_NOT_KEYFRAME = 4,   /* replaces exit(0) in returnOptimizationStructure, imgpyramidrgbd.h:113-117 */
    REVO_ERR_NOT_ORTHOGONAL = 5, /* replaces Sophus ENSURE abort(), so3.hpp:419-424 */
    REVO_ERR_BAD_LEVEL = 6,      /* replaces assert(lvl < size), imgpyramidrgbd.h:59-88 */
    REVO_ERR_BUFFER_TOO_SMALL = 7,
    REVO_ERR_UNSUPPORTED = 8,
    REVO_ERR_COMM = 9
};

typedef struct revo_ctx revo_ctx; /* owns device, stream, scratch memory */
typedef struct revo_pyr revo_pyr; /* one ImgPyramidRGBD (datastructures/imgpyramidrgbd.h:27) */

/* ---- POD mirrors of the reference's settings --------------------------- */

/* Camera -- datastructures/camerapyr.h:90-111 (level-0 intrinsics; level l is
 * derived as fx,fy,cx,cy * 2^-l, w,h = floor(w * 2^-l): camerapyr.h:98-103,139-144). */
typedef struct revo_camera {
    float fx, fy, cx, cy;
    int32_t width, height;
} revo_camera;

/* ImgPyramidSettings -- datastructures/camerapyr.h:27-89 (hot-path fields, same defaults). */
typedef struct revo_pyr_config {
    int32_t n_levels;         /* PYR_MIN_LVL - PYR_MAX_LVL + 1, camerapyr.h:68-71 (default 3) */
    int32_t canny_threshold1; /* 150, camerapyr.h:40 */
    int32_t canny_threshold2; /* 100, camerapyr.h:41 */
    float depth_min;          /* 0.1, camerapyr.h:43 */
    float depth_max;          /* 5.2, camerapyr.h:44 */
    int32_t use_edge_hist;    /* true, camerapyr.h:63 */
    float n_percentage;       /* 0.3, camerapyr.h:64 */
    int32_t patch0;           /* distPatchSizes[0] = 20 (imgpyramidrgbd.cpp:50); level l uses patch0 >> l */
} revo_pyr_config;

/* OptimizerSettings -- system/optimizer.h:42-112 (fields read by trackFrames). */
typedef struct revo_opt_config {
    float lambda_success_fac;                /* 0.5   optimizer.h:53 */
    float lambda_fail_fac;                   /* 2.0   optimizer.h:54 */
    float lambda_initial[REVO_MAX_LEVELS];   /* 0     optimizer.h:63 */
    float step_size_min[REVO_MAX_LEVELS];    /* 1e-16 optimizer.h:55 */
    float convergence_eps[REVO_MAX_LEVELS];  /* 0.999 optimizer.h:65 */
    int32_t max_its_per_lvl[REVO_MAX_LEVELS];/* 100   optimizer.h:56 */
    float edge_distance_lvl[REVO_MAX_LEVELS];/* {30,20,10,5,5,5} optimizer.h:59 */
    float huber_edge;                        /* 0.3   optimizer.h:75 */
    int32_t use_edge_filter;                 /* OptimizerSettings ctor: false (optimizer.h:80);
                                                TrackerSettings default: true (tracker.h:46) -> default 1 here */
    int32_t max_lm_tries;                    /* 0 = reference termination rules only; >0 caps the total
                                                number of LM tries per level (fixed-iteration test mode) */
} revo_opt_config;

/* TrackerSettings + pyramid level range -- system/tracker.h:31-50, camerapyr.h:45-46. */
typedef struct revo_tracker_config {
    int32_t check_init_values; /* CHECK_INIT_VALUES, tracker.h:43 (default true) */
    int32_t pyr_min_lvl;       /* coarsest level, PYR_MIN_LVL (default 2) */
    int32_t pyr_max_lvl;       /* finest level,   PYR_MAX_LVL (default 0) */
    revo_opt_config opt;
} revo_tracker_config;

/* Optimizer::ResidualInfo -- system/optimizer.h:117-139 (of the LAST evaluation). */
typedef struct revo_residual_info {
    int32_t good_pts_edges;
    int32_t bad_pts_edges;
    float sum_error_unweighted;
    float sum_error_weighted;
} revo_residual_info;

/* TrackerNew::TrackerStatus -- system/tracker.h:60-65 (same numeric values). */
enum {
    REVO_TRACKER_STATE_OK = 0,
    REVO_TRACKER_STATE_LOST = 1,
    REVO_TRACKER_STATE_NEW_KF = 2,
    REVO_TRACKER_STATE_UNKNOWN = 3
};

/* Result of tracking one frame pair (TrackerNew::trackFrames, tracker.cpp:294-353). */
typedef struct revo_track_result {
    float R[9];                       /* column-major */
    float t[3];
    float error;                      /* last accepted mean weighted error of the finest level */
    int32_t status;                   /* REVO_TRACKER_STATE_* */
    int32_t rc;                       /* REVO_OK or REVO_ERR_NOT_ORTHOGONAL for this pair */
    revo_residual_info res;           /* of the last evaluation (tracker.cpp:351 uses good/bad) */
    int32_t n_evals[REVO_MAX_LEVELS]; /* fused evaluations ("GN iterations") per level */
    int32_t n_pts[REVO_MAX_LEVELS];   /* 3-D edge points of the current frame per level (return3DEdges(l).cols()) */
    int32_t used_identity_init;       /* 1 if checkInitializationValues reset (R,t) to identity */
} revo_track_result;

/* One entry of the optional LM trace (parity/debug aid; not in the reference). */
typedef struct revo_trace_entry {
    float error;
    float lambda;
    int32_t accepted;
    int32_t good, bad;
    int32_t level;
} revo_trace_entry;

/* Which array revo_pyr_download returns -- the accessors of imgpyramidrgbd.h:45-117. */
enum {
    REVO_ARRAY_GRAY = 0,        /* returnGray(lvl):           u8  h*w              */
    REVO_ARRAY_DEPTH = 1,       /* returnDepth(lvl):          f32 h*w              */
    REVO_ARRAY_EDGES = 2,       /* returnEdges(lvl):          u8  h*w {0,255} (after fill-in) */
    REVO_ARRAY_EDGES_ORIG = 3,  /* returnOrigEdges(lvl):      u8  h*w (Canny output) */
    REVO_ARRAY_HIST = 4,        /* histPyr[lvl]:              u8  (h/P)*(w/P)      */
    REVO_ARRAY_EDGES3D = 5,     /* return3DEdges(lvl):        f32 4*N, reference (column-major scan) order */
    REVO_ARRAY_DT = 6,          /* returnDistTransform(lvl):  f32 h*w (keyframes)  */
    REVO_ARRAY_OPTSTRUCT = 7,   /* returnOptimizationStructure(lvl): f32 4*h*w (keyframes) */
    REVO_ARRAY_EDGES3D_DEVICE_ORDER = 8 /* the tile-major list the tracker actually iterates */
};

/* ---- defaults ----------------------------------------------------------- */
REVO_API void revo_pyr_config_default(revo_pyr_config *cfg);
REVO_API void revo_opt_config_default(revo_opt_config *cfg);
REVO_API void revo_tracker_config_default(revo_tracker_config *cfg);

/* ---- context ------------------------------------------------------------ */
REVO_API int revo_ctx_create(int device, revo_ctx **ctx_out);
REVO_API int revo_ctx_destroy(revo_ctx *ctx);
REVO_API int revo_ctx_synchronize(revo_ctx *ctx);
REVO_API const char *revo_strerror(int code);
REVO_API const char *revo_last_error(revo_ctx *ctx); /* text of the last CUDA failure */
/* cudaStream_t of the context as an integer (for callers that enqueue their own work/events). */
REVO_API uint64_t revo_ctx_stream(revo_ctx *ctx);
/* number of kernels this library has launched on the context so far */
REVO_API uint64_t revo_ctx_launch_count(revo_ctx *ctx);
/* Device time (CUDA events on the context stream) of the most recent completed pyramid-construction,
 * keyframe-promotion and tracking-kernel launches, in milliseconds (0 if none yet). Synchronises. */
REVO_API int revo_ctx_last_timings(revo_ctx *ctx, float *pyramid_ms, float *keyframe_ms, float *track_kernel_ms);
/* Device time of the most recent host-to-device upload of a frame batch (CUDA events on the copy stream). Synchronises it. */
REVO_API int revo_ctx_last_upload_ms(revo_ctx *ctx, float *upload_ms);

/* ---- ImgPyramidRGBD ------------------------------------------------------ */
/* ImgPyramidRGBD(settings, camPyr, rgb, depth, ts) -- imgpyramidrgbd.cpp:43-96.
 * bgr: 8UC3 / 8UC4 (channels = 3|4), row stride in bytes; depth: 32FC1 metres (0/NaN invalid),
 * row stride in bytes.  Asynchronous on the context stream (host buffers are staged first). */
REVO_API int revo_pyr_create(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0,
                             const uint8_t *bgr, size_t bgr_stride, int channels,
                             const float *depth, size_t depth_stride, double timestamp, revo_pyr **pyr_out);
/* n frames of identical geometry in one go (tightly packed: frame i at bgr + i*h*w*channels,
 * depth + i*h*w).  pyr_out receives n handles.  One set of batched kernel launches. */
REVO_API int revo_pyr_create_batch(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, int n,
                                   const uint8_t *bgr, int channels, const float *depth,
                                   const double *timestamps, revo_pyr **pyr_out);
/* Same with the depth in the sensor / dataset wire format: 16-bit raw values (TUM: 16-bit PNG, 5000 units per metre).
 * metres = float(raw) * depth_scale with depth_scale = 1.0f / DEPTH_SCALE_FACTOR, i.e. exactly
 * depth.convertTo(depth, CV_32FC1, 1.0f / depthScaleFactor) of the reference's reader (io/iowrapperRGBD.cpp:327), done on the
 * device: 2 instead of 4 bytes per pixel cross the host link. */
REVO_API int revo_pyr_create_batch_u16(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, int n,
                                       const uint8_t *bgr, int channels, const uint16_t *depth_raw, float depth_scale,
                                       const double *timestamps, revo_pyr **pyr_out);
/* makeKeyframe() -- imgpyramidrgbd.cpp:231-252: exact L2 EDT + {gx,gy,dt,0} structure, all levels. Idempotent. */
REVO_API int revo_pyr_make_keyframe(revo_ctx *ctx, revo_pyr *pyr);
REVO_API int revo_pyr_make_keyframe_batch(revo_ctx *ctx, int n, revo_pyr *const *pyrs);
REVO_API int revo_pyr_destroy(revo_ctx *ctx, revo_pyr *pyr);
REVO_API int revo_pyr_destroy_batch(revo_ctx *ctx, int n, revo_pyr *const *pyrs);
REVO_API int revo_pyr_is_keyframe(const revo_pyr *pyr);
REVO_API int revo_pyr_level_camera(const revo_pyr *pyr, int lvl, revo_camera *cam_out); /* cameraPyr->at(lvl) */
REVO_API double revo_pyr_timestamp(const revo_pyr *pyr);
/* return3DEdges(lvl).cols() -- synchronises the context stream. */
REVO_API int revo_pyr_num_edges(revo_ctx *ctx, const revo_pyr *pyr, int lvl, int *n_out);
/* Accessors: copies array `which` of level lvl to host memory dst (synchronous).
 * bytes_out (optional) receives the byte size; REVO_ERR_BUFFER_TOO_SMALL if dst_bytes is short. */
REVO_API int revo_pyr_download(revo_ctx *ctx, const revo_pyr *pyr, int lvl, int which, void *dst, size_t dst_bytes,
                               size_t *bytes_out);
/* Test hook: overwrite one level of a pyramid with caller-provided arrays (host pointers; any may be
 * NULL to keep the existing content): the 3-D edge list (n x float4), the distance transform (h*w) and
 * the lookup structure (h*w float4).  Marks the pyramid a keyframe when opt4 is given. */
REVO_API int revo_pyr_upload_level(revo_ctx *ctx, revo_pyr *pyr, int lvl, const float *pts4, int n,
                                   const float *dt, const float *opt4);

/* generateColoredPcl(lvl, clrPcl, densePcl) -- imgpyramidrgbd.cpp:279-327: the viewer's coloured cloud of level lvl, one column
 * (X, Y, Z, 1, r, g, b, 1) per pixel with a valid depth (and, unless dense, an edge label), in the reference's column-major scan
 * order; out receives N columns of 8 floats (Eigen::MatrixXf(8, N)::data()).  bgr: the full-resolution colour image the pyramid
 * was built from (host or device pointer, `channels` = 3|4): the reference keeps its own clone (rgbFullSize), here the caller
 * keeps it; it is brought to level lvl by cv::pyrDown like the reference does (levels > 2 give an empty cloud, as there).
 * *n_out receives N; with out == NULL only the count is produced.  REVO_ERR_BUFFER_TOO_SMALL if capacity_points < N (the
 * reference sizes its matrix cam.area / 5 for the edge cloud and writes past its end when more points turn up). */
REVO_API int revo_pyr_colored_pcl(revo_ctx *ctx, const revo_pyr *pyr, int lvl, int dense, const uint8_t *bgr, int channels,
                                  float *out, size_t capacity_points, int *n_out);

/* TrackerNew::addOldPclAndPose(pcl, worldPose, ts) -- system/tracker.cpp:209-224 keeps a COPY of return3DEdges(histogramLevel)
 * per past frame: out[i] becomes a handle that owns only a copy of pyrs[i]'s level-`lvl` 3-D edge list (cameras and sizes of all
 * levels are kept; every other array is absent, downloads of them answer REVO_ERR_UNSUPPORTED), so that the vote's history does
 * not keep whole frame batches alive.  Valid as `past` of revo_track_quality*; release with revo_pyr_destroy. */
REVO_API int revo_pyr_copy_points_batch(revo_ctx *ctx, int n, revo_pyr *const *pyrs, int lvl, revo_pyr **out);

/* ---- Optimizer / TrackerNew --------------------------------------------- */
/* One fused evaluation at pose (R,t): PASS A + PASS B of system/optimizer.cpp:74-234.
 * record32: [0..20] upper triangle of sum(w v v^T) in LGS6 slot order (0,0..5),(1,1..5),..,(5,5)
 * (utils/LGSX.h:212-314), [21..26] sum(w r v), [27] sum(w r^2), [28] sum(r^2), [29] good, [30] bad, [31] 0. */
REVO_API int revo_eval(revo_ctx *ctx, const revo_opt_config *cfg, const revo_pyr *ref, const revo_pyr *cur, int lvl,
                       const float *R9, const float *t3, double *record32);
/* Optimizer::trackFrames -- system/optimizer.cpp:235-311 (one level).  Returns the error in *err. */
REVO_API int revo_track_level(revo_ctx *ctx, const revo_opt_config *cfg, const revo_pyr *ref, const revo_pyr *cur,
                              int lvl, float *R9_inout, float *t3_inout, revo_residual_info *res, float *err,
                              int *n_evals);
/* TrackerNew::trackFrames -- system/tracker.cpp:294-353 (init check + coarse-to-fine). */
REVO_API int revo_track(revo_ctx *ctx, const revo_tracker_config *cfg, const revo_pyr *ref, const revo_pyr *cur,
                        float *R9_inout, float *t3_inout, revo_track_result *result);
/* n independent pairs in ONE persistent kernel launch (BASELINE config 4).  R9s/t3s: n*9 / n*3 floats in,
 * results[i] out.  trace (optional): n*trace_cap entries, trace_counts n ints. */
REVO_API int revo_track_batch(revo_ctx *ctx, const revo_tracker_config *cfg, int n, revo_pyr *const *refs,
                              revo_pyr *const *curs, const float *R9s, const float *t3s, revo_track_result *results,
                              revo_trace_entry *trace, int trace_cap, int *trace_counts);
/* TrackerNew::assessTrackingQuality -- system/tracker.cpp:118-201 (Schenk & Fraundorfer, IROS 2017): the 3-D edge points of
 * up to n_frames_voting past frames (their level-`hist_level` lists, tracker.cpp:173,259 of system.cpp feed
 * return3DEdges(histogramLevel)) are projected into the current frame with  inv(estimated_pose) * past_world_pose[i];
 * every past frame marks each pixel at most once (M_i), M = sum M_i; over the pixels of the current frame with a valid
 * depth, histogram[M] counts pixels and overlaps[M] those that are Canny edges (returnOrigEdges).  The vote:
 * overlap_measure = sum_{k>0} overlaps[k] * {0, 1, 1.25, 1.5}[k];  OK if overlap_measure >= overlaps[0] or fewer than 3 past
 * frames took part, else NEW_KF.  Poses are column-major 4x4 floats (Eigen::Matrix4f::data()). */
typedef struct revo_quality_result {
    int32_t histogram[4];
    int32_t overlaps[4];
    float overlap_measure;
    int32_t status;        /* REVO_TRACKER_STATE_OK or REVO_TRACKER_STATE_NEW_KF */
    int32_t out_of_bounds; /* projections that left the image (the reference only logs it) */
    int32_t n_frames;      /* past frames that took part: min(n_past, n_frames_voting) */
} revo_quality_result;
REVO_API int revo_track_quality(revo_ctx *ctx, const revo_pyr *cur, int hist_level, int n_past, revo_pyr *const *past,
                                const float *past_world_poses16, const float *estimated_pose16, int n_frames_voting,
                                revo_quality_result *out);
/* The same vote for n streams in one launch pair (MultiStreamREVO's keyframe policy): curs[i] is voted on by n_past[i] (<= 3
 * used) past frames past[3*i + f] with world poses past_world_poses16[(3*i + f)*16 ..], under estimated_poses16[16*i ..].
 * All current frames must have one size.  One device->host read of n x 16 counters. */
REVO_API int revo_track_quality_batch(revo_ctx *ctx, int n, revo_pyr *const *curs, int hist_level, const int *n_past,
                                      revo_pyr *const *past, const float *past_world_poses16, const float *estimated_poses16,
                                      int n_frames_voting, revo_quality_result *out);

/* Launch-shape override for revo_track_batch (0 = automatic): CTAs per pair (cluster size 1,2,4,8,16)
 * and threads per CTA. */
REVO_API int revo_ctx_set_track_shape(revo_ctx *ctx, int ctas_per_pair, int threads_per_cta);
/* Cap on the clusters of the persistent tracking kernel that are resident at a time (0 = as many as the device holds: 74 clusters
 * of 8 CTAs on B200).  For pipelines that build the pyramids of the next frames on a second context while this one tracks: the
 * tracker owns every register of an SM at full residency, so a build kernel can only start when a tracker CTA retires; with a few
 * cluster slots left free both run side by side and fill each other's stalls.  Timed alone the kernel is fastest uncapped. */
REVO_API int revo_ctx_set_track_max_clusters(revo_ctx *ctx, int max_clusters);
/* Kept for ABI compatibility: there is ONE tracking engine (one thread-block cluster per pair, track.cu).  engine 0 / 1 are
 * accepted, anything else answers REVO_ERR_UNSUPPORTED (the round-1 alternatives -- task queue, ping-pong clusters -- measured
 * slower and are no longer built); chunk_points is ignored. */
REVO_API int revo_ctx_set_track_engine(revo_ctx *ctx, int engine, int chunk_points);
/* Pre-size the device memory pool: later stream-ordered slab allocations up to `bytes` in total are served
 * from cached memory instead of the driver (call once before a steady-state stream starts). */
REVO_API int revo_ctx_reserve(revo_ctx *ctx, size_t bytes);

/* ---- pose helpers for a Sophus::SE3f adapter (host arithmetic, no device needed) ------------- */
/* Unit quaternion (x, y, z, w: Eigen / Sophus coefficient order) <-> column-major 3x3 rotation as the tracking calls take
 * it.  revo_quat_to_R9 normalises q (REVO_ERR_INVALID_ARG for a zero or non-finite quaternion); revo_R9_to_quat applies
 * the tracker's own test (||R R^T - I||_F < 1e-5, det > 0: the Sophus ENSUREs of so3.hpp:419-424) and returns
 * REVO_ERR_NOT_ORTHOGONAL otherwise; the conversion is Eigen's (Shepperd), as SO3(Matrix3) does. */
REVO_API int revo_quat_to_R9(const float *q_xyzw, float *R9_out);
REVO_API int revo_R9_to_quat(const float *R9, float *q_xyzw_out);

/* ---- single pair split over several GPUs (BASELINE config 5) ------------- */
/* One process per GPU.  Every rank builds the same pyramids; rank r evaluates the r-th contiguous
 * chunk of each level's point list and the ranks exchange the 32-float record once per evaluation
 * through peer memory over NVLink (fused into the persistent kernel).  Setup: each rank calls
 * revo_split_export to obtain an opaque handle blob, the host exchanges blobs (e.g. torch.distributed
 * all_gather), then every rank calls revo_split_open with all blobs in rank order. */
#define REVO_SPLIT_HANDLE_BYTES 128
REVO_API int revo_split_export(revo_ctx *ctx, int rank, int world, void *handle_out /* REVO_SPLIT_HANDLE_BYTES */);
REVO_API int revo_split_open(revo_ctx *ctx, const void *handles /* world * REVO_SPLIT_HANDLE_BYTES */);
REVO_API int revo_track_split(revo_ctx *ctx, const revo_tracker_config *cfg, const revo_pyr *ref, const revo_pyr *cur,
                              float *R9_inout, float *t3_inout, revo_track_result *result);

#ifdef __cplusplus
}
#endif
#endif /* REVO_B200_H */
