// internal.h -- shared declarations of the revo_b200 CUDA library (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <string>
#include <vector>

#include "../../include/revo_b200.h"

namespace revo {

// One (frame, level) of an ImgPyramidRGBD as the kernels see it (device-visible POD).
// Mirrors the per-level members of datastructures/imgpyramidrgbd.h:186-213.
struct ImgLevel {
    uint8_t *gray;        // grayPyr[l]       h*w
    float *depth;         // depthPyr[l]      h*w
    uint8_t *edges;       // edgesPyr[l]      h*w {0,255} (class map 0/1/2 while Canny runs)
    uint8_t *edges_orig;  // edgesOrigPyr[l]  h*w
    uint8_t *hist;        // histPyr[l]       (h/P)*(w/P)
    float4 *pts;          // edges3DPyr[l]    tile-major order, capacity pts_cap
    int *n_pts;           // number of valid entries of pts (device scalar)
    int *nz_patches;      // countNonZero(hist) (device scalar)
    int *tile_off;        // per-tile exclusive offsets of the compaction (n_tiles + 1)
    int *labels;          // scratch h*w int32: union-find labels (Canny) / column distances (EDT)
    uint8_t *flags;       // scratch w0*h0 bytes per frame: integer patch counters of the histogram (K5)
    float *dt;            // dtPyr[l]         h*w   (keyframes, else nullptr)
    uint2 *opt;           // optimizationStructure[l] in the device layout: 8-byte texels in 4x4 tiles (see opt_texel_index; keyframes)
    int w, h;
    int pts_cap;
    int patch;            // distPatchSizes[l]
    int hist_w, hist_h;
    float fx, fy, cx, cy; // Camera at this level (camerapyr.h:98-103)
};

// Device layout of the lookup structure (optimizationStructure of the reference: one float4 {gx, gy, dt, .} per pixel,
// imgpyramidrgbd.cpp:255-276): one 8-byte TEXEL per pixel -- dt as float32 (the residual is exact) and the gradient as two
// snorm16 (step 1/32764) -- stored in 4x4-pixel TILES of 128 bytes = one L2 line, tiles row-major.  The tracker reads the
// 2x2 texels around a projected point; with tiles a line serves a 4x4 neighbourhood in every direction, which halves the
// number of L2 lines a frame pair keeps busy against a row-major 32-byte record per pixel (profiles/r1_lookup_layout_study.txt)
// and is what lets twice as many pairs be in flight.  Index of texel (x, y) in uint2 units; tw = tiles per row.
__host__ __device__ inline unsigned opt_texel_index(int x, int y, int tw)
{
    return (((unsigned)(y >> 2) * (unsigned)tw + (unsigned)(x >> 2)) << 4) + ((unsigned)(y & 3) << 2) + (unsigned)(x & 3);
}
inline int opt_tiles_per_row(int w) { return (w + 3) >> 2; }
inline size_t opt_bytes(int w, int h) { return (size_t)opt_tiles_per_row(w) * (size_t)((h + 3) >> 2) * 128; }

// Point-list tile: one warp <-> one 8x4 pixel tile (row-major inside, tiles row-major).
constexpr int kTileW = 8;
constexpr int kTileH = 4;

struct Slab;    // one device allocation shared by the frames of a batch
struct KfSlab;  // one device allocation shared by the keyframe structures promoted together

}  // namespace revo

// The opaque handle types of the C ABI.
struct revo_pyr {
    revo::Slab *slab;
    int index_in_slab;
    int n_levels;
    revo_pyr_config cfg;
    revo_camera cam0;
    double timestamp;
    revo::ImgLevel lv[REVO_MAX_LEVELS];    // host copy of the device descriptors
    revo::KfSlab *kf_slab;                 // keyframe allocation (dt + pair structure of all levels), shared by a batch
    bool is_keyframe;
};

struct revo_ctx {
    int device;
    cudaStream_t stream;
    cudaStream_t copy_stream;   // uploads of host inputs (so that the H2D of the next batch overlaps the kernels of this one)
    cudaDeviceProp prop;
    std::string last_error;
    uint64_t launches;
    // scratch
    void *scratch;        // generic device scratch (descriptor tables, staging of uploads)
    size_t scratch_bytes;
    void *pinned;         // pinned, device-mapped host staging of the pair descriptors (run_track)
    size_t pinned_bytes;
    void *pinned_kf;      // same for the descriptor tables of keyframe promotion
    size_t pinned_kf_bytes;
    cudaEvent_t pinned_kf_read;   // recorded after the kernel that reads pinned_kf
    bool pinned_kf_busy;
    // double-buffered device staging of uploaded host bgr frames: the upload of batch k+2 must not wait for the build of k+1
    void *stage[2];
    size_t stage_bytes[2];
    cudaEvent_t stage_consumed[2];   // recorded on the main stream after the gray kernel that read the buffer
    bool stage_used[2];
    int stage_next;
    int track_ctas_per_pair;
    int track_threads;
    int track_max_clusters;   // cap on resident clusters of k_track (0 = all the device holds), revo_ctx_set_track_max_clusters
    // pyramid construction: the Canny / compaction chains of the levels run on their own streams between a fork after the gray /
    // depth pyramids and a join before the build-complete event (the small levels hide under level 0)
    cudaStream_t lvl_stream[REVO_MAX_LEVELS], depth_stream;
    cudaEvent_t lvl_fork, lvl_done[REVO_MAX_LEVELS], lvl_canny0, lvl_fill[REVO_MAX_LEVELS], lvl_gray[REVO_MAX_LEVELS], lvl_depth[REVO_MAX_LEVELS];
    cudaEvent_t ev[8];      // pyramid begin/end, keyframe begin/end, track kernel begin/end, upload begin/end (copy stream)
    bool ev_valid[4];
    // split mode (multi-GPU single pair)
    int split_rank, split_world;
    void *split_local;                 // this rank's mailbox (device memory, IPC-exported)
    void *split_peers[16];             // mapped mailboxes of all ranks (own entry = split_local)
    unsigned long long split_seq;
};

namespace revo {

struct Slab {
    void *mem;
    size_t bytes;
    int n_frames;
    int live;             // pyramids still alive
    cudaStream_t stream;  // stream the frames were built on
    cudaEvent_t ready;    // recorded on `stream` when the build is complete; other streams wait on it before reading
    ImgLevel *d_desc[REVO_MAX_LEVELS];  // device descriptor tables, n_frames entries each (inside mem)
};

struct KfSlab {
    void *mem;
    int live;
};

// error helper: records the failure text in ctx and returns REVO_ERR_CUDA
int cuda_fail(revo_ctx *ctx, cudaError_t e, const char *what);
#define REVO_CUDA(ctx, call)                                            \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return revo::cuda_fail((ctx), e__, #call); \
    } while (0)

// ---- pyramid.cu ----------------------------------------------------------
// All launchers are asynchronous on ctx->stream and batched over n frames (d_desc: device table).
int launch_gray(revo_ctx *ctx, const uint8_t *d_bgr, size_t stride, int ch, size_t frame_bytes, const ImgLevel *d_desc,
                int n, int w, int h);
int launch_depth_u16(revo_ctx *ctx, const uint16_t *d_raw, size_t frame_px, float scale, const ImgLevel *d_desc, int n, int px);
int launch_pyrdown_depth(revo_ctx *ctx, const ImgLevel *d_src, const ImgLevel *d_dst, int n, int w_dst, int h_dst,
                         int w_src, int h_src);
int launch_pyrdown(revo_ctx *ctx, const ImgLevel *d_src, const ImgLevel *d_dst, int n, int w_dst, int h_dst, int w_src, int h_src);
int launch_depth_half(revo_ctx *ctx, const ImgLevel *d_src, const ImgLevel *d_dst, int n, int w_dst, int h_dst, int w_src);
// gray_tmap: host pointer to a CUtensorMap made by make_gray_tensor_map (nullptr = plain loads)
// Also produces the patch histogram (hist, nz_patches) of the Canny output; d_counts0/counts_stride: the per-frame
// scratch (ImgLevel::flags of frame 0, byte stride between frames) used for the integer counters.
int launch_canny(revo_ctx *ctx, const ImgLevel *d_desc, int n, int w, int h, int low, int high, const void *gray_tmap, int patch,
                 void *d_counts0, size_t counts_stride);
// 3-D (x, y, frame) tensor map over the u8 gray images of one level of a slab; false if TMA cannot be used
bool make_gray_tensor_map(void *tmap_out /* 128 bytes, 64-aligned */, const uint8_t *base, int w, int h, int n_frames,
                          size_t frame_stride);
int launch_hist_fill(revo_ctx *ctx, const ImgLevel *d_desc, const ImgLevel *d_top, int n, int w, int h, int patch,
                     int patch_low, bool do_fill, float n_percentage);
int launch_compact(revo_ctx *ctx, const ImgLevel *d_desc, int n, int w, int h, float dmin, float dmax);
int launch_keyframe(revo_ctx *ctx, const ImgLevel *d_desc, int n, int w, int h);
// generateColoredPcl: channel-wise cv::pyrDown of a colour image; the cloud of one level (count -> scan -> [scatter if d_out])
int launch_pyrdown_color(revo_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, int ws, int hs, int ch);
int launch_colored_pcl(revo_ctx *ctx, const ImgLevel *d_desc_one, int w, int h, int dense, float dmin, float dmax, const uint8_t *d_bgr, int ch,
                       float *d_out, int cap, int *d_n, int *d_col_off);
// tracking-quality vote (revo_track_quality): the past frames' 3-D lists with the transform into the current frame
struct QualityFrame {
    const float4 *pts;
    const int *n_pts;
    float R[9], T[3];     // column-major R, as Eigen::Matrix3f
};
struct QualityArgs {      // one current frame and the (up to 3) past frames that vote on it
    QualityFrame fr[4];
    int n_frames;
    float fx, fy, cx, cy;
    int w, h;
    const float *depth;   // level `hist_level` of the current frame
    const uint8_t *edges; // returnOrigEdges(hist_level)
};
// n votes in one launch pair (all current frames of one size w x h).  d_args: n QualityArgs on the device; d_mbits: n planes of
// (w*h+3)/4 words; d_counters: n x 16 ints = histogram[4], overlaps[4], out_of_bounds, ...
int launch_quality(revo_ctx *ctx, const QualityArgs *d_args, int n, int w, int h, float dmin, float dmax, unsigned *d_mbits, int *d_counters);
struct PointListCopy {
    const float4 *src;
    const int *src_n;
    float4 *dst;
    int *dst_n;
    int cap;
};
int launch_copy_point_lists(revo_ctx *ctx, const PointListCopy *d_tab, int n);
int launch_opt_struct_f4(revo_ctx *ctx, const float *d_dt, int w, int h, float4 *d_out);
int launch_opt_pack_from_f4(revo_ctx *ctx, const float4 *d_in, int w, int h, uint2 *d_out);
// reference-order (column-major scan) 3-D edge list into d_out (capacity w*h float4); *d_n receives the count
int launch_edges3d_reference_order(revo_ctx *ctx, const ImgLevel *d_desc_one, int w, int h, float dmin, float dmax,
                                   float4 *d_out, int *d_n, int *d_col_off);

// ---- track.cu --------------------------------------------------------------
struct LevelIn {
    const float4 *pts;
    const int *n_pts;
    const uint2 *opt;    // tiled texel layout (opt_texel_index): {dt f32 | snorm16 gx, gy} per pixel
    float fx, fy, cx, cy;
    int w, h;
};
struct PairDesc {
    LevelIn lvl[REVO_MAX_LEVELS];
    const float *ref_dt_min;  // returnDistTransform(min_lvl) of the reference frame
    float R[9];
    float t[3];
};
struct TrackParams {
    revo_tracker_config cfg;
    int mode;            // 0 = full trackFrames, 1 = single level (Optimizer::trackFrames), 2 = one evaluation
    int level;           // for modes 1,2
    int trace_cap;
    int profile;         // 1: thread 0 accumulates clock64() cycles per phase (REVO_TRACK_PROF)
    int speculate;       // 1: the reject-successor of every LM try is computed while the record is exchanged (default)
    // split mode
    int split_rank, split_world;
    unsigned long long split_seq0;
    void *split_peers[16];
};
int launch_track(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm,
                 revo_track_result *d_results, double *d_records, revo_trace_entry *d_trace, int *d_trace_counts,
                 int *d_work_counter);

int launch_stage_in(revo_ctx *ctx, const void *src_mapped_host, void *dst, size_t bytes);

}  // namespace revo
