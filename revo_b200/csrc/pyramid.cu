// pyramid.cu -- ImgPyramidRGBD construction and keyframe promotion on the GPU.
//
// Replaces (reference file:line, fabianschenk/REVO):
//   K1 gray            cv::cvtColor(BGRA2GRAY)            datastructures/imgpyramidrgbd.cpp:53
//   K2 Canny           cv::Canny(g, e, 150, 100, 3, true) datastructures/imgpyramidrgbd.cpp:184
//   K3 pyrDown         cv::pyrDown                        datastructures/imgpyramidrgbd.cpp:82
//   K4 depth /2        FilterSubsampleWithHoles           datastructures/imgpyramidrgbd.h:218-249
//   K5 hist + fill-in  generateDistHistogram/fillInEdges  datastructures/imgpyramidrgbd.cpp:146-172,111-145
//   K6 3-D edge list   loop in addLevelEdge               datastructures/imgpyramidrgbd.cpp:199-226
//   K7 exact L2 EDT    cv::distanceTransform(L2,PRECISE)  datastructures/imgpyramidrgbd.cpp:241
//   K8 lookup struct   buildOptimizationStructure         datastructures/imgpyramidrgbd.cpp:255-276
//
// All kernels are batched over frames (blockIdx.z = frame) and integer/byte
// exact against OpenCV 4.13 (see oracle/revo_oracle.c for the CPU restatement
// these are tested against).  HBM-bound byte work: no tensor cores.
#include "internal.h"

namespace revo {

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

#define LAUNCH_CHECK(ctx)                                   \
    do {                                                    \
        (ctx)->launches++;                                  \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return cuda_fail((ctx), e__, __func__); \
    } while (0)

// ---------------------------------------------------------------------------
// K1: BGR(A) -> gray, Y = (3735 B + 19235 G + 9798 R + 16384) >> 15  (OpenCV 4.x)
// 4 pixels per thread: 3 x 32-bit loads (BGR) / 1 x 128-bit load (BGRA), one 32-bit store.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r)
{
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

__global__ void __launch_bounds__(256) k_gray(const uint8_t *__restrict__ bgr, size_t stride, int ch, size_t frame_bytes,
                                              const ImgLevel *__restrict__ desc, int w, int h)
{
    const int f = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= h || x0 >= w) return;
    const uint8_t *row = bgr + (size_t)f * frame_bytes + (size_t)y * stride;
    uint8_t *out = desc[f].gray + (size_t)y * w;
    const uint8_t *p = row + (size_t)x0 * ch;
    if (x0 + 4 <= w && (((uintptr_t)p) & 3) == 0 && (((uintptr_t)(out + x0)) & 3) == 0) {
        uint32_t y0, y1, y2, y3;
        if (ch == 3) {
            const uint32_t a = __ldg((const uint32_t *)p), b = __ldg((const uint32_t *)p + 1), c = __ldg((const uint32_t *)p + 2);
            // a = B0 G0 R0 B1 | b = G1 R1 B2 G2 | c = R2 B3 G3 R3   (little endian)
            y0 = gray_of(a & 255, (a >> 8) & 255, (a >> 16) & 255);
            y1 = gray_of(a >> 24, b & 255, (b >> 8) & 255);
            y2 = gray_of((b >> 16) & 255, b >> 24, c & 255);
            y3 = gray_of((c >> 8) & 255, (c >> 16) & 255, c >> 24);
        } else {
            const uint32_t *q = (const uint32_t *)p;
            const uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
            y0 = gray_of(a & 255, (a >> 8) & 255, (a >> 16) & 255);
            y1 = gray_of(b & 255, (b >> 8) & 255, (b >> 16) & 255);
            y2 = gray_of(c & 255, (c >> 8) & 255, (c >> 16) & 255);
            y3 = gray_of(d & 255, (d >> 8) & 255, (d >> 16) & 255);
        }
        *(uint32_t *)(out + x0) = y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
    } else {
        for (int k = 0; k < 4 && x0 + k < w; ++k) {
            const uint8_t *q = p + k * ch;
            out[x0 + k] = (uint8_t)gray_of(q[0], q[1], q[2]);
        }
    }
}

// K0: raw 16-bit depth -> metres, float(z) * scale with one rounding: what cv::Mat::convertTo(CV_32FC1, 1.0f / DEPTH_SCALE_FACTOR)
// computes in the reference's reader (io/iowrapperRGBD.cpp:327).  8 pixels per thread (one 128-bit load, two 128-bit stores).
__global__ void __launch_bounds__(256) k_depth_u16(const uint16_t *__restrict__ raw, size_t frame_px, float scale,
                                                   const ImgLevel *__restrict__ desc, int px)
{
    const int f = blockIdx.z;
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= px) return;
    const uint16_t *src = raw + (size_t)f * frame_px + i0;
    float *dst = desc[f].depth + i0;
    if (i0 + 8 <= px && ((((uintptr_t)src) & 15) == 0) && ((((uintptr_t)dst) & 15) == 0)) {
        const uint4 v = __ldg((const uint4 *)src);
        const unsigned wv[4] = {v.x, v.y, v.z, v.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[2 * k] = __fmul_rn((float)(wv[k] & 0xffffu), scale);
            o[2 * k + 1] = __fmul_rn((float)(wv[k] >> 16), scale);
        }
        *(float4 *)dst = make_float4(o[0], o[1], o[2], o[3]);
        *(float4 *)(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    } else {
        for (int k = 0; k < 8 && i0 + k < px; ++k) dst[k] = __fmul_rn((float)src[k], scale);
    }
}

int launch_depth_u16(revo_ctx *ctx, const uint16_t *d_raw, size_t frame_px, float scale, const ImgLevel *d_desc, int n, int px)
{
    dim3 grid(cdiv(cdiv(px, 8), 256), 1, n);
    k_depth_u16<<<grid, 256, 0, ctx->stream>>>(d_raw, frame_px, scale, d_desc, px);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

// 3-channel fast path: 16 pixels per thread = three 128-bit loads (48 bytes, 16-byte aligned when the row is) and one 128-bit
// store: four times the bytes in flight per thread of k_gray, whose top stall is the load (long scoreboard 15.6 warps per issue).
__global__ void __launch_bounds__(256) k_gray16(const uint8_t *__restrict__ bgr, size_t stride, size_t frame_bytes,
                                                const ImgLevel *__restrict__ desc, int w, int h)
{
    const int f = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= h || x0 >= w) return;
    const uint4 *p = (const uint4 *)(bgr + (size_t)f * frame_bytes + (size_t)y * stride + (size_t)x0 * 3);
    const uint4 A = __ldg(p), B = __ldg(p + 1), C = __ldg(p + 2);
    const uint32_t v[12] = {A.x, A.y, A.z, A.w, B.x, B.y, B.z, B.w, C.x, C.y, C.z, C.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = v[3 * q], b = v[3 * q + 1], c = v[3 * q + 2];
        // a = B0 G0 R0 B1 | b = G1 R1 B2 G2 | c = R2 B3 G3 R3   (little endian)
        const uint32_t y0 = gray_of(a & 255, (a >> 8) & 255, (a >> 16) & 255);
        const uint32_t y1 = gray_of(a >> 24, b & 255, (b >> 8) & 255);
        const uint32_t y2 = gray_of((b >> 16) & 255, b >> 24, c & 255);
        const uint32_t y3 = gray_of((c >> 8) & 255, (c >> 16) & 255, c >> 24);
        o[q] = y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
    }
    *(uint4 *)(desc[f].gray + (size_t)y * w + x0) = make_uint4(o[0], o[1], o[2], o[3]);
}

int launch_gray(revo_ctx *ctx, const uint8_t *d_bgr, size_t stride, int ch, size_t frame_bytes, const ImgLevel *d_desc,
                int n, int w, int h)
{
    static const int no16 = getenv("REVO_GRAY_NO16") ? atoi(getenv("REVO_GRAY_NO16")) : 0;      // A/B switch
    // every frame's image and gray plane are 256-byte aligned (slab chunks, staging buffers); rows stay 16-byte aligned when
    // the width is a multiple of 16 and the input rows are tight or 16-byte pitched
    if (ch == 3 && !no16 && (w & 15) == 0 && (stride & 15) == 0 && (frame_bytes & 15) == 0 && (((uintptr_t)d_bgr) & 15) == 0) {
        dim3 block(8, 32), grid(cdiv(w / 16, 8), cdiv(h, 32), n);
        k_gray16<<<grid, block, 0, ctx->stream>>>(d_bgr, stride, frame_bytes, d_desc, w, h);
        LAUNCH_CHECK(ctx);
        return REVO_OK;
    }
    dim3 block(32, 8), grid(cdiv(cdiv(w, 4), 32), cdiv(h, 8), n);
    k_gray<<<grid, block, 0, ctx->stream>>>(d_bgr, stride, ch, frame_bytes, d_desc, w, h);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

// ---------------------------------------------------------------------------
// K3: pyrDown 8U: separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8.
// 32x8 outputs per CTA; input tile 67x19 staged in shared memory.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// Register-tiled: one thread -> 4 horizontally adjacent outputs of one row.  Per input row (5 of them) the 11 bytes it
// needs come from four aligned 32-bit loads (columns 2x-4 .. 2x+11); the horizontal [1 4 6 4 1] of an output is one
// DP4A on a PRMT-aligned word plus one byte; the 2.5-fold vertical reuse of input rows between neighbouring output rows
// is served by L1.  BORDER_REFLECT_101 without divergence: the first thread of a row synthesises its left halo word
// from its own first word (columns -2,-1 = columns 2,1), the last one takes column ws from column ws-2.
__device__ __forceinline__ int reflect101_near(int i, int n)   // |overshoot| <= 2 < n
{
    return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i);
}

__global__ void __launch_bounds__(256) k_pyrdown(const ImgLevel *__restrict__ src, const ImgLevel *__restrict__ dst, int ws,
                                                 int hs, int wd, int hd)
{
    const int f = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= wd || y >= hd) return;
    const uint8_t *__restrict__ in = src[f].gray;
    uint8_t *__restrict__ out = dst[f].gray;
    const int c0 = 2 * x0 - 4;   // column of byte 0 of the 16-byte window
    // vector path: ws a multiple of 8 (so wd = ws/2 is a multiple of 4 and every thread owns 4 outputs), >= 16, aligned
    const bool vec = ((ws & 7) == 0) && ws >= 16 && ((((uintptr_t)in) & 3) == 0) && hs >= 4;
    int hsum[5][4];
    if (vec) {
        const bool left = c0 < 0, right = c0 + 16 > ws;    // at most one of them (ws >= 16)
        constexpr unsigned kW = 0x04060401u;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const uint32_t *q = (const uint32_t *)(in + (size_t)reflect101_near(2 * y - 2 + r, hs) * ws + c0);
            const uint32_t w1 = __ldg(q + 1), w2 = __ldg(q + 2);
            uint32_t w0, w3;
            if (left) w0 = __byte_perm(w1, 0u, 0x1200); else w0 = __ldg(q);            // bytes 2,3 = columns 2,1
            if (right) w3 = (w2 >> 16) & 255u; else w3 = __ldg(q + 3);                // byte 0 = column ws-2
            // output k is centred on byte 2k+4: bytes 2k+2 .. 2k+5 times (1,4,6,4) + byte 2k+6
            hsum[r][0] = (int)__dp4a(__byte_perm(w0, w1, 0x5432), kW, (w1 >> 16) & 255u);
            hsum[r][1] = (int)__dp4a(w1, kW, w2 & 255u);
            hsum[r][2] = (int)__dp4a(__byte_perm(w1, w2, 0x5432), kW, (w2 >> 16) & 255u);
            hsum[r][3] = (int)__dp4a(w2, kW, w3 & 255u);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const uint8_t *__restrict__ row = in + (size_t)reflect101(2 * y - 2 + r, hs) * ws;
            int b[16];
#pragma unroll
            for (int i = 2; i <= 12; ++i) b[i] = row[reflect101(c0 + i, ws)];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 2 * k + 4;
                hsum[r][k] = b[i - 2] + 4 * b[i - 1] + 6 * b[i] + 4 * b[i + 1] + b[i + 2];
            }
        }
    }
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int v = hsum[0][k] + 4 * hsum[1][k] + 6 * hsum[2][k] + 4 * hsum[3][k] + hsum[4][k];
        o |= (uint32_t)((v + 128) >> 8) << (8 * k);
    }
    uint8_t *op = out + (size_t)y * wd + x0;
    if (x0 + 4 <= wd && ((((uintptr_t)op) & 3) == 0)) {
        *(uint32_t *)op = o;
    } else {
        for (int k = 0; k < 4 && x0 + k < wd; ++k) op[k] = (uint8_t)(o >> (8 * k));
    }
}

// K4: FilterSubsampleWithHoles: mean of the >0 entries of each 2x2 block (NaN excluded by the compare).
__device__ __forceinline__ float depth_half_of(float a, float b, float c, float d)
{
    float acc = 0.f, n = 0.f;
    if (a > 0.0f) { acc = __fadd_rn(acc, a); n += 1.f; }
    if (b > 0.0f) { acc = __fadd_rn(acc, b); n += 1.f; }
    if (c > 0.0f) { acc = __fadd_rn(acc, c); n += 1.f; }
    if (d > 0.0f) { acc = __fadd_rn(acc, d); n += 1.f; }
    if (n > 0.f) acc = __fdiv_rn(acc, n);
    return acc;
}

// one thread -> 4 outputs of one row: two float4 loads from each of the two input rows, one float4 store
__global__ void __launch_bounds__(256) k_depth_half(const ImgLevel *__restrict__ src, const ImgLevel *__restrict__ dst, int ws,
                                                    int wd, int hd)
{
    const int f = blockIdx.z;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= wd || y >= hd) return;
    const float *__restrict__ in = src[f].depth;
    float *__restrict__ out = dst[f].depth + (size_t)y * wd + x0;
    const float *r0 = in + (size_t)(2 * y) * ws + 2 * x0, *r1 = r0 + ws;
    if (x0 + 4 <= wd && ((ws & 3) == 0) && ((wd & 3) == 0) && ((((uintptr_t)in) & 15) == 0) && ((((uintptr_t)dst[f].depth) & 15) == 0)) {
        const float4 a0 = __ldg((const float4 *)r0), a1 = __ldg((const float4 *)r0 + 1);
        const float4 b0 = __ldg((const float4 *)r1), b1 = __ldg((const float4 *)r1 + 1);
        *(float4 *)out = make_float4(depth_half_of(a0.x, a0.y, b0.x, b0.y), depth_half_of(a0.z, a0.w, b0.z, b0.w),
                                     depth_half_of(a1.x, a1.y, b1.x, b1.y), depth_half_of(a1.z, a1.w, b1.z, b1.w));
    } else {
        for (int k = 0; k < 4 && x0 + k < wd; ++k) out[k] = depth_half_of(r0[2 * k], r0[2 * k + 1], r1[2 * k], r1[2 * k + 1]);
    }
}

int launch_pyrdown(revo_ctx *ctx, const ImgLevel *d_src, const ImgLevel *d_dst, int n, int w_dst, int h_dst, int w_src, int h_src)
{
    dim3 block(32, 8), grid(cdiv(cdiv(w_dst, 4), 32), cdiv(h_dst, 8), n);
    k_pyrdown<<<grid, block, 0, ctx->stream>>>(d_src, d_dst, w_src, h_src, w_dst, h_dst);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

int launch_depth_half(revo_ctx *ctx, const ImgLevel *d_src, const ImgLevel *d_dst, int n, int w_dst, int h_dst, int w_src)
{
    dim3 block(32, 8), grid(cdiv(cdiv(w_dst, 4), 32), cdiv(h_dst, 8), n);
    k_depth_half<<<grid, block, 0, ctx->stream>>>(d_src, d_dst, w_src, w_dst, h_dst);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

int launch_pyrdown_depth(revo_ctx *ctx, const ImgLevel *d_src, const ImgLevel *d_dst, int n, int w_dst, int h_dst,
                         int w_src, int h_src)
{
    int rc = launch_pyrdown(ctx, d_src, d_dst, n, w_dst, h_dst, w_src, h_src);
    if (!rc) rc = launch_depth_half(ctx, d_src, d_dst, n, w_dst, h_dst, w_src);
    return rc;
}

// K2 (Canny) lives in canny.cu.

// ---------------------------------------------------------------------------
// K5: fill-in from the level above (the patch histogram itself -- u8 counts that wrap like cv::Mat_<uchar>::operator++,
// number of non-empty patches -- is accumulated by the Canny output kernels and finalised by k_hist_finalize, canny.cu).
// ---------------------------------------------------------------------------
// fillInEdges: this-level pixel (ox,oy) <- top pixel (2ox+1, 2oy+1) when the patch of the top pixel has
// fewer than 0.05 P^2 edge pixels AT THIS LEVEL and the whole level has < n_percentage non-empty patches.
__global__ void __launch_bounds__(256) k_fill_in(const ImgLevel *__restrict__ desc, const ImgLevel *__restrict__ top, int w, int h,
                                                 int P, int P_low, float n_percentage)
{
    const int f = blockIdx.z;
    const ImgLevel L = desc[f];
    const float frac = __fdiv_rn((float)(*L.nz_patches), (float)(L.hist_w * L.hist_h));
    if (!(frac < n_percentage)) return;      // the usual case: the whole CTA leaves (few CTAs: 8 rows per thread)
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    if (ox >= w) return;
    const int wt = top[f].w, ht = top[f].h;
    const int xx = 2 * ox + 1;
    if (xx >= wt) return;
    const int px = xx / P_low;
    if (px >= L.hist_w) return;
    for (int oy = (blockIdx.y * blockDim.y + threadIdx.y) * 8, k = 0; k < 8 && oy < h; ++k, ++oy) {
        const int yy = 2 * oy + 1;
        if (yy >= ht) break;
        const int py = yy / P_low;
        if (py >= L.hist_h) break;
        if ((double)L.hist[(size_t)py * L.hist_w + px] < (double)(P * P) * 0.05) {
            if (top[f].edges[(size_t)yy * wt + xx]) L.edges[(size_t)oy * w + ox] = 255;
        }
    }
}

int launch_hist_fill(revo_ctx *ctx, const ImgLevel *d_desc, const ImgLevel *d_top, int n, int w, int h, int patch,
                     int patch_low, bool do_fill, float n_percentage)
{
    const int hist_w = w / patch, hist_h = h / patch;
    if (hist_w > 0 && hist_h > 0) {
        // the histogram itself is produced by the Canny output kernel (canny.cu: k_canny_final + k_hist_finalize)
        if (do_fill) {
            dim3 block(32, 8), g2(cdiv(w, 32), cdiv(h, 64), n);
            k_fill_in<<<g2, block, 0, ctx->stream>>>(d_desc, d_top, w, h, patch, patch_low, n_percentage);
            LAUNCH_CHECK(ctx);
        }
    }
    return REVO_OK;
}

// ---------------------------------------------------------------------------
// K6: 3-D edge list.  One warp per 8x4 tile; deterministic tile-major order
// (count -> exclusive scan -> scatter).  X = Z (x - cx) / fx exactly as the reference.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool edge_point_ok(const ImgLevel &L, int x, int y, int w, int h, float dmin, float dmax, float &Z)
{
    if (x >= w || y >= h) return false;
    // edge test first: ~94 % of the pixels stop here and never touch the 4-byte depth plane
    if (L.edges[(size_t)y * w + x] == 0) return false;
    Z = L.depth[(size_t)y * w + x];
    return isfinite(Z) && Z > dmin && Z < dmax;
}

__global__ void __launch_bounds__(256) k_tile_count(const ImgLevel *__restrict__ desc, int w, int h, int tiles_x, int n_tiles,
                                                    float dmin, float dmax)
{
    const int f = blockIdx.z;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const int lane = threadIdx.x & 31;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    float Z;
    const bool ok = edge_point_ok(desc[f], tx * kTileW + (lane & 7), ty * kTileH + (lane >> 3), w, h, dmin, dmax, Z);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) desc[f].tile_off[tile] = __popc(m);
}

// in-place exclusive scan of tile_off[0..n_tiles) ; tile_off[n_tiles] = total ; n_pts = min(total, cap)
__global__ void __launch_bounds__(1024) k_tile_scan(const ImgLevel *__restrict__ desc, int n_tiles)
{
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int f = blockIdx.x;
    int *off = desc[f].tile_off;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_tiles ? off[i] : 0;
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int ws = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) ws += t;
            }
            warp_sums[lane] = ws;   // inclusive
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (wid ? warp_sums[wid - 1] : 0) + s - v;
        if (i < n_tiles) off[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        off[n_tiles] = carry_s;
        *desc[f].n_pts = min(carry_s, desc[f].pts_cap);
    }
}

__global__ void __launch_bounds__(256) k_tile_scatter(const ImgLevel *__restrict__ desc, int w, int h, int tiles_x, int n_tiles,
                                                      float dmin, float dmax)
{
    const int f = blockIdx.z;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tile >= n_tiles) return;
    const int lane = threadIdx.x & 31;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    const ImgLevel &L = desc[f];
    const int x = tx * kTileW + (lane & 7), y = ty * kTileH + (lane >> 3);
    float Z;
    const bool ok = edge_point_ok(L, x, y, w, h, dmin, dmax, Z);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (!ok) return;
    const int o = L.tile_off[tile] + __popc(m & ((1u << lane) - 1u));
    if (o >= L.pts_cap) return;
    const float X = __fdiv_rn(__fmul_rn(Z, __fsub_rn((float)x, L.cx)), L.fx);
    const float Y = __fdiv_rn(__fmul_rn(Z, __fsub_rn((float)y, L.cy)), L.fy);
    L.pts[o] = make_float4(X, Y, Z, 1.0f);
}

// Batched variant (n >= 8 frames): two streaming kernels, one warp per GROUP of 32 horizontally adjacent tiles
// (a 256 x 4 pixel strip, so every edge-map row segment a warp touches is one coalesced 256-byte read).
//   k_group_mask : lane = tile; the four 8-byte row segments of the tile -> 32-bit mask of its edge pixels, depth is
//                  fetched only for set bits (~7 % of the pixels), the validity mask goes to tile_off[tile] and the
//                  number of points of the group to gcnt[group] (scratch: the frame's label plane, free after Canny);
//   k_group_scatter: group offset = sum of the counts of all preceding groups (a few hundred ints, warp-reduced),
//                  exclusive scan over the 32 tiles, then every lane writes the points of its set bits.
// Same deterministic tile-major order as the three-kernel path.
constexpr int kGroupTiles = 32;

__device__ __forceinline__ unsigned tile_edge_mask(const ImgLevel &L, int tx, int ty, int w, int h, float dmin, float dmax)
{
    const int x0 = tx * kTileW, y0 = ty * kTileH;
    unsigned m = 0;
    const bool fast = (x0 + kTileW <= w) && ((w & 7) == 0) && ((((uintptr_t)L.edges) & 7) == 0);
#pragma unroll
    for (int r = 0; r < kTileH; ++r) {
        const int y = y0 + r;
        if (y >= h) break;
        unsigned long long e8 = 0;
        if (fast) {
            e8 = *(const unsigned long long *)(L.edges + (size_t)y * w + x0);
        } else {
            for (int c = 0; c < kTileW && x0 + c < w; ++c) e8 |= (unsigned long long)L.edges[(size_t)y * w + x0 + c] << (8 * c);
        }
        if (!e8) continue;
#pragma unroll
        for (int c = 0; c < kTileW; ++c)
            if ((e8 >> (8 * c)) & 0xffull) m |= 1u << (r * kTileW + c);
    }
    // depth test only where an edge pixel is (isfinite, dmin < Z < dmax: imgpyramidrgbd.cpp:210-214); four loads in flight
    // per lane (the loop is latency-bound: one dependent global load per set bit otherwise)
    unsigned keep = 0;
    const float *__restrict__ dp = L.depth + (size_t)y0 * w + x0;
    for (unsigned mm = m; mm;) {
        int b[4];
        float Z[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            b[k] = mm ? __ffs(mm) - 1 : -1;
            mm &= mm - 1;       // 0 stays 0
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) Z[k] = b[k] >= 0 ? __ldg(dp + (size_t)(b[k] >> 3) * w + (b[k] & 7)) : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (b[k] >= 0 && isfinite(Z[k]) && Z[k] > dmin && Z[k] < dmax) keep |= 1u << b[k];
    }
    return keep;
}

__global__ void __launch_bounds__(256) k_group_mask(const ImgLevel *__restrict__ desc, int w, int h, int tiles_x, int tiles_y, int groups_x,
                                                    float dmin, float dmax)
{
    const int f = blockIdx.z;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= groups_x * tiles_y) return;
    const int lane = threadIdx.x & 31;
    const ImgLevel &L = desc[f];
    const int ty = g / groups_x, tx = (g - ty * groups_x) * kGroupTiles + lane;
    unsigned keep = 0;
    if (tx < tiles_x) {
        keep = tile_edge_mask(L, tx, ty, w, h, dmin, dmax);
        ((unsigned *)L.tile_off)[ty * tiles_x + tx] = keep;
    }
    const int cnt = __reduce_add_sync(0xffffffffu, __popc(keep));
    if (lane == 0) L.labels[g] = cnt;
}

__global__ void __launch_bounds__(256) k_group_scatter(const ImgLevel *__restrict__ desc, int w, int h, int tiles_x, int tiles_y,
                                                       int groups_x)
{
    const int f = blockIdx.z;
    const int n_groups = groups_x * tiles_y;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const int lane = threadIdx.x & 31;
    const ImgLevel &L = desc[f];
    const int *__restrict__ gcnt = L.labels;
    int before = 0;
    for (int i = lane; i < g; i += 32) before += gcnt[i];
    before = __reduce_add_sync(0xffffffffu, before);
    const int ty = g / groups_x, tx = (g - ty * groups_x) * kGroupTiles + lane;
    const unsigned keep = tx < tiles_x ? ((const unsigned *)L.tile_off)[ty * tiles_x + tx] : 0u;
    const int c = __popc(keep);
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    int o = before + incl - c;
    const int x0 = tx * kTileW, y0 = ty * kTileH;
    const float *__restrict__ dp = L.depth + (size_t)y0 * w + x0;
    for (unsigned mm = keep; mm;) {      // four depth loads in flight per lane
        int b[4];
        float Z[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            b[k] = mm ? __ffs(mm) - 1 : -1;
            mm &= mm - 1;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) Z[k] = b[k] >= 0 ? __ldg(dp + (size_t)(b[k] >> 3) * w + (b[k] & 7)) : 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (b[k] < 0 || o >= L.pts_cap) continue;
            const int x = x0 + (b[k] & 7), y = y0 + (b[k] >> 3);
            const float X = __fdiv_rn(__fmul_rn(Z[k], __fsub_rn((float)x, L.cx)), L.fx);
            const float Y = __fdiv_rn(__fmul_rn(Z[k], __fsub_rn((float)y, L.cy)), L.fy);
            L.pts[o++] = make_float4(X, Y, Z[k], 1.0f);
        }
    }
    if (g == n_groups - 1) {
        const int total = before + __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) *L.n_pts = min(total, L.pts_cap);
    }
}

int launch_compact(revo_ctx *ctx, const ImgLevel *d_desc, int n, int w, int h, float dmin, float dmax)
{
    const int tiles_x = cdiv(w, kTileW), tiles_y = cdiv(h, kTileH), n_tiles = tiles_x * tiles_y;
    const int groups_x = cdiv(tiles_x, kGroupTiles), n_groups = groups_x * tiles_y;
    // the group counts live in the label plane of the frame (w0*h0 ints); it always holds n_groups <= w*h/4 + h ints
    if (n >= 8 && (size_t)n_groups <= (size_t)w * h) {
        dim3 grid(cdiv(n_groups, 8), 1, n);
        k_group_mask<<<grid, 256, 0, ctx->stream>>>(d_desc, w, h, tiles_x, tiles_y, groups_x, dmin, dmax);
        LAUNCH_CHECK(ctx);
        k_group_scatter<<<grid, 256, 0, ctx->stream>>>(d_desc, w, h, tiles_x, tiles_y, groups_x);
        LAUNCH_CHECK(ctx);
        return REVO_OK;
    }
    dim3 grid(cdiv(n_tiles, 8), 1, n);
    k_tile_count<<<grid, 256, 0, ctx->stream>>>(d_desc, w, h, tiles_x, n_tiles, dmin, dmax);
    LAUNCH_CHECK(ctx);
    k_tile_scan<<<n, 1024, 0, ctx->stream>>>(d_desc, n_tiles);
    LAUNCH_CHECK(ctx);
    k_tile_scatter<<<grid, 256, 0, ctx->stream>>>(d_desc, w, h, tiles_x, n_tiles, dmin, dmax);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

// Reference order (xx outer, yy inner -- imgpyramidrgbd.cpp:203-205) for the return3DEdges accessor.
__global__ void k_col_count(const ImgLevel *__restrict__ desc, int w, int h, float dmin, float dmax, int *col_off)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    int c = 0;
    float Z;
    for (int y = 0; y < h; ++y) c += edge_point_ok(desc[0], x, y, w, h, dmin, dmax, Z);
    col_off[x] = c;
}
__global__ void k_col_scan(int *col_off, int w, int *n_out)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int s = 0;
        for (int x = 0; x < w; ++x) { const int c = col_off[x]; col_off[x] = s; s += c; }
        *n_out = s;
    }
}
__global__ void k_col_scatter(const ImgLevel *__restrict__ desc, int w, int h, float dmin, float dmax, const int *col_off,
                              float4 *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const ImgLevel &L = desc[0];
    int o = col_off[x];
    float Z;
    for (int y = 0; y < h; ++y)
        if (edge_point_ok(L, x, y, w, h, dmin, dmax, Z)) {
            const float X = __fdiv_rn(__fmul_rn(Z, __fsub_rn((float)x, L.cx)), L.fx);
            const float Y = __fdiv_rn(__fmul_rn(Z, __fsub_rn((float)y, L.cy)), L.fy);
            out[o++] = make_float4(X, Y, Z, 1.0f);
        }
}

int launch_edges3d_reference_order(revo_ctx *ctx, const ImgLevel *d_desc_one, int w, int h, float dmin, float dmax,
                                   float4 *d_out, int *d_n, int *d_col_off)
{
    k_col_count<<<cdiv(w, 128), 128, 0, ctx->stream>>>(d_desc_one, w, h, dmin, dmax, d_col_off);
    LAUNCH_CHECK(ctx);
    k_col_scan<<<1, 32, 0, ctx->stream>>>(d_col_off, w, d_n);
    LAUNCH_CHECK(ctx);
    k_col_scatter<<<cdiv(w, 128), 128, 0, ctx->stream>>>(d_desc_one, w, h, dmin, dmax, d_col_off, d_out);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

// ---------------------------------------------------------------------------
// generateColoredPcl (imgpyramidrgbd.cpp:279-327): the viewer's coloured cloud of one level, columns
// (X, Y, Z, 1, r, g, b, 1) in the reference's column-major scan order.  The colour image of the level is cv::pyrDown of the
// full-resolution one, channel by channel (k_pyrdown_color: plain version of K3, not on the hot path); the compaction is
// K6's reference-order path (count per column -> scan -> scatter) with the depth test of isPointOkDepth and, unless the
// dense cloud is asked for, the edge label.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pyrdown_color(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int ws, int hs, int wd, int hd,
                                                       int ch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= wd) return;
    constexpr int kw[5] = {1, 4, 6, 4, 1};
    for (int c = 0; c < ch; ++c) {
        int s = 0;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const uint8_t *row = in + (size_t)reflect101(2 * y - 2 + r, hs) * ws * ch;
            int hsum = 0;
#pragma unroll
            for (int k = 0; k < 5; ++k) hsum += kw[k] * row[(size_t)reflect101(2 * x - 2 + k, ws) * ch + c];
            s += kw[r] * hsum;
        }
        out[((size_t)y * wd + x) * ch + c] = (uint8_t)((s + 128) >> 8);
    }
}

__device__ __forceinline__ bool pcl_point_ok(const ImgLevel &L, int x, int y, int w, bool dense, float dmin, float dmax, float &Z)
{
    Z = L.depth[(size_t)y * w + x];
    if (!(isfinite(Z) && Z > dmin && Z < dmax)) return false;      // isPointOkDepth, imgpyramidrgbd.h:163-166
    return dense || L.edges[(size_t)y * w + x] > 0;
}
__global__ void k_pcl_col_count(const ImgLevel *__restrict__ desc, int w, int h, int dense, float dmin, float dmax, int *col_off)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    int c = 0;
    float Z;
    for (int y = 0; y < h; ++y) c += pcl_point_ok(desc[0], x, y, w, dense != 0, dmin, dmax, Z);
    col_off[x] = c;
}
__global__ void k_pcl_col_scatter(const ImgLevel *__restrict__ desc, int w, int h, int dense, float dmin, float dmax, const int *col_off,
                                  const uint8_t *__restrict__ bgr, int ch, float *__restrict__ out, int cap)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const ImgLevel &L = desc[0];
    int o = col_off[x];
    float Z;
    for (int y = 0; y < h; ++y)
        if (pcl_point_ok(L, x, y, w, dense != 0, dmin, dmax, Z)) {
            if (o >= cap) return;
            const float X = __fdiv_rn(__fmul_rn(Z, __fsub_rn((float)x, L.cx)), L.fx);
            const float Y = __fdiv_rn(__fmul_rn(Z, __fsub_rn((float)y, L.cy)), L.fy);
            const uint8_t *c = bgr + ((size_t)y * w + x) * ch;
            float4 *q = (float4 *)(out + (size_t)o * 8);
            q[0] = make_float4(X, Y, Z, 1.0f);
            q[1] = make_float4(__fdiv_rn((float)c[2], 255.0f), __fdiv_rn((float)c[1], 255.0f), __fdiv_rn((float)c[0], 255.0f), 1.0f);
            ++o;
        }
}

int launch_pyrdown_color(revo_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, int ws, int hs, int ch)
{
    const int wd = (ws + 1) / 2, hd = (hs + 1) / 2;
    k_pyrdown_color<<<dim3(cdiv(wd, 256), hd), 256, 0, ctx->stream>>>(d_in, d_out, ws, hs, wd, hd, ch);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

int launch_colored_pcl(revo_ctx *ctx, const ImgLevel *d_desc_one, int w, int h, int dense, float dmin, float dmax, const uint8_t *d_bgr, int ch,
                       float *d_out, int cap, int *d_n, int *d_col_off)
{
    k_pcl_col_count<<<cdiv(w, 128), 128, 0, ctx->stream>>>(d_desc_one, w, h, dense, dmin, dmax, d_col_off);
    LAUNCH_CHECK(ctx);
    k_col_scan<<<1, 32, 0, ctx->stream>>>(d_col_off, w, d_n);
    LAUNCH_CHECK(ctx);
    if (d_out) {
        k_pcl_col_scatter<<<cdiv(w, 128), 128, 0, ctx->stream>>>(d_desc_one, w, h, dense, dmin, dmax, d_col_off, d_bgr, ch, d_out, cap);
        LAUNCH_CHECK(ctx);
    }
    return REVO_OK;
}

// ---------------------------------------------------------------------------
// K7: exact Euclidean distance transform to the nearest edge pixel, out = sqrtf(d2).
//  (a) column pass: vertical distance g(x,y) to the nearest edge in the column (u16, kEdtInf if none).  A column is cut into
//      segments of <= 64 rows; a thread owns (column, segment), packs the segment's edge bytes into a 64-bit mask, exchanges
//      "first / last edge row of my segment" with the other segments of its column through shared memory, and gets every row's
//      distance to the nearest edge above / below with clz / ffs on the mask: no sequential dependency along the column (the
//      first version walked all h rows twice per thread and was latency-bound), 1 byte read + 2 bytes written per pixel.
//  (b) row pass: d2(x,y) = min_j (x-j)^2 + g(j,y)^2 by an outward search that stops once r^2 >= best
//      (exact; typical DT values are small so the search is short).
// K8: {0.5(dt[i-1]-dt[i+1]), 0.5(dt[i-w]-dt[i+w]), dt[i], 0} for rows 1..h-2, zeros elsewhere (linear indices: column 0 takes
//     its left neighbour from the end of the previous row, like the reference); texels are stored tile by tile (full lines).
//     (Doing (b) and K8 in one kernel for a band of rows -- distances of band + halo rows in shared memory, dt never read
//     back -- was measured slower than the two kernels, 1.78 vs 1.66 ms per 256 promotions: 38 KB of shared memory per CTA cost
//     more occupancy in the latency-bound search than the saved dt read is worth.)
// ---------------------------------------------------------------------------
constexpr int kEdtInf = 1 << 14;
// value of every pixel when the edge map is empty: what OpenCV's own trueDistTrans returns (cv2 4.13, IPP off)
constexpr float kEdtEmpty = 65536.0f;

__global__ void __launch_bounds__(1024) k_edt_cols(const ImgLevel *__restrict__ desc, int w, int h, int seg_rows)
{
    __shared__ short first_s[32][32];                      // [segment][column of the block]: first / last edge row, -1 = none
    __shared__ short last_s[32][32];
    const int f = blockIdx.z;
    const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5, n_seg = blockDim.x >> 5;
    const int x = blockIdx.x * 32 + lane;
    const int y0 = seg * seg_rows, y1 = min(h, y0 + seg_rows);
    const bool in = x < w;
    const uint8_t *__restrict__ e = desc[f].edges;
    unsigned long long m = 0;                               // bit k = edge at row y0 + k
    if (in)
        for (int y = y0; y < y1; ++y) m |= (unsigned long long)(e[(size_t)y * w + x] != 0) << (y - y0);
    first_s[seg][lane] = m ? (short)(y0 + __ffsll((long long)m) - 1) : (short)-1;
    last_s[seg][lane] = m ? (short)(y0 + 63 - __clzll((long long)m)) : (short)-1;
    __syncthreads();
    if (!in) return;
    int above = -kEdtInf, below = 2 * kEdtInf;              // nearest edge rows outside the segment
    for (int s2 = seg - 1; s2 >= 0; --s2)
        if (last_s[s2][lane] >= 0) { above = last_s[s2][lane]; break; }
    for (int s2 = seg + 1; s2 < n_seg; ++s2)
        if (first_s[s2][lane] >= 0) { below = first_s[s2][lane]; break; }
    unsigned short *__restrict__ g = (unsigned short *)desc[f].labels;
    for (int y = y0; y < y1; ++y) {
        const int k = y - y0;
        const unsigned long long lo = m & (~0ull >> (63 - k));      // edges at rows <= y inside the segment
        const unsigned long long hi = m >> k;                        // edges at rows >= y
        const int up = lo ? k - (63 - __clzll((long long)lo)) : y - above;
        const int dn = hi ? __ffsll((long long)hi) - 1 : below - y;
        g[(size_t)y * w + x] = (unsigned short)min(min(up, dn), kEdtInf);
    }
}

// one row of (b): squared distance of pixel x from the column distances of its row.  [jmin, jmax]: the columns of the row whose
// column distance is finite -- a column without any edge can never win, so the search never leaves that span (an empty or
// nearly empty edge map would otherwise cost O(w) steps per pixel)
__device__ __forceinline__ float edt_row_px(const unsigned short *__restrict__ grow, int x, int jmin, int jmax)
{
    if (jmin > jmax) return kEdtEmpty;
    const int g0 = grow[x];
    int best = g0 * g0;
    const int rmax = max(x - jmin, jmax - x);
    for (int r = 1; r <= rmax && r * r < best; ++r) {
        const int r2 = r * r;
        if (x - r >= jmin) { const int gl = grow[x - r]; best = min(best, r2 + gl * gl); }
        if (x + r <= jmax) { const int gr = grow[x + r]; best = min(best, r2 + gr * gr); }
    }
    return best >= kEdtInf * kEdtInf ? kEdtEmpty : sqrtf((float)best);
}

__global__ void __launch_bounds__(256) k_edt_rows(const ImgLevel *__restrict__ desc, int w, int h)
{
    extern __shared__ unsigned short grow[];
    __shared__ int span_lo;
    __shared__ int span_hi;
    const int f = blockIdx.z, y = blockIdx.x;
    const unsigned short *__restrict__ g = (const unsigned short *)desc[f].labels + (size_t)y * w;
    if (threadIdx.x == 0) { span_lo = w; span_hi = -1; }
    __syncthreads();
    int lo = w, hi = -1;
    for (int i = threadIdx.x; i < w; i += blockDim.x) {
        const unsigned short v = g[i];
        grow[i] = v;
        if (v < kEdtInf) { lo = min(lo, i); hi = max(hi, i); }
    }
    if (hi >= 0) { atomicMin(&span_lo, lo); atomicMax(&span_hi, hi); }
    __syncthreads();
    const int jmin = span_lo, jmax = span_hi;
    float *__restrict__ out = desc[f].dt + (size_t)y * w;
    for (int x = threadIdx.x; x < w; x += blockDim.x) out[x] = edt_row_px(grow, x, jmin, jmax);
}

// The reference's {gx, gy, dt, .} float4 texel (imgpyramidrgbd.cpp:255-276) at linear index i; zeros in rows 0 and h-1.
__device__ __forceinline__ float4 opt_texel(const float *__restrict__ dt, size_t i, int w, int h)
{
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i >= (size_t)w && i < (size_t)w * (h - 1)) {
        o.x = __fmul_rn(0.5f, __fsub_rn(dt[i - 1], dt[i + 1]));
        o.y = __fmul_rn(0.5f, __fsub_rn(dt[i - w], dt[i + w]));
        o.z = dt[i];
    }
    return o;
}

// Gradient components lie in [-1, 1] (|dt[a] - dt[b]| <= 2 for pixels two apart): 16-bit fixed point, step 1/32764 (a multiple of 4, so the frequent exact values 0, +-1/4, +-1/2, +-1 carry no rounding bias).
__device__ __forceinline__ uint32_t pack_grad(float gx, float gy)
{
    const int qx = __float2int_rn(fminf(fmaxf(gx, -1.f), 1.f) * 32764.f);
    const int qy = __float2int_rn(fminf(fmaxf(gy, -1.f), 1.f) * 32764.f);
    return ((uint32_t)qx & 0xffffu) | ((uint32_t)qy << 16);
}

// K8 (device layout, internal.h: opt_texel_index): one 8-byte texel {dt float32 | snorm16 gx, gy} per pixel in 4x4 tiles.
// dt stays float32, only the Jacobian direction is quantised (1.5e-5 absolute).  The reference's float4 array is produced on
// demand for the accessor (k_opt_struct_f4).
__device__ __forceinline__ uint2 pack_texel(const float4 t)
{
    return make_uint2(__float_as_uint(t.z), pack_grad(t.x, t.y));
}

__global__ void __launch_bounds__(256) k_opt_struct(const ImgLevel *__restrict__ desc, int w, int h)
{
    // a CTA covers 64 columns x 4 rows = 16 tiles; 16 consecutive threads write the 16 texels (128 bytes = one line) of a tile
    const int f = blockIdx.z;
    const int tile = threadIdx.x >> 4, k = threadIdx.x & 15;
    const int x = blockIdx.x * 64 + tile * 4 + (k & 3), y = blockIdx.y * 4 + (k >> 2);
    if (x >= w || y >= h) return;
    const float *__restrict__ dt = desc[f].dt;
    desc[f].opt[opt_texel_index(x, y, (w + 3) >> 2)] = pack_texel(opt_texel(dt, (size_t)y * w + x, w, h));
}

// the reference layout, for returnOptimizationStructure(): out[i] = {gx, gy, dt, 0}
__global__ void __launch_bounds__(256) k_opt_struct_f4(const float *__restrict__ dt, int w, int h, float4 *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)w * h) return;
    out[i] = opt_texel(dt, i, w, h);
}

// test hook: caller-provided float4 structure -> device layout
__global__ void __launch_bounds__(256) k_opt_pack_from_f4(const float4 *__restrict__ in, int w, int h, uint2 *__restrict__ out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    out[opt_texel_index(x, y, (w + 3) >> 2)] = pack_texel(in[(size_t)y * w + x]);
}

int launch_opt_struct_f4(revo_ctx *ctx, const float *d_dt, int w, int h, float4 *d_out)
{
    k_opt_struct_f4<<<cdiv(w * h, 256), 256, 0, ctx->stream>>>(d_dt, w, h, d_out);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

int launch_opt_pack_from_f4(revo_ctx *ctx, const float4 *d_in, int w, int h, uint2 *d_out)
{
    k_opt_pack_from_f4<<<dim3(cdiv(w, 256), h), 256, 0, ctx->stream>>>(d_in, w, h, d_out);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

// ---------------------------------------------------------------------------
// Tracking-quality vote (TrackerNew::assessTrackingQuality, system/tracker.cpp:118-201).
//  k_quality_scatter: one thread per (past frame, 3-D point): newPt = R pt + T, u = fx x / z + cx, v = fy y / z + cy in the
//      reference's float operation order; in-bounds projections set bit `frame` of the pixel's byte (atomicOr on the
//      containing word: "prevent coinciding reprojections" -- a frame counts a pixel once), M = popcount.
//  k_quality_hist: one thread per pixel of the current frame: valid depth -> histogram[M]++, and overlaps[M]++ if the pixel
//      is a Canny edge; block-level shared counters, one global atomic per counter and block.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_quality_scatter(const QualityArgs *__restrict__ args, unsigned *__restrict__ mbits_all, size_t words,
                                                         int *__restrict__ counters_all)
{
    const QualityArgs &a = args[blockIdx.z];
    if ((int)blockIdx.y >= a.n_frames) return;
    const QualityFrame &F = a.fr[blockIdx.y];
    unsigned *mbits = mbits_all + words * blockIdx.z;
    int *counters = counters_all + 16 * blockIdx.z;
    const int n = *F.n_pts;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = __ldg(F.pts + i);
        // Eigen: R * pt + T (column-major accumulation order), then tracker.cpp:157-158
        const float X = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(F.R[0], p.x), __fmul_rn(F.R[3], p.y)), __fmul_rn(F.R[6], p.z)), F.T[0]);
        const float Y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(F.R[1], p.x), __fmul_rn(F.R[4], p.y)), __fmul_rn(F.R[7], p.z)), F.T[1]);
        const float Z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(F.R[2], p.x), __fmul_rn(F.R[5], p.y)), __fmul_rn(F.R[8], p.z)), F.T[2]);
        const float u = __fadd_rn(__fdiv_rn(__fmul_rn(a.fx, X), Z), a.cx);
        const float v = __fadd_rn(__fdiv_rn(__fmul_rn(a.fy, Y), Z), a.cy);
        if (u >= 0.f && u < (float)a.w && v >= 0.f && v < (float)a.h) {
            const int px = (int)floorf(v) * a.w + (int)floorf(u);
            atomicOr(mbits + (px >> 2), 1u << ((px & 3) * 8 + blockIdx.y));
        } else {
            atomicAdd(counters + 8, 1);
        }
    }
}

__global__ void __launch_bounds__(256) k_quality_hist(const QualityArgs *__restrict__ args, const unsigned *__restrict__ mbits_all, size_t words,
                                                      int n_px, float dmin, float dmax, int *__restrict__ counters_all)
{
    __shared__ int sc[8];
    if (threadIdx.x < 8) sc[threadIdx.x] = 0;
    __syncthreads();
    const QualityArgs &a = args[blockIdx.y];
    const unsigned *mbits = mbits_all + words * blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_px) {
        const float Z = a.depth[i];
        if (isfinite(Z) && Z > dmin && Z < dmax) {     // ImgPyramidRGBD::isPointOkDepth
            const int val = __popc((mbits[i >> 2] >> ((i & 3) * 8)) & 0xffu);
            atomicAdd(&sc[val & 3], 1);
            if (a.edges[i]) atomicAdd(&sc[4 + (val & 3)], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 8 && sc[threadIdx.x]) atomicAdd(counters_all + 16 * blockIdx.y + threadIdx.x, sc[threadIdx.x]);
}

int launch_quality(revo_ctx *ctx, const QualityArgs *d_args, int n, int w, int h, float dmin, float dmax, unsigned *d_mbits, int *d_counters)
{
    const size_t words = ((size_t)w * h + 3) / 4;
    REVO_CUDA(ctx, cudaMemsetAsync(d_mbits, 0, words * 4 * n, ctx->stream));
    REVO_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 16 * sizeof(int) * (size_t)n, ctx->stream));
    k_quality_scatter<<<dim3(n > 16 ? 8 : 64, 3, n), 256, 0, ctx->stream>>>(d_args, d_mbits, words, d_counters);
    LAUNCH_CHECK(ctx);
    k_quality_hist<<<dim3(cdiv(w * h, 256), n), 256, 0, ctx->stream>>>(d_args, d_mbits, words, w * h, dmin, dmax, d_counters);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

// Copies of 3-D edge lists (TrackerNew::addOldPclAndPose keeps `return3DEdges(histogramLevel)` by value, tracker.cpp:209-224):
// block (x, f) copies a share of list f and block (0, f) its count.
__global__ void __launch_bounds__(256) k_copy_point_lists(const PointListCopy *__restrict__ tab)
{
    const PointListCopy c = tab[blockIdx.y];
    const int n = min(*c.src_n, c.cap);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) c.dst[i] = __ldg(c.src + i);
    if (blockIdx.x == 0 && threadIdx.x == 0) *c.dst_n = n;
}
int launch_copy_point_lists(revo_ctx *ctx, const PointListCopy *d_tab, int n)
{
    k_copy_point_lists<<<dim3(4, n), 256, 0, ctx->stream>>>(d_tab);
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

int launch_keyframe(revo_ctx *ctx, const ImgLevel *d_desc, int n, int w, int h)
{
    {
        // segments of <= 64 rows, one warp each (h <= 2048)
        int n_seg = cdiv(h, 64);
        n_seg = n_seg < 1 ? 1 : n_seg;
        if (n_seg > 32) return REVO_ERR_UNSUPPORTED;
        dim3 grid(cdiv(w, 32), 1, n);
        k_edt_cols<<<grid, n_seg * 32, 0, ctx->stream>>>(d_desc, w, h, cdiv(h, n_seg));
        LAUNCH_CHECK(ctx);
    }
    {
        dim3 grid(h, 1, n);
        k_edt_rows<<<grid, 256, w * sizeof(unsigned short), ctx->stream>>>(d_desc, w, h);
        LAUNCH_CHECK(ctx);
    }
    {
        dim3 grid(cdiv(w, 64), cdiv(h, 4), n);
        k_opt_struct<<<grid, 256, 0, ctx->stream>>>(d_desc, w, h);
        LAUNCH_CHECK(ctx);
    }
    return REVO_OK;
}

}  // namespace revo
