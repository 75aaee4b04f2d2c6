// canny.cu -- K2: cv::Canny(gray, edges, t1, t2, 3, true)  (datastructures/imgpyramidrgbd.cpp:184), bit-exact.
//
// Three kernels per pyramid level, batched over frames:
//  (A) k_canny_tile: one CTA per 128x16 pixel tile.  The gray tile (+2 halo) is fetched with ONE TMA bulk-tensor
//      copy (cp.async.bulk.tensor.3d, zero fill outside the image; BORDER_REPLICATE is patched in shared
//      memory), 3x3 Sobel -> mag = dx^2 + dy^2 (zero outside the image) -> non-maximum suppression with OpenCV's
//      TG22 fixed-point sector test -> class map (0 none / 1 weak / 2 strong).  The hysteresis inside the tile
//      is resolved right there: union-find over the tile's candidates in shared memory (8-connectivity), with
//      the "strong" flag folded into the key (strong keys are smaller, roots are minima) so a component's root
//      tells whether it holds a strong pixel.  Writes the class map and one global label per candidate
//      (= key of its tile-local root in global coordinates).
//  (B) k_canny_merge: only candidates on tile borders: union with candidate neighbours in adjacent tiles
//      (lock-free atomicMin union-find in global memory; trees are at most a few tiles deep).
//  (C) k_canny_final: 16 pixels per thread: 255 where the candidate's root key is strong, else 0, into both
//      edges and edges_orig.
// The result is the unique fixed point of OpenCV's hysteresis (every 8-connected component of candidates that
// contains a strong pixel), independent of thread order.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "internal.h"

namespace revo {

#define LAUNCH_CHECK(ctx)                                   \
    do {                                                    \
        (ctx)->launches++;                                  \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return cuda_fail((ctx), e__, __func__); \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

constexpr int CT_W = 128, CT_H = 16;                 // output tile
// TMA box: origin (x0 - CT_XO, y0 - 2).  The innermost start coordinate of a bulk-tensor copy must be a multiple
// of 16 bytes (measured on B200: any other start raises "illegal instruction"), so the 2-pixel halo is fetched as
// a 16-pixel apron on both sides; rows may start anywhere.  Out-of-image cells are zero-filled by the TMA unit.
constexpr int CT_XO = 16;
constexpr int CT_BW = CT_W + 2 * CT_XO, CT_BH = CT_H + 4;
#define GRAY(r, c) g[(r)][(c) + CT_XO - 2]   // (r, c) relative to (y0 - 2, x0 - 2), as the stencil code indexes
constexpr int CT_MW = CT_W + 8;
constexpr int kWeakBit = 0x40000000;
constexpr int kIdxMask = 0x3fffffff;
constexpr int kNoLabel = 0x7fffffff;

// ---- TMA / mbarrier PTX ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// ---- union-find on keys ----------------------------------------------------------------------------
// key = (weak ? kWeakBit : 0) | index of the pixel; a pixel's label is the key of its parent; roots point
// to themselves.  Linking always attaches the larger key under the smaller, so a root is the minimum key of
// its component and is "strong" (bit clear) iff the component contains a strong pixel.
template <typename LoadFn>
__device__ __forceinline__ int uf_find_key(LoadFn load, int key)
{
    int p = load(key & kIdxMask);
    while (p != key) {
        key = p;
        p = load(key & kIdxMask);
    }
    return key;
}

__device__ __forceinline__ void uf_union_smem(int *lab, int ka, int kb)
{
    auto ld = [&](int i) { return ((volatile int *)lab)[i]; };
    while (true) {
        ka = uf_find_key(ld, ka);
        kb = uf_find_key(ld, kb);
        if (ka == kb) return;
        if (ka < kb) { const int t = ka; ka = kb; kb = t; }
        const int old = atomicMin(lab + (ka & kIdxMask), kb);
        if (old == ka) return;
        ka = old;
    }
}

__device__ __forceinline__ void uf_union_gmem(int *lab, int ka, int kb)
{
    auto ld = [&](int i) { return __ldcg(lab + i); };
    while (true) {
        ka = uf_find_key(ld, ka);
        kb = uf_find_key(ld, kb);
        if (ka == kb) return;
        if (ka < kb) { const int t = ka; ka = kb; kb = t; }
        const int old = atomicMin(lab + (ka & kIdxMask), kb);
        if (old == ka) return;
        ka = old;
    }
}

// ---- (A) ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_canny_tile(const __grid_constant__ CUtensorMap tm_gray, const int use_tma,
                                                    const ImgLevel *__restrict__ desc, int w, int h, int low, int high)
{
    __shared__ alignas(128) uint8_t g[CT_BH][CT_BW];
    __shared__ alignas(16) int mag[CT_H + 2][CT_MW];   // pixel column tx lives at array column tx + 4
    __shared__ alignas(16) int lab[CT_H * CT_W];
    __shared__ alignas(16) uint8_t cls[CT_H][CT_W];
    __shared__ alignas(8) uint64_t bar;

    const int f = blockIdx.z;
    const ImgLevel L = desc[f];
    const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    const int tid = threadIdx.x;

    // ---- gray tile with halo 2
    if (use_tma) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, CT_BW * CT_BH);
            tma_load_3d(&g[0][0], &tm_gray, x0 - CT_XO, y0 - 2, f, &bar);
        }
        mbar_wait(&bar, 0);
        // BORDER_REPLICATE: cells outside the image take the value of the clamped cell (always inside this box)
        const bool edge_tile = (x0 == 0) || (y0 == 0) || (x0 + CT_W + 2 > w) || (y0 + CT_H + 2 > h);
        if (edge_tile) {
            for (int i = tid; i < CT_BH * (CT_W + 4); i += 256) {
                const int r = i / (CT_W + 4), c = i - r * (CT_W + 4);
                const int gy = y0 - 2 + r, gx = x0 - 2 + c;
                if (gx < 0 || gx >= w || gy < 0 || gy >= h) {
                    const int sy = min(max(gy, 0), h - 1) - (y0 - 2), sx = min(max(gx, 0), w - 1) - (x0 - 2);
                    GRAY(r, c) = GRAY(sy, sx);
                }
            }
        }
    } else {
        for (int i = tid; i < CT_BH * (CT_W + 4); i += 256) {
            const int r = i / (CT_W + 4), c = i - r * (CT_W + 4);
            const int yy = min(max(y0 + r - 2, 0), h - 1), xx = min(max(x0 + c - 2, 0), w - 1);
            GRAY(r, c) = L.gray[(size_t)yy * w + xx];
        }
    }
    __syncthreads();

    // ---- Sobel + squared magnitude, register tiled: thread -> 8 consecutive pixels of one row.
    // Three rows x 24 bytes of the gray tile are read as 9 LDS.64; dx, dy and mag of the 8 pixels stay in registers.
    const int ty = tid >> 4, tx0 = (tid & 15) * 8;
    int dxs[8], dys[8], mid[10];
    {
        unsigned long long rw[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int q = 0; q < 3; ++q) rw[r][q] = *(const unsigned long long *)&g[ty + 1 + r][tx0 + 8 + 8 * q];
        auto px = [&](int r, int j) -> int { return (int)((rw[r][j >> 3] >> ((j & 7) * 8)) & 0xffull); };
        const bool row_in = (y0 + ty) < h;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int l = k + 7, c = k + 8, rr = k + 9;
            const int gx = (px(0, rr) - px(0, l)) + 2 * (px(1, rr) - px(1, l)) + (px(2, rr) - px(2, l));
            const int gy = (px(2, l) + 2 * px(2, c) + px(2, rr)) - (px(0, l) + 2 * px(0, c) + px(0, rr));
            dxs[k] = gx;
            dys[k] = gy;
            mid[k + 1] = (row_in && (x0 + tx0 + k) < w) ? gx * gx + gy * gy : 0;
        }
        *(int4 *)&mag[ty + 1][tx0 + 4] = make_int4(mid[1], mid[2], mid[3], mid[4]);
        *(int4 *)&mag[ty + 1][tx0 + 8] = make_int4(mid[5], mid[6], mid[7], mid[8]);
    }
    // halo ring of the magnitude tile (pixel rows -1 and CT_H, pixel columns -1 and CT_W)
    for (int i = tid; i < 2 * (CT_W + 2) + 2 * CT_H; i += 256) {
        int py, pxx;
        if (i < CT_W + 2) { py = -1; pxx = i - 1; }
        else if (i < 2 * (CT_W + 2)) { py = CT_H; pxx = i - (CT_W + 2) - 1; }
        else if (i < 2 * (CT_W + 2) + CT_H) { py = i - 2 * (CT_W + 2); pxx = -1; }
        else { py = i - 2 * (CT_W + 2) - CT_H; pxx = CT_W; }
        const int yy = y0 + py, xx = x0 + pxx;
        int m = 0;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
            const int r = py + 1, c = pxx + 1;   // top-left of the 3x3 window in GRAY coordinates
            const int gx = (GRAY(r, c + 2) - GRAY(r, c)) + 2 * (GRAY(r + 1, c + 2) - GRAY(r + 1, c)) + (GRAY(r + 2, c + 2) - GRAY(r + 2, c));
            const int gy = (GRAY(r + 2, c) - GRAY(r, c)) + 2 * (GRAY(r + 2, c + 1) - GRAY(r, c + 1)) + (GRAY(r + 2, c + 2) - GRAY(r, c + 2));
            m = gx * gx + gy * gy;
        }
        mag[py + 1][pxx + 4] = m;
    }
    __syncthreads();

    // ---- non-maximum suppression from registers (rows above / below: 4 LDS.128 each)
    int c8[8];
    {
        int up[16], dn[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int4 u = *(const int4 *)&mag[ty][tx0 + 4 * q];
            const int4 d = *(const int4 *)&mag[ty + 2][tx0 + 4 * q];
            up[4 * q] = u.x; up[4 * q + 1] = u.y; up[4 * q + 2] = u.z; up[4 * q + 3] = u.w;
            dn[4 * q] = d.x; dn[4 * q + 1] = d.y; dn[4 * q + 2] = d.z; dn[4 * q + 3] = d.w;
        }
        mid[0] = mag[ty + 1][tx0 + 3];
        mid[9] = mag[ty + 1][tx0 + 12];
        // pixel k sits at array column k + 4 of up/dn and k + 1 of mid
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int m = mid[k + 1];
            int cls_k = 0;
            if (m > low) {
                const int xs = dxs[k], ys = dys[k];
                const int ax = abs(xs), ay = abs(ys) << 15;
                const int tg22x = ax * 13573;
                bool cand;
                if (ay < tg22x) {
                    cand = (m > mid[k]) && (m >= mid[k + 2]);
                } else {
                    const int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x) cand = (m > up[k + 4]) && (m >= dn[k + 4]);
                    else {
                        const bool neg = (xs ^ ys) < 0;          // s = -1: compare (y-1, x+1) and (y+1, x-1)
                        const int a = neg ? up[k + 5] : up[k + 3];
                        const int b2 = neg ? dn[k + 3] : dn[k + 5];
                        cand = (m > a) && (m > b2);
                    }
                }
                if (cand) cls_k = (m > high) ? 2 : 1;
            }
            c8[k] = cls_k;
        }
    }

    // ---- labels: every horizontal run inside the thread's 8 pixels starts flat (all point to the run's minimum key)
    int key[8];
    {
        int runmin[8];
        const int li0 = ty * CT_W + tx0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            key[k] = (c8[k] == 2 ? 0 : kWeakBit) | (li0 + k);
            runmin[k] = c8[k] ? ((k > 0 && c8[k - 1]) ? min(runmin[k - 1], key[k]) : key[k]) : kNoLabel;
        }
#pragma unroll
        for (int k = 6; k >= 0; --k)
            if (c8[k] && c8[k + 1]) runmin[k] = min(runmin[k], runmin[k + 1]);
        *(int4 *)&lab[li0] = make_int4(runmin[0], runmin[1], runmin[2], runmin[3]);
        *(int4 *)&lab[li0 + 4] = make_int4(runmin[4], runmin[5], runmin[6], runmin[7]);
        unsigned lo4 = 0, hi4 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { lo4 |= (unsigned)c8[k] << (8 * k); hi4 |= (unsigned)c8[k + 4] << (8 * k); }
        *(uint2 *)&cls[ty][tx0] = make_uint2(lo4, hi4);
    }
    __syncthreads();

    // ---- hysteresis inside the tile: runs are linked to their W neighbour and to the row above (N, else NW / NE)
    {
        const int li0 = ty * CT_W + tx0;
        if (c8[0] && tx0 > 0) {
            const int cw = cls[ty][tx0 - 1];
            if (cw) uf_union_smem(lab, key[0], (cw == 2 ? 0 : kWeakBit) | (li0 - 1));
        }
        if (ty > 0) {
            int cu[10];
            cu[0] = tx0 > 0 ? cls[ty - 1][tx0 - 1] : 0;
            {
                const uint2 v = *(const uint2 *)&cls[ty - 1][tx0];
#pragma unroll
                for (int k = 0; k < 4; ++k) { cu[1 + k] = (v.x >> (8 * k)) & 255; cu[5 + k] = (v.y >> (8 * k)) & 255; }
            }
            cu[9] = tx0 + 8 < CT_W ? cls[ty - 1][tx0 + 8] : 0;
            const int ui0 = li0 - CT_W;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (!c8[k]) continue;
                if (cu[k + 1]) {
                    uf_union_smem(lab, key[k], (cu[k + 1] == 2 ? 0 : kWeakBit) | (ui0 + k));
                } else {
                    if (cu[k]) uf_union_smem(lab, key[k], (cu[k] == 2 ? 0 : kWeakBit) | (ui0 + k - 1));
                    if (cu[k + 2]) uf_union_smem(lab, key[k], (cu[k + 2] == 2 ? 0 : kWeakBit) | (ui0 + k + 1));
                }
            }
        }
    }
    __syncthreads();

    // ---- write the class map (class of the pixel's tile-local ROOT: 2 = its component holds a strong pixel, 1 = weak so
    // far) and, for every candidate, the key of its tile-local root in global coordinates
    const int y = y0 + ty;
    if (y < h && x0 + tx0 < w) {
        auto ld = [&](int i) { return lab[i]; };
        unsigned lo4 = 0, hi4 = 0;
        int *lrow = L.labels + (size_t)y * w + x0 + tx0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (!c8[k] || x0 + tx0 + k >= w) continue;
            const int root = uf_find_key(ld, key[k]);
            const int ri = root & kIdxMask;
            const int gidx = (y0 + ri / CT_W) * w + (x0 + (ri % CT_W));
            lrow[k] = (root & kWeakBit) | gidx;
            const unsigned rc = (root & kWeakBit) ? 1u : 2u;
            if (k < 4) lo4 |= rc << (8 * k); else hi4 |= rc << (8 * (k - 4));
        }
        uint8_t *erow = L.edges + (size_t)y * w + x0 + tx0;
        if (x0 + tx0 + 8 <= w && ((((uintptr_t)erow) & 7) == 0)) {
            *(uint2 *)erow = make_uint2(lo4, hi4);
        } else {
            for (int k = 0; k < 8 && x0 + tx0 + k < w; ++k) erow[k] = (uint8_t)(((k < 4 ? lo4 : hi4) >> (8 * (k & 3))) & 255u);
        }
    }
}

// ---- (B): candidates on the top row / left column / right column of every tile ----------------------------
__global__ void __launch_bounds__(192) k_canny_merge(const ImgLevel *__restrict__ desc, int w, int h)
{
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
    const int t = threadIdx.x;
    int lx, ly;
    if (t < CT_W) { lx = t; ly = 0; }                                   // top row
    else if (t < CT_W + CT_H - 1) { lx = 0; ly = t - CT_W + 1; }         // left column (below the corner)
    else if (t < CT_W + 2 * (CT_H - 1)) { lx = CT_W - 1; ly = t - (CT_W + CT_H - 1) + 1; }   // right column
    else return;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= w || y >= h) return;
    const uint8_t *__restrict__ cls = desc[f].edges;
    int *lab = desc[f].labels;
    const int p = y * w + x;
    const int cp = cls[p];
    if (!cp) return;
    const int key = (cp == 2 ? 0 : kWeakBit) | p;
    auto other_tile = [&](int qx, int qy) { return (qx / CT_W != x / CT_W) || (qy / CT_H != y / CT_H); };
    auto try_union = [&](int qx, int qy) {
        if (qx < 0 || qx >= w || qy < 0 || qy >= h || !other_tile(qx, qy)) return;
        const int q = qy * w + qx;
        const int cq = cls[q];
        if (cq && !(cp == 2 && cq == 2)) uf_union_gmem(lab, key, (cq == 2 ? 0 : kWeakBit) | q);   // strong-strong: nothing to learn
    };
    try_union(x - 1, y);
    try_union(x - 1, y - 1);
    try_union(x, y - 1);
    try_union(x + 1, y - 1);
}

// ---- (C) ---------------------------------------------------------------------------------------------------
// Also accumulates the patch histogram of generateDistHistogram (imgpyramidrgbd.cpp:146-172) for the edge pixels it
// emits (integer counters in the per-frame scratch, finalised to the reference's wrapping u8 by k_hist_finalize).
__device__ __forceinline__ void hist_add(const ImgLevel &L, int *cnt, int p, int w, int P)
{
    const int y = p / w, x = p - y * w;
    const int py = y / P, px = x / P;
    if (py < L.hist_h && px < L.hist_w) atomicAdd(cnt + py * L.hist_w + px, 1);
}

__global__ void __launch_bounds__(256) k_canny_final(const ImgLevel *__restrict__ desc, int w, int h, int P)
{
    const int f = blockIdx.z;
    const size_t n = (size_t)w * h;
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= n) return;
    uint8_t *e = desc[f].edges, *eo = desc[f].edges_orig;
    const int *lab = desc[f].labels;
    int *cnt = (int *)desc[f].flags;
    auto ld = [&](int i) { return __ldcg(lab + i); };
    if (i0 + 16 <= n && ((((uintptr_t)(e + i0)) & 15) == 0) && ((((uintptr_t)(eo + i0)) & 15) == 0)) {
        uint4 v = *(const uint4 *)(e + i0);
        uint32_t wv[4] = {v.x, v.y, v.z, v.w};
        if (v.x | v.y | v.z | v.w) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (!wv[q]) continue;
                uint32_t o = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t c = (wv[q] >> (8 * b)) & 255u;
                    if (c == 2) o |= 255u << (8 * b);        // its tile-local component already holds a strong pixel
                    else if (c) {
                        const int p = (int)i0 + q * 4 + b;
                        const int root = uf_find_key(ld, kWeakBit | p);
                        if (!(root & kWeakBit)) o |= 255u << (8 * b);
                    }
                }
                wv[q] = o;
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if ((o >> (8 * b)) & 255u) hist_add(desc[f], cnt, (int)i0 + q * 4 + b, w, P);
            }
            v = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            *(uint4 *)(e + i0) = v;
        }
        *(uint4 *)(eo + i0) = v;
    } else {
        for (size_t i = i0; i < i0 + 16 && i < n; ++i) {
            const int c = e[i];
            uint8_t o = 0;
            if (c == 2) o = 255;
            else if (c) {
                const int root = uf_find_key(ld, kWeakBit | (int)i);
                o = (root & kWeakBit) ? 0 : 255;
            }
            e[i] = o;
            eo[i] = o;
            if (o) hist_add(desc[f], cnt, (int)i, w, P);
        }
    }
}

// ---- host: tensor map for the gray images of one level of a slab ---------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn()
{
    static PFN_tmapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_tmapEncodeTiled)p;
        (void)cudaGetLastError();
    }
    return fn;
}

bool make_gray_tensor_map(void *tmap_out, const uint8_t *base, int w, int h, int n_frames, size_t frame_stride)
{
    PFN_tmapEncodeTiled fn = get_encode_fn();
    if (!fn || (w % 16) != 0 || (frame_stride % 16) != 0 || (((uintptr_t)base) & 15)) return false;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n_frames};
    cuuint64_t strides[2] = {(cuuint64_t)w, (cuuint64_t)frame_stride};
    cuuint32_t box[3] = {CT_BW, CT_BH, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn((CUtensorMap *)tmap_out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void *)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// counts -> wrapping u8 histogram + number of non-empty patches (countNonZero(dist), imgpyramidrgbd.cpp:160)
__global__ void __launch_bounds__(256) k_hist_finalize(const ImgLevel *__restrict__ desc)
{
    __shared__ int red[8];
    const ImgLevel &L = desc[blockIdx.x];
    const int *cnt = (const int *)L.flags;
    const int n = L.hist_w * L.hist_h;
    int nz = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint8_t v = (uint8_t)(cnt[i] & 255);
        L.hist[i] = v;
        nz += v != 0;
    }
    nz = __reduce_add_sync(0xffffffffu, nz);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nz;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < 8; ++k) s += red[k];
        *L.nz_patches = s;
    }
}

// ===================================================================================================================
// Canny v4: bit-mask pipeline (default).  Same result as the tile / union-find path above, bit for bit:
//  (1) k_canny_nms : streaming, no shared memory.  A lane owns 4 pixel columns and slides down a strip of rows keeping
//      three rows of squared gradient magnitudes in registers; the 3x3 Sobel of a pixel is five DP4A (u8 x s8 dot
//      products) on byte-aligned windows built with PRMT from the lane's own 32-bit load and its neighbours' (two
//      shuffles); non-maximum suppression with OpenCV's TG22 fixed-point sector test; the result is TWO BITS per pixel,
//      written as two bit masks (candidates C, strong S; 64 pixels per 64-bit word, LSB = leftmost) into the frame's
//      label plane: 1/16 of the bytes of a class map, and the form the hysteresis wants.
//  (2) k_canny_hyst: one CTA per image.  Hysteresis = S <- every candidate 8-connected to a strong pixel.  A warp owns
//      a band of rows, a lane owns a 64-bit word of a row; a row is flooded in O(1) word operations with the carry
//      trick  up = (((C + S) ^ C) & C) | S  (and its bit-reversed twin for the other direction), lanes exchange their
//      edge bits by shuffle; sweeping a band down and up propagates any distance inside the band, bands exchange
//      boundary rows between sweeps (block barrier) until no bit changes: exact, typically 2-4 sweeps.
//  (3) k_canny_expand: S -> edges / edges_orig bytes (0 / 255) + the integer patch counters of the histogram.
// ===================================================================================================================
constexpr int NMS_RS = 33;      // output rows per warp strip (a multiple of the 3-fold register rotation)
constexpr int NMS_PX = 4;       // pixels per lane

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ unsigned rep_byte(unsigned b) { return (b & 0xffu) * 0x01010101u; }

// Requires w % 4 == 0, w >= 8 and a 4-byte aligned gray plane (launch_canny checks).
__device__ __forceinline__ void canny_nms_body(const ImgLevel *__restrict__ desc, int w, int h, int low, int high, int wp32,
                                               int rows_per_strip /* multiple of 3 */)
{
    const int f = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y0 = (blockIdx.y * 4 + warp) * rows_per_strip;
    if (y0 >= h) return;
    const int xw = blockIdx.x * 32 * NMS_PX;          // first column of the warp
    const int x = xw + lane * NMS_PX;                 // first column of the lane
    const uint8_t *__restrict__ g = desc[f].gray;
    unsigned *__restrict__ maskC = (unsigned *)desc[f].labels;
    unsigned *__restrict__ maskS = maskC + (size_t)wp32 * h;
    constexpr int W_DX1 = 0x000100FF, W_DX2 = 0x000200FE, W_DYM = 0x00FFFEFF, W_DYP = 0x00010201;

    // per-lane constants of the BORDER_REPLICATE column handling
    const bool beyond = x >= w;                        // lane entirely right of the image: replicates column w-1
    const int xl = beyond ? w - 4 : x;                 // column of the word this lane loads
    const bool halo_l_in = x >= 4, halo_r_in = x + 4 < w;
    // magnitude columns x-1 .. x+4 that lie inside the image (magnitudes outside are 0), as AND masks
    int cm[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) cm[c] = (x - 1 + c >= 0 && x - 1 + c < w) ? -1 : 0;
    unsigned own_px = 0;                               // the lane's own pixels inside the image (4 bits)
#pragma unroll
    for (int k = 0; k < 4; ++k) own_px |= (x + k < w ? 1u : 0u) << k;
    const int wi = (xw >> 5) + (lane >> 3);
    const bool writer = (lane & 7) == 0 && wi < wp32;

    // A gray row is fetched one step ahead of its use (raw own word + the halo word of lanes 0 / 31), then turned into
    // aligned windows: aw[c] = bytes of columns (x-2+c .. x+1+c), c = 0..5  <->  pixel column x-1+c
    const bool edge_lane = lane == 0 || lane == 31;
    const int xh = lane == 0 ? x - 4 : x + 4;          // column of the halo word an edge lane loads (if inside the image)
    const bool halo_in = lane == 0 ? halo_l_in : halo_r_in;
    auto issue_row = [&](int y, unsigned &v, unsigned &hv) {
        const uint8_t *row = g + (size_t)min(max(y, 0), h - 1) * w;
        v = __ldg((const unsigned *)(row + xl));
        hv = 0;
        if (edge_lane && halo_in) hv = __ldg((const unsigned *)(row + xh));
    };
    auto finish_row = [&](unsigned v, unsigned hv, unsigned (&aw)[6]) {
        const unsigned w1 = beyond ? rep_byte(v >> 24) : v;
        unsigned w0 = __shfl_up_sync(0xffffffffu, w1, 1), w2 = __shfl_down_sync(0xffffffffu, w1, 1);
        if (lane == 0) w0 = halo_l_in ? hv : rep_byte(w1);
        if (lane == 31) w2 = halo_r_in ? hv : rep_byte(w1 >> 24);
        aw[0] = __byte_perm(w0, w1, 0x5432);
        aw[1] = __byte_perm(w0, w1, 0x6543);
        aw[2] = w1;
        aw[3] = __byte_perm(w1, w2, 0x4321);
        aw[4] = __byte_perm(w1, w2, 0x5432);
        aw[5] = __byte_perm(w1, w2, 0x6543);
    };
    auto load_row = [&](int y, unsigned (&aw)[6]) {
        unsigned v, hv;
        issue_row(y, v, hv);
        finish_row(v, hv, aw);
    };
    // squared magnitude of row ym for the 6 columns x-1..x+4 (0 outside the image) and dx, dy of the lane's own 4
    auto mag_row = [&](int ym, const unsigned (&r0)[6], const unsigned (&r1)[6], const unsigned (&r2)[6], int (&mg)[6], int (&dxo)[4],
                       int (&dyo)[4]) {
        const int rowm = (ym >= 0 && ym < h) ? -1 : 0;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int dx = dp4a_us(r0[c], W_DX1, dp4a_us(r1[c], W_DX2, dp4a_us(r2[c], W_DX1, 0)));
            const int dy = dp4a_us(r2[c], W_DYP, dp4a_us(r0[c], W_DYM, 0));
            mg[c] = (dx * dx + dy * dy) & cm[c] & rowm;
            if (c >= 1 && c <= 4) { dxo[c - 1] = dx; dyo[c - 1] = dy; }
        }
    };
    // non-maximum suppression of row yn (middle magnitudes mm, rows above / below mu / md) -> 4 candidate / strong bits
    auto nms_row = [&](int yn, const int (&mu)[6], const int (&mm)[6], const int (&md)[6], const int (&dxs)[4], const int (&dys)[4]) {
        if (yn >= h) return;                              // warp-uniform
        unsigned cb = 0, sb = 0;
        const int mx = max(max(mm[1], mm[2]), max(mm[3], mm[4]));
        if (__any_sync(0xffffffffu, mx > low)) {          // most 128-pixel row segments of a real image hold no candidate
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int m = mm[k + 1];
                const int xs = dxs[k], ys = dys[k];
                const int ax = abs(xs), ay = abs(ys) << 15;
                const int tg22x = ax * 13573;
                const int tg67x = tg22x + (ax << 16);
                const bool horiz = ay < tg22x, vert = ay > tg67x;
                const bool neg = (xs ^ ys) < 0;           // s = -1: compare (y-1, x+1) and (y+1, x-1)
                const int a = horiz ? mm[k] : (vert ? mu[k + 1] : (neg ? mu[k + 2] : mu[k]));
                const int b = horiz ? mm[k + 2] : (vert ? md[k + 1] : (neg ? md[k] : md[k + 2]));
                // horizontal / vertical: m > a && m >= b ; diagonal: m > a && m > b   <=>   m > b - (horiz || vert)
                const bool cand = (m > low) && (m > a) && (m > b - ((horiz || vert) ? 1 : 0));
                cb |= (cand ? 1u : 0u) << k;
                sb |= ((cand && m > high) ? 1u : 0u) << k;
            }
            cb &= own_px;
            sb &= own_px;
        }
        // OR over the 8 lanes of a mask word (xor butterflies stay inside the aligned group of 8)
        unsigned wc = cb << (4 * (lane & 7)), wsx = sb << (4 * (lane & 7));
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            wc |= __shfl_xor_sync(0xffffffffu, wc, d);
            wsx |= __shfl_xor_sync(0xffffffffu, wsx, d);
        }
        if (writer) {
            maskC[(size_t)yn * wp32 + wi] = wc;
            maskS[(size_t)yn * wp32 + wi] = wsx;
        }
    };

    unsigned ra[6], rb[6], rc[6];
    int ma[6], mb[6], mc[6];
    int dxa[4], dya[4], dxb[4], dyb[4], dxc[4], dyc[4];
    // prologue: magnitudes of row y0-1 need gray rows y0-2 .. y0
    load_row(y0 - 2, ra);
    load_row(y0 - 1, rb);
    load_row(y0, rc);
    mag_row(y0 - 1, ra, rb, rc, ma, dxa, dya);       // slot a: row y0-1
    load_row(y0 + 1, ra);
    mag_row(y0, rb, rc, ra, mb, dxb, dyb);           // slot b: row y0
    // steady state, unrolled by three so that the register slots rotate without moves:
    //   gray rows held: (rc, ra) = (y, y+1) ; magnitudes held: (ma, mb) = (y-1, y) ; row y+2 is in flight (pv, ph)
    const int y_end = min(y0 + rows_per_strip, h);
    unsigned pv, ph, qv, qh;
    issue_row(y0 + 2, pv, ph);
    for (int y = y0; y < y_end; y += 3) {
        issue_row(y + 3, qv, qh);
        finish_row(pv, ph, rb);
        mag_row(y + 1, rc, ra, rb, mc, dxc, dyc);
        nms_row(y, ma, mb, mc, dxb, dyb);
        issue_row(y + 4, pv, ph);
        finish_row(qv, qh, rc);
        mag_row(y + 2, ra, rb, rc, ma, dxa, dya);
        nms_row(y + 1, mb, mc, ma, dxc, dyc);
        issue_row(y + 5, qv, qh);
        finish_row(pv, ph, ra);
        mag_row(y + 3, rb, rc, ra, mb, dxb, dyb);
        nms_row(y + 2, mc, ma, mb, dxa, dya);
        pv = qv; ph = qh;
    }
}
// The same body at three register budgets (A/B switch REVO_NMS_MINBLOCKS: 4 = the compiler's 111 registers, 5 = at most 96 (default:
// the kernel's top stall is the gray-row load, 20 instead of 16 resident warps are worth 2 % of the build), 6 = at most 80 with a
// few spilled words (no better than 4)).
__global__ void __launch_bounds__(128) k_canny_nms(const ImgLevel *__restrict__ desc, int w, int h, int low, int high, int wp32, int rows_per_strip)
{
    canny_nms_body(desc, w, h, low, high, wp32, rows_per_strip);     // warp-synchronous inside (__shfl_*_sync, __any_sync)
}
__global__ void __launch_bounds__(128, 5) k_canny_nms_mb5(const ImgLevel *__restrict__ desc, int w, int h, int low, int high, int wp32, int rows_per_strip)
{
    canny_nms_body(desc, w, h, low, high, wp32, rows_per_strip);     // warp-synchronous inside (__shfl_*_sync, __any_sync)
}
__global__ void __launch_bounds__(128, 6) k_canny_nms_mb6(const ImgLevel *__restrict__ desc, int w, int h, int low, int high, int wp32, int rows_per_strip)
{
    canny_nms_body(desc, w, h, low, high, wp32, rows_per_strip);     // warp-synchronous inside (__shfl_*_sync, __any_sync)
}

// 16 mask bits -> 16 edge bytes (0 / 255) of row y from column x0 on, in edges and (where it is a plane of its own) edges_orig
__device__ __forceinline__ void expand_store16(const ImgLevel &L, int w, int y, int x0, unsigned bits)
{
    uint8_t *e = L.edges + (size_t)y * w + x0, *eo = L.edges_orig + (size_t)y * w + x0;
    const bool two = eo != e;
    unsigned o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const unsigned nib = (bits >> (4 * q)) & 15u;
        // 4 bits -> 4 bytes of 0 / 255
        o[q] = ((nib & 1u) * 0xffu) | (((nib >> 1) & 1u) * 0xff00u) | (((nib >> 2) & 1u) * 0xff0000u) | (((nib >> 3) & 1u) * 0xff000000u);
    }
    if (x0 + 16 <= w && ((((uintptr_t)e) & 15) == 0) && ((((uintptr_t)eo) & 15) == 0)) {
        const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
        *(uint4 *)e = v;
        if (two) *(uint4 *)eo = v;
    } else {
        for (int k = 0; k < 16 && x0 + k < w; ++k) {
            const uint8_t v = (bits >> k) & 1u ? 255 : 0;
            e[k] = v;
            if (two) eo[k] = v;
        }
    }
}

// ---- (2) hysteresis on the bit masks --------------------------------------------------------------------------
// Word type W: a lane owns one W of a row.  32-bit words serve rows of up to 1024 pixels with single-instruction
// arithmetic; 64-bit words rows of up to 2048 pixels.  The masks are the same bytes either way (little endian).
template <typename W> struct WordOps;
template <> struct WordOps<unsigned> {
    static constexpr int kBits = 32;
    static __device__ __forceinline__ unsigned rev(unsigned v) { return __brev(v); }
};
template <> struct WordOps<unsigned long long> {
    static constexpr int kBits = 64;
    static __device__ __forceinline__ unsigned long long rev(unsigned long long v) { return __brevll(v); }
};

// Flood towards higher bit positions over the whole row (lane = word, lane 0 = leftmost pixels): the row is one long
// integer, up = (((C + S) ^ C) & C) | S with the carries between the lanes' words resolved by carry look-ahead on two
// ballots: G = lanes whose word overflows, P = lanes whose word is all ones (would pass a carry on); the lanes that
// receive a carry are ((G << 1) + P) ^ P.  Constant time, whatever the length of a run.
template <typename W>
__device__ __forceinline__ W flood_up_row(W c, W s, int lane)
{
    const W sum = c + s;
    const unsigned G = __ballot_sync(0xffffffffu, sum < c);
    const unsigned P = __ballot_sync(0xffffffffu, sum == (W)~(W)0);
    const unsigned cin = (((G << 1) + P) ^ P);
    const W tot = sum + (W)((cin >> lane) & 1u);
    return (((tot ^ c) & c) | s);
}

// all candidate bits of the row connected to a seed bit, in both directions
template <typename W>
__device__ __forceinline__ W flood_row(W c, W s, int lane)
{
    s &= c;
    const W up = flood_up_row<W>(c, s, lane);
    // the other direction: the same on the mirrored row (bits reversed inside the words, lane order reversed)
    const W rc = WordOps<W>::rev(__shfl_sync(0xffffffffu, c, 31 - lane));
    const W rs = WordOps<W>::rev(__shfl_sync(0xffffffffu, s, 31 - lane));
    const W dn = WordOps<W>::rev(__shfl_sync(0xffffffffu, flood_up_row<W>(rc, rs, lane), 31 - lane));
    return up | dn;
}

// seeds a strong row hands to the row next to it: the bits themselves and their left / right neighbours (8-connectivity)
template <typename W>
__device__ __forceinline__ W spread_row(W p, int lane)
{
    constexpr int kTop = WordOps<W>::kBits - 1;
    const unsigned from_left = __shfl_up_sync(0xffffffffu, (unsigned)(p >> kTop), 1);
    const unsigned from_right = __shfl_down_sync(0xffffffffu, (unsigned)(p & (W)1), 1);
    W o = p | (p << 1) | (p >> 1);
    if (lane > 0) o |= (W)from_left;
    if (lane < 31) o |= (W)from_right << kTop;
    return o;
}

// Fallback for images whose masks do not fit shared memory: blind down / up sweeps per band on the masks in L2.
__global__ void __launch_bounds__(1024) k_canny_hyst(const ImgLevel *__restrict__ desc, int w, int h, int wp64)
{
    typedef unsigned long long W;
    const ImgLevel &L = desc[blockIdx.x];
    const W *__restrict__ C = (const W *)L.labels;
    W *S = (W *)L.labels + (size_t)wp64 * h;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int R = (h + n_warps - 1) / n_warps;
    const int y_lo = warp * R, y_hi = min(h, y_lo + R);
    const bool act = lane < wp64;
    auto ldS = [&](int y) -> W { return act ? __ldcg(S + (size_t)y * wp64 + lane) : (W)0; };
    auto ldC = [&](int y) -> W { return act ? __ldg(C + (size_t)y * wp64 + lane) : (W)0; };
    while (true) {
        int changed = 0;
        if (y_lo < y_hi) {
            W prev = y_lo > 0 ? ldS(y_lo - 1) : (W)0;
            W c = ldC(y_lo), s0 = ldS(y_lo);
            for (int y = y_lo; y < y_hi; ++y) {           // down
                W cn = 0, sn = 0;
                if (y + 1 < y_hi) { cn = ldC(y + 1); sn = ldS(y + 1); }
                const W s2 = flood_row<W>(c, s0 | spread_row<W>(prev, lane), lane);
                if (s2 != s0) { S[(size_t)y * wp64 + lane] = s2; changed = 1; }
                prev = s2; c = cn; s0 = sn;
            }
            W nxt = y_hi < h ? ldS(y_hi) : (W)0;
            for (int y = y_hi - 1; y >= y_lo; --y) {      // up (prev = last row of the band, just computed)
                const W cy = ldC(y), sy = (y == y_hi - 1) ? prev : ldS(y);
                const W s2 = flood_row<W>(cy, sy | spread_row<W>(nxt, lane), lane);
                if (s2 != sy) { S[(size_t)y * wp64 + lane] = s2; changed = 1; }
                nxt = s2;
            }
        }
        if (!__syncthreads_or(changed)) break;
    }
}

// Default variant: both masks of the image staged in shared memory (2 * ceil(w/64) * 8 * h bytes: 77 KB at VGA) and a
// DIRTY-ROW worklist instead of blind sweeps.  A row is dirty when one of its two neighbour rows has gained strong bits
// that hand it a seed it does not have yet (all rows are dirty at the start).  A warp walks the dirty rows of its band
// downwards, then upwards; visiting a row floods it from its own strong bits and the 8-connected bits of both neighbour
// rows, and a row that changed marks its neighbours.  Marks that cross a band boundary are MESSAGES to the neighbour warp
// (a flag per direction in shared memory); there is no block-wide barrier between rounds: a warp that runs out of dirty rows
// sleeps on its two flags.  Termination is detected with a token count: every active warp and every raised flag holds one
// token, a sender adds the token before it raises the flag (and gives it back if the flag was already up), a receiver that
// was asleep takes the flag's token over, one that was awake destroys it, a warp that falls asleep destroys its own; zero
// tokens = nobody awake and nothing in flight, for good.  (The round barrier of the previous version was 44 % of this
// kernel's warp time: profiles/r2_pyramid_kernels_ncu.txt.)  wp = row pitch in words of type W.
template <typename W>
__global__ void __launch_bounds__(1024) k_canny_hyst_smem(const ImgLevel *__restrict__ desc, int w, int h, int wp)
{
    extern __shared__ unsigned long long hs_mem[];
    __shared__ unsigned flag_from_above[32];             // "your first row got a new seed" (from the warp above)
    __shared__ unsigned flag_from_below[32];             // "your last row got a new seed" (from the warp below)
    __shared__ int tokens;
    const ImgLevel &L = desc[blockIdx.x];
    const size_t nw = (size_t)wp * h;
    W *gC = (W *)L.labels, *gS = gC + nw;
    W *C = (W *)hs_mem, *S = C + nw;
    for (size_t i = threadIdx.x; i < nw; i += blockDim.x) { C[i] = gC[i]; S[i] = gS[i]; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    if (threadIdx.x < 32) { flag_from_above[threadIdx.x] = 0; flag_from_below[threadIdx.x] = 0; }
    if (threadIdx.x == 0) tokens = n_warps;
    __syncthreads();
    const int R = (h + n_warps - 1) / n_warps;          // <= 32 (launcher)
    const int y_lo = warp * R, y_hi = min(h, y_lo + R);
    const int n_rows = y_hi > y_lo ? y_hi - y_lo : 0;
    const bool act = lane < wp;
    const int col = act ? lane : 0;
    unsigned dirty = n_rows >= 32 ? 0xffffffffu : ((1u << n_rows) - 1u);
    unsigned fresh = dirty;                               // rows not yet flooded from their own strong bits
    auto row = [&](const W *M, int y) -> W { return act ? ((const volatile W *)M)[(size_t)y * wp + col] : (W)0; };
    // raise a flag of a neighbour warp (lane 0): token first, handed back if the flag was already up
    auto send = [&](unsigned *flag) {
        __threadfence_block();                            // the strong bits just written are visible before the flag
        if (lane == 0) {
            atomicAdd(&tokens, 1);
            if (atomicExch(flag, 1u)) atomicSub(&tokens, 1);
        }
    };
    bool awake = true;                                    // holds a token
    while (true) {
        // mail (lane 0 looks, everybody learns)
        unsigned mail = 0;
        if (lane == 0) {
            if (n_rows > 0) {
                if (atomicExch(&flag_from_above[warp], 0u)) mail |= 1u;
                if (atomicExch(&flag_from_below[warp], 0u)) mail |= 2u;
            }
        }
        mail = __shfl_sync(0xffffffffu, mail, 0);
        if (mail) {
            __threadfence_block();
            const int n_tok = ((mail & 1u) ? 1 : 0) + ((mail & 2u) ? 1 : 0);
            // asleep: one of the flags' tokens becomes this warp's own; every other one is destroyed
            if (lane == 0 && n_tok - (awake ? 0 : 1) > 0) atomicSub(&tokens, n_tok - (awake ? 0 : 1));
            awake = true;
            if (mail & 1u) dirty |= 1u;
            if (mail & 2u) dirty |= 1u << (n_rows - 1);
        }
        if (!dirty) {
            if (awake) {
                if (lane == 0) atomicSub(&tokens, 1);
                awake = false;
            }
            int t = 0;
            if (lane == 0) t = *(volatile int *)&tokens;
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t == 0) break;
            __nanosleep(40);
            continue;
        }
        for (int pass = 0; pass < 2; ++pass) {
            int pos = pass == 0 ? 0 : 32;                 // down: next row >= pos ; up: next row < pos
            while (true) {
                const unsigned m = pass == 0 ? (pos < 32 ? dirty & ~((1u << pos) - 1u) : 0u)
                                             : (pos > 0 ? dirty & (pos >= 32 ? 0xffffffffu : ((1u << pos) - 1u)) : 0u);
                if (!m) break;
                const int r = pass == 0 ? __ffs(m) - 1 : 31 - __clz(m);
                pos = pass == 0 ? r + 1 : r;
                dirty &= ~(1u << r);
                const int y = y_lo + r;
                const W c = row(C, y), s0 = row(S, y);
                W nb = 0;
                if (y > 0) nb = row(S, y - 1);
                if (y + 1 < h) nb |= row(S, y + 1);
                const W seeds = spread_row<W>(nb, lane) & c & ~s0;
                const bool first = (fresh >> r) & 1u;
                fresh &= ~(1u << r);
                if (!first && !__any_sync(0xffffffffu, seeds != 0)) continue;
                const W s2 = flood_row<W>(c, s0 | seeds, lane);
                const W delta = s2 & ~s0;
                if (!__any_sync(0xffffffffu, delta != 0)) continue;
                if (delta) ((volatile W *)S)[(size_t)y * wp + col] = s2;
                __syncwarp();
                // a neighbour row must be (re)visited only if the new bits hand it a seed it does not have yet
                const W sp = spread_row<W>(delta, lane);
                if (y > 0 && __any_sync(0xffffffffu, (sp & row(C, y - 1) & ~row(S, y - 1)) != 0)) {
                    if (r > 0) dirty |= 1u << (r - 1);
                    else send(&flag_from_below[warp - 1]);
                }
                if (y + 1 < h && __any_sync(0xffffffffu, (sp & row(C, y + 1) & ~row(S, y + 1)) != 0)) {
                    if (r + 1 < n_rows) dirty |= 1u << (r + 1);
                    else send(&flag_from_above[warp + 1]);
                }
            }
        }
    }
    __syncthreads();
    // (Writing the edge bytes and the patch histogram from here -- one kernel less, no mask round trip -- was measured: the
    // CTAs of a launch finish together, so the stores of 512 threads per image do not hide under anybody's flood, and the build
    // got 0.12 ms per 256 frames SLOWER than with the streaming k_canny_expand.  profiles/r2_pyramid_kernels_ncu.txt)
    for (size_t i = threadIdx.x; i < nw; i += blockDim.x) gS[i] = S[i];
}

// ---- (3) bit mask -> byte maps + patch counters ------------------------------------------------------------------
// One thread -> 16 pixels of a row (half a mask word): one 128-bit store per byte map.  Threads are numbered over the
// (row, 16-pixel chunk) pairs of the image, so no thread is idle whatever the width.  On the levels that never run the edge
// fill-in (level 0, levels >= 3) edges_orig IS edges (one plane, one store).
__global__ void __launch_bounds__(256) k_canny_expand(const ImgLevel *__restrict__ desc, int w, int h, int wp32, int P, int chunks_x)
{
    const int f = blockIdx.z;
    const ImgLevel &L = desc[f];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = t / chunks_x;
    const int x0 = (t - y * chunks_x) * 16;
    if (y >= h) return;
    const unsigned *__restrict__ maskS = (const unsigned *)L.labels + (size_t)wp32 * h;
    const unsigned bits = (__ldcg(maskS + (size_t)y * wp32 + (x0 >> 5)) >> (x0 & 31)) & 0xffffu;
    expand_store16(L, w, y, x0, bits);
    if (bits && L.hist_w > 0 && L.hist_h > 0) {
        int *cnt = (int *)L.flags;
        const int py = y / P;
        if (py < L.hist_h)
            for (unsigned mm = bits; mm;) {
                const int b = __ffs(mm) - 1;
                mm &= mm - 1;
                const int px = (x0 + b) / P;
                if (px < L.hist_w) atomicAdd(cnt + py * L.hist_w + px, 1);
            }
    }
}

static int launch_canny_bits(revo_ctx *ctx, const ImgLevel *d_desc, int n, int w, int h, int low, int high, int patch,
                             void *d_counts0, size_t counts_stride)
{
    const int hist_w = w / patch, hist_h = h / patch;
    const int wp64 = cdiv(w, 64), wp32 = 2 * wp64;
    // a warp owns a band of up to 30 rows (16 warps at VGA: two CTAs per SM, so 256 images run in ONE wave on 148 SMs)
    static const int band = getenv("REVO_HYST_BAND") ? atoi(getenv("REVO_HYST_BAND")) : 30;
    int warps = cdiv(h, band);
    warps = warps < 1 ? 1 : (warps > 32 ? 32 : warps);
    const size_t smem = (size_t)2 * wp64 * 8 * h;
    const bool in_smem = smem <= 200 * 1024 && cdiv(h, warps) <= 32;
    if (hist_w > 0 && hist_h > 0)
        REVO_CUDA(ctx, cudaMemset2DAsync(d_counts0, counts_stride, 0, (size_t)hist_w * hist_h * sizeof(int), (size_t)n, ctx->stream));
    {
        // rows per warp strip: long strips amortise the 2-row prologue, short ones keep the small levels parallel
        const int rs = h >= 400 ? NMS_RS : (h >= 200 ? 18 : 9);
        dim3 grid(cdiv(w, 32 * NMS_PX), cdiv(cdiv(h, rs), 4), n);
        static const int mb = getenv("REVO_NMS_MINBLOCKS") ? atoi(getenv("REVO_NMS_MINBLOCKS")) : 5;
        if (mb == 5)
            k_canny_nms_mb5<<<grid, 128, 0, ctx->stream>>>(d_desc, w, h, low, high, wp32, rs);
        else if (mb == 6)
            k_canny_nms_mb6<<<grid, 128, 0, ctx->stream>>>(d_desc, w, h, low, high, wp32, rs);
        else
            k_canny_nms<<<grid, 128, 0, ctx->stream>>>(d_desc, w, h, low, high, wp32, rs);
        LAUNCH_CHECK(ctx);
    }
    {
        if (in_smem) {
            if (wp32 <= 32) {
                REVO_CUDA(ctx, cudaFuncSetAttribute(k_canny_hyst_smem<unsigned>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                k_canny_hyst_smem<unsigned><<<n, warps * 32, smem, ctx->stream>>>(d_desc, w, h, wp32);
            } else {
                REVO_CUDA(ctx, cudaFuncSetAttribute(k_canny_hyst_smem<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    200 * 1024));
                k_canny_hyst_smem<unsigned long long><<<n, warps * 32, smem, ctx->stream>>>(d_desc, w, h, wp64);
            }
        } else {
            k_canny_hyst<<<n, warps * 32, 0, ctx->stream>>>(d_desc, w, h, wp64);
        }
        LAUNCH_CHECK(ctx);
    }
    {
        const int chunks_x = cdiv(w, 16);
        dim3 grid(cdiv(chunks_x * h, 256), 1, n);
        k_canny_expand<<<grid, 256, 0, ctx->stream>>>(d_desc, w, h, wp32, patch, chunks_x);
        LAUNCH_CHECK(ctx);
    }
    if (hist_w > 0 && hist_h > 0) {
        k_hist_finalize<<<n, 256, 0, ctx->stream>>>(d_desc);
        LAUNCH_CHECK(ctx);
    }
    return REVO_OK;
}

int launch_canny(revo_ctx *ctx, const ImgLevel *d_desc, int n, int w, int h, int low, int high, const void *gray_tmap, int patch,
                 void *d_counts0, size_t counts_stride)
{
    // bit-mask pipeline: rows of up to 2048 pixels (a lane owns one 64-bit word of a row in the hysteresis); the two
    // masks must fit the frame's label plane (2 * ceil(w/64) * 8 * h bytes <= 4 * w0 * h0: always)
    static const int force_tile = getenv("REVO_CANNY_TILE") ? atoi(getenv("REVO_CANNY_TILE")) : 0;
    if (!force_tile && w <= 2048 && w >= 8 && (w & 3) == 0) return launch_canny_bits(ctx, d_desc, n, w, h, low, high, patch, d_counts0, counts_stride);
    const int hist_w = w / patch, hist_h = h / patch;
    if (hist_w > 0 && hist_h > 0)
        REVO_CUDA(ctx, cudaMemset2DAsync(d_counts0, counts_stride, 0, (size_t)hist_w * hist_h * sizeof(int), (size_t)n, ctx->stream));
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    const int use_tma = gray_tmap != nullptr;
    if (use_tma) memcpy(&tm, gray_tmap, sizeof(tm));
    dim3 grid(cdiv(w, CT_W), cdiv(h, CT_H), n);
    k_canny_tile<<<grid, 256, 0, ctx->stream>>>(tm, use_tma, d_desc, w, h, low, high);
    LAUNCH_CHECK(ctx);
    k_canny_merge<<<grid, 192, 0, ctx->stream>>>(d_desc, w, h);
    LAUNCH_CHECK(ctx);
    dim3 g3(cdiv(cdiv(w * h, 16), 256), 1, n);
    k_canny_final<<<g3, 256, 0, ctx->stream>>>(d_desc, w, h, patch);
    LAUNCH_CHECK(ctx);
    if (hist_w > 0 && hist_h > 0) {
        k_hist_finalize<<<n, 256, 0, ctx->stream>>>(d_desc);
        LAUNCH_CHECK(ctx);
    }
    return REVO_OK;
}

}  // namespace revo
