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

int launch_canny(revo_ctx *ctx, const ImgLevel *d_desc, int n, int w, int h, int low, int high, const void *gray_tmap, int patch,
                 void *d_counts0, size_t counts_stride)
{
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
