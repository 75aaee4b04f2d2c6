"""Test infrastructure: runs the tracking KERNELS of this repo on the CPU.

The kernel source text (track.cu: k_track; optionally scratch/experiments/track_lean.cu: k_track_lean) and the device helpers
(track_common.cuh, the quad-record packing of pyramid.cu) are extracted from the CUDA files at test time and compiled with g++
against a small emulation layer: one OS thread per CUDA thread of a CTA, pthread barriers for __syncthreads / the warp
shuffles, `static` storage for __shared__, a one-CTA "cluster", host versions of the few PTX helpers (256-bit gather,
rcp.approx, shared-memory loads of the experiment).  CTAs run one after the other (the persistent kernels pull pairs from a
work counter, so that is a valid schedule for clusters of one CTA).  What this does NOT cover: the distributed-shared-memory
exchange of clusters with more than one CTA and the multi-GPU mailboxes -- those paths need the hardware.
"""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PRELUDE = r'''
#include <pthread.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <type_traits>
#include <vector>
#include "revo_b200.h"

struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace emu {
struct D3 { unsigned x, y, z; };
static thread_local D3 tidx;
static D3 bidx, bdim, gdim;
static float *dyn_smem;
static pthread_barrier_t cta_bar, warp_bar[32];
static float shfl_slot[32][32];
}
#define threadIdx emu::tidx
#define blockIdx emu::bidx
#define blockDim emu::bdim
#define gridDim emu::gdim
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
static inline void __syncthreads() { pthread_barrier_wait(&emu::cta_bar); }
static inline void __syncwarp() { pthread_barrier_wait(&emu::warp_bar[emu::tidx.x >> 5]); }
static inline float __shfl_xor_sync(unsigned, float v, int m)
{
    const int lane = emu::tidx.x & 31, w = emu::tidx.x >> 5;
    emu::shfl_slot[w][lane] = v;
    pthread_barrier_wait(&emu::warp_bar[w]);
    const float r = emu::shfl_slot[w][lane ^ m];
    pthread_barrier_wait(&emu::warp_bar[w]);
    return r;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicAdd(int *p, int v) { return __sync_fetch_and_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __sync_fetch_and_add(p, v); }
static inline long long clock64() { return 0; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline int __float2int_rn(float a) { return (int)lrintf(a); }
static inline double __drcp_rn(double x) { return 1.0 / x; }
static inline float __fdividef(float a, float b) { return a / b; }
namespace cooperative_groups {
struct cluster_group {
    unsigned num_blocks() const { return 1; }
    unsigned block_rank() const { return 0; }
    void sync() const { __syncthreads(); }
    template <class T> T *map_shared_rank(T *p, int) const { return p; }
};
static inline cluster_group this_cluster() { return cluster_group(); }
}
namespace cg = cooperative_groups;

namespace revo {
// host versions of the PTX helpers of track_common.cuh / track.cu
static inline void ldg_quad(const uint4 *p, uint4 &r0, uint4 &r1)
{
    const uint32_t *q = (const uint32_t *)p;
    r0 = make_uint4(q[0], q[1], q[4], q[5]);
    r1 = make_uint4(q[2], q[3], q[6], q[7]);
}
template <int kHint> static inline void ldg_quad_h(const uint4 *p, uint4 &r0, uint4 &r1) { ldg_quad(p, r0, r1); }
static inline float rcp_approx(float x) { return 1.0f / x; }
static inline uint32_t smem_u32(const void *p) { return (uint32_t)((const char *)p - (const char *)emu::dyn_smem); }   // only meaningful for the dynamic buffer
static inline void mbar_init(uint64_t *, uint32_t) {}
static inline void mbar_expect_tx(uint64_t *, uint32_t) {}
static inline void mbar_wait(uint64_t *, uint32_t) { std::abort(); }                       // clusters of one CTA never exchange
static inline void st_async_b64(void *, unsigned, unsigned long long, uint64_t *) { std::abort(); }
static inline void st_release_sys(unsigned long long *, unsigned long long) { std::abort(); }   // multi-GPU split: not emulated
static inline unsigned long long ld_acquire_sys(const unsigned long long *) { std::abort(); }
static inline void __threadfence_system() {}
template <int kThreads> static inline void lds3(uint32_t addr, float &x, float &y, float &z)
{
    const char *b = (const char *)emu::dyn_smem + addr;
    std::memcpy(&x, b, 4); std::memcpy(&y, b + kThreads * 4, 4); std::memcpy(&z, b + kThreads * 8, 4);
}
template <int kThreads> static inline void sts3(uint32_t addr, float x, float y, float z)
{
    char *b = (char *)emu::dyn_smem + addr;
    std::memcpy(b, &x, 4); std::memcpy(b + kThreads * 4, &y, 4); std::memcpy(b + kThreads * 8, &z, 4);
}
static inline float2 ffma2(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
static inline float2 fmul2(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float pin(float x) { return x; }
}
'''

RUNNER = r'''
namespace emu {
template <class F> static void run_grid(int n_ctas, int threads, size_t dyn_bytes, F kernel)
{
    std::vector<float> smem(dyn_bytes / 4 + 64);
    dyn_smem = smem.data();
    bdim = D3{(unsigned)threads, 1, 1};
    gdim = D3{(unsigned)n_ctas, 1, 1};
    for (int c = 0; c < n_ctas; ++c) {
        bidx = D3{(unsigned)c, 0, 0};
        pthread_barrier_init(&cta_bar, nullptr, threads);
        for (int w = 0; w < threads / 32; ++w) pthread_barrier_init(&warp_bar[w], nullptr, 32);
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t)
            th.emplace_back([=]() { tidx = D3{(unsigned)t, 0, 0}; kernel(); });
        for (auto &x : th) x.join();
        pthread_barrier_destroy(&cta_bar);
        for (int w = 0; w < threads / 32; ++w) pthread_barrier_destroy(&warp_bar[w]);
    }
}
}

// One frame pair through a tracking kernel.  variant: 0 = k_track<128,4> (library), 1 = k_track_lean<128,4,0,false>,
// 2 = k_track_lean<128,4,0,true> (packed accumulation); the lean variants exist only when the experiment source was given.
extern "C" int emu_track_pairs(int variant, int n_pairs, int n_ctas, int n_levels, const float *const *pts, const int *n_pts,
                               const float *const *dt, const int *w, const int *h, const float *cam4, const float *R9s, const float *t3s,
                               const revo_tracker_config *cfg, int mode, int level, int pcap, revo_track_result *results, double *records)
{
    using namespace revo;
    std::vector<std::vector<uint4>> opt((size_t)n_pairs * n_levels);
    std::vector<PairDesc> pairs(n_pairs);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < n_pairs; ++p) {
        std::memset(&pairs[p], 0, sizeof(PairDesc));
        for (int l = 0; l < n_levels; ++l) {
            const int k = p * n_levels + l;
            const size_t npx = (size_t)w[k] * h[k];
            opt[k].resize(2 * npx);
            for (size_t i = 0; i < npx; ++i) {              // k_opt_struct, pyramid.cu
                const float4 a = opt_texel(dt[k], i, w[k], h[k]);
                const float4 b = (i + 1 < npx) ? opt_texel(dt[k], i + 1, w[k], h[k]) : z;
                const float4 c = (i + w[k] < npx) ? opt_texel(dt[k], i + w[k], w[k], h[k]) : z;
                const float4 d = (i + w[k] + 1 < npx) ? opt_texel(dt[k], i + w[k] + 1, w[k], h[k]) : z;
                store_quad(opt[k].data(), i, a, b, c, d);
            }
            LevelIn &L = pairs[p].lvl[l];
            L.pts = (const float4 *)pts[k]; L.n_pts = &n_pts[k]; L.opt = opt[k].data();
            L.fx = cam4[4 * k]; L.fy = cam4[4 * k + 1]; L.cx = cam4[4 * k + 2]; L.cy = cam4[4 * k + 3]; L.w = w[k]; L.h = h[k];
        }
        pairs[p].ref_dt_min = dt[p * n_levels + (mode == 0 ? cfg->pyr_min_lvl : level)];
        std::memcpy(pairs[p].R, R9s + 9 * p, sizeof(float) * 9);
        std::memcpy(pairs[p].t, t3s + 3 * p, sizeof(float) * 3);
    }
    TrackParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.cfg = *cfg; prm.mode = mode; prm.level = level;
    alignas(16) int work_counter[64] = {0};
    constexpr int T = 128;
    const size_t dyn = (size_t)pcap * T * 12;
    const PairDesc *d_pairs = pairs.data();
    int *wc = work_counter;
    if (variant == 0) {
        emu::run_grid(n_ctas, T, dyn, [=]() { k_track<T, 4>(d_pairs, n_pairs, prm, results, records, nullptr, nullptr, wc, pcap); });
    }
#ifdef EMU_WITH_LEAN
    else if (variant == 1) {
        emu::run_grid(n_ctas, T, dyn, [=]() { k_track_lean<T, 4, 0, false>(d_pairs, n_pairs, prm, results, records, nullptr, nullptr, wc, pcap); });
    } else if (variant == 2) {
        emu::run_grid(n_ctas, T, dyn, [=]() { k_track_lean<T, 4, 0, true>(d_pairs, n_pairs, prm, results, records, nullptr, nullptr, wc, pcap); });
    }
#endif
    else return 1;
    return 0;
}
'''


def _strip_functions(text, names):
    """Remove top-level function definitions (column-0 `__device__ ... name(` up to the closing brace at column 0)."""
    for n in names:
        m = re.search(r"^(template <[^>]*>\n)?__device__[^\n]*\b" + re.escape(n) + r"\(", text, re.M)
        assert m, n
        line_end = text.index("\n", m.end())
        if text[m.start():line_end].rstrip().endswith("}"):       # one-liner
            text = text[:m.start()] + text[line_end + 1:]
            continue
        j = text.index("\n}\n", m.start())
        text = text[:m.start()] + text[j + 3:]
    return text


def _struct(text, name):
    m = re.search(r"^struct " + name + r" \{", text, re.M)
    assert m, name
    j = text.index("\n};", m.start())
    return text[m.start():j + 4]


def _kernel(text, name):
    m = re.search(r"^template <[^>]*>\n__global__ void __launch_bounds__\([^)]*\)\n" + name + r"\(", text, re.M)
    assert m, name
    j = text.index("\n}\n", m.start())
    k = text[m.start():j + 3]
    k = k.replace("extern __shared__ float s_pts[];", "float *s_pts = emu::dyn_smem;")
    k = re.sub(r'\n[^\n]*asm volatile\("fence\.mbarrier_init[^\n]*\n', "\n", k)
    assert "asm" not in k, "unexpected inline PTX left in " + name
    return k


def build(out_dir, with_lean=False):
    """Compile the emulated kernels into a shared library and return the ctypes handle."""
    rd = lambda *p: open(os.path.join(ROOT, *p)).read()      # noqa: E731
    common, pyr, internal, track = rd("revo_b200", "csrc", "track_common.cuh"), rd("revo_b200", "csrc", "pyramid.cu"), \
        rd("revo_b200", "csrc", "internal.h"), rd("revo_b200", "csrc", "track.cu")
    body = common[common.index("namespace revo {") + len("namespace revo {"):common.index("}  // namespace revo")]
    body = _strip_functions(body, ["ldg_quad", "rcp_approx", "smem_u32", "mbar_init", "mbar_expect_tx", "mbar_wait", "st_async_b64"])
    assert "asm" not in body
    grab = lambda t, pat: t[re.search(pat, t, re.M).start():t.index("\n}\n", re.search(pat, t, re.M).start()) + 3]      # noqa: E731
    parts = ["namespace revo {", _struct(internal, "LevelIn"), _struct(internal, "PairDesc"), _struct(internal, "TrackParams"), body,
             grab(pyr, r"^__device__ __forceinline__ float4 opt_texel"), grab(pyr, r"^__device__ __forceinline__ uint32_t pack_grad"),
             grab(pyr, r"^__device__ __forceinline__ void store_quad"), _struct(track, "Mailbox"), _kernel(track, "k_track")]
    flags = []
    if with_lean:
        lean = rd("scratch", "experiments", "track_lean.cu")
        parts += [grab(lean, r"^__device__ __forceinline__ ProjB project_l"), grab(lean, r"^__device__ __forceinline__ void finish_point_l"),
                  _struct(lean, "PackedAcc"), grab(lean, r"^__device__ __forceinline__ void finish_point_p"), _kernel(lean, "k_track_lean")]
        flags = ["-DEMU_WITH_LEAN"]
    parts.append("}  // namespace revo")
    src, lib = os.path.join(out_dir, "cuda_emu.cpp"), os.path.join(out_dir, "libcuda_emu.so")
    open(src, "w").write(PRELUDE + "\n".join(parts) + RUNNER)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-mfma", "-ffp-contract=fast", "-shared", "-fPIC", "-pthread", *flags,
                    "-I", os.path.join(ROOT, "include"), src, "-o", lib], check=True)
    return C.CDLL(lib)
