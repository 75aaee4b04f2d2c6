"""Test infrastructure: runs the KERNELS of this repo on the CPU.

build(): the tracking kernels (track.cu: k_track; optionally scratch/experiments/track_lean.cu); build_canny(): the bit-mask Canny
pipeline of canny.cu with its launch code; build_pyramid(): every kernel of pyramid.cu + the Canny pipeline with the launch
sequence of capi.cu:create_batch_impl / launch_keyframe, and the tracking-quality vote.  The kernel source text and the device
helpers are extracted from the CUDA files at test time (`kernel<<<grid, block, smem>>>(args)` becomes a call of the emulated
launcher; kernels without barriers / warp collectives run their threads sequentially) and compiled with g++
against a small emulation layer: one OS thread per CUDA thread, pthread barriers for __syncthreads / the warp shuffles /
cluster.sync(), a per-CTA arena for the __shared__ variables (same offsets in every CTA, so that distributed shared memory is
an offset into the peer's arena), mbarriers with transaction counts as 64-bit atomics, st.async as store + complete_tx, host
versions of the other PTX helpers (256-bit gather, rcp.approx, shared-memory loads of the experiment).  The CTAs of a cluster
run concurrently, clusters one after the other (the persistent kernels pull pairs from a work counter, so that is a valid
schedule).  What this does NOT cover: the multi-GPU mailboxes, and of course timing.
"""
import ctypes as C
import os
import re
import shlex
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXTRA_FLAGS = shlex.split(os.environ.get("REVO_EMU_CXXFLAGS", ""))      # e.g. "-fsanitize=thread -g" (scratch/tools/emu_sanitize.py)

PRELUDE = r'''
#include <pthread.h>
#include <sched.h>
#include <algorithm>
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
struct uint2 { uint32_t x, y; };
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

namespace emu {
struct D3 { unsigned x, y, z; };
constexpr int kMaxWarps = 32, kArenaBytes = 64 * 1024, kMaxSlots = 128;
struct Cluster;
struct Cta {                        // one thread block: its barriers, shuffle slots and "shared memory"
    int rank;
    Cluster *cluster;
    D3 bidx;
    pthread_barrier_t bar, warp_bar[kMaxWarps];
    float shfl_slot[kMaxWarps][32];
    alignas(8) unsigned long long coll_slot[kMaxWarps][32];
    int or_flag;
    alignas(64) char arena[kArenaBytes];      // the __shared__ variables of the kernel, same offsets in every CTA
    std::vector<float> dyn;                   // dynamic shared memory
};
struct Cluster {
    int n_ctas;
    std::vector<Cta *> cta;
    pthread_barrier_t bar;                    // cluster.sync()
    pthread_mutex_t mu;
    size_t slot_off[kMaxSlots];
    bool slot_set[kMaxSlots];
    size_t used;
};
static thread_local D3 tidx;
static thread_local Cta *cta;
static D3 bdim, gdim;
// storage of the k-th __shared__ declaration of the kernel (first caller of the cluster fixes the offset)
static inline void *smem_slot(int k, size_t bytes, size_t align)
{
    if (k < 0 || k >= kMaxSlots) std::abort();
    Cluster *cl = cta->cluster;
    pthread_mutex_lock(&cl->mu);
    if (!cl->slot_set[k]) {
        cl->used = (cl->used + align - 1) / align * align;
        cl->slot_off[k] = cl->used;
        cl->used += bytes;
        cl->slot_set[k] = true;
        if (cl->used > (size_t)kArenaBytes) std::abort();
    }
    const size_t off = cl->slot_off[k];
    pthread_mutex_unlock(&cl->mu);
    return cta->arena + off;
}
// the same shared-memory address in CTA `rank` of the cluster (distributed shared memory)
template <class T> static inline T *map_rank(T *p, unsigned rank)
{
    const char *base = cta->arena;
    const ptrdiff_t off = (const char *)p - base;
    if (off < 0 || off >= (ptrdiff_t)kArenaBytes) std::abort();
    return (T *)(cta->cluster->cta[rank]->arena + off);
}
}
#define threadIdx emu::tidx
#define blockIdx emu::cta->bidx
#define blockDim emu::bdim
#define gridDim emu::gdim
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
static inline void __syncthreads() { pthread_barrier_wait(&emu::cta->bar); }
static inline void __syncwarp() { pthread_barrier_wait(&emu::cta->warp_bar[emu::tidx.x >> 5]); }
static inline float __shfl_xor_sync(unsigned, float v, int m)
{
    const int lane = emu::tidx.x & 31, w = emu::tidx.x >> 5;
    emu::cta->shfl_slot[w][lane] = v;
    pthread_barrier_wait(&emu::cta->warp_bar[w]);
    const float r = emu::cta->shfl_slot[w][lane ^ m];
    pthread_barrier_wait(&emu::cta->warp_bar[w]);
    return r;
}
namespace emu {
static inline int lin_tid() { return (int)(tidx.x + bdim.x * (tidx.y + bdim.y * tidx.z)); }
// all lanes of the warp publish a value, then read what they need: the building block of every warp collective
template <class T, class F> static inline auto warp_collective(T v, F pick) -> decltype(pick((const T *)nullptr))
{
    static_assert(sizeof(T) <= 8, "slot size");
    const int w = lin_tid() >> 5, lane = lin_tid() & 31;
    T *slots = (T *)cta->coll_slot[w];
    slots[lane] = v;
    pthread_barrier_wait(&cta->warp_bar[w]);
    T copy[32];
    for (int i = 0; i < 32; ++i) copy[i] = slots[i];
    auto r = pick((const T *)copy);
    pthread_barrier_wait(&cta->warp_bar[w]);
    return r;
}
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emu::warp_collective(v, [=](const T *s) { return s[src & 31]; }); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
    const int lane = emu::lin_tid() & 31;
    return emu::warp_collective(v, [=](const T *s) { return lane >= (int)d ? s[lane - d] : s[lane]; });
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d)
{
    const int lane = emu::lin_tid() & 31;
    return emu::warp_collective(v, [=](const T *s) { return lane + (int)d < 32 ? s[lane + d] : s[lane]; });
}
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int m)
{
    const int lane = emu::lin_tid() & 31;
    return emu::warp_collective(v, [=](const unsigned *s) { return s[lane ^ m]; });
}
static inline unsigned __ballot_sync(unsigned, bool p)
{
    return emu::warp_collective((unsigned)p, [](const unsigned *s) { unsigned b = 0; for (int i = 0; i < 32; ++i) b |= (s[i] ? 1u : 0u) << i; return b; });
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline int __reduce_add_sync(unsigned, int v)
{
    return emu::warp_collective(v, [](const int *s) { int t = 0; for (int i = 0; i < 32; ++i) t += s[i]; return t; });
}
static inline int __syncthreads_or(int p)
{
    if (p) __atomic_store_n(&emu::cta->or_flag, 1, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&emu::cta->bar);
    const int r = __atomic_load_n(&emu::cta->or_flag, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&emu::cta->bar);
    if (emu::lin_tid() == 0) __atomic_store_n(&emu::cta->or_flag, 0, __ATOMIC_SEQ_CST);
    pthread_barrier_wait(&emu::cta->bar);
    return r;
}
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline unsigned long long __brevll(unsigned long long v) { unsigned long long r = 0; for (int i = 0; i < 64; ++i) r |= ((v >> i) & 1ull) << (63 - i); return r; }
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __ffsll(long long v) { return v ? __builtin_ctzll((unsigned long long)v) + 1 : 0; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c)
{
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
    return c;
}
using std::isfinite;
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    const unsigned long long ab = ((unsigned long long)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((ab >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
static inline unsigned atomicExch(unsigned *p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T __ldcg(const T *p) { return *p; }
using std::max;
using std::min;
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicAdd(int *p, int v) { return __sync_fetch_and_add(p, v); }
static inline int atomicSub(int *p, int v) { return __sync_fetch_and_sub(p, v); }
static inline int atomicMin(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
static inline int atomicMax(int *p, int v) { int o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
static inline void __nanosleep(unsigned) { sched_yield(); }
static inline void __threadfence_block() { __sync_synchronize(); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __sync_fetch_and_add(p, v); }
static inline long long clock64() { return 0; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __float2int_rn(float a) { return (int)lrintf(a); }
static inline double __drcp_rn(double x) { return 1.0 / x; }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline int __double2loint(double v) { uint64_t u; std::memcpy(&u, &v, 8); return (int)(uint32_t)u; }
static inline int __double2hiint(double v) { uint64_t u; std::memcpy(&u, &v, 8); return (int)(uint32_t)(u >> 32); }
static inline double __hiloint2double(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double v; std::memcpy(&v, &u, 8); return v; }
static inline float __fdividef(float a, float b) { return a / b; }
namespace cooperative_groups {
struct cluster_group {
    unsigned num_blocks() const { return (unsigned)emu::cta->cluster->n_ctas; }
    unsigned block_rank() const { return (unsigned)emu::cta->rank; }
    void sync() const { pthread_barrier_wait(&emu::cta->cluster->bar); }
    template <class T> T *map_shared_rank(T *p, int r) const { return emu::map_rank(p, (unsigned)r); }
};
static inline cluster_group this_cluster() { return cluster_group(); }
}
namespace cg = cooperative_groups;

namespace revo {
// host versions of the PTX helpers of track_common.cuh / track.cu
template <int kHint> static inline float4 ldg_point(const float4 *p) { return *p; }
template <int kHint> static inline uint2 ldg_texel(const uint2 *p, unsigned long long) { return *p; }
static inline unsigned long long l2_policy_evict_last() { return 0; }
static inline int opt_tiles_per_row_dev(int w) { return (w + 3) >> 2; }
static inline float rcp_approx(float x) { return 1.0f / x; }
static inline uint32_t smem_u32(const void *p) { return (uint32_t)((const char *)p - (const char *)emu::cta->dyn.data()); }   // only meaningful for the dynamic buffer
// mbarrier with transaction count in one 64-bit word: [31:0] pending transaction bytes (signed: completions may come before the
// expectation), [39:32] pending arrivals, [47:40] arrival count of a phase, [48] phase parity
static inline uint64_t mb_pack(int32_t tx, unsigned pend, unsigned cnt, unsigned ph) { return (uint32_t)tx | ((uint64_t)pend << 32) | ((uint64_t)cnt << 40) | ((uint64_t)ph << 48); }
static inline void mb_update(uint64_t *bar, int32_t dtx, int darrive)
{
    uint64_t o = __atomic_load_n(bar, __ATOMIC_SEQ_CST), n;
    do {
        int32_t tx = (int32_t)(uint32_t)o + dtx;
        unsigned pend = (unsigned)((o >> 32) & 0xff) - (unsigned)darrive, cnt = (unsigned)((o >> 40) & 0xff), ph = (unsigned)((o >> 48) & 1);
        if (pend == 0 && tx == 0) { ph ^= 1; pend = cnt; }
        n = mb_pack(tx, pend, cnt, ph);
    } while (!__atomic_compare_exchange_n(bar, &o, n, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
}
static inline void mbar_init(uint64_t *bar, uint32_t count) { __atomic_store_n(bar, mb_pack(0, count, count, 0), __ATOMIC_SEQ_CST); }
static inline void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { mb_update(bar, (int32_t)bytes, 1); }       // arrive.expect_tx
static inline void mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (((__atomic_load_n(bar, __ATOMIC_SEQ_CST) >> 48) & 1) == parity) sched_yield();
}
// st.async...mbarrier::complete_tx::bytes.b64: 8 bytes into CTA dst_rank, then 8 bytes of its barrier's transaction count
static inline void st_async_b64(void *local_ptr, unsigned dst_rank, unsigned long long v, uint64_t *local_bar)
{
    __atomic_store_n((unsigned long long *)emu::map_rank((char *)local_ptr, dst_rank), v, __ATOMIC_SEQ_CST);
    mb_update(emu::map_rank(local_bar, dst_rank), -8, 0);
}
static inline void st_release_sys(unsigned long long *, unsigned long long) { std::abort(); }   // multi-GPU split: not emulated
static inline unsigned long long ld_acquire_sys(const unsigned long long *) { std::abort(); }
static inline void __threadfence_system() {}
template <int kThreads> static inline void lds3(uint32_t addr, float &x, float &y, float &z)
{
    const char *b = (const char *)emu::cta->dyn.data() + addr;
    std::memcpy(&x, b, 4); std::memcpy(&y, b + kThreads * 4, 4); std::memcpy(&z, b + kThreads * 8, 4);
}
template <int kThreads> static inline void sts3(uint32_t addr, float x, float y, float z)
{
    char *b = (char *)emu::cta->dyn.data() + addr;
    std::memcpy(b, &x, 4); std::memcpy(b + kThreads * 4, &y, 4); std::memcpy(b + kThreads * 8, &z, 4);
}
static inline float2 ffma2(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
static inline float2 fmul2(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float pin(float x) { return x; }
}
'''

RUNNER = r'''
// @GENERIC_BEGIN
namespace emu {
// Clusters run one after the other; the CTAs of a cluster run concurrently (one OS thread per CUDA thread).
template <class F> static void run_grid(int n_clusters, int ctas_per_cluster, int threads, size_t dyn_bytes, F kernel)
{
    bdim = D3{(unsigned)threads, 1, 1};
    gdim = D3{(unsigned)(n_clusters * ctas_per_cluster), 1, 1};
    for (int c = 0; c < n_clusters; ++c) {
        Cluster cl;
        cl.n_ctas = ctas_per_cluster; cl.used = 0;
        std::memset(cl.slot_set, 0, sizeof(cl.slot_set));
        pthread_mutex_init(&cl.mu, nullptr);
        pthread_barrier_init(&cl.bar, nullptr, threads * ctas_per_cluster);
        std::vector<Cta *> ctas;
        for (int r = 0; r < ctas_per_cluster; ++r) {
            Cta *b = new Cta();
            b->rank = r; b->cluster = &cl; b->bidx = D3{(unsigned)(c * ctas_per_cluster + r), 0, 0};
            std::memset(b->arena, 0, sizeof(b->arena));
            b->dyn.assign(dyn_bytes / 4 + 64, 0.f);
            pthread_barrier_init(&b->bar, nullptr, threads);
            for (int w = 0; w < threads / 32; ++w) pthread_barrier_init(&b->warp_bar[w], nullptr, 32);
            ctas.push_back(b);
        }
        cl.cta = ctas;
        std::vector<std::thread> th;
        for (int r = 0; r < ctas_per_cluster; ++r)
            for (int t = 0; t < threads; ++t)
                th.emplace_back([=]() { cta = ctas[r]; tidx = D3{(unsigned)t, 0, 0}; kernel(); });
        for (auto &x : th) x.join();
        for (Cta *b : ctas) {
            pthread_barrier_destroy(&b->bar);
            for (int w = 0; w < threads / 32; ++w) pthread_barrier_destroy(&b->warp_bar[w]);
            delete b;
        }
        pthread_barrier_destroy(&cl.bar);
        pthread_mutex_destroy(&cl.mu);
    }
}
}

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
namespace emu {
// kernel<<<grid, block, smem>>>: thread blocks one after the other, one OS thread per CUDA thread (no clusters)
template <class F> static void launch(dim3 grid, dim3 block, size_t dyn_bytes, F kernel)
{
    const int threads = (int)(block.x * block.y * block.z), n_warps = (threads + 31) / 32;
    bdim = D3{block.x, block.y, block.z};
    gdim = D3{grid.x, grid.y, grid.z};
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                Cluster cl;
                cl.n_ctas = 1; cl.used = 0;
                std::memset(cl.slot_set, 0, sizeof(cl.slot_set));
                pthread_mutex_init(&cl.mu, nullptr);
                pthread_barrier_init(&cl.bar, nullptr, threads);
                Cta *b = new Cta();
                b->rank = 0; b->cluster = &cl; b->bidx = D3{bx, by, bz};
                std::memset(b->arena, 0, sizeof(b->arena));
                b->dyn.assign(dyn_bytes / 4 + 64, 0.f);
                pthread_barrier_init(&b->bar, nullptr, threads);
                for (int w = 0; w < n_warps; ++w) pthread_barrier_init(&b->warp_bar[w], nullptr, std::min(32, threads - 32 * w));
                cl.cta.assign(1, b);
                std::vector<std::thread> th;
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx)
                            th.emplace_back([=]() { cta = b; tidx = D3{tx, ty, tz}; kernel(); });
                for (auto &x : th) x.join();
                pthread_barrier_destroy(&b->bar);
                for (int w = 0; w < n_warps; ++w) pthread_barrier_destroy(&b->warp_bar[w]);
                delete b;
                pthread_barrier_destroy(&cl.bar);
                pthread_mutex_destroy(&cl.mu);
            }
}
}

namespace emu {
// kernels without barriers or warp collectives: the threads of a block simply run one after the other
template <class F> static void launch_seq(dim3 grid, dim3 block, size_t dyn_bytes, F kernel)
{
    bdim = D3{block.x, block.y, block.z};
    gdim = D3{grid.x, grid.y, grid.z};
    Cluster cl;
    cl.n_ctas = 1;
    pthread_mutex_init(&cl.mu, nullptr);
    Cta *b = new Cta();
    b->rank = 0; b->cluster = &cl;
    cl.cta.assign(1, b);
    cta = b;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                cl.used = 0;
                std::memset(cl.slot_set, 0, sizeof(cl.slot_set));
                b->bidx = D3{bx, by, bz};
                b->dyn.assign(dyn_bytes / 4 + 64, 0.f);
                for (unsigned tz = 0; tz < block.z; ++tz)
                    for (unsigned ty = 0; ty < block.y; ++ty)
                        for (unsigned tx = 0; tx < block.x; ++tx) { tidx = D3{tx, ty, tz}; kernel(); }
            }
    cta = nullptr;
    delete b;
    pthread_mutex_destroy(&cl.mu);
}
}
// @GENERIC_END
// Frame pairs through the tracking kernel k_track<128, 4, true>.  variant: 0 = with the speculative reject-successor (lane 0 of
// warp 1, the default of the library), 1 = without it.
extern "C" int emu_track_pairs(int variant, int n_pairs, int n_clusters, int ctas_per_pair, int n_levels, const float *const *pts, const int *n_pts,
                               const float *const *dt, const int *w, const int *h, const float *cam4, const float *R9s, const float *t3s,
                               const revo_tracker_config *cfg, int mode, int level, int pcap, revo_track_result *results, double *records)
{
    using namespace revo;
    std::vector<std::vector<uint2>> opt((size_t)n_pairs * n_levels);
    std::vector<PairDesc> pairs(n_pairs);
    for (int p = 0; p < n_pairs; ++p) {
        std::memset(&pairs[p], 0, sizeof(PairDesc));
        for (int l = 0; l < n_levels; ++l) {
            const int k = p * n_levels + l;
            const int tw = (w[k] + 3) >> 2, th = (h[k] + 3) >> 2;
            opt[k].resize((size_t)tw * th * 16);
            for (int y = 0; y < h[k]; ++y)                  // k_opt_struct, pyramid.cu
                for (int x = 0; x < w[k]; ++x)
                    opt[k][opt_texel_index(x, y, tw)] = pack_texel(opt_texel(dt[k], (size_t)y * w[k] + x, w[k], h[k]));
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
    if (variant != 0 && variant != 1) return 1;
    prm.speculate = variant == 0 ? 1 : 0;
    emu::run_grid(n_clusters, ctas_per_pair, T, dyn, [=]() { k_track<T, 4, 1>(d_pairs, n_pairs, prm, results, records, nullptr, nullptr, wc, pcap); });
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
    k = k.replace("extern __shared__ float s_pts[];", "float *s_pts = emu::cta->dyn.data();")
    k = re.sub(r'\n[^\n]*asm volatile\("fence\.mbarrier_init[^\n]*\n', "\n", k)
    assert "asm" not in k, "unexpected inline PTX left in " + name
    # `__shared__ [__align__(n)] T name[dims];`  ->  a reference into the CTA's arena (same offset in every CTA of the cluster)
    slot = [0]

    def repl(mm):
        align, typ, var, dims = mm.group(2) or "8", mm.group(3), mm.group(4), mm.group(5) or ""
        i = slot[0]
        slot[0] += 1
        if dims:
            return (f"{mm.group(1)}{typ} (&{var}){dims} = *reinterpret_cast<{typ} (*){dims}>(emu::smem_slot({i}, sizeof({typ}{dims}), {align}));")
        return f"{mm.group(1)}{typ} &{var} = *reinterpret_cast<{typ} *>(emu::smem_slot({i}, sizeof({typ}), {align}));"

    k = re.sub(r"^(\s*)__shared__ (?:__align__\((\d+)\) )?([A-Za-z_][A-Za-z0-9_]*) ([A-Za-z_][A-Za-z0-9_]*)((?:\[[^\]]+\])*);", repl, k, flags=re.M)
    assert "__shared__" not in k, "unhandled __shared__ declaration in " + name
    return k


def build(out_dir, with_lean=False):
    """Compile the emulated kernels into a shared library and return the ctypes handle."""
    rd = lambda *p: open(os.path.join(ROOT, *p)).read()      # noqa: E731
    common, pyr, internal, track = rd("revo_b200", "csrc", "track_common.cuh"), rd("revo_b200", "csrc", "pyramid.cu"), \
        rd("revo_b200", "csrc", "internal.h"), rd("revo_b200", "csrc", "track.cu")
    body = common[common.index("namespace revo {") + len("namespace revo {"):common.index("}  // namespace revo")]
    body = _strip_functions(body, ["ldg_texel", "l2_policy_evict_last", "ldg_point", "rcp_approx", "smem_u32", "mbar_init", "mbar_expect_tx", "mbar_wait", "st_async_b64"])
    assert "asm" not in body
    grab = lambda t, pat: t[re.search(pat, t, re.M).start():t.index("\n}\n", re.search(pat, t, re.M).start()) + 3]      # noqa: E731
    parts = ["namespace revo {", _struct(internal, "LevelIn"), _struct(internal, "PairDesc"), _struct(internal, "TrackParams"), body,
             grab(internal, r"^__host__ __device__ inline unsigned opt_texel_index"),
             grab(pyr, r"^__device__ __forceinline__ float4 opt_texel"), grab(pyr, r"^__device__ __forceinline__ uint32_t pack_grad"),
             grab(pyr, r"^__device__ __forceinline__ uint2 pack_texel"), _struct(track, "Mailbox"), _kernel(track, "k_track")]
    flags = []
    parts.append("}  // namespace revo")
    src, lib = os.path.join(out_dir, "cuda_emu.cpp"), os.path.join(out_dir, "libcuda_emu.so")
    open(src, "w").write(PRELUDE + "\n".join(parts) + RUNNER)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-mfma", "-ffp-contract=fast", "-shared", "-fPIC", "-pthread", *flags, *EXTRA_FLAGS,
                    "-I", os.path.join(ROOT, "include"), src, "-o", lib], check=True)
    return C.CDLL(lib)


# ---------------------------------------------------------------------------------------------------------------------
# Canny (bit-mask pipeline of canny.cu: NMS -> hysteresis -> expand -> histogram) on the same layer
# ---------------------------------------------------------------------------------------------------------------------
CANNY_SHIMS = r'''
namespace revo {
static inline int dp4a_us(unsigned a, int b, int c)      // dp4a.u32.s32: unsigned bytes of a x signed bytes of b
{
    for (int i = 0; i < 4; ++i) c += (int)((a >> (8 * i)) & 0xffu) * (int)(int8_t)((b >> (8 * i)) & 0xff);
    return c;
}
}
struct revo_ctx { int stream; uint64_t launches; };
#define REVO_CUDA(ctx, expr) do { (void)(expr); } while (0)
#define LAUNCH_CHECK(ctx) do { (ctx)->launches++; } while (0)
#define REVO_OK 0
static inline int cudaMemset2DAsync(void *p, size_t pitch, int v, size_t width, size_t height, int)
{
    for (size_t r = 0; r < height; ++r) std::memset((char *)p + r * pitch, v, width);
    return 0;
}
static inline int cudaMemsetAsync(void *p, int v, size_t bytes, int) { std::memset(p, v, bytes); return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class K> static inline int cudaFuncSetAttribute(K, int, int) { return 0; }
'''

CANNY_DRIVER = r'''
// cv::Canny(gray, edges, t1, t2, 3, true) + the patch histogram of generateDistHistogram through the bit-mask kernels
extern "C" int emu_canny_bits(const uint8_t *gray, int w, int h, int low_sq, int high_sq, int patch, uint8_t *edges, uint8_t *edges_orig,
                              uint8_t *hist, int *nz_patches)
{
    using namespace revo;
    std::vector<int> labels((size_t)w * h + 64, 0);
    std::vector<uint8_t> flags((size_t)w * h + 256, 0);
    ImgLevel L;
    std::memset(&L, 0, sizeof(L));
    L.gray = (uint8_t *)gray; L.edges = edges; L.edges_orig = edges_orig; L.hist = hist; L.nz_patches = nz_patches;
    L.labels = labels.data(); L.flags = flags.data(); L.w = w; L.h = h; L.patch = patch; L.hist_w = w / patch; L.hist_h = h / patch;
    revo_ctx ctx{0, 0};
    return launch_canny_bits(&ctx, &L, 1, w, h, low_sq, high_sq, patch, flags.data(), flags.size());
}
'''


def _split_top(text):
    out, depth, cur = [], 0, ""
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


_COLLECTIVES = ("__syncthreads", "__shfl", "__ballot", "__any_sync", "__reduce", "__syncwarp", "flood_row", "spread_row", "flood_up_row")


def _kernel_bodies(text):
    """name -> body text of every __global__ function of `text`."""
    out = {}
    for m in re.finditer(r"^__global__ void (?:__launch_bounds__\([^)]*\)\s*)?([A-Za-z_][A-Za-z0-9_]*)\(", text, re.M):
        out[m.group(1)] = text[m.start():text.index("\n}\n", m.start()) + 3]
    return out


def _launches(text):
    """kernel<<<grid, block[, smem[, stream]]>>>(args);  ->  emu::launch(grid, block, smem, [=]() { kernel(args); });
    kernels whose body has no barrier / warp collective run their threads sequentially (emu::launch_seq)."""
    bodies = _kernel_bodies(text)

    def repl(m):
        cfg = _split_top(m.group(2))
        smem = cfg[2] if len(cfg) > 2 else "0"
        base = re.sub(r"<.*", "", m.group(1))
        seq = base in bodies and not any(c in bodies[base] for c in _COLLECTIVES)
        return f"emu::{'launch_seq' if seq else 'launch'}({cfg[0]}, {cfg[1]}, {smem}, [=]() {{ {m.group(1)}({m.group(3)}); }});"
    return re.sub(r"([A-Za-z_][A-Za-z0-9_]*(?:<[^<>;]*>)?)<<<(.*?)>>>\((.*?)\);", repl, text, flags=re.S)


def _device_text(text):
    """__shared__ declarations -> arena slots, dynamic shared memory -> the CTA's dynamic buffer."""
    text = re.sub(r"extern __shared__ ([A-Za-z_ ]+?) ([A-Za-z_][A-Za-z0-9_]*)\[\];", r"\1 *\2 = (\1 *)emu::cta->dyn.data();", text)
    slot = [16]                                    # the tracking kernels use the first slots

    def repl(mm):
        align, typ, var, dims = mm.group(2) or "8", mm.group(3), mm.group(4), mm.group(5) or ""
        i = slot[0]
        slot[0] += 1
        if dims:
            return f"{mm.group(1)}{typ} (&{var}){dims} = *reinterpret_cast<{typ} (*){dims}>(emu::smem_slot({i}, sizeof({typ}{dims}), {align}));"
        return f"{mm.group(1)}{typ} &{var} = *reinterpret_cast<{typ} *>(emu::smem_slot({i}, sizeof({typ}), {align}));"

    text = re.sub(r"^(\s*)__shared__ (?:__align__\((\d+)\) )?([A-Za-z_][A-Za-z0-9_]*) ([A-Za-z_][A-Za-z0-9_]*)((?:\[[^\]]+\])*);", repl, text, flags=re.M)
    assert "__shared__" not in text
    return text


def build_canny(out_dir):
    rd = lambda *p: open(os.path.join(ROOT, *p)).read()      # noqa: E731
    canny, internal = rd("revo_b200", "csrc", "canny.cu"), rd("revo_b200", "csrc", "internal.h")
    a = canny.index("// counts -> wrapping u8 histogram")
    b = canny.index("\n}\n", canny.index("static int launch_canny_bits")) + 3
    body = canny[a:b]
    body = _strip_functions(body, ["dp4a_us"])
    body = body.replace("#pragma unroll", "")
    body = _launches(_device_text(body))
    assert "asm" not in body and "<<<" not in body
    generic = RUNNER[RUNNER.index("// @GENERIC_BEGIN"):RUNNER.index("// @GENERIC_END")]
    src_text = (PRELUDE + CANNY_SHIMS + generic + "namespace revo {\nstatic inline int cdiv(int a, int b) { return (a + b - 1) / b; }\n"
                + _struct(internal, "ImgLevel") + "\n" + body + "}  // namespace revo\n" + CANNY_DRIVER)
    src, lib = os.path.join(out_dir, "canny_emu.cpp"), os.path.join(out_dir, "libcanny_emu.so")
    open(src, "w").write(src_text)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-pthread", *EXTRA_FLAGS, "-I", os.path.join(ROOT, "include"), src, "-o", lib],
                   check=True)
    return C.CDLL(lib)


# ---------------------------------------------------------------------------------------------------------------------
# The whole pyramid construction (pyramid.cu + the bit-mask Canny) with the launch sequence of capi.cu:create_batch_impl
# ---------------------------------------------------------------------------------------------------------------------
PYRAMID_DRIVER = r'''
struct EmuLevelOut {          // caller-provided host buffers of one level
    uint8_t *gray; float *depth; uint8_t *edges, *edges_orig, *hist; float *pts; int *n_pts, *nz_patches;
    float *dt; uint32_t *opt; float *pts_ref; int *n_ref; float *opt_f4;
    int w, h, patch, cap, n_tiles;
    float fx, fy, cx, cy;
};

// n identical frames (n >= 8 takes the group compaction of launch_compact, n < 8 the tile one); outputs of frame 0
extern "C" int emu_pyramid(const uint8_t *bgr, int channels, const float *depth, int n_frames, int n_levels, EmuLevelOut *out, int low_sq,
                           int high_sq, float dmin, float dmax, int use_edge_hist, float n_percentage, int keyframe)
{
    using namespace revo;
    revo_ctx ctx{0, 0};
    const int w0 = out[0].w, h0 = out[0].h, n = n_frames;
    std::vector<std::vector<ImgLevel>> desc(n_levels, std::vector<ImgLevel>(n));
    std::vector<std::vector<uint8_t>> store;
    auto alloc = [&](size_t bytes) { store.emplace_back(bytes + 256, 0); return store.back().data(); };
    std::vector<uint8_t *> labels(n), flags(n);
    for (int f = 0; f < n; ++f) { labels[f] = alloc((size_t)w0 * h0 * 4); flags[f] = alloc((size_t)w0 * h0); }
    for (int l = 0; l < n_levels; ++l)
        for (int f = 0; f < n; ++f) {
            const EmuLevelOut &o = out[l];
            ImgLevel &L = desc[l][f];
            std::memset(&L, 0, sizeof(L));
            const size_t px = (size_t)o.w * o.h;
            const bool first = f == 0;
            L.gray = first ? o.gray : alloc(px); L.depth = first ? o.depth : (float *)alloc(px * 4);
            L.edges = first ? o.edges : alloc(px); L.edges_orig = first ? o.edges_orig : alloc(px);
            L.hist = first ? o.hist : alloc(px); L.pts = (float4 *)(first ? (uint8_t *)o.pts : alloc((size_t)o.cap * 16));
            L.n_pts = first ? o.n_pts : (int *)alloc(8); L.nz_patches = first ? o.nz_patches : (int *)alloc(8);
            L.tile_off = (int *)alloc(((size_t)o.n_tiles + 1) * 4);
            L.labels = (int *)labels[f]; L.flags = flags[f];
            L.dt = first ? o.dt : (float *)alloc(px * 4); L.opt = (uint2 *)(first ? (uint8_t *)o.opt : alloc(px * 8 + 4096));
            L.w = o.w; L.h = o.h; L.pts_cap = o.cap; L.patch = o.patch; L.hist_w = o.w / o.patch; L.hist_h = o.h / o.patch;
            L.fx = o.fx; L.fy = o.fy; L.cx = o.cx; L.cy = o.cy;
        }
    // n copies of the input frame, tightly packed (what revo_pyr_create_batch takes)
    const size_t bgr_frame = (size_t)w0 * h0 * channels;
    std::vector<uint8_t> bgr_n(bgr_frame * n);
    for (int f = 0; f < n; ++f) {
        std::memcpy(bgr_n.data() + bgr_frame * f, bgr, bgr_frame);
        std::memcpy(desc[0][f].depth, depth, (size_t)w0 * h0 * 4);
    }
    // the launch sequence of create_batch_impl (capi.cu)
    int rc = launch_gray(&ctx, bgr_n.data(), (size_t)w0 * channels, channels, bgr_frame, desc[0].data(), n, w0, h0);
    for (int l = 0; l < n_levels && !rc; ++l) {
        const EmuLevelOut &o = out[l];
        if (l > 0) rc = launch_pyrdown_depth(&ctx, desc[l - 1].data(), desc[l].data(), n, o.w, o.h, out[l - 1].w, out[l - 1].h);
        if (!rc) rc = launch_canny_bits(&ctx, desc[l].data(), n, o.w, o.h, low_sq, high_sq, o.patch, flags[0], 0);
        const bool fill = use_edge_hist && l >= 1 && l <= 2;
        if (!rc) rc = launch_hist_fill(&ctx, desc[l].data(), l > 0 ? desc[l - 1].data() : nullptr, n, o.w, o.h, o.patch,
                                      l > 0 ? out[l - 1].patch : o.patch, fill, n_percentage);
        if (!rc) rc = launch_compact(&ctx, desc[l].data(), n, o.w, o.h, dmin, dmax);
    }
    for (int l = 0; l < n_levels && !rc; ++l) {
        const EmuLevelOut &o = out[l];
        std::vector<int> col_off(o.w + 2);
        rc = launch_edges3d_reference_order(&ctx, desc[l].data(), o.w, o.h, dmin, dmax, (float4 *)o.pts_ref, o.n_ref, col_off.data());
        if (!rc && keyframe) rc = launch_keyframe(&ctx, desc[l].data(), n, o.w, o.h);
        if (!rc && keyframe) rc = launch_opt_struct_f4(&ctx, o.dt, o.w, o.h, (float4 *)o.opt_f4);
    }
    return rc;
}

// TrackerNew::assessTrackingQuality on the device (k_quality_scatter + k_quality_hist): the 16 counters of launch_quality
extern "C" int emu_quality(int n_frames, const float *const *pts, const int *n_pts, const float *R9s, const float *T3s, float fx, float fy,
                           float cx, float cy, int w, int h, const float *depth, const uint8_t *edges, float dmin, float dmax, int *counters16)
{
    using namespace revo;
    revo_ctx ctx{0, 0};
    QualityArgs a;
    std::memset(&a, 0, sizeof(a));
    a.n_frames = n_frames; a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.w = w; a.h = h; a.depth = depth; a.edges = edges;
    for (int f = 0; f < n_frames; ++f) {
        a.fr[f].pts = (const float4 *)pts[f]; a.fr[f].n_pts = &n_pts[f];
        std::memcpy(a.fr[f].R, R9s + 9 * f, sizeof(float) * 9);
        std::memcpy(a.fr[f].T, T3s + 3 * f, sizeof(float) * 3);
    }
    // two votes in one launch pair: the vote asked for and an empty one (no past frames), whose counters must stay apart
    QualityArgs two[2] = {a, a};
    two[1].n_frames = 0;
    const size_t words = ((size_t)w * h + 3) / 4;
    std::vector<unsigned> mbits(2 * words + 16);
    std::vector<int> counters(32);
    int rc = launch_quality(&ctx, two, 2, w, h, dmin, dmax, mbits.data(), counters.data());
    std::memcpy(counters16, counters.data(), 16 * sizeof(int));
    for (int k = 1; k < 4; ++k)
        if (counters[16 + k] || counters[20 + k]) rc = rc ? rc : -100;      // nothing can overlap without past frames
    return rc;
}

// ImgPyramidRGBD::generateColoredPcl on the device: colour pyrDown to the level + count / scan / scatter (launch_colored_pcl)
extern "C" int emu_colored_pcl(const uint8_t *bgr0, int channels, int w0, int h0, int lvl, const float *depth, const uint8_t *edges, int w, int h,
                               float fx, float fy, float cx, float cy, int dense, float dmin, float dmax, float *out, int cap, int *n_out)
{
    using namespace revo;
    revo_ctx ctx{0, 0};
    std::vector<uint8_t> a(bgr0, bgr0 + (size_t)w0 * h0 * channels), b(a.size());
    int cw = w0, chh = h0, rc = 0;
    for (int l = 0; l < lvl && !rc; ++l) {
        rc = launch_pyrdown_color(&ctx, a.data(), b.data(), cw, chh, channels);
        cw = (cw + 1) / 2; chh = (chh + 1) / 2;
        a.swap(b);
    }
    if (rc || cw != w || chh != h) return rc ? rc : -101;
    ImgLevel L;
    std::memset(&L, 0, sizeof(L));
    L.depth = const_cast<float *>(depth); L.edges = const_cast<uint8_t *>(edges); L.w = w; L.h = h; L.fx = fx; L.fy = fy; L.cx = cx; L.cy = cy;
    std::vector<int> col(w + 4);
    return launch_colored_pcl(&ctx, &L, w, h, dense, dmin, dmax, a.data(), channels, out, cap, n_out, col.data());
}
'''


def build_pyramid(out_dir):
    rd = lambda *p: open(os.path.join(ROOT, *p)).read()      # noqa: E731
    canny, pyr, internal = rd("revo_b200", "csrc", "canny.cu"), rd("revo_b200", "csrc", "pyramid.cu"), rd("revo_b200", "csrc", "internal.h")
    a = canny.index("// counts -> wrapping u8 histogram")
    b = canny.index("\n}\n", canny.index("static int launch_canny_bits")) + 3
    cbody = _strip_functions(canny[a:b], ["dp4a_us"])
    pa = pyr.index("// ---------------------------------------------------------------------------\n// K1: BGR(A) -> gray")
    pb = pyr.index("}  // namespace revo")
    pbody = pyr[pa:pb]
    # the memset of the histogram counters in launch_canny_bits takes the per-frame pitch of the slab: here every frame has its
    # own counter buffer, so clear them through the descriptors instead
    cbody = cbody.replace("REVO_CUDA(ctx, cudaMemset2DAsync(d_counts0, counts_stride, 0, (size_t)hist_w * hist_h * sizeof(int), (size_t)n, ctx->stream));",
                          "for (int f_ = 0; f_ < n; ++f_) std::memset(d_desc[f_].flags, 0, (size_t)hist_w * hist_h * sizeof(int));")
    body = (cbody + pbody).replace("#pragma unroll", "")
    body = _launches(_device_text(body))
    assert "asm" not in body and "<<<" not in body
    generic = RUNNER[RUNNER.index("// @GENERIC_BEGIN"):RUNNER.index("// @GENERIC_END")]
    src_text = (PRELUDE + CANNY_SHIMS + generic + "namespace revo {\nstatic inline int cdiv(int a, int b) { return (a + b - 1) / b; }\n"
                + "constexpr int kTileW = 8;\nconstexpr int kTileH = 4;\n"
                + _struct(internal, "ImgLevel") + "\n" + internal[internal.index("__host__ __device__ inline unsigned opt_texel_index"):internal.index("inline int opt_tiles_per_row")]
                + _struct(internal, "QualityFrame") + "\n" + _struct(internal, "QualityArgs") + "\n" + _struct(internal, "PointListCopy") + "\n"
                + body + "}  // namespace revo\n" + PYRAMID_DRIVER)
    src, lib = os.path.join(out_dir, "pyramid_emu.cpp"), os.path.join(out_dir, "libpyramid_emu.so")
    open(src, "w").write(src_text)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-pthread", *EXTRA_FLAGS, "-I", os.path.join(ROOT, "include"), src, "-o", lib],
                   check=True)
    return C.CDLL(lib)
