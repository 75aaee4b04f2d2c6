
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
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace emu {
struct D3 { unsigned x, y, z; };
constexpr int kMaxWarps = 32, kArenaBytes = 64 * 1024, kMaxSlots = 32;
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
static inline void ldg_quad(const uint4 *p, uint4 &r0, uint4 &r1)
{
    const uint32_t *q = (const uint32_t *)p;
    r0 = make_uint4(q[0], q[1], q[4], q[5]);
    r1 = make_uint4(q[2], q[3], q[6], q[7]);
}
template <int kHint> static inline void ldg_quad_h(const uint4 *p, uint4 &r0, uint4 &r1) { ldg_quad(p, r0, r1); }
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

// ---- fake CUDA runtime: device memory is host memory, everything is synchronous -------------------------------------
#include <new>
#include <string>
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801, cudaErrorInvalidValue = 1 };
struct FakeStream { int id; };
struct FakeEvent { int id; };
typedef FakeStream *cudaStream_t;
typedef FakeEvent *cudaEvent_t;
typedef void *cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaMemPoolAttrReleaseThreshold = 4, cudaIpcMemLazyEnablePeerAccess = 1,
       cudaEnableDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; };
struct cudaDeviceProp { int multiProcessorCount; char name[256]; };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributeNonPortableClusterSizeAllowed = 9 };
static inline const char *cudaGetErrorString(cudaError_t) { return "fake CUDA runtime"; }
static inline const char *cudaGetErrorName(cudaError_t) { return "cudaErrorFake"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { std::memset(p, 0, sizeof(*p)); p->multiProcessorCount = 2; return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void *) { return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new FakeStream{0}; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new FakeEvent{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new FakeEvent{0}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::calloc(1, n + 64); return *p ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset2DAsync(void *p, size_t pitch, int v, size_t width, size_t height, cudaStream_t)
{
    for (size_t r = 0; r < height; ++r) std::memset((char *)p + r * pitch, v, width);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t width, size_t height, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < height; ++r) std::memmove((char *)d + r * dp, (const char *)s + r * sp, width);
    return cudaSuccess;
}
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
static inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaErrorNotSupported; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
// launches with attributes (thread-block clusters)
enum cudaLaunchAttributeID { cudaLaunchAttributeClusterDimension = 4 };
struct cudaLaunchAttribute {
    cudaLaunchAttributeID id;
    struct { struct { unsigned x, y, z; } clusterDim; } val;
};
struct cudaLaunchConfig_t {
    dim3 gridDim_, blockDim_;   // (gridDim / blockDim are macros of the emulation layer)
    size_t dynamicSmemBytes;
    cudaStream_t stream;
    cudaLaunchAttribute *attrs;
    unsigned numAttrs;
};
template <class K> static inline cudaError_t cudaOccupancyMaxActiveClusters(int *n, K, const cudaLaunchConfig_t *) { *n = 2; return cudaSuccess; }
template <class... P, class... A> static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *cfg, void (*kern)(P...), A... args)
{
    int C = 1;
    for (unsigned i = 0; i < cfg->numAttrs; ++i)
        if (cfg->attrs[i].id == cudaLaunchAttributeClusterDimension) C = (int)cfg->attrs[i].val.clusterDim.x;
    emu::run_grid((int)cfg->gridDim_.x / C, C, (int)cfg->blockDim_.x, cfg->dynamicSmemBytes, [=]() { kern(args...); });
    return cudaSuccess;
}
namespace revo {
static inline int dp4a_us(unsigned a, int b, int c)      // dp4a.u32.s32: unsigned bytes of a x signed bytes of b
{
    for (int i = 0; i < 4; ++i) c += (int)((a >> (8 * i)) & 0xffu) * (int)(int8_t)((b >> (8 * i)) & 0xff);
    return c;
}
}
// internal.h -- shared declarations of the revo_b200 CUDA library (not part of the C ABI).




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
    uint4 *opt;           // optimizationStructure[l] in the device QUAD layout (see k_opt_struct), 2 x uint4 per pixel (keyframes)
    int w, h;
    int pts_cap;
    int patch;            // distPatchSizes[l]
    int hist_w, hist_h;
    float fx, fy, cx, cy; // Camera at this level (camerapyr.h:98-103)
};

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
    int track_engine;        // 0 = automatic, 1 = one cluster per pair (track.cu), 2 = task queue (track_queue.cu), 3 = ping-pong clusters (track_pp.cu)
    int track_chunk_points;  // queue engine: minimum points per task (0 = automatic)
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
// tracking-quality vote (revo_track_quality): the past frames' 3-D lists with the transform into the current frame
struct QualityFrame {
    const float4 *pts;
    const int *n_pts;
    float R[9], T[3];     // column-major R, as Eigen::Matrix3f
};
struct QualityArgs {
    QualityFrame fr[4];
    int n_frames;
    float fx, fy, cx, cy;
    int w, h;
};
// d_counters: 16 ints = histogram[4], overlaps[4], out_of_bounds, ...
int launch_quality(revo_ctx *ctx, const QualityArgs &args, const float *d_depth, const uint8_t *d_edges, float dmin, float dmax,
                   unsigned *d_mbits, int *d_counters);
int launch_opt_struct_f4(revo_ctx *ctx, const float *d_dt, int w, int h, float4 *d_out);
int launch_opt_pack_from_f4(revo_ctx *ctx, const float4 *d_in, int w, int h, uint4 *d_out);
// reference-order (column-major scan) 3-D edge list into d_out (capacity w*h float4); *d_n receives the count
int launch_edges3d_reference_order(revo_ctx *ctx, const ImgLevel *d_desc_one, int w, int h, float dmin, float dmax,
                                   float4 *d_out, int *d_n, int *d_col_off);

// ---- track.cu --------------------------------------------------------------
struct LevelIn {
    const float4 *pts;
    const int *n_pts;
    const uint4 *opt;    // quad layout, 32 B per pixel: dt of (x,y),(x+1,y),(x,y+1),(x+1,y+1) | snorm16 gx|gy of the same four
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
    // split mode
    int split_rank, split_world;
    unsigned long long split_seq0;
    void *split_peers[16];
};
int launch_track(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm,
                 revo_track_result *d_results, double *d_records, revo_trace_entry *d_trace, int *d_trace_counts,
                 int *d_work_counter);

int launch_stage_in(revo_ctx *ctx, const void *src_mapped_host, void *dst, size_t bytes);

// ---- track_pp.cu: cluster engine with warp-specialised CTAs working on two pairs at once ---------------
int launch_track_pp(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                    double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, int *d_work_counter);

// ---- track_queue.cu ----------------------------------------------------------
// Task-queue engine: device workspace size for n_pairs (ring + pair states + partial tables) and the launcher.
size_t track_queue_workspace_bytes(int n_pairs, int grid_cap, unsigned *cap_out);
int launch_track_queue(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                       double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, void *d_ws, size_t ws_bytes);

}  // namespace revo

namespace revo {
// engines / paths that are not part of the CPU build
int launch_track_pp(revo_ctx *ctx, const PairDesc *, int, const TrackParams &, revo_track_result *, double *, revo_trace_entry *, int *, int *)
{
    ctx->last_error = "ping-pong engine: not in the emulated build";
    return REVO_ERR_UNSUPPORTED;
}
int launch_track_queue(revo_ctx *ctx, const PairDesc *, int, const TrackParams &, revo_track_result *, double *, revo_trace_entry *, int *, void *, size_t)
{
    ctx->last_error = "task-queue engine: not in the emulated build";
    return REVO_ERR_UNSUPPORTED;
}
size_t track_queue_workspace_bytes(int, int, unsigned *) { return 256; }
bool make_gray_tensor_map(void *, const uint8_t *, int, int, int, size_t) { return false; }
}
// track_common.cuh -- device helpers shared by the two tracking engines (track.cu: one cluster per pair;
// track_queue.cu: chip-wide task queue): record layout, SE3 / 6x6 solver in double, the fused PASS A + PASS B
// per-point work and the transposing warp reduction.
//
// Reference (fabianschenk/REVO): system/optimizer.cpp:74-311, system/optimizer.h:156-185, utils/LGSX.h:196-398,
// thirdparty/Sophus/sophus/se3.hpp:317-321,723-748, so3.hpp:335-352,419-424,531-564, system/tracker.cpp:357-393.



namespace revo {

constexpr unsigned kFull = 0xffffffffu;

// ---- record layout ---------------------------------------------------------
// [0..20] sum w v_i v_j (i<=j, LGS6 slot order), [21..26] sum w r v_i, [27] sum w r^2, [28] sum r^2,
// [29] good, [30] bad, [31] unused.
constexpr int kRecA = 0, kRecB = 21, kRecSW = 27, kRecSU = 28, kRecGood = 29, kRecBad = 30;

struct Ctrl {
    // written by thread 0 of every CTA (identically), read by all threads
    float R[9];
    float t[3];
    int level_done;
    int pair_skip;
    int next_pair;
};

struct LMState {
    double q[4], t[3];    // accepted pose (Sophus SE3: unit quaternion xyzw + translation)
    double qn[4], tn[3];  // trial pose
    double A[21], b[6], n;
    double inc[6];
    float lastErr, last_residual, lambda;
    int iteration, incTry, tries;
};

// ---- small double-precision SE3 / solver helpers (thread 0 only) --------------
__device__ __forceinline__ void quat_to_R(const double *q, double *R /* col-major */)
{
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[3] = txy - twz;       R[6] = txz + twy;
    R[1] = txy + twz;       R[4] = 1 - (txx + tzz); R[7] = tyz - twx;
    R[2] = txz - twy;       R[5] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Eigen quaternion-from-matrix (Shepperd), as SO3(Matrix3) does (so3.hpp:419). R col-major float.
__device__ inline void quat_from_R(const float *Rf, double *q)
{
    double R[9];
    for (int i = 0; i < 9; ++i) R[i] = Rf[i];
#define RMAT(i, j) R[(j) * 3 + (i)]
    double t = RMAT(0, 0) + RMAT(1, 1) + RMAT(2, 2);
    if (t > 0) {
        t = sqrt(t + 1.0);
        q[3] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (RMAT(2, 1) - RMAT(1, 2)) * t;
        q[1] = (RMAT(0, 2) - RMAT(2, 0)) * t;
        q[2] = (RMAT(1, 0) - RMAT(0, 1)) * t;
    } else {
        int i = 0;
        if (RMAT(1, 1) > RMAT(0, 0)) i = 1;
        if (RMAT(2, 2) > RMAT(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(RMAT(i, i) - RMAT(j, j) - RMAT(k, k) + 1.0);
        q[i] = 0.5 * t;
        t = 0.5 / t;
        q[3] = (RMAT(k, j) - RMAT(j, k)) * t;
        q[j] = (RMAT(j, i) + RMAT(i, j)) * t;
        q[k] = (RMAT(k, i) + RMAT(i, k)) * t;
    }
#undef RMAT
}

// ||R R^T - I||_F < 1e-5 and det > 0: the Sophus ENSUREs of so3.hpp:419-424 (float epsilon, common.hpp:152).
__device__ inline bool rotation_ok(const float *Rf)
{
    double n2 = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)Rf[k * 3 + i] * (double)Rf[k * 3 + j];
            s -= (i == j) ? 1.0 : 0.0;
            n2 += s * s;
        }
    const double det = (double)Rf[0] * ((double)Rf[4] * Rf[8] - (double)Rf[7] * Rf[5]) -
                       (double)Rf[3] * ((double)Rf[1] * Rf[8] - (double)Rf[7] * Rf[2]) +
                       (double)Rf[6] * ((double)Rf[1] * Rf[5] - (double)Rf[4] * Rf[2]);
    return (sqrt(n2) < 1e-5) && (det > 0);
}

// Sophus::SE3::exp (se3.hpp:723-748, so3.hpp:531-564) in double.  The four coefficients sin(t/2)/t, cos(t/2),
// (1 - cos t)/t^2 and (t - sin t)/t^3 are even functions of t; for the increments of a tracker (t < 0.5 rad, in practice
// < 0.05) they are evaluated as power series in t^2 (8 terms: truncation < 1e-17 relative) as four independent Horner
// chains: no sqrt, sincos or division on the serial critical path of an evaluation.  Larger angles take the closed form.
__device__ __forceinline__ void se3_exp(const double *xi, double *q, double *t)
{
    const double ox = xi[3], oy = xi[4], oz = xi[5];
    const double s = ox * ox + oy * oy + oz * oz;   // theta^2
    double imag, re, c1, c2;
    if (s < 1e-10) {   // theta < Sophus::Constants<float>::epsilon() = 1e-5
        const double t4 = s * s;
        imag = 0.5 - (1.0 / 48.0) * s + (1.0 / 3840.0) * t4;
        re = 1.0 - (1.0 / 8.0) * s + (1.0 / 384.0) * t4;
        // V = R(q) there (se3.hpp:735-737) = I + 2 re imag Om + 2 imag^2 Om^2
        c1 = 2.0 * re * imag;
        c2 = 2.0 * imag * imag;
    } else if (s < 0.25) {
        // coefficients: 1/(2^(2k+1) (2k+1)!), 1/(4^k (2k)!), 1/(2k+2)!, 1/(2k+3)!  with alternating sign
        imag = 1.0 / 42849873690624000.0;
        re = 1.0 / 1428329123020800.0;
        c1 = 1.0 / 20922789888000.0;
        c2 = 1.0 / 355687428096000.0;
        imag = imag * -s + 1.0 / 51011754393600.0;     re = re * -s + 1.0 / 1961990553600.0;
        c1 = c1 * -s + 1.0 / 87178291200.0;             c2 = c2 * -s + 1.0 / 1307674368000.0;
        imag = imag * -s + 1.0 / 81749606400.0;         re = re * -s + 1.0 / 3715891200.0;
        c1 = c1 * -s + 1.0 / 479001600.0;               c2 = c2 * -s + 1.0 / 6227020800.0;
        imag = imag * -s + 1.0 / 185794560.0;           re = re * -s + 1.0 / 10321920.0;
        c1 = c1 * -s + 1.0 / 3628800.0;                 c2 = c2 * -s + 1.0 / 39916800.0;
        imag = imag * -s + 1.0 / 645120.0;              re = re * -s + 1.0 / 46080.0;
        c1 = c1 * -s + 1.0 / 40320.0;                   c2 = c2 * -s + 1.0 / 362880.0;
        imag = imag * -s + 1.0 / 3840.0;                re = re * -s + 1.0 / 384.0;
        c1 = c1 * -s + 1.0 / 720.0;                     c2 = c2 * -s + 1.0 / 5040.0;
        imag = imag * -s + 1.0 / 48.0;                  re = re * -s + 1.0 / 8.0;
        c1 = c1 * -s + 1.0 / 24.0;                      c2 = c2 * -s + 1.0 / 120.0;
        imag = imag * -s + 0.5;                         re = re * -s + 1.0;
        c1 = c1 * -s + 0.5;                             c2 = c2 * -s + 1.0 / 6.0;
    } else {
        const double theta = sqrt(s);
        double sn, cs;
        sincos(0.5 * theta, &sn, &cs);
        const double inv_t = __drcp_rn(theta), inv_t2 = inv_t * inv_t;
        imag = sn * inv_t;
        re = cs;
        c1 = 2.0 * sn * sn * inv_t2;                        // (1 - cos t) / t^2
        c2 = (theta - 2.0 * sn * cs) * inv_t2 * inv_t;      // (t - sin t) / t^3
    }
    q[0] = imag * ox; q[1] = imag * oy; q[2] = imag * oz; q[3] = re;
    // V = I + c1 Om + c2 Om^2 ; Om = hat(omega), Om^2 = omega omega^T - |omega|^2 I
    const double v00 = 1 + c2 * (ox * ox - s), v01 = -c1 * oz + c2 * ox * oy, v02 = c1 * oy + c2 * ox * oz;
    const double v10 = c1 * oz + c2 * ox * oy, v11 = 1 + c2 * (oy * oy - s), v12 = -c1 * ox + c2 * oy * oz;
    const double v20 = -c1 * oy + c2 * ox * oz, v21 = c1 * ox + c2 * oy * oz, v22 = 1 + c2 * (oz * oz - s);
    t[0] = v00 * xi[0] + v01 * xi[1] + v02 * xi[2];
    t[1] = v10 * xi[0] + v11 * xi[1] + v12 * xi[2];
    t[2] = v20 * xi[0] + v21 * xi[1] + v22 * xi[2];
}

// (qa,ta) * (qb,tb) with Sophus' renormalisation (se3.hpp:317-321, so3.hpp:335-352)
__device__ __forceinline__ void se3_mul(const double *qa, const double *ta, const double *qb, const double *tb, double *q, double *t)
{
    double ux = qa[1] * tb[2] - qa[2] * tb[1], uy = qa[2] * tb[0] - qa[0] * tb[2], uz = qa[0] * tb[1] - qa[1] * tb[0];
    ux += ux; uy += uy; uz += uz;
    const double cx = qa[1] * uz - qa[2] * uy, cy = qa[2] * ux - qa[0] * uz, cz = qa[0] * uy - qa[1] * ux;
    t[0] = ta[0] + (tb[0] + qa[3] * ux + cx);
    t[1] = ta[1] + (tb[1] + qa[3] * uy + cy);
    t[2] = ta[2] + (tb[2] + qa[3] * uz + cz);
    const double ax = qa[0], ay = qa[1], az = qa[2], aw = qa[3], bx = qb[0], by = qb[1], bz = qb[2], bw = qb[3];
    double w = aw * bw - ax * bx - ay * by - az * bz;
    double x = aw * bx + ax * bw + ay * bz - az * by;
    double y = aw * by + ay * bw + az * bx - ax * bz;
    double z = aw * bz + az * bw + ax * by - ay * bx;
    const double sn = x * x + y * y + z * z + w * w;
    if (sn != 1.0) {
        const double s = 2.0 * __drcp_rn(1.0 + sn);
        x *= s; y *= s; z *= s; w *= s;
    }
    q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

// Solve (A/n with diag * lam1) x = b/n for the symmetric positive (semi-)definite 6x6 normal equations
// (system/optimizer.cpp:258-262, "A.ldlt().solve(b)").  LDL^T in double, fully unrolled so that everything
// stays in registers; no pivoting (the matrix is a damped sum of outer products; Eigen's diagonal pivoting
// only changes rounding, which double precision makes irrelevant at the float tolerance of this path).
// Non-positive / non-finite pivots are treated like Eigen's pseudo-inverse of D: that component becomes 0.
__device__ __forceinline__ void solve6(const double *Au /* 21 upper slots */, const double *b, double inv_n, double lam1, double *x)
{
    double a[6][6];
    {
        int s = 0;

        for (int i = 0; i < 6; ++i)

            for (int j = i; j < 6; ++j) a[j][i] = Au[s++] * inv_n;   // lower triangle
    }
    double y[6], invd[6];

    for (int i = 0; i < 6; ++i) { a[i][i] *= lam1; y[i] = b[i] * inv_n; }

    for (int k = 0; k < 6; ++k) {
        const double dk = a[k][k];
        const double id = (dk > 0.0 && dk < 1e300) ? __drcp_rn(dk) : 0.0;
        invd[k] = id;

        for (int j = k + 1; j < 6; ++j) {
            const double ljk = a[j][k] * id;

            for (int i = j; i < 6; ++i) a[i][j] -= a[i][k] * ljk;
        }

        for (int i = k + 1; i < 6; ++i) a[i][k] *= id;   // L
    }

    for (int i = 1; i < 6; ++i)

        for (int j = 0; j < i; ++j) y[i] -= a[i][j] * y[j];

    for (int i = 0; i < 6; ++i) y[i] *= invd[i];

    for (int i = 4; i >= 0; --i)

        for (int j = i + 1; j < 6; ++j) y[i] -= a[j][i] * y[j];

    for (int i = 0; i < 6; ++i) x[i] = y[i];
}

// ---- one step of the Levenberg-Marquardt state machine (thread-serial) ------------------------------------------
// Optimizer::trackFrames, system/optimizer.cpp:243-306, restated as "consume the record of the evaluation that just
// finished, decide, and name the next pose to evaluate".  `first`: the record was taken at the level's start pose
// (optimizer.cpp:246-249); otherwise at the trial pose (lm.qn, lm.tn).  Returns true when the level is finished;
// R_out/t_out then hold the accepted pose (:308-309), else the next trial pose exp(inc) * referenceToFrame (:266).
// *traced is set when an LM try was judged (te, if not null, receives it).
__device__ __forceinline__ bool lm_step(LMState &lm, const double *rec, const revo_opt_config &oc, int lvl, bool first,
                                        float *R_out, float *t_out, revo_trace_entry *te, bool *traced)
{
    const float err = (float)(rec[kRecSW] / rec[kRecGood]);    // :190
    bool propose = false, done = false;
    *traced = false;
    if (first) {
        lm.lastErr = err;
        lm.last_residual = err;
        lm.lambda = oc.lambda_initial[lvl];
        lm.iteration = 0; lm.incTry = 0; lm.tries = 0;
        for (int i = 0; i < 21; ++i) lm.A[i] = rec[kRecA + i];
        for (int i = 0; i < 6; ++i) lm.b[i] = rec[kRecB + i];
        lm.n = rec[kRecGood];
        propose = true;
    } else {
        const bool accepted = err < lm.lastErr;                // :273
        *traced = true;
        if (te) {
            te->error = err; te->lambda = lm.lambda; te->accepted = accepted ? 1 : 0;
            te->good = (int)rec[kRecGood]; te->bad = (int)rec[kRecBad]; te->level = lvl;
        }
        if (accepted) {
            for (int i = 0; i < 4; ++i) lm.q[i] = lm.qn[i];
            for (int i = 0; i < 3; ++i) lm.t[i] = lm.tn[i];
            for (int i = 0; i < 21; ++i) lm.A[i] = rec[kRecA + i];
            for (int i = 0; i < 6; ++i) lm.b[i] = rec[kRecB + i];
            lm.n = rec[kRecGood];
            if (err / lm.lastErr > oc.convergence_eps[lvl]) lm.iteration = oc.max_its_per_lvl[lvl];   // :279-283
            lm.last_residual = lm.lastErr = err;
            if (lm.lambda <= 0.2f) lm.lambda = 0.f; else lm.lambda *= oc.lambda_success_fac;          // :286-289
            lm.iteration++;     // for-loop increment after the break (:291)
            lm.incTry = 0;
            propose = true;
        } else {
            double dot = 0;
            for (int i = 0; i < 6; ++i) dot += lm.inc[i] * lm.inc[i];
            if (!((float)dot > oc.step_size_min[lvl])) {                                               // :294
                done = true;
            } else {
                if (lm.lambda == 0.f) lm.lambda = 0.2f;                                                // :300-303
                else {                                                                                 // pow(fail_fac, incTry)
                    float pw = 1.f;
                    for (int k = 0; k < lm.incTry; ++k) pw *= oc.lambda_fail_fac;
                    lm.lambda *= pw;
                }
                propose = true;
            }
        }
    }
    if (propose && !done) {
        if (lm.iteration >= oc.max_its_per_lvl[lvl]) done = true;
        else if (oc.max_lm_tries > 0 && lm.tries >= oc.max_lm_tries) done = true;
    }
    if (propose && !done) {
        // solve (A/n with diag *(1+lambda)) inc = (sum w r v)/n     :258-262
        solve6(lm.A, lm.b, __drcp_rn(lm.n), (double)(1.f + lm.lambda), lm.inc);
        lm.incTry++; lm.tries++;
        double qe[4], te3[3];
        se3_exp(lm.inc, qe, te3);
        se3_mul(qe, te3, lm.q, lm.t, lm.qn, lm.tn);              // :266 exp(inc) * referenceToFrame
        double Rn[9];
        quat_to_R(lm.qn, Rn);
        for (int i = 0; i < 9; ++i) R_out[i] = (float)Rn[i];
        for (int i = 0; i < 3; ++i) t_out[i] = (float)lm.tn[i];
    }
    if (done) {
        // next level (or the result) starts from the accepted pose      :308-309
        double Ra[9];
        quat_to_R(lm.q, Ra);
        for (int i = 0; i < 9; ++i) R_out[i] = (float)Ra[i];
        for (int i = 0; i < 3; ++i) t_out[i] = (float)lm.t[i];
    }
    return done;
}

// ---- per-point work: PASS A + PASS B fused ---------------------------------------
// One 256-bit load (LDG.E.ENL2.256 on sm_100a) of the 32-byte QUAD record of pixel (ix,iy): the four distance-transform
// values and the four packed gradients the bilinear fetch of optimizer.h:173-185 needs.  Returned as the two row
// records r0 = {dt(x,y), dt(x+1,y), g(x,y), g(x+1,y)}, r1 = the same for row y+1.  One gather and one address per point
// instead of two (or four texel fetches); on its own this measured neutral -- the gather phase is bound neither by L1
// wavefronts nor by per-thread memory parallelism (profiles/r1_k_track_v6_hotspots.txt) -- but it is the cheapest fetch.

// snorm16 pair -> floats (scale folded in by the caller)
__device__ __forceinline__ void unpack_grad(uint32_t g, float &gx, float &gy)
{
    gx = (float)(short)(g & 0xffffu);
    gy = (float)((int)g >> 16);
}

// ---- branch-free per-point work (all engines) -------------------------------------------------------------------
// optimizer.cpp:93-131 + calculateWarpUpdate (:204-228) + LGS6::update (LGSX.h:392-398).  A point that does not exist,
// projects out of bounds or fails the edge filter runs through the same straight-line code with weight 0 (its texel fetch is redirected to texel 0 and
// its projection is zeroed so that no inf/NaN can reach the sums).  Straight-line code lets the compiler interleave
// the arithmetic of one point with the address computation and gathers of the next, and no lane ever waits for a
// divergent neighbour.  The two divisions are single MUFU.RCP (<= 1 ulp, far inside the float tolerance of the path).

struct ProjB {
    float a, b, iz, dx, dy;   // a = Wx/Wz, b = Wy/Wz (0 when invalid)
    const uint4 *bp;
    bool exists, valid;
};

struct LevelConst {           // per-level constants of an evaluation, kept in registers
    float fx, fy, cx, cy, umax, vmax;
    int w;
    const uint4 *opt;
};

__device__ __forceinline__ ProjB project_b(bool exists, const float4 p, const LevelConst &L, const float *__restrict__ R,
                                           const float *__restrict__ t)
{
    ProjB o;
    const float Wx = R[0] * p.x + R[3] * p.y + R[6] * p.z + t[0];
    const float Wy = R[1] * p.x + R[4] * p.y + R[7] * p.z + t[1];
    const float Wz = R[2] * p.x + R[5] * p.y + R[8] * p.z + t[2];
    const float iz = rcp_approx(Wz);
    const float a = Wx * iz, b = Wy * iz;
    const float u = a * L.fx + L.cx;
    const float v = b * L.fy + L.cy;
    const bool inb = (u > 1.f && v > 1.f && u < L.umax && v < L.vmax);   // NaN-safe (optimizer.cpp:100)
    o.exists = exists;
    o.valid = exists && inb;
    const int ix = o.valid ? (int)u : 0, iy = o.valid ? (int)v : 0;
    o.dx = o.valid ? u - (float)ix : 0.f;
    o.dy = o.valid ? v - (float)iy : 0.f;
    o.a = o.valid ? a : 0.f;
    o.b = o.valid ? b : 0.f;
    o.iz = o.valid ? iz : 0.f;
    o.bp = L.opt + 2u * (unsigned)(iy * L.w + ix);
    return o;
}

__device__ __forceinline__ void finish_point_b(const ProjB &P, const uint4 r0, const uint4 r1, const LevelConst &L, float edge_dist,
                                               bool use_filter, float huber, float (&acc)[32])
{
    // getInterpolatedElement43, optimizer.h:173-185
    const float dxdy = P.dx * P.dy;
    const float w11 = dxdy, w01 = P.dy - dxdy, w10 = P.dx - dxdy, w00 = 1.f - P.dx - P.dy + dxdy;
    float gx00, gy00, gx10, gy10, gx01, gy01, gx11, gy11;
    unpack_grad(r0.z, gx00, gy00); unpack_grad(r0.w, gx10, gy10);
    unpack_grad(r1.z, gx01, gy01); unpack_grad(r1.w, gx11, gy11);
    constexpr float kq = 1.0f / 32764.0f;
    const float gx = (w11 * gx11 + w01 * gx01 + w10 * gx10 + w00 * gx00) * (kq * L.fx);   // optimizer.cpp:119
    const float gy = (w11 * gy11 + w01 * gy01 + w10 * gy10 + w00 * gy00) * (kq * L.fy);   // optimizer.cpp:120
    const float r = w11 * __uint_as_float(r1.y) + w01 * __uint_as_float(r1.x) + w10 * __uint_as_float(r0.y) + w00 * __uint_as_float(r0.x);
    const bool pass = P.valid && !(use_filter && r > edge_dist);                   // optimizer.cpp:100,112
    const float hub = huber * rcp_approx(fmaxf(r, huber));                          // optimizer.h:159: r <= huber ? 1 : huber / r
    const float wr = pass ? ((r <= huber) ? 1.f : hub) : 0.f;
    const float rs = pass ? r : 0.f;
    acc[kRecGood] += pass ? 1.f : 0.f;
    acc[kRecBad] += (P.exists && !pass) ? 1.f : 0.f;
    // calculateWarpUpdate, optimizer.cpp:204-228, factored through a = x/z, b = y/z, s = a gx + b gy
    const float z = P.iz, a = P.a, b = P.b;
    const float s = a * gx + b * gy;
    float J[6];
    J[0] = z * gx;
    J[1] = z * gy;
    J[2] = -(s * z);
    J[3] = -(b * s + gy);
    J[4] = a * s + gx;
    J[5] = a * gy - b * gx;
    // LGS6::update, LGSX.h:392-398 (upper triangle only; A is symmetric)
    int k = 0;

    for (int i = 0; i < 6; ++i) {
        const float wi = wr * J[i];

        for (int j = i; j < 6; ++j) acc[k++] += wi * J[j];
    }
    const float rw = rs * wr;

    for (int i = 0; i < 6; ++i) acc[kRecB + i] += rw * J[i];
    acc[kRecSW] += rw * rs;     // optimizer.cpp:131
    acc[kRecSU] += rs * rs;
}

// evalCostFunction (tracker.cpp:357-393) for one pose
__device__ __forceinline__ float cost_point(float X, float Y, float Z, const LevelIn &L, const float *__restrict__ dt, float edge_dist,
                                            bool use_filter)
{
    const float nx = L.fx * X / Z + L.cx;    // tracker.cpp:378-379
    const float ny = L.fy * Y / Z + L.cy;
    if (nx >= 0.f && nx < (float)L.w && ny >= 0.f && ny < (float)L.h) {
        const float r = __ldg(dt + (size_t)floorf(ny) * L.w + (size_t)floorf(nx));
        if (use_filter && r > edge_dist) return 0.f;
        return r;
    }
    return 0.f;
}

// After the call lane L holds the warp total of v[L].
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane)
{

    for (int half = 16; half >= 1; half >>= 1) {
        const bool hi = (lane & half) != 0;

        for (int i = 0; i < half; ++i) {
            const float send = hi ? v[i] : v[i + half];
            const float keep = hi ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(kFull, send, half);
        }
    }
    return v[0];
}


// ---- mbarrier / st.async PTX (cluster exchange without a cluster-wide fence) ---------------------------------
// 8 bytes into the shared memory of CTA `dst_rank` of this cluster (same offset as `local_ptr`), completing 8 bytes of
// the transaction count of that CTA's mbarrier (same offset as `local_bar`): STAS.64 on sm_100a.

}  // namespace revo

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
    emu::launch_seq(grid, 256, 0, [=]() { k_depth_u16(d_raw, frame_px, scale, d_desc, px); });
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

int launch_gray(revo_ctx *ctx, const uint8_t *d_bgr, size_t stride, int ch, size_t frame_bytes, const ImgLevel *d_desc,
                int n, int w, int h)
{
    dim3 block(32, 8), grid(cdiv(cdiv(w, 4), 32), cdiv(h, 8), n);
    emu::launch_seq(grid, block, 0, [=]() { k_gray(d_bgr, stride, ch, frame_bytes, d_desc, w, h); });
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

        for (int r = 0; r < 5; ++r) {
            const uint8_t *__restrict__ row = in + (size_t)reflect101(2 * y - 2 + r, hs) * ws;
            int b[16];

            for (int i = 2; i <= 12; ++i) b[i] = row[reflect101(c0 + i, ws)];

            for (int k = 0; k < 4; ++k) {
                const int i = 2 * k + 4;
                hsum[r][k] = b[i - 2] + 4 * b[i - 1] + 6 * b[i] + 4 * b[i + 1] + b[i + 2];
            }
        }
    }
    uint32_t o = 0;

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

int launch_pyrdown_depth(revo_ctx *ctx, const ImgLevel *d_src, const ImgLevel *d_dst, int n, int w_dst, int h_dst,
                         int w_src, int h_src)
{
    dim3 block(32, 8), grid(cdiv(cdiv(w_dst, 4), 32), cdiv(h_dst, 8), n);
    emu::launch_seq(grid, block, 0, [=]() { k_pyrdown(d_src, d_dst, w_src, h_src, w_dst, h_dst); });
    LAUNCH_CHECK(ctx);
    emu::launch_seq(grid, block, 0, [=]() { k_depth_half(d_src, d_dst, w_src, w_dst, h_dst); });
    LAUNCH_CHECK(ctx);
    return REVO_OK;
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
            emu::launch_seq(g2, block, 0, [=]() { k_fill_in(d_desc, d_top, w, h, patch, patch_low, n_percentage); });
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
    int (&warp_sums)[32] = *reinterpret_cast<int (*)[32]>(emu::smem_slot(16, sizeof(int[32]), 8));
    int &carry_s = *reinterpret_cast<int *>(emu::smem_slot(17, sizeof(int), 8));
    const int f = blockIdx.x;
    int *off = desc[f].tile_off;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_tiles ? off[i] : 0;
        int s = v;

        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int ws = warp_sums[lane];

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

        for (int c = 0; c < kTileW; ++c)
            if ((e8 >> (8 * c)) & 0xffull) m |= 1u << (r * kTileW + c);
    }
    // depth test only where an edge pixel is (isfinite, dmin < Z < dmax: imgpyramidrgbd.cpp:210-214)
    unsigned keep = 0;
    for (unsigned mm = m; mm;) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        const float Z = L.depth[(size_t)(y0 + (b >> 3)) * w + x0 + (b & 7)];
        if (isfinite(Z) && Z > dmin && Z < dmax) keep |= 1u << b;
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

    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    int o = before + incl - c;
    const int x0 = tx * kTileW, y0 = ty * kTileH;
    for (unsigned mm = keep; mm; ++o) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        if (o >= L.pts_cap) break;
        const int x = x0 + (b & 7), y = y0 + (b >> 3);
        const float Z = L.depth[(size_t)y * w + x];
        const float X = __fdiv_rn(__fmul_rn(Z, __fsub_rn((float)x, L.cx)), L.fx);
        const float Y = __fdiv_rn(__fmul_rn(Z, __fsub_rn((float)y, L.cy)), L.fy);
        L.pts[o] = make_float4(X, Y, Z, 1.0f);
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
        emu::launch(grid, 256, 0, [=]() { k_group_mask(d_desc, w, h, tiles_x, tiles_y, groups_x, dmin, dmax); });
        LAUNCH_CHECK(ctx);
        emu::launch(grid, 256, 0, [=]() { k_group_scatter(d_desc, w, h, tiles_x, tiles_y, groups_x); });
        LAUNCH_CHECK(ctx);
        return REVO_OK;
    }
    dim3 grid(cdiv(n_tiles, 8), 1, n);
    emu::launch(grid, 256, 0, [=]() { k_tile_count(d_desc, w, h, tiles_x, n_tiles, dmin, dmax); });
    LAUNCH_CHECK(ctx);
    emu::launch(n, 1024, 0, [=]() { k_tile_scan(d_desc, n_tiles); });
    LAUNCH_CHECK(ctx);
    emu::launch(grid, 256, 0, [=]() { k_tile_scatter(d_desc, w, h, tiles_x, n_tiles, dmin, dmax); });
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
    emu::launch_seq(cdiv(w, 128), 128, 0, [=]() { k_col_count(d_desc_one, w, h, dmin, dmax, d_col_off); });
    LAUNCH_CHECK(ctx);
    emu::launch_seq(1, 32, 0, [=]() { k_col_scan(d_col_off, w, d_n); });
    LAUNCH_CHECK(ctx);
    emu::launch_seq(cdiv(w, 128), 128, 0, [=]() { k_col_scatter(d_desc_one, w, h, dmin, dmax, d_col_off, d_out); });
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

// ---------------------------------------------------------------------------
// K7: exact Euclidean distance transform to the nearest edge pixel, out = sqrtf(d2).
//  (a) column pass: vertical distance g(x,y) to the nearest edge in the column (int, kEdtInf if none)
//  (b) row pass: d2(x,y) = min_j (x-j)^2 + g(j,y)^2 by an outward search that stops once r^2 >= best
//      (exact; typical DT values are small so the search is short).
// K8: {0.5(dt[i-1]-dt[i+1]), 0.5(dt[i-w]-dt[i+w]), dt[i], 0} for rows 1..h-2, zeros elsewhere.
// ---------------------------------------------------------------------------
constexpr int kEdtInf = 1 << 14;
// value of every pixel when the edge map is empty: what OpenCV's own trueDistTrans returns (cv2 4.13, IPP off)
constexpr float kEdtEmpty = 65536.0f;

__global__ void k_edt_cols(const ImgLevel *__restrict__ desc, int w, int h)
{
    const int f = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w) return;
    const uint8_t *__restrict__ e = desc[f].edges;
    int *__restrict__ g = desc[f].labels;
    int d = kEdtInf;
    for (int y = 0; y < h; ++y) {
        d = e[(size_t)y * w + x] ? 0 : min(d + 1, kEdtInf);
        g[(size_t)y * w + x] = d;
    }
    d = kEdtInf;
    for (int y = h - 1; y >= 0; --y) {
        const int cur = g[(size_t)y * w + x];
        d = min(cur, min(d + 1, kEdtInf));
        if (d < cur) g[(size_t)y * w + x] = d;
    }
}

__global__ void __launch_bounds__(256) k_edt_rows(const ImgLevel *__restrict__ desc, int w, int h)
{
    int *grow = (int *)emu::cta->dyn.data();
    const int f = blockIdx.z, y = blockIdx.x;
    const int *__restrict__ g = desc[f].labels + (size_t)y * w;
    for (int i = threadIdx.x; i < w; i += blockDim.x) grow[i] = g[i];
    __syncthreads();
    float *__restrict__ out = desc[f].dt + (size_t)y * w;
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
        const int g0 = grow[x];
        int best = g0 * g0;
        const int rmax = max(x, w - 1 - x);
        for (int r = 1; r <= rmax && r * r < best; ++r) {
            const int r2 = r * r;
            if (x - r >= 0) { const int gl = grow[x - r]; best = min(best, r2 + gl * gl); }
            if (x + r < w) { const int gr = grow[x + r]; best = min(best, r2 + gr * gr); }
        }
        out[x] = best >= kEdtInf * kEdtInf ? kEdtEmpty : sqrtf((float)best);
    }
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

// K8 (device layout): one 32-byte QUAD record per pixel (x,y) holding everything the bilinear fetch of the tracker
// needs for a point that projects into [x,x+1) x [y,y+1): {dt(x,y), dt(x+1,y), dt(x,y+1), dt(x+1,y+1)} as float32 and
// the snorm16 (gx,gy) of the same four texels.  A residual evaluation then costs ONE 256-bit gather per edge point
// instead of four 16-byte texel fetches; dt stays float32, only the Jacobian direction is quantised (1.5e-5 absolute).
// The reference's float4 array is produced on demand for the accessor (k_opt_struct_f4).
__device__ __forceinline__ void store_quad(uint4 *__restrict__ out, size_t i, const float4 a, const float4 b, const float4 c, const float4 d)
{
    out[2 * i] = make_uint4(__float_as_uint(a.z), __float_as_uint(b.z), __float_as_uint(c.z), __float_as_uint(d.z));
    out[2 * i + 1] = make_uint4(pack_grad(a.x, a.y), pack_grad(b.x, b.y), pack_grad(c.x, c.y), pack_grad(d.x, d.y));
}

__global__ void __launch_bounds__(256) k_opt_struct(const ImgLevel *__restrict__ desc, int w, int h)
{
    const int f = blockIdx.z;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)w * h;
    if (i >= n) return;
    const float *__restrict__ dt = desc[f].dt;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a = opt_texel(dt, i, w, h);
    const float4 b = (i + 1 < n) ? opt_texel(dt, i + 1, w, h) : z;
    const float4 c = (i + w < n) ? opt_texel(dt, i + w, w, h) : z;
    const float4 d = (i + w + 1 < n) ? opt_texel(dt, i + w + 1, w, h) : z;
    store_quad(desc[f].opt, i, a, b, c, d);
}

// the reference layout, for returnOptimizationStructure(): out[i] = {gx, gy, dt, 0}
__global__ void __launch_bounds__(256) k_opt_struct_f4(const float *__restrict__ dt, int w, int h, float4 *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)w * h) return;
    out[i] = opt_texel(dt, i, w, h);
}

// test hook: caller-provided float4 structure -> device quad layout
__global__ void __launch_bounds__(256) k_opt_pack_from_f4(const float4 *__restrict__ in, int w, int h, uint4 *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n = (size_t)w * h;
    if (i >= n) return;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 a = in[i];
    const float4 b = (i + 1 < n) ? in[i + 1] : z;
    const float4 c = (i + w < n) ? in[i + w] : z;
    const float4 d = (i + w + 1 < n) ? in[i + w + 1] : z;
    store_quad(out, i, a, b, c, d);
}

int launch_opt_struct_f4(revo_ctx *ctx, const float *d_dt, int w, int h, float4 *d_out)
{
    emu::launch_seq(cdiv(w * h, 256), 256, 0, [=]() { k_opt_struct_f4(d_dt, w, h, d_out); });
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

int launch_opt_pack_from_f4(revo_ctx *ctx, const float4 *d_in, int w, int h, uint4 *d_out)
{
    emu::launch_seq(cdiv(w * h, 256), 256, 0, [=]() { k_opt_pack_from_f4(d_in, w, h, d_out); });
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
__global__ void __launch_bounds__(256) k_quality_scatter(const QualityArgs a, unsigned *__restrict__ mbits, int *__restrict__ counters)
{
    const QualityFrame &F = a.fr[blockIdx.y];
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

__global__ void __launch_bounds__(256) k_quality_hist(const float *__restrict__ depth, const uint8_t *__restrict__ edges,
                                                      const unsigned *__restrict__ mbits, int n_px, float dmin, float dmax,
                                                      int *__restrict__ counters)
{
    int (&sc)[8] = *reinterpret_cast<int (*)[8]>(emu::smem_slot(18, sizeof(int[8]), 8));
    if (threadIdx.x < 8) sc[threadIdx.x] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_px) {
        const float Z = depth[i];
        if (isfinite(Z) && Z > dmin && Z < dmax) {     // ImgPyramidRGBD::isPointOkDepth
            const int val = __popc((mbits[i >> 2] >> ((i & 3) * 8)) & 0xffu);
            atomicAdd(&sc[val & 3], 1);
            if (edges[i]) atomicAdd(&sc[4 + (val & 3)], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x < 8 && sc[threadIdx.x]) atomicAdd(counters + threadIdx.x, sc[threadIdx.x]);
}

int launch_quality(revo_ctx *ctx, const QualityArgs &a, const float *d_depth, const uint8_t *d_edges, float dmin, float dmax,
                   unsigned *d_mbits, int *d_counters)
{
    const int w = a.w, h = a.h;
    const size_t words = ((size_t)w * h + 3) / 4;
    REVO_CUDA(ctx, cudaMemsetAsync(d_mbits, 0, words * 4, ctx->stream));
    REVO_CUDA(ctx, cudaMemsetAsync(d_counters, 0, 16 * sizeof(int), ctx->stream));
    if (a.n_frames > 0) {
        dim3 grid(64, a.n_frames);
        emu::launch_seq(grid, 256, 0, [=]() { k_quality_scatter(a, d_mbits, d_counters); });
        LAUNCH_CHECK(ctx);
    }
    emu::launch(cdiv(w * h, 256), 256, 0, [=]() { k_quality_hist(d_depth, d_edges, d_mbits, w * h, dmin, dmax, d_counters); });
    LAUNCH_CHECK(ctx);
    return REVO_OK;
}

int launch_keyframe(revo_ctx *ctx, const ImgLevel *d_desc, int n, int w, int h)
{
    {
        dim3 grid(cdiv(w, 64), 1, n);
        emu::launch_seq(grid, 64, 0, [=]() { k_edt_cols(d_desc, w, h); });
        LAUNCH_CHECK(ctx);
    }
    {
        dim3 grid(h, 1, n);
        emu::launch(grid, 256, w * sizeof(int), [=]() { k_edt_rows(d_desc, w, h); });
        LAUNCH_CHECK(ctx);
    }
    {
        dim3 grid(cdiv(w * h, 256), 1, n);
        emu::launch_seq(grid, 256, 0, [=]() { k_opt_struct(d_desc, w, h); });
        LAUNCH_CHECK(ctx);
    }
    return REVO_OK;
}

}  // namespace revo

namespace revo {
// counts -> wrapping u8 histogram + number of non-empty patches (countNonZero(dist), imgpyramidrgbd.cpp:160)
__global__ void __launch_bounds__(256) k_hist_finalize(const ImgLevel *__restrict__ desc)
{
    int (&red)[8] = *reinterpret_cast<int (*)[8]>(emu::smem_slot(19, sizeof(int[8]), 8));
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


__device__ __forceinline__ unsigned rep_byte(unsigned b) { return (b & 0xffu) * 0x01010101u; }

// Requires w % 4 == 0, w >= 8 and a 4-byte aligned gray plane (launch_canny checks).
__global__ void __launch_bounds__(128) k_canny_nms(const ImgLevel *__restrict__ desc, int w, int h, int low, int high, int wp32,
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

    for (int c = 0; c < 6; ++c) cm[c] = (x - 1 + c >= 0 && x - 1 + c < w) ? -1 : 0;
    unsigned own_px = 0;                               // the lane's own pixels inside the image (4 bits)

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
// rows, and a row that changed marks its neighbours (across band boundaries through a per-warp word in shared memory,
// taken at the next block barrier).  The first round costs about two floods per row, every later round only touches
// the handful of rows a weak chain is still creeping along.  wp = row pitch in words of type W.
template <typename W>
__global__ void __launch_bounds__(1024) k_canny_hyst_smem(const ImgLevel *__restrict__ desc, int w, int h, int wp)
{
    unsigned long long *hs_mem = (unsigned long long *)emu::cta->dyn.data();
    unsigned (&incoming)[2][32] = *reinterpret_cast<unsigned (*)[2][32]>(emu::smem_slot(20, sizeof(unsigned[2][32]), 8));
    const ImgLevel &L = desc[blockIdx.x];
    const size_t nw = (size_t)wp * h;
    W *gC = (W *)L.labels, *gS = gC + nw;
    W *C = (W *)hs_mem, *S = C + nw;
    for (size_t i = threadIdx.x; i < nw; i += blockDim.x) { C[i] = gC[i]; S[i] = gS[i]; }
    if (threadIdx.x < 64) incoming[threadIdx.x >> 5][threadIdx.x & 31] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int R = (h + n_warps - 1) / n_warps;          // <= 32 (launcher)
    const int y_lo = warp * R, y_hi = min(h, y_lo + R);
    const int n_rows = y_hi > y_lo ? y_hi - y_lo : 0;
    const bool act = lane < wp;
    const int col = act ? lane : 0;
    unsigned dirty = n_rows >= 32 ? 0xffffffffu : ((1u << n_rows) - 1u);
    unsigned fresh = dirty;                               // rows not yet flooded from their own strong bits
    auto row = [&](const W *M, int y) -> W { return act ? M[(size_t)y * wp + col] : (W)0; };
    for (int it = 0;; ++it) {
        const int par = it & 1;
        unsigned inc = 0;
        if (lane == 0) inc = atomicExch(&incoming[par][warp], 0u);
        dirty |= __shfl_sync(0xffffffffu, inc, 0);
        int sent = 0;
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
                if (delta) S[(size_t)y * wp + col] = s2;
                // a neighbour row must be (re)visited only if the new bits hand it a seed it does not have yet
                const W sp = spread_row<W>(delta, lane);
                if (y > 0 && __any_sync(0xffffffffu, (sp & row(C, y - 1) & ~row(S, y - 1)) != 0)) {
                    if (r > 0) dirty |= 1u << (r - 1);
                    else { if (lane == 0) atomicOr(&incoming[par ^ 1][warp - 1], 1u << (R - 1)); sent = 1; }
                }
                if (y + 1 < h && __any_sync(0xffffffffu, (sp & row(C, y + 1) & ~row(S, y + 1)) != 0)) {
                    if (r + 1 < n_rows) dirty |= 1u << (r + 1);
                    else { if (lane == 0) atomicOr(&incoming[par ^ 1][warp + 1], 1u); sent = 1; }
                }
            }
        }
        if (!__syncthreads_or((dirty != 0u) || sent)) break;
    }
    for (size_t i = threadIdx.x; i < nw; i += blockDim.x) gS[i] = S[i];
}

// ---- (3) bit mask -> byte maps + patch counters ------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_canny_expand(const ImgLevel *__restrict__ desc, int w, int h, int wp32, int P)
{
    const int f = blockIdx.z;
    const ImgLevel &L = desc[f];
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (y >= h || x0 >= w) return;
    const unsigned *__restrict__ maskS = (const unsigned *)L.labels + (size_t)wp32 * h;
    const unsigned bits = (__ldcg(maskS + (size_t)y * wp32 + (x0 >> 5)) >> (x0 & 31)) & 0xffffu;
    uint8_t *e = L.edges + (size_t)y * w + x0, *eo = L.edges_orig + (size_t)y * w + x0;
    unsigned o[4];

    for (int q = 0; q < 4; ++q) {
        const unsigned nib = (bits >> (4 * q)) & 15u;
        // 4 bits -> 4 bytes of 0 / 255
        o[q] = ((nib & 1u) * 0xffu) | (((nib >> 1) & 1u) * 0xff00u) | (((nib >> 2) & 1u) * 0xff0000u) | (((nib >> 3) & 1u) * 0xff000000u);
    }
    if (x0 + 16 <= w && ((((uintptr_t)e) & 15) == 0) && ((((uintptr_t)eo) & 15) == 0)) {
        const uint4 v = make_uint4(o[0], o[1], o[2], o[3]);
        *(uint4 *)e = v;
        *(uint4 *)eo = v;
    } else {
        for (int k = 0; k < 16 && x0 + k < w; ++k) {
            const uint8_t v = (bits >> k) & 1u ? 255 : 0;
            e[k] = v; eo[k] = v;
        }
    }
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
    if (hist_w > 0 && hist_h > 0)
        REVO_CUDA(ctx, cudaMemset2DAsync(d_counts0, counts_stride, 0, (size_t)hist_w * hist_h * sizeof(int), (size_t)n, ctx->stream));
    const int wp64 = cdiv(w, 64), wp32 = 2 * wp64;
    {
        // rows per warp strip: long strips amortise the 2-row prologue, short ones keep the small levels parallel
        const int rs = h >= 400 ? NMS_RS : (h >= 200 ? 18 : 9);
        dim3 grid(cdiv(w, 32 * NMS_PX), cdiv(cdiv(h, rs), 4), n);
        emu::launch(grid, 128, 0, [=]() { k_canny_nms(d_desc, w, h, low, high, wp32, rs); });
        LAUNCH_CHECK(ctx);
    }
    {
        int warps = cdiv(h, 8);
        warps = warps < 1 ? 1 : (warps > 32 ? 32 : warps);
        const size_t smem = (size_t)2 * wp64 * 8 * h;
        if (smem <= 200 * 1024 && cdiv(h, warps) <= 32) {
            if (wp32 <= 32) {
                REVO_CUDA(ctx, cudaFuncSetAttribute(k_canny_hyst_smem<unsigned>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                emu::launch(n, warps * 32, smem, [=]() { k_canny_hyst_smem<unsigned>(d_desc, w, h, wp32); });
            } else {
                REVO_CUDA(ctx, cudaFuncSetAttribute(k_canny_hyst_smem<unsigned long long>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    200 * 1024));
                emu::launch(n, warps * 32, smem, [=]() { k_canny_hyst_smem<unsigned long long>(d_desc, w, h, wp64); });
            }
        } else {
            emu::launch(n, warps * 32, 0, [=]() { k_canny_hyst(d_desc, w, h, wp64); });
        }
        LAUNCH_CHECK(ctx);
    }
    {
        dim3 block(32, 8), grid(cdiv(cdiv(w, 16), 32), cdiv(h, 8), n);
        emu::launch_seq(grid, block, 0, [=]() { k_canny_expand(d_desc, w, h, wp32, patch); });
        LAUNCH_CHECK(ctx);
    }
    if (hist_w > 0 && hist_h > 0) {
        emu::launch(n, 256, 0, [=]() { k_hist_finalize(d_desc); });
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
    ctx->last_error = "tile / TMA Canny fallback: not in the emulated build";
    return REVO_ERR_UNSUPPORTED;
}

}  // namespace revo

// track.cu -- K9/K10: the coarse-to-fine Gauss-Newton / Levenberg-Marquardt edge alignment as ONE
// persistent kernel.
//
// Replaces (reference file:line, fabianschenk/REVO):
//   TrackerNew::trackFrames / checkInitializationValues / evalCostFunction   system/tracker.cpp:294-353, 265-283, 357-393
//   Optimizer::trackFrames (LM loop)                                          system/optimizer.cpp:235-311
//   Optimizer::calcErrorAndBuffers (PASS A) + getInterpolatedElement43       system/optimizer.cpp:74-191, optimizer.h:173-185
//   Optimizer::calculateWarpUpdate (PASS B) + LGS6::update/finish             system/optimizer.cpp:192-234, utils/LGSX.h:320-326,392-398
//   Eigen LDLT 6x6 solve, Sophus::SE3f exp / product                          system/optimizer.cpp:258-266
//
// Design (B200): a frame pair is owned by one thread-block CLUSTER (1..16 CTAs); clusters pull pairs from a
// global work counter (persistent kernel).  PASS A and PASS B are fused: every evaluation at a pose warps each
// 3-D edge point, fetches the 4 {gx,gy,dt} texels, forms the residual, Huber weight and 1x6 Jacobian and
// accumulates the 21+6 normal-equation terms + 4 statistics in registers -- the 7 SoA buffers of the reference
// never exist.  Two points are in flight per thread (their 8 texel gathers are issued back to back) to cover the
// dependent pts -> texel latency.  The 32-value record is reduced with a transposing warp-shuffle tree, across
// warps through shared memory, across the CTAs of the cluster through distributed shared memory (one cluster
// barrier per evaluation), and -- when one pair is split over several GPUs -- across GPUs through peer-mapped
// mailboxes over NVLink inside the same kernel.  Every CTA then runs the identical 6x6 LDL^T solve, SE3 update
// and accept/reject test redundantly (bitwise-equal inputs, so no broadcast is needed): all levels and all LM
// iterations of a pair run without a host round trip.  No tensor cores: there is no dense contraction here.


namespace cg = cooperative_groups;

namespace revo {

// ---- mailbox for the multi-GPU split ----------------------------------------
struct Mailbox {
    double data[2][16][32];
    unsigned long long flag[2][16];
};



// ---- the kernel ---------------------------------------------------------------
// Dynamic shared memory: the thread-private cache of the level's 3-D points, float[3][pcap][kThreads] (x, y, z planes):
// thread t keeps the first `pcap` of ITS points of the current level there for all evaluations of the level, so an
// evaluation starts with shared-memory reads instead of an L2 round trip and re-reads no list bytes from L2 / HBM.
template <int kThreads, int kMinBlocks>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_track(const PairDesc *__restrict__ pairs, int n_pairs, const TrackParams prm, revo_track_result *__restrict__ results,
        double *__restrict__ records, revo_trace_entry *__restrict__ trace, int *__restrict__ trace_counts,
        int *__restrict__ work_counter, int pcap)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int kWarps = kThreads / 32;

    float *s_pts = emu::cta->dyn.data();
    float *const sx = s_pts + tid, *const sy = sx + (size_t)pcap * kThreads, *const sz = sy + (size_t)pcap * kThreads;

    float (&warp_part)[kWarps][32] = *reinterpret_cast<float (*)[kWarps][32]>(emu::smem_slot(21, sizeof(float[kWarps][32]), 8));
    double (&cta_part)[2][16][32] = *reinterpret_cast<double (*)[2][16][32]>(emu::smem_slot(22, sizeof(double[2][16][32]), 16));   // [parity][source rank]: partials pushed by the CTAs of the cluster
    double (&total)[2][32] = *reinterpret_cast<double (*)[2][32]>(emu::smem_slot(23, sizeof(double[2][32]), 8));                        // split mode: CTA 0 publishes the cross-GPU total here
    double (&rec)[32] = *reinterpret_cast<double (*)[32]>(emu::smem_slot(24, sizeof(double[32]), 8));
    uint64_t (&xbar)[2] = *reinterpret_cast<uint64_t (*)[2]>(emu::smem_slot(25, sizeof(uint64_t[2]), 8));              // transaction barriers of the partial exchange (one per parity)
    Ctrl &ctrl = *reinterpret_cast<Ctrl *>(emu::smem_slot(26, sizeof(Ctrl), 8));
    LMState &lm = *reinterpret_cast<LMState *>(emu::smem_slot(27, sizeof(LMState), 8));

    const revo_opt_config &oc = prm.cfg.opt;
    const bool use_filter = oc.use_edge_filter != 0;
    const int world = prm.split_world > 1 ? prm.split_world : 1;
    const int n_members = world * C;
    const int member = (world > 1 ? prm.split_rank : 0) * C + crank;
    unsigned seq = 0;   // evaluation counter of this cluster (drives the double buffers)

    if (tid == 0) {
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
    }
    if (C > 1) cluster.sync(); else __syncthreads();

    // ---- reduction of a per-thread accumulator to `rec` (identical in every CTA of the cluster / every rank)
    auto reduce_record = [&](float (&acc)[32]) {
        const float mine = warp_transpose_reduce(acc, lane);
        warp_part[wid][lane] = mine;
        __syncthreads();
        const int par = seq & 1;
        if (world == 1) {
            // Every CTA pushes its 32-double partial into slot [its rank] of every CTA of the cluster (st.async over
            // distributed shared memory, 8 bytes per lane and destination) and waits on its OWN transaction barrier for
            // the C x 256 bytes of this evaluation: one-sided, no cluster barrier, no fence.  Two parities suffice: a CTA
            // can run at most one evaluation ahead of the slowest CTA of its cluster.
            if (wid == 0) {
                double s = 0;

                for (int w = 0; w < kWarps; ++w) s += (double)warp_part[w][lane];
                if (C == 1) {
                    rec[lane] = s;
                } else {
                    if (lane == 0) mbar_expect_tx(&xbar[par], (uint32_t)C * 256u);
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
                    for (int r = 0; r < C; ++r) st_async_b64(&cta_part[par][crank][lane], (unsigned)r, bits, &xbar[par]);
                    mbar_wait(&xbar[par], (seq >> 1) & 1u);
                    double tot = 0;
                    for (int r = 0; r < C; ++r) tot += cta_part[par][r][lane];   // rank order: deterministic
                    rec[lane] = tot;
                }
            }
        } else {
            if (tid < 32) {
                double s = 0;

                for (int w = 0; w < kWarps; ++w) s += (double)warp_part[w][tid];
                cta_part[par][0][tid] = s;
            }
            cluster.sync();
            // cross-GPU exchange: CTA 0 of each rank pushes the rank partial into every rank's mailbox,
            // then waits for all `world` partials of this evaluation and sums them in rank order.
            const unsigned long long fl = prm.split_seq0 + (unsigned long long)seq + 1ull;
            if (crank == 0 && tid < 32) {
                double s = 0;
                for (int r = 0; r < C; ++r) s += *cluster.map_shared_rank(&cta_part[par][0][tid], r);
                for (int g = 0; g < world; ++g) {
                    Mailbox *mb = (Mailbox *)prm.split_peers[g];
                    mb->data[par][prm.split_rank][tid] = s;
                }
                
                __syncwarp();
                if (tid < world) st_release_sys(&((Mailbox *)prm.split_peers[tid])->flag[par][prm.split_rank], fl);
                Mailbox *mine_mb = (Mailbox *)prm.split_peers[prm.split_rank];
                if (tid < world) {
                    // watchdog (~5 s): a peer that never launches must not hang this GPU; the result is then invalid
                    // (ranks disagree), which the caller's cross-rank check catches
                    const long long t0 = clock64();
                    while (ld_acquire_sys(&mine_mb->flag[par][tid]) < fl) {
                        if (clock64() - t0 > 10000000000ll) break;
                    }
                }
                __syncwarp();
                double tot = 0;
                for (int g = 0; g < world; ++g) tot += ((volatile double *)mine_mb->data[par][g])[tid];
                total[par][tid] = tot;
            }
            cluster.sync();
            if (tid < 32) rec[tid] = *cluster.map_shared_rank(&total[par][tid], 0);
        }
        seq++;
        __syncthreads();
    };

    long long prof_gather = 0, prof_reduce = 0, prof_serial = 0, prof_evals = 0;   // thread 0: cycles per phase
    int pair = cluster_id;
    while (pair < n_pairs) {
        const PairDesc &P = pairs[pair];
        const int min_lvl = prm.mode == 0 ? prm.cfg.pyr_min_lvl : prm.level;
        const int max_lvl = prm.mode == 0 ? prm.cfg.pyr_max_lvl : prm.level;
        int evals_lvl[REVO_MAX_LEVELS] = {0, 0, 0, 0, 0, 0};
        int used_identity = 0;
        int ntrace = 0;

        if (tid == 0) {
            for (int i = 0; i < 9; ++i) ctrl.R[i] = P.R[i];
            for (int i = 0; i < 3; ++i) ctrl.t[i] = P.t[i];
            ctrl.pair_skip = rotation_ok(P.R) ? 0 : 1;
            ctrl.level_done = 0;
        }
        __syncthreads();
        const bool skip = ctrl.pair_skip != 0;
        if (skip) {
            if (crank == 0 && tid == 0) {
                revo_track_result &o = results[pair];
                for (int i = 0; i < 9; ++i) o.R[i] = P.R[i];
                for (int i = 0; i < 3; ++i) o.t[i] = P.t[i];
                o.error = INFINITY;
                o.status = REVO_TRACKER_STATE_UNKNOWN;
                o.rc = REVO_ERR_NOT_ORTHOGONAL;
                o.res.good_pts_edges = o.res.bad_pts_edges = 0;
                o.res.sum_error_unweighted = o.res.sum_error_weighted = 0.f;
                for (int l = 0; l < REVO_MAX_LEVELS; ++l) { o.n_evals[l] = 0; o.n_pts[l] = 0; }
                o.used_identity_init = 0;
                if (trace_counts) trace_counts[pair] = 0;
            }
        } else {
            // ---- checkInitializationValues (tracker.cpp:265-283): cost at identity vs cost at (R,t), coarsest level
            if (prm.mode == 0 && prm.cfg.check_init_values) {
                const LevelIn L = P.lvl[min_lvl];
                const int n = *L.n_pts;
                const int lo = (int)((long long)n * member / n_members), hi = (int)((long long)n * (member + 1) / n_members);
                float acc[32];

                for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                const float ed = oc.edge_distance_lvl[min_lvl];
                float R[9], t[3];

                for (int i = 0; i < 9; ++i) R[i] = ctrl.R[i];

                for (int i = 0; i < 3; ++i) t[i] = ctrl.t[i];
                for (int i = lo + tid; i < hi; i += kThreads) {
                    const float4 p = __ldg(L.pts + i);
                    acc[0] += cost_point(p.x, p.y, p.z, L, P.ref_dt_min, ed, use_filter);
                    const float X = R[0] * p.x + R[3] * p.y + R[6] * p.z + t[0];
                    const float Y = R[1] * p.x + R[4] * p.y + R[7] * p.z + t[1];
                    const float Z = R[2] * p.x + R[5] * p.y + R[8] * p.z + t[2];
                    acc[1] += cost_point(X, Y, Z, L, P.ref_dt_min, ed, use_filter);
                }
                reduce_record(acc);
                if (tid == 0) {
                    if ((float)rec[0] < (float)rec[1]) {   // tracker.cpp:277
                        for (int i = 0; i < 9; ++i) ctrl.R[i] = (i % 4 == 0) ? 1.f : 0.f;
                        for (int i = 0; i < 3; ++i) ctrl.t[i] = 0.f;
                        ctrl.pair_skip = 2;   // marker: identity init used
                    }
                }
                __syncthreads();
                used_identity = ctrl.pair_skip == 2;
                __syncthreads();
            }

            if (tid == 0) {
                quat_from_R(ctrl.R, lm.q);
                for (int i = 0; i < 3; ++i) lm.t[i] = ctrl.t[i];
                lm.last_residual = INFINITY;
            }
            float last_good = 0.f, last_bad = 0.f, last_sw = 0.f, last_su = 0.f;

            for (int lvl = min_lvl; lvl >= max_lvl; --lvl) {
                const LevelIn Lin = P.lvl[lvl];
                const int n = *Lin.n_pts;
                // block-cyclic split of the list over the CTAs of the cluster (and the ranks of a GPU split): member m takes
                // the blocks m, m + M, m + 2M, ... of kThreads points -- balanced (the exchange waits for the slowest CTA)
                // and the cluster as a whole still sweeps the tile-major list front to back
                const int stride = n_members * kThreads;
                const int first_idx = member * kThreads + tid;
                const int n_iter = (n + stride - 1) / stride;          // uniform over the cluster
                const int n_cached = n_iter < pcap ? n_iter : pcap;
                const float4 *__restrict__ pts = Lin.pts;
                LevelConst L;
                L.fx = Lin.fx; L.fy = Lin.fy; L.cx = Lin.cx; L.cy = Lin.cy;
                L.umax = (float)(Lin.w - 2); L.vmax = (float)(Lin.h - 2); L.w = Lin.w; L.opt = Lin.opt;
                const float ed = oc.edge_distance_lvl[lvl];
                const float huber = oc.huber_edge;
                // this thread's points of the level -> its private columns of the shared-memory cache
                for (int k = 0; k < n_cached; ++k) {
                    const int i = first_idx + k * stride;
                    const float4 p = i < n ? __ldg(pts + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    sx[k * kThreads] = p.x; sy[k * kThreads] = p.y; sz[k * kThreads] = p.z;
                }
                auto fetch = [&](int k, bool &exists) -> float4 {
                    const int i = first_idx + k * stride;
                    exists = i < n;
                    if (k < n_cached) return make_float4(sx[k * kThreads], sy[k * kThreads], sz[k * kThreads], 1.f);
                    return __ldg(pts + (exists ? i : 0));
                };
                bool first = true;
                __syncthreads();
                while (true) {
                    float R[9], t[3];

                    for (int i = 0; i < 9; ++i) R[i] = ctrl.R[i];

                    for (int i = 0; i < 3; ++i) t[i] = ctrl.t[i];
                    float acc[32];

                    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                    const long long c_begin = prm.profile ? clock64() : 0;
                    // Software pipeline, two register sets (A/B): while point k is being finished the 256-bit gather of
                    // point k+1 is in flight.  All loop conditions are uniform over the CTA; points that do not exist,
                    // project out of bounds or fail the edge filter run the same straight-line code with weight 0.
                    if (n_iter > 0) {
                        bool eA, eB;
                        float4 p = fetch(0, eA);
                        ProjB A = project_b(eA, p, L, R, t), B;
                        uint4 a0, a1, b0, b1;
                        ldg_quad(A.bp, a0, a1);
                        int k = 0;
                        while (true) {
                            const bool hasB = k + 1 < n_iter;
                            if (hasB) {
                                p = fetch(k + 1, eB);
                                B = project_b(eB, p, L, R, t);
                                ldg_quad(B.bp, b0, b1);
                            }
                            finish_point_b(A, a0, a1, L, ed, use_filter, huber, acc);
                            if (!hasB) break;
                            const bool hasA = k + 2 < n_iter;
                            if (hasA) {
                                p = fetch(k + 2, eA);
                                A = project_b(eA, p, L, R, t);
                                ldg_quad(A.bp, a0, a1);
                            }
                            finish_point_b(B, b0, b1, L, ed, use_filter, huber, acc);
                            if (!hasA) break;
                            k += 2;
                        }
                    }
                    const long long c_gather = prm.profile ? clock64() : 0;
                    reduce_record(acc);
                    const long long c_reduce = prm.profile ? clock64() : 0;
                    evals_lvl[lvl]++;
                    last_good = (float)rec[kRecGood]; last_bad = (float)rec[kRecBad];
                    last_sw = (float)rec[kRecSW]; last_su = (float)rec[kRecSU];

                    if (prm.mode == 2) {   // single evaluation: export the record
                        if (crank == 0 && tid < 32 && records) records[(size_t)pair * 32 + tid] = rec[tid];
                        break;
                    }

                    if (tid == 0) {
                        // Optimizer::trackFrames LM logic, optimizer.cpp:243-306 (track_common.cuh: lm_step)
                        revo_trace_entry te;
                        bool traced;
                        const bool done = lm_step(lm, rec, oc, lvl, first, ctrl.R, ctrl.t, &te, &traced);
                        if (traced) {
                            if (trace && crank == 0 && ntrace < prm.trace_cap) trace[(size_t)pair * prm.trace_cap + ntrace] = te;
                            ntrace++;
                        }
                        ctrl.level_done = done ? 1 : 0;
                    }
                    first = false;
                    __syncthreads();
                    if (prm.profile && tid == 0) {
                        const long long c_end = clock64();
                        prof_gather += c_gather - c_begin; prof_reduce += c_reduce - c_gather; prof_serial += c_end - c_reduce;
                        prof_evals++;
                    }
                    if (ctrl.level_done) break;
                }
                __syncthreads();
            }

            if (crank == 0 && tid == 0 && prm.mode != 2) {
                revo_track_result &o = results[pair];
                for (int i = 0; i < 9; ++i) o.R[i] = ctrl.R[i];
                for (int i = 0; i < 3; ++i) o.t[i] = ctrl.t[i];
                o.error = lm.last_residual;
                o.res.good_pts_edges = (int)last_good;
                o.res.bad_pts_edges = (int)last_bad;
                o.res.sum_error_weighted = last_sw;
                o.res.sum_error_unweighted = last_su;
                // tracker.cpp:351: good/bad < 4 -> NEW_KF (double division; bad == 0 -> inf -> OK)
                o.status = ((double)last_good / (double)last_bad < 4.0) ? REVO_TRACKER_STATE_NEW_KF : REVO_TRACKER_STATE_OK;
                o.rc = REVO_OK;
                for (int l = 0; l < REVO_MAX_LEVELS; ++l) {
                    o.n_evals[l] = evals_lvl[l];
                    o.n_pts[l] = (l >= max_lvl && l <= min_lvl) ? *P.lvl[l].n_pts : 0;
                }
                o.used_identity_init = used_identity;
                if (trace_counts) trace_counts[pair] = ntrace < prm.trace_cap ? ntrace : prm.trace_cap;
            }
        }
        // ---- next pair from the global work counter (cluster rank 0 fetches, everybody reads it over DSMEM)
        __syncthreads();
        if (crank == 0 && tid == 0) ctrl.next_pair = n_clusters + atomicAdd(work_counter, 1);
        if (C > 1) cluster.sync(); else __syncthreads();
        pair = *cluster.map_shared_rank(&ctrl.next_pair, 0);
        if (C > 1) cluster.sync(); else __syncthreads();
    }
    if (prm.profile && tid == 0 && crank == 0) {   // phase cycle counters behind the work counter (read back when REVO_TRACK_PROF is set)
        unsigned long long *prof = (unsigned long long *)(work_counter + 2);
        atomicAdd(prof + 0, (unsigned long long)prof_gather);
        atomicAdd(prof + 1, (unsigned long long)prof_reduce);
        atomicAdd(prof + 2, (unsigned long long)prof_serial);
        atomicAdd(prof + 3, (unsigned long long)prof_evals);
    }
    if (C > 1 || world > 1) cluster.sync();   // nobody may exit while a peer can still write into its shared memory
}

// Descriptor upload without the copy engine: the host writes the pair descriptors into pinned, device-mapped memory and
// this kernel pulls them into the workspace.  A cudaMemcpyAsync would queue behind the multi-hundred-megabyte frame
// uploads of the next batches on the same H2D engine and stall the tracker for milliseconds.
__global__ void __launch_bounds__(256) k_stage_in(const uint4 *__restrict__ src_mapped_host, uint4 *__restrict__ dst, size_t n16)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src_mapped_host[i];
}

int launch_stage_in(revo_ctx *ctx, const void *src_mapped_host, void *dst, size_t bytes)
{
    const size_t n16 = (bytes + 15) / 16;
    const int blocks = (int)((n16 + 255) / 256 < 64 ? (n16 + 255) / 256 : 64);
    emu::launch_seq(blocks < 1 ? 1 : blocks, 256, 0, [=]() { k_stage_in((const uint4 *)src_mapped_host, (uint4 *)dst, n16); });
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "k_stage_in launch");
    ctx->launches++;
    return REVO_OK;
}

// ---- launcher -------------------------------------------------------------------
template <int kThreads, int kMinBlocks>
static int launch_track_t(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, int ctas_per_pair,
                          revo_track_result *d_results, double *d_records, revo_trace_entry *d_trace, int *d_trace_counts,
                          int *d_work_counter)
{
    auto kern = k_track<kThreads, kMinBlocks>;
    if (ctas_per_pair > 8) REVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // points per thread cached in shared memory: ~half of the SM's shared memory over the resident CTAs (the rest stays L1)
    const int env_pcap = getenv("REVO_TRACK_PCAP") ? atoi(getenv("REVO_TRACK_PCAP")) : -1;
    int pcap = env_pcap >= 0 ? env_pcap : (int)((112 * 1024 / kMinBlocks) / (12 * kThreads));
    if (pcap > 64) pcap = 64;
    const size_t dyn = (size_t)pcap * kThreads * 12;
    REVO_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = ctas_per_pair;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim_ = dim3(kThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = ctx->stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent: as many clusters as can be co-resident, never more than there are pairs
    cfg.gridDim_ = dim3(ctas_per_pair);
    int max_clusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
    if (e != cudaSuccess || max_clusters < 1) {
        (void)cudaGetLastError();
        max_clusters = ctx->prop.multiProcessorCount / ctas_per_pair;
        if (max_clusters < 1) max_clusters = 1;
    }
    // optional cap on the number of pairs in flight (their lookup structures should stay L2-resident)
    const int env_maxc = getenv("REVO_TRACK_MAX_CLUSTERS") ? atoi(getenv("REVO_TRACK_MAX_CLUSTERS")) : 0;
    if (env_maxc > 0 && max_clusters > env_maxc) max_clusters = env_maxc;
    const int n_clusters = n_pairs < max_clusters ? n_pairs : max_clusters;
    cfg.gridDim_ = dim3(n_clusters * ctas_per_pair);
    REVO_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, d_pairs, n_pairs, prm, d_results, d_records, d_trace, d_trace_counts,
                                      d_work_counter, pcap));
    ctx->launches++;
    return REVO_OK;
}

int launch_track(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                 double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, int *d_work_counter)
{
    if (n_pairs <= 0) return REVO_OK;
    // Default shape (measured on B200, scratch/track_bench.py): clusters of 8 CTAs; 128-thread CTAs (4 per SM, so that an
    // SM interleaves four different pairs) once more than one wave of 256-thread clusters would be needed.
    const int slots256 = 2 * ctx->prop.multiProcessorCount;
    int C = ctx->track_ctas_per_pair > 0 ? ctx->track_ctas_per_pair : 8;
    const int T = ctx->track_threads > 0 ? ctx->track_threads : ((long long)n_pairs * C > slots256 ? 128 : 256);
#define REVO_TRACK_ARGS ctx, d_pairs, n_pairs, prm, C, d_results, d_records, d_trace, d_trace_counts, d_work_counter
    switch (T) {
        case 128: return launch_track_t<128, 4>(REVO_TRACK_ARGS);
        case 512: return launch_track_t<512, 1>(REVO_TRACK_ARGS);
        case 1024: return launch_track_t<1024, 1>(REVO_TRACK_ARGS);
        default: return launch_track_t<256, 2>(REVO_TRACK_ARGS);
    }
#undef REVO_TRACK_ARGS
}

}  // namespace revo
// capi.cu -- implementation of the C ABI declared in include/revo_b200.h.
// Host-side orchestration only: memory layout in HBM, descriptor tables, launch sequencing.



using namespace revo;

namespace revo {

int cuda_fail(revo_ctx *ctx, cudaError_t e, const char *what)
{
    if (ctx) {
        char buf[512];
        snprintf(buf, sizeof(buf), "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
        ctx->last_error = buf;
    }
    (void)cudaGetLastError();
    return REVO_ERR_CUDA;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static bool is_device_ptr(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

static int ensure_scratch(revo_ctx *ctx, size_t bytes)
{
    if (ctx->scratch_bytes >= bytes) return REVO_OK;
    if (ctx->scratch) {
        REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        REVO_CUDA(ctx, cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    bytes = align_up(bytes + bytes / 4, 1 << 20);
    REVO_CUDA(ctx, cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
    return REVO_OK;
}

static int ensure_pinned(revo_ctx *ctx, size_t bytes)
{
    if (ctx->pinned_bytes >= bytes) return REVO_OK;
    if (ctx->pinned) {
        REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        REVO_CUDA(ctx, cudaFreeHost(ctx->pinned));
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
    }
    bytes = align_up(bytes * 2, 1 << 16);
    REVO_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
    ctx->pinned_bytes = bytes;
    return REVO_OK;
}

// Camera(fx,fy,cx,cy,w,h,scale) -- camerapyr.h:98-103 with scale = 1.0f/pow(2,lvl) (:141)
static void level_camera(const revo_camera &c0, int lvl, revo_camera *out)
{
    if (lvl == 0) { *out = c0; return; }
    const float scale = 1.0f / (float)pow(2.0, (double)lvl);
    out->fx = c0.fx * scale; out->fy = c0.fy * scale; out->cx = c0.cx * scale; out->cy = c0.cy * scale;
    out->width = (int32_t)((float)c0.width * scale);
    out->height = (int32_t)((float)c0.height * scale);
}

}  // namespace revo

// ---------------------------------------------------------------------------------------------------
// defaults
// ---------------------------------------------------------------------------------------------------
extern "C" {

void revo_pyr_config_default(revo_pyr_config *c)
{
    c->n_levels = 3; c->canny_threshold1 = 150; c->canny_threshold2 = 100;
    c->depth_min = 0.1f; c->depth_max = 5.2f; c->use_edge_hist = 1; c->n_percentage = 0.3f; c->patch0 = 20;
}

void revo_opt_config_default(revo_opt_config *c)
{
    const float ed[6] = {30, 20, 10, 5, 5, 5};
    c->lambda_success_fac = 0.5f; c->lambda_fail_fac = 2.0f;
    for (int l = 0; l < REVO_MAX_LEVELS; ++l) {
        c->lambda_initial[l] = 0.f; c->step_size_min[l] = 1e-16f; c->convergence_eps[l] = 0.999f;
        c->max_its_per_lvl[l] = 100; c->edge_distance_lvl[l] = ed[l];
    }
    c->huber_edge = 0.3f; c->use_edge_filter = 1; c->max_lm_tries = 0;
}

void revo_tracker_config_default(revo_tracker_config *c)
{
    c->check_init_values = 1; c->pyr_min_lvl = 2; c->pyr_max_lvl = 0;
    revo_opt_config_default(&c->opt);
}

const char *revo_strerror(int code)
{
    switch (code) {
        case REVO_OK: return "ok";
        case REVO_ERR_INVALID_ARG: return "invalid argument";
        case REVO_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
        case REVO_ERR_CUDA: return "CUDA error (see revo_last_error)";
        case REVO_ERR_NOT_KEYFRAME: return "optimization structure not built (makeKeyframe was not called)";
        case REVO_ERR_NOT_ORTHOGONAL: return "R is not a rotation matrix";
        case REVO_ERR_BAD_LEVEL: return "pyramid level out of range";
        case REVO_ERR_BUFFER_TOO_SMALL: return "destination buffer too small";
        case REVO_ERR_UNSUPPORTED: return "unsupported configuration";
        case REVO_ERR_COMM: return "multi-GPU setup error";
        default: return "unknown error";
    }
}

// ---------------------------------------------------------------------------------------------------
// pose helpers (host arithmetic)
// ---------------------------------------------------------------------------------------------------
int revo_quat_to_R9(const float *q, float *R)
{
    if (!q || !R) return REVO_ERR_INVALID_ARG;
    const double n2 = (double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2] + (double)q[3] * q[3];
    if (!(n2 > 0.0) || !std::isfinite(n2)) return REVO_ERR_INVALID_ARG;
    const double s = 1.0 / std::sqrt(n2);
    const double x = q[0] * s, y = q[1] * s, z = q[2] * s, w = q[3] * s;
    R[0] = (float)(1 - 2 * (y * y + z * z)); R[3] = (float)(2 * (x * y - z * w));     R[6] = (float)(2 * (x * z + y * w));
    R[1] = (float)(2 * (x * y + z * w));     R[4] = (float)(1 - 2 * (x * x + z * z)); R[7] = (float)(2 * (y * z - x * w));
    R[2] = (float)(2 * (x * z - y * w));     R[5] = (float)(2 * (y * z + x * w));     R[8] = (float)(1 - 2 * (x * x + y * y));
    return REVO_OK;
}

int revo_R9_to_quat(const float *Rf, float *q)
{
    if (!Rf || !q) return REVO_ERR_INVALID_ARG;
    double R[9];
    for (int i = 0; i < 9; ++i) R[i] = Rf[i];
    auto M = [&](int i, int j) { return R[j * 3 + i]; };
    double n2 = 0;   // ||R^T R - I||_F^2
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double d = -(i == j ? 1.0 : 0.0);
            for (int k = 0; k < 3; ++k) d += M(k, i) * M(k, j);
            n2 += d * d;
        }
    const double det = M(0, 0) * (M(1, 1) * M(2, 2) - M(1, 2) * M(2, 1)) - M(0, 1) * (M(1, 0) * M(2, 2) - M(1, 2) * M(2, 0)) +
                       M(0, 2) * (M(1, 0) * M(2, 1) - M(1, 1) * M(2, 0));
    if (!(std::sqrt(n2) < 1e-5) || !(det > 0)) return REVO_ERR_NOT_ORTHOGONAL;
    double o[4];
    double t = M(0, 0) + M(1, 1) + M(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        o[3] = 0.5 * t;
        t = 0.5 / t;
        o[0] = (M(2, 1) - M(1, 2)) * t;
        o[1] = (M(0, 2) - M(2, 0)) * t;
        o[2] = (M(1, 0) - M(0, 1)) * t;
    } else {
        int i = 0;
        if (M(1, 1) > M(0, 0)) i = 1;
        if (M(2, 2) > M(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(M(i, i) - M(j, j) - M(k, k) + 1.0);
        o[i] = 0.5 * t;
        t = 0.5 / t;
        o[3] = (M(k, j) - M(j, k)) * t;
        o[j] = (M(j, i) + M(i, j)) * t;
        o[k] = (M(k, i) + M(i, k)) * t;
    }
    for (int c = 0; c < 4; ++c) q[c] = (float)o[c];
    return REVO_OK;
}

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
int revo_ctx_create(int device, revo_ctx **out)
{
    if (!out) return REVO_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        (void)cudaGetLastError();
        return REVO_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) return REVO_ERR_INVALID_ARG;
    revo_ctx *ctx = new (std::nothrow) revo_ctx();
    if (!ctx) return REVO_ERR_INVALID_ARG;
    ctx->device = device;
    ctx->launches = 0;
    ctx->scratch = nullptr; ctx->scratch_bytes = 0; ctx->pinned = nullptr; ctx->pinned_bytes = 0; ctx->pinned_kf = nullptr; ctx->pinned_kf_bytes = 0; ctx->pinned_kf_busy = false;
    for (int i = 0; i < 2; ++i) { ctx->stage[i] = nullptr; ctx->stage_bytes[i] = 0; ctx->stage_used[i] = false; }
    ctx->stage_next = 0;
    ctx->track_ctas_per_pair = 0; ctx->track_threads = 0; ctx->track_engine = 0; ctx->track_chunk_points = 0;
    for (auto &v : ctx->ev_valid) v = false;
    ctx->split_rank = 0; ctx->split_world = 1; ctx->split_local = nullptr; ctx->split_seq = 0;
    for (auto &p : ctx->split_peers) p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        (void)cudaGetLastError();
        delete ctx;
        return REVO_ERR_CUDA;
    }
    // keep freed stream-ordered allocations cached in the pool (no trimming at synchronisation points)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    for (auto &e : ctx->ev) cudaEventCreate(&e);
    cudaEventCreateWithFlags(&ctx->pinned_kf_read, cudaEventDisableTiming);
    for (auto &e : ctx->stage_consumed) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    (void)cudaGetLastError();
    *out = ctx;
    return REVO_OK;
}

int revo_ctx_destroy(revo_ctx *ctx)
{
    if (!ctx) return REVO_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 16; ++i)
        if (ctx->split_peers[i] && ctx->split_peers[i] != ctx->split_local) cudaIpcCloseMemHandle(ctx->split_peers[i]);
    if (ctx->split_local) cudaFree(ctx->split_local);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pinned_kf) cudaFreeHost(ctx->pinned_kf);
    cudaEventDestroy(ctx->pinned_kf_read);
    for (int i = 0; i < 2; ++i) { if (ctx->stage[i]) cudaFree(ctx->stage[i]); cudaEventDestroy(ctx->stage_consumed[i]); }
    for (auto &e : ctx->ev) cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    (void)cudaGetLastError();
    delete ctx;
    return REVO_OK;
}

int revo_ctx_synchronize(revo_ctx *ctx)
{
    if (!ctx) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

const char *revo_last_error(revo_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t revo_ctx_stream(revo_ctx *ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }
uint64_t revo_ctx_launch_count(revo_ctx *ctx) { return ctx ? ctx->launches : 0; }

int revo_ctx_last_timings(revo_ctx *ctx, float *pyramid_ms, float *keyframe_ms, float *track_kernel_ms)
{
    if (!ctx) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float *out[3] = {pyramid_ms, keyframe_ms, track_kernel_ms};
    for (int i = 0; i < 3; ++i) {
        if (!out[i]) continue;
        *out[i] = 0.f;
        if (ctx->ev_valid[i]) REVO_CUDA(ctx, cudaEventElapsedTime(out[i], ctx->ev[2 * i], ctx->ev[2 * i + 1]));
    }
    return REVO_OK;
}

int revo_ctx_last_upload_ms(revo_ctx *ctx, float *upload_ms)
{
    if (!ctx || !upload_ms) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    *upload_ms = 0.f;
    if (ctx->ev_valid[3]) REVO_CUDA(ctx, cudaEventElapsedTime(upload_ms, ctx->ev[6], ctx->ev[7]));
    return REVO_OK;
}

int revo_ctx_set_track_shape(revo_ctx *ctx, int ctas_per_pair, int threads_per_cta)
{
    if (!ctx) return REVO_ERR_INVALID_ARG;
    if (ctas_per_pair != 0 && ctas_per_pair != 1 && ctas_per_pair != 2 && ctas_per_pair != 4 && ctas_per_pair != 8 &&
        ctas_per_pair != 16)
        return REVO_ERR_INVALID_ARG;
    if (threads_per_cta != 0 && threads_per_cta != 128 && threads_per_cta != 256 && threads_per_cta != 512 &&
        threads_per_cta != 1024)
        return REVO_ERR_INVALID_ARG;
    ctx->track_ctas_per_pair = ctas_per_pair;
    ctx->track_threads = threads_per_cta;
    return REVO_OK;
}

int revo_ctx_set_track_engine(revo_ctx *ctx, int engine, int chunk_points)
{
    if (!ctx || engine < 0 || engine > 3 || chunk_points < 0) return REVO_ERR_INVALID_ARG;
    ctx->track_engine = engine;
    ctx->track_chunk_points = chunk_points;
    return REVO_OK;
}

// Grow the device's stream-ordered memory pool to at least `bytes` of cached, reusable memory now (one allocation +
// free on the context stream), so that later slab allocations of a steady-state stream never reach the driver.
int revo_ctx_reserve(revo_ctx *ctx, size_t bytes)
{
    if (!ctx) return REVO_ERR_INVALID_ARG;
    if (bytes == 0) return REVO_OK;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    void *p = nullptr;
    REVO_CUDA(ctx, cudaMallocAsync(&p, bytes, ctx->stream));
    REVO_CUDA(ctx, cudaFreeAsync(p, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

// ---------------------------------------------------------------------------------------------------
// pyramid construction
// ---------------------------------------------------------------------------------------------------
struct LevelGeom {
    int w, h, patch, hist_w, hist_h, cap, n_tiles;
    revo_camera cam;
};

static int level_geometry(const revo_pyr_config *cfg, const revo_camera *cam0, LevelGeom *g)
{
    if (cfg->n_levels < 1 || cfg->n_levels > REVO_MAX_LEVELS) return REVO_ERR_INVALID_ARG;
    if (cam0->width < 8 || cam0->height < 8) return REVO_ERR_INVALID_ARG;
    for (int l = 0; l < cfg->n_levels; ++l) {
        level_camera(*cam0, l, &g[l].cam);
        g[l].w = g[l].cam.width; g[l].h = g[l].cam.height;
        if (g[l].w < 4 || g[l].h < 4) return REVO_ERR_UNSUPPORTED;
        // the reference is only self-consistent for even sizes (pyrDown -> (n+1)/2, depth/Camera -> n/2:
        // imgpyramidrgbd.cpp:79-85, camerapyr.h:100); refuse the others instead of reading out of bounds
        if (l + 1 < cfg->n_levels && ((g[l].w & 1) || (g[l].h & 1))) return REVO_ERR_UNSUPPORTED;
        g[l].patch = std::max(1, cfg->patch0 >> l);
        g[l].hist_w = g[l].w / g[l].patch; g[l].hist_h = g[l].h / g[l].patch;
        g[l].cap = g[l].w * g[l].h / 2 + 1024;
        g[l].n_tiles = ((g[l].w + kTileW - 1) / kTileW) * ((g[l].h + kTileH - 1) / kTileH);
    }
    return REVO_OK;
}

// depth16 != nullptr: raw 16-bit depth (units of depth_scale metres), converted on the device (K0 k_depth_u16)
static int create_batch_impl(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, int n, const uint8_t *bgr,
                             int channels, const float *depth, const uint16_t *depth16, float depth_scale,
                             const double *timestamps, revo_pyr **pyr_out)
{
    if (!ctx || !cfg || !cam0 || !bgr || (!depth && !depth16) || !pyr_out || n < 1 || (channels != 3 && channels != 4))
        return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    LevelGeom g[REVO_MAX_LEVELS];
    int rc = level_geometry(cfg, cam0, g);
    if (rc) return rc;
    const int NL = cfg->n_levels;
    const int w0 = g[0].w, h0 = g[0].h;

    // ---- slab layout: [array][level][frame], every chunk 256-byte aligned ------------------------
    size_t off = 0;
    auto take = [&](size_t per_frame) { size_t o = off; off += align_up(per_frame, 256) * (size_t)n; return o; };
    size_t o_desc[REVO_MAX_LEVELS], o_gray[REVO_MAX_LEVELS], o_depth[REVO_MAX_LEVELS], o_edges[REVO_MAX_LEVELS],
        o_eorig[REVO_MAX_LEVELS], o_hist[REVO_MAX_LEVELS], o_pts[REVO_MAX_LEVELS], o_toff[REVO_MAX_LEVELS];
    for (int l = 0; l < NL; ++l) { o_desc[l] = off; off += align_up(sizeof(ImgLevel) * (size_t)n, 256); }
    const size_t o_counters = off; off += align_up(sizeof(int) * 2 * NL * (size_t)n, 256);   // n_pts, nz_patches
    for (int l = 0; l < NL; ++l) {
        const size_t px = (size_t)g[l].w * g[l].h;
        o_gray[l] = take(px); o_depth[l] = take(px * 4); o_edges[l] = take(px); o_eorig[l] = take(px);
        o_hist[l] = take((size_t)std::max(1, g[l].hist_w * g[l].hist_h));
        o_pts[l] = take((size_t)g[l].cap * 16);
        o_toff[l] = take(((size_t)g[l].n_tiles + 1) * 4);
    }
    const size_t o_labels = take((size_t)w0 * h0 * 4);
    const size_t o_flags = take((size_t)w0 * h0);
    const size_t total = off;

    // Host inputs are uploaded on the context's COPY stream (slab allocation, descriptor tables, bgr staging and depth),
    // the kernels run on the main stream behind an event: the H2D of the next batch overlaps the kernels of this one.
    const bool host_in = !is_device_ptr(bgr);
    cudaStream_t up = host_in ? ctx->copy_stream : ctx->stream;

    Slab *slab = new (std::nothrow) Slab();
    if (!slab) return REVO_ERR_INVALID_ARG;
    slab->n_frames = n; slab->live = n; slab->bytes = total; slab->mem = nullptr;
    slab->stream = ctx->stream; slab->ready = nullptr;
    cudaError_t e = cudaMallocAsync(&slab->mem, total, up);
    if (e != cudaSuccess) { delete slab; return cuda_fail(ctx, e, "cudaMallocAsync(slab)"); }
    uint8_t *base = (uint8_t *)slab->mem;
    auto chunk = [&](size_t o, size_t per_frame, int f) { return base + o + align_up(per_frame, 256) * (size_t)f; };

    std::vector<revo_pyr *> pyrs(n);
    std::vector<ImgLevel> host_desc((size_t)NL * n);
    for (int f = 0; f < n; ++f) {
        revo_pyr *p = new revo_pyr();
        p->slab = slab; p->index_in_slab = f; p->n_levels = NL; p->cfg = *cfg; p->cam0 = *cam0;
        p->timestamp = timestamps ? timestamps[f] : 0.0; p->kf_slab = nullptr; p->is_keyframe = false;
        for (int l = 0; l < NL; ++l) {
            ImgLevel &L = p->lv[l];
            const size_t px = (size_t)g[l].w * g[l].h;
            L.gray = chunk(o_gray[l], px, f);
            L.depth = (float *)chunk(o_depth[l], px * 4, f);
            L.edges = chunk(o_edges[l], px, f);
            L.edges_orig = chunk(o_eorig[l], px, f);
            L.hist = chunk(o_hist[l], (size_t)std::max(1, g[l].hist_w * g[l].hist_h), f);
            L.pts = (float4 *)chunk(o_pts[l], (size_t)g[l].cap * 16, f);
            L.n_pts = (int *)(base + o_counters) + ((size_t)f * NL + l) * 2;
            L.nz_patches = L.n_pts + 1;
            L.tile_off = (int *)chunk(o_toff[l], ((size_t)g[l].n_tiles + 1) * 4, f);
            L.labels = (int *)chunk(o_labels, (size_t)w0 * h0 * 4, f);
            L.flags = chunk(o_flags, (size_t)w0 * h0, f);
            L.dt = nullptr; L.opt = nullptr;
            L.w = g[l].w; L.h = g[l].h; L.pts_cap = g[l].cap; L.patch = g[l].patch;
            L.hist_w = g[l].hist_w; L.hist_h = g[l].hist_h;
            L.fx = g[l].cam.fx; L.fy = g[l].cam.fy; L.cx = g[l].cam.cx; L.cy = g[l].cam.cy;
            host_desc[(size_t)l * n + f] = L;
        }
        pyrs[f] = p;
    }
    int stage_slot = -1;     // which of the context's two bgr staging buffers this call uploads into
    auto fail = [&](int code) {
        for (auto *p : pyrs) delete p;
        cudaStreamSynchronize(ctx->stream);   // error path only: nothing may still be using the slab
        cudaFreeAsync(slab->mem, up);
        delete slab;
        return code;
    };
    for (int l = 0; l < NL; ++l) {
        slab->d_desc[l] = (ImgLevel *)(base + o_desc[l]);
        e = cudaMemcpyAsync(slab->d_desc[l], &host_desc[(size_t)l * n], sizeof(ImgLevel) * (size_t)n, cudaMemcpyHostToDevice, up);
        if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memcpy(desc)"));
    }
    e = cudaMemsetAsync(base + o_counters, 0, sizeof(int) * 2 * NL * (size_t)n, up);
    if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memset(counters)"));

    // ---- inputs -------------------------------------------------------------------------------------
    const size_t bgr_frame = (size_t)w0 * h0 * channels;
    const uint8_t *d_bgr = bgr;
    if (host_in) {
        // persistent double buffer: no allocator dependency between this upload and the kernels of the previous batch
        stage_slot = ctx->stage_next;
        ctx->stage_next ^= 1;
        const size_t bgr_bytes = align_up(bgr_frame * (size_t)n, 256);
        const size_t need = bgr_bytes + (depth16 ? (size_t)w0 * h0 * 2 * (size_t)n : 0);
        if (ctx->stage_bytes[stage_slot] < need) {
            if (ctx->stage[stage_slot]) {
                cudaStreamSynchronize(ctx->stream);
                cudaStreamSynchronize(up);
                cudaFree(ctx->stage[stage_slot]);
                ctx->stage[stage_slot] = nullptr; ctx->stage_bytes[stage_slot] = 0; ctx->stage_used[stage_slot] = false;
            }
            e = cudaMalloc(&ctx->stage[stage_slot], need);
            if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "cudaMalloc(bgr staging)"));
            ctx->stage_bytes[stage_slot] = need;
        }
        if (ctx->stage_used[stage_slot]) cudaStreamWaitEvent(up, ctx->stage_consumed[stage_slot], 0);   // gray of two batches ago
        cudaEventRecord(ctx->ev[6], up);
        e = cudaMemcpyAsync(ctx->stage[stage_slot], bgr, bgr_frame * (size_t)n, cudaMemcpyHostToDevice, up);
        if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memcpy(bgr)"));
        d_bgr = (const uint8_t *)ctx->stage[stage_slot];
        if (depth16) {
            uint8_t *d16 = (uint8_t *)ctx->stage[stage_slot] + bgr_bytes;
            e = cudaMemcpyAsync(d16, depth16, (size_t)w0 * h0 * 2 * (size_t)n, cudaMemcpyHostToDevice, up);
            if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memcpy(depth16)"));
            depth16 = (const uint16_t *)d16;
        }
    }
    if (!depth16) {
        const size_t fb = (size_t)w0 * h0 * 4;
        e = cudaMemcpy2DAsync(base + o_depth[0], align_up(fb, 256), depth, fb, fb, (size_t)n, cudaMemcpyDefault, up);
        if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "memcpy(depth)"));
    }
    if (host_in) {
        cudaEventRecord(ctx->ev[7], up);
        ctx->ev_valid[3] = true;
        cudaEvent_t uploaded;
        e = cudaEventCreateWithFlags(&uploaded, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventRecord(uploaded, up);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, uploaded, 0);
        if (e != cudaSuccess) return fail(cuda_fail(ctx, e, "upload event"));
        cudaEventDestroy(uploaded);   // released by the runtime once the wait has been satisfied
    }

    // ---- the pyramid (imgpyramidrgbd.cpp:43-96) -------------------------------------------------
    const double t1 = cfg->canny_threshold1, t2 = cfg->canny_threshold2;
    double lo = std::min(t1, t2), hi = std::max(t1, t2);
    lo = std::min(32767.0, lo); hi = std::min(32767.0, hi);
    if (lo > 0) lo *= lo;
    if (hi > 0) hi *= hi;
    const int low = (int)floor(lo), high = (int)floor(hi);

    cudaEventRecord(ctx->ev[0], ctx->stream);
    rc = launch_gray(ctx, d_bgr, (size_t)w0 * channels, channels, bgr_frame, slab->d_desc[0], n, w0, h0);
    if (!rc && depth16) rc = launch_depth_u16(ctx, depth16, (size_t)w0 * h0, depth_scale, slab->d_desc[0], n, w0 * h0);
    if (stage_slot >= 0) {
        cudaEventRecord(ctx->stage_consumed[stage_slot], ctx->stream);
        ctx->stage_used[stage_slot] = true;
    }
    for (int l = 0; l < NL && !rc; ++l) {
        if (l > 0) rc = launch_pyrdown_depth(ctx, slab->d_desc[l - 1], slab->d_desc[l], n, g[l].w, g[l].h, g[l - 1].w, g[l - 1].h);
        if (!rc) {
            alignas(64) unsigned char tmap[128];
            const bool tma = make_gray_tensor_map(tmap, base + o_gray[l], g[l].w, g[l].h, n, align_up((size_t)g[l].w * g[l].h, 256));
            rc = launch_canny(ctx, slab->d_desc[l], n, g[l].w, g[l].h, low, high, tma ? tmap : nullptr, g[l].patch, base + o_flags,
                              align_up((size_t)w0 * h0, 256));
        }
        // fill-in is only defined for the reference's 3 patch sizes (levels 1,2); see SURVEY D5
        const bool fill = cfg->use_edge_hist && l >= 1 && l <= 2;
        if (!rc) rc = launch_hist_fill(ctx, slab->d_desc[l], l > 0 ? slab->d_desc[l - 1] : nullptr, n, g[l].w, g[l].h,
                                      g[l].patch, l > 0 ? g[l - 1].patch : g[l].patch, fill, cfg->n_percentage);
        if (!rc) rc = launch_compact(ctx, slab->d_desc[l], n, g[l].w, g[l].h, cfg->depth_min, cfg->depth_max);
    }
    if (rc) return fail(rc);
    cudaEventRecord(ctx->ev[1], ctx->stream);
    ctx->ev_valid[0] = true;
    if (cudaEventCreateWithFlags(&slab->ready, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(slab->ready, ctx->stream);
    for (int f = 0; f < n; ++f) pyr_out[f] = pyrs[f];
    return REVO_OK;
}

int revo_pyr_create_batch(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, int n, const uint8_t *bgr,
                          int channels, const float *depth, const double *timestamps, revo_pyr **pyr_out)
{
    if (!depth) return REVO_ERR_INVALID_ARG;
    return create_batch_impl(ctx, cfg, cam0, n, bgr, channels, depth, nullptr, 0.f, timestamps, pyr_out);
}

int revo_pyr_create_batch_u16(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, int n, const uint8_t *bgr,
                              int channels, const uint16_t *depth_raw, float depth_scale, const double *timestamps,
                              revo_pyr **pyr_out)
{
    if (!depth_raw) return REVO_ERR_INVALID_ARG;
    return create_batch_impl(ctx, cfg, cam0, n, bgr, channels, nullptr, depth_raw, depth_scale, timestamps, pyr_out);
}

int revo_pyr_create(revo_ctx *ctx, const revo_pyr_config *cfg, const revo_camera *cam0, const uint8_t *bgr, size_t bgr_stride,
                    int channels, const float *depth, size_t depth_stride, double timestamp, revo_pyr **pyr_out)
{
    if (!ctx || !cfg || !cam0 || !bgr || !depth || !pyr_out) return REVO_ERR_INVALID_ARG;
    const size_t tight_bgr = (size_t)cam0->width * channels, tight_d = (size_t)cam0->width * 4;
    if (bgr_stride == 0) bgr_stride = tight_bgr;
    if (depth_stride == 0) depth_stride = tight_d;
    if (bgr_stride == tight_bgr && depth_stride == tight_d)
        return revo_pyr_create_batch(ctx, cfg, cam0, 1, bgr, channels, depth, &timestamp, pyr_out);
    // strided inputs (cv::Mat ROI): repack rows on the host, then take the tight path
    if (is_device_ptr(bgr) || is_device_ptr(depth)) return REVO_ERR_UNSUPPORTED;
    std::vector<uint8_t> b(tight_bgr * cam0->height);
    std::vector<float> d((size_t)cam0->width * cam0->height);
    for (int y = 0; y < cam0->height; ++y) {
        memcpy(b.data() + tight_bgr * y, bgr + bgr_stride * y, tight_bgr);
        memcpy((uint8_t *)d.data() + tight_d * y, (const uint8_t *)depth + depth_stride * y, tight_d);
    }
    int rc = revo_pyr_create_batch(ctx, cfg, cam0, 1, b.data(), channels, d.data(), &timestamp, pyr_out);
    if (!rc) cudaStreamSynchronize(ctx->stream);   // the temporaries die here
    return rc;
}

// A pyramid built on another context's stream: order this context's stream after the build (no host sync).
static void wait_for_build(revo_ctx *ctx, const revo_pyr *p)
{
    if (p && p->slab && p->slab->ready && p->slab->stream != ctx->stream) cudaStreamWaitEvent(ctx->stream, p->slab->ready, 0);
}

// One stream-ordered allocation for the keyframe structures (dt 4 B/px + quad structure 32 B/px, all levels) of
// every pyramid in `ps` that does not have them yet.
static int alloc_keyframe_mem(revo_ctx *ctx, revo_pyr *const *ps, int n)
{
    auto bytes_of = [](const revo_pyr *p) {
        size_t b = 0;
        for (int l = 0; l < p->n_levels; ++l) b += align_up((size_t)p->lv[l].w * p->lv[l].h * 36, 256);
        return b;
    };
    size_t total = 0;
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (!ps[i]->kf_slab) { total += bytes_of(ps[i]); ++m; }
    if (!m) return REVO_OK;
    KfSlab *ks = new (std::nothrow) KfSlab();
    if (!ks) return REVO_ERR_INVALID_ARG;
    ks->mem = nullptr; ks->live = m;
    cudaError_t e = cudaMallocAsync(&ks->mem, total, ctx->stream);
    if (e != cudaSuccess) { delete ks; return cuda_fail(ctx, e, "cudaMallocAsync(keyframe)"); }
    uint8_t *mem = (uint8_t *)ks->mem;
    for (int i = 0; i < n; ++i) {
        revo_pyr *p = ps[i];
        if (p->kf_slab) continue;
        p->kf_slab = ks;
        for (int l = 0; l < p->n_levels; ++l) {
            const size_t px = (size_t)p->lv[l].w * p->lv[l].h;
            p->lv[l].opt = (uint4 *)mem;
            p->lv[l].dt = (float *)(mem + px * 32);
            mem += align_up(px * 36, 256);
        }
    }
    return REVO_OK;
}

int revo_pyr_make_keyframe_batch(revo_ctx *ctx, int n, revo_pyr *const *pyrs)
{
    if (!ctx || !pyrs || n < 0) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<revo_pyr *> todo;
    for (int i = 0; i < n; ++i) {
        if (!pyrs[i]) return REVO_ERR_INVALID_ARG;
        if (!pyrs[i]->is_keyframe && std::find(todo.begin(), todo.end(), pyrs[i]) == todo.end()) todo.push_back(pyrs[i]);
    }
    if (todo.empty()) return REVO_OK;
    const int m = (int)todo.size();
    const int NL = todo[0]->n_levels;
    for (auto *p : todo)
        if (p->n_levels != NL || p->lv[0].w != todo[0]->lv[0].w || p->lv[0].h != todo[0]->lv[0].h) return REVO_ERR_INVALID_ARG;
    for (size_t i = 0; i < todo.size(); ++i)
        if (i == 0 || todo[i]->slab != todo[i - 1]->slab) wait_for_build(ctx, todo[i]);
    {
        int rc = alloc_keyframe_mem(ctx, todo.data(), m);
        if (rc) return rc;
    }
    // temporary descriptor tables (with dt/opt set) in a stream-ordered allocation, filled from pinned mapped host memory by
    // a kernel (not by the copy engine, which may be busy for milliseconds with the frame uploads of the next batches)
    const size_t tab_bytes = sizeof(ImgLevel) * (size_t)NL * m;
    if (ctx->pinned_kf_busy) { REVO_CUDA(ctx, cudaEventSynchronize(ctx->pinned_kf_read)); ctx->pinned_kf_busy = false; }
    if (ctx->pinned_kf_bytes < tab_bytes + 16) {
        if (ctx->pinned_kf) REVO_CUDA(ctx, cudaFreeHost(ctx->pinned_kf));
        ctx->pinned_kf = nullptr;
        ctx->pinned_kf_bytes = align_up(2 * tab_bytes + 16, 1 << 16);
        REVO_CUDA(ctx, cudaMallocHost(&ctx->pinned_kf, ctx->pinned_kf_bytes));
    }
    ImgLevel *host = (ImgLevel *)ctx->pinned_kf;
    for (int l = 0; l < NL; ++l)
        for (int i = 0; i < m; ++i) host[(size_t)l * m + i] = todo[i]->lv[l];
    ImgLevel *d_tab = nullptr;
    REVO_CUDA(ctx, cudaMallocAsync((void **)&d_tab, tab_bytes + 16, ctx->stream));
    {
        int rc = launch_stage_in(ctx, host, d_tab, tab_bytes);
        if (rc) return rc;
        cudaEventRecord(ctx->pinned_kf_read, ctx->stream);
        ctx->pinned_kf_busy = true;
    }
    int rc = REVO_OK;
    cudaEventRecord(ctx->ev[2], ctx->stream);
    for (int l = 0; l < NL && !rc; ++l) rc = launch_keyframe(ctx, d_tab + (size_t)l * m, m, todo[0]->lv[l].w, todo[0]->lv[l].h);
    cudaEventRecord(ctx->ev[3], ctx->stream);
    ctx->ev_valid[1] = true;
    cudaFreeAsync(d_tab, ctx->stream);
    if (rc) return rc;
    for (auto *p : todo) p->is_keyframe = true;
    return REVO_OK;
}

int revo_pyr_make_keyframe(revo_ctx *ctx, revo_pyr *pyr) { return revo_pyr_make_keyframe_batch(ctx, 1, &pyr); }

static void destroy_one(revo_ctx *ctx, revo_pyr *pyr)
{
    if (KfSlab *k = pyr->kf_slab) {
        if (--k->live == 0) {
            cudaFreeAsync(k->mem, ctx->stream);
            delete k;
        }
    }
    Slab *s = pyr->slab;
    if (s && --s->live == 0) {
        if (s->ready) cudaEventDestroy(s->ready);
        cudaFreeAsync(s->mem, ctx->stream);
        delete s;
    }
    delete pyr;
}

int revo_pyr_destroy(revo_ctx *ctx, revo_pyr *pyr)
{
    if (!pyr) return REVO_OK;
    if (!ctx) return REVO_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    destroy_one(ctx, pyr);
    (void)cudaGetLastError();
    return REVO_OK;
}

int revo_pyr_destroy_batch(revo_ctx *ctx, int n, revo_pyr *const *pyrs)
{
    if (!ctx || (n > 0 && !pyrs)) return REVO_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < n; ++i)
        if (pyrs[i]) destroy_one(ctx, pyrs[i]);
    (void)cudaGetLastError();
    return REVO_OK;
}

int revo_pyr_is_keyframe(const revo_pyr *pyr) { return pyr && pyr->is_keyframe; }
double revo_pyr_timestamp(const revo_pyr *pyr) { return pyr ? pyr->timestamp : 0.0; }

int revo_pyr_level_camera(const revo_pyr *pyr, int lvl, revo_camera *cam_out)
{
    if (!pyr || !cam_out) return REVO_ERR_INVALID_ARG;
    if (lvl < 0 || lvl >= pyr->n_levels) return REVO_ERR_BAD_LEVEL;
    const ImgLevel &L = pyr->lv[lvl];
    cam_out->fx = L.fx; cam_out->fy = L.fy; cam_out->cx = L.cx; cam_out->cy = L.cy; cam_out->width = L.w; cam_out->height = L.h;
    return REVO_OK;
}

int revo_pyr_num_edges(revo_ctx *ctx, const revo_pyr *pyr, int lvl, int *n_out)
{
    if (!ctx || !pyr || !n_out) return REVO_ERR_INVALID_ARG;
    if (lvl < 0 || lvl >= pyr->n_levels) return REVO_ERR_BAD_LEVEL;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    REVO_CUDA(ctx, cudaMemcpyAsync(n_out, pyr->lv[lvl].n_pts, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

int revo_pyr_download(revo_ctx *ctx, const revo_pyr *pyr, int lvl, int which, void *dst, size_t dst_bytes, size_t *bytes_out)
{
    if (!ctx || !pyr) return REVO_ERR_INVALID_ARG;
    if (lvl < 0 || lvl >= pyr->n_levels) return REVO_ERR_BAD_LEVEL;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    const ImgLevel &L = pyr->lv[lvl];
    const size_t px = (size_t)L.w * L.h;
    const void *src = nullptr;
    size_t bytes = 0;
    switch (which) {
        case REVO_ARRAY_GRAY: src = L.gray; bytes = px; break;
        case REVO_ARRAY_DEPTH: src = L.depth; bytes = px * 4; break;
        case REVO_ARRAY_EDGES: src = L.edges; bytes = px; break;
        case REVO_ARRAY_EDGES_ORIG: src = L.edges_orig; bytes = px; break;
        case REVO_ARRAY_HIST: src = L.hist; bytes = (size_t)L.hist_w * L.hist_h; break;
        case REVO_ARRAY_DT:
            if (!pyr->is_keyframe) return REVO_ERR_NOT_KEYFRAME;
            src = L.dt; bytes = px * 4; break;
        case REVO_ARRAY_OPTSTRUCT: {
            // the reference's float4 layout is materialised on demand from the distance transform
            if (!pyr->is_keyframe) return REVO_ERR_NOT_KEYFRAME;
            int rc = ensure_scratch(ctx, px * 16);
            if (rc) return rc;
            rc = launch_opt_struct_f4(ctx, L.dt, L.w, L.h, (float4 *)ctx->scratch);
            if (rc) return rc;
            src = ctx->scratch; bytes = px * 16; break;
        }
        case REVO_ARRAY_EDGES3D_DEVICE_ORDER: {
            int n = 0;
            int rc = revo_pyr_num_edges(ctx, pyr, lvl, &n);
            if (rc) return rc;
            src = L.pts; bytes = (size_t)n * 16; break;
        }
        case REVO_ARRAY_EDGES3D: {
            // reference order: column-major scan (imgpyramidrgbd.cpp:203-205)
            const size_t need = px * 16 + (size_t)(L.w + 4) * 4 + sizeof(ImgLevel) + 1024;
            int rc = ensure_scratch(ctx, need);
            if (rc) return rc;
            uint8_t *s = (uint8_t *)ctx->scratch;
            float4 *d_out = (float4 *)s;
            int *d_col = (int *)(s + align_up(px * 16, 256));
            int *d_n = d_col + L.w + 1;
            ImgLevel *d_desc = (ImgLevel *)(s + align_up(px * 16, 256) + align_up((size_t)(L.w + 4) * 4, 256));
            REVO_CUDA(ctx, cudaMemcpyAsync(d_desc, &L, sizeof(ImgLevel), cudaMemcpyHostToDevice, ctx->stream));
            rc = launch_edges3d_reference_order(ctx, d_desc, L.w, L.h, pyr->cfg.depth_min, pyr->cfg.depth_max, d_out, d_n, d_col);
            if (rc) return rc;
            int n = 0;
            REVO_CUDA(ctx, cudaMemcpyAsync(&n, d_n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            src = d_out; bytes = (size_t)n * 16; break;
        }
        default: return REVO_ERR_INVALID_ARG;
    }
    if (bytes_out) *bytes_out = bytes;
    if (!dst) return REVO_OK;   // size query
    if (dst_bytes < bytes) return REVO_ERR_BUFFER_TOO_SMALL;
    if (bytes) REVO_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

int revo_pyr_upload_level(revo_ctx *ctx, revo_pyr *pyr, int lvl, const float *pts4, int n, const float *dt, const float *opt4)
{
    if (!ctx || !pyr) return REVO_ERR_INVALID_ARG;
    if (lvl < 0 || lvl >= pyr->n_levels) return REVO_ERR_BAD_LEVEL;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    ImgLevel &L = pyr->lv[lvl];
    const size_t px = (size_t)L.w * L.h;
    if (pts4) {
        if (n < 0 || n > L.pts_cap) return REVO_ERR_BUFFER_TOO_SMALL;
        if (n) REVO_CUDA(ctx, cudaMemcpyAsync(L.pts, pts4, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
        REVO_CUDA(ctx, cudaMemcpyAsync(L.n_pts, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (dt || opt4) {
        int rc = alloc_keyframe_mem(ctx, &pyr, 1);
        if (rc) return rc;
        if (dt) REVO_CUDA(ctx, cudaMemcpyAsync(L.dt, dt, px * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (opt4) {
            rc = ensure_scratch(ctx, px * 16);
            if (rc) return rc;
            REVO_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, opt4, px * 16, cudaMemcpyHostToDevice, ctx->stream));
            rc = launch_opt_pack_from_f4(ctx, (const float4 *)ctx->scratch, L.w, L.h, L.opt);
            if (rc) return rc;
            pyr->is_keyframe = true;
        }
    }
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return REVO_OK;
}

// ---------------------------------------------------------------------------------------------------
// tracking
// ---------------------------------------------------------------------------------------------------
static int fill_pair(const revo_pyr *ref, const revo_pyr *cur, int min_lvl, int max_lvl, const float *R9, const float *t3,
                     PairDesc *d)
{
    if (!ref || !cur) return REVO_ERR_INVALID_ARG;
    if (!ref->is_keyframe) return REVO_ERR_NOT_KEYFRAME;
    if (min_lvl < max_lvl || max_lvl < 0 || min_lvl >= cur->n_levels || min_lvl >= ref->n_levels) return REVO_ERR_BAD_LEVEL;
    memset(d, 0, sizeof(*d));
    for (int l = max_lvl; l <= min_lvl; ++l) {
        const ImgLevel &c = cur->lv[l], &r = ref->lv[l];
        if (c.w != r.w || c.h != r.h) return REVO_ERR_INVALID_ARG;
        LevelIn &L = d->lvl[l];
        L.pts = c.pts; L.n_pts = c.n_pts; L.opt = r.opt;
        // calcErrorAndBuffers takes the camera from the reference frame (optimizer.cpp:80)
        L.fx = r.fx; L.fy = r.fy; L.cx = r.cx; L.cy = r.cy; L.w = r.w; L.h = r.h;
    }
    d->ref_dt_min = ref->lv[min_lvl].dt;
    memcpy(d->R, R9, sizeof(float) * 9);
    memcpy(d->t, t3, sizeof(float) * 3);
    return REVO_OK;
}

static int run_track(revo_ctx *ctx, TrackParams &prm, int n, revo_pyr *const *refs, revo_pyr *const *curs, const float *R9s,
                     const float *t3s, revo_track_result *results, double *records, revo_trace_entry *trace, int trace_cap,
                     int *trace_counts)
{
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    const int min_lvl = prm.mode == 0 ? prm.cfg.pyr_min_lvl : prm.level;
    const int max_lvl = prm.mode == 0 ? prm.cfg.pyr_max_lvl : prm.level;
    // descriptors are written straight into pinned, device-mapped host memory (see k_stage_in)
    {
        int rc = ensure_pinned(ctx, align_up(sizeof(PairDesc) * (size_t)n, 16));
        if (rc) return rc;
    }
    PairDesc *host = (PairDesc *)ctx->pinned;
    for (int i = 0; i < n; ++i) {
        int rc = fill_pair(refs[i], curs[i], min_lvl, max_lvl, R9s + 9 * (size_t)i, t3s + 3 * (size_t)i, &host[i]);
        if (rc) return rc;
        // one device-side wait per distinct batch (the pairs of a batch share two slabs)
        if (i == 0 || refs[i]->slab != refs[i - 1]->slab) wait_for_build(ctx, refs[i]);
        if (i == 0 || curs[i]->slab != curs[i - 1]->slab) wait_for_build(ctx, curs[i]);
    }
    if (!trace) trace_cap = 0;
    prm.trace_cap = trace_cap;
    prm.profile = getenv("REVO_TRACK_PROF") != nullptr;
    // device workspace: pairs | results | records | trace | trace counts
    const size_t b_pairs = align_up(sizeof(PairDesc) * (size_t)n, 256);
    const size_t b_res = align_up(sizeof(revo_track_result) * (size_t)n, 256);
    const size_t b_rec = align_up(sizeof(double) * 32 * (size_t)n, 256);
    const size_t b_tr = align_up(sizeof(revo_trace_entry) * (size_t)trace_cap * n, 256);
    const size_t b_tc = align_up(sizeof(int) * (size_t)n, 256);
    uint8_t *ws = nullptr;
    // engine: one cluster per pair (track.cu; measured faster at every batch size, scratch/track_bench.py) unless the
    // caller / REVO_TRACK_ENGINE asks for the task queue (track_queue.cu); a pair split over several GPUs always uses
    // the cluster engine (the peer mailboxes live there)
    const int env_engine = getenv("REVO_TRACK_ENGINE") ? atoi(getenv("REVO_TRACK_ENGINE")) : 0;
    int engine = ctx->track_engine ? ctx->track_engine : env_engine;
    if (prm.split_world > 1) engine = 1;
    if (engine != 2 && engine != 3) engine = 1;
    const size_t b_q = engine == 2 ? align_up(track_queue_workspace_bytes(n, 8 * ctx->prop.multiProcessorCount, nullptr), 256) : 0;
    REVO_CUDA(ctx, cudaMallocAsync((void **)&ws, b_pairs + b_res + b_rec + b_tr + b_tc + 256 + b_q, ctx->stream));
    PairDesc *d_pairs = (PairDesc *)ws;
    revo_track_result *d_res = (revo_track_result *)(ws + b_pairs);
    double *d_rec = (double *)(ws + b_pairs + b_res);
    revo_trace_entry *d_tr = trace_cap ? (revo_trace_entry *)(ws + b_pairs + b_res + b_rec) : nullptr;
    int *d_tc = (int *)(ws + b_pairs + b_res + b_rec + b_tr);
    int *d_wc = (int *)(ws + b_pairs + b_res + b_rec + b_tr + b_tc);
    uint8_t *d_q = ws + b_pairs + b_res + b_rec + b_tr + b_tc + 256;
    int rc = REVO_OK;
    rc = launch_stage_in(ctx, host, d_pairs, sizeof(PairDesc) * (size_t)n);
    cudaError_t e = cudaMemsetAsync(d_res, 0, b_res + b_rec, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_wc, 0, 256, ctx->stream);
    if (e != cudaSuccess) rc = cuda_fail(ctx, e, "track upload");
    cudaEventRecord(ctx->ev[4], ctx->stream);
    if (!rc) {
        if (engine == 2) rc = launch_track_queue(ctx, d_pairs, n, prm, d_res, d_rec, d_tr, d_tc, d_q, b_q);
        else if (engine == 3) rc = launch_track_pp(ctx, d_pairs, n, prm, d_res, d_rec, d_tr, d_tc, d_wc);
        else rc = launch_track(ctx, d_pairs, n, prm, d_res, d_rec, d_tr, d_tc, d_wc);
    }
    cudaEventRecord(ctx->ev[5], ctx->stream);
    ctx->ev_valid[2] = true;
    if (!rc && results) {
        e = cudaMemcpyAsync(results, d_res, sizeof(revo_track_result) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "results download");
    }
    if (!rc && records) {
        e = cudaMemcpyAsync(records, d_rec, sizeof(double) * 32 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "records download");
    }
    if (!rc && trace && trace_cap) {
        e = cudaMemcpyAsync(trace, d_tr, sizeof(revo_trace_entry) * (size_t)trace_cap * n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && trace_counts)
            e = cudaMemcpyAsync(trace_counts, d_tc, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) rc = cuda_fail(ctx, e, "trace download");
    }
    unsigned long long prof[4] = {0, 0, 0, 0};
    const bool want_prof = getenv("REVO_TRACK_PROF") != nullptr;
    int q_ctl[4] = {0, 0, 0, 0};   // head, tail, pairs_done, abort of the queue engine
    if (!rc && engine == 2) cudaMemcpyAsync(q_ctl, d_q, sizeof(q_ctl), cudaMemcpyDeviceToHost, ctx->stream);
    std::vector<unsigned long long> q_prof;
    if (!rc && engine == 2 && want_prof) {
        q_prof.resize(9 + (size_t)n);
        cudaMemcpyAsync(q_prof.data(), d_q + 16, q_prof.size() * 8, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (!rc && want_prof && engine == 1) cudaMemcpyAsync(prof, d_wc + 2, sizeof(prof), cudaMemcpyDeviceToHost, ctx->stream);
    cudaFreeAsync(ws, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess && !rc) rc = cuda_fail(ctx, e, "track kernel");
    if (!rc && engine == 2 && (q_ctl[3] != 0 || q_ctl[2] != n)) {
        char buf[160];
        snprintf(buf, sizeof(buf), "track queue watchdog: abort=%d pairs_done=%d/%d head=%u tail=%u", q_ctl[3], q_ctl[2], n,
                 (unsigned)q_ctl[0], (unsigned)q_ctl[1]);
        ctx->last_error = buf;
        rc = REVO_ERR_CUDA;
    }
    if (!rc && want_prof && engine == 2) {
        const unsigned long long *st = q_prof.data() + 1;
        std::vector<unsigned long long> fin(q_prof.begin() + 9, q_prof.end());
        std::sort(fin.begin(), fin.end());
        const double nt = (double)std::max<unsigned long long>(st[4], 1), nl = (double)std::max<unsigned long long>(st[5], 1);
        fprintf(stderr, "[k_track_queue prof] pairs %d ctas %llu tasks %llu evals %llu | cycles/task: pop %.0f gather %.0f partial %.0f | "
                        "cycles/last-arrival %.0f | pair finish us: min %.0f p25 %.0f p50 %.0f p75 %.0f p90 %.0f max %.0f\n",
                n, st[6], st[4], st[5], st[0] / nt, st[1] / nt, st[2] / nt, st[3] / nl, fin.front() * 1e-3, fin[fin.size() / 4] * 1e-3,
                fin[fin.size() / 2] * 1e-3, fin[fin.size() * 3 / 4] * 1e-3, fin[fin.size() * 9 / 10] * 1e-3, fin.back() * 1e-3);
    }
    if (!rc && want_prof && prof[3])
        fprintf(stderr, "[k_track prof] pairs %d evals %llu  cycles/eval: gather %.0f reduce %.0f serial+sync %.0f\n", n, prof[3],
                (double)prof[0] / prof[3], (double)prof[1] / prof[3], (double)prof[2] / prof[3]);
    return rc;
}

int revo_track_batch(revo_ctx *ctx, const revo_tracker_config *cfg, int n, revo_pyr *const *refs, revo_pyr *const *curs,
                     const float *R9s, const float *t3s, revo_track_result *results, revo_trace_entry *trace, int trace_cap,
                     int *trace_counts)
{
    if (!ctx || !cfg || !refs || !curs || !R9s || !t3s || !results || n < 0) return REVO_ERR_INVALID_ARG;
    if (n == 0) return REVO_OK;
    TrackParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.cfg = *cfg; prm.mode = 0; prm.split_world = 1;
    return run_track(ctx, prm, n, refs, curs, R9s, t3s, results, nullptr, trace, trace_cap, trace_counts);
}

int revo_track(revo_ctx *ctx, const revo_tracker_config *cfg, const revo_pyr *ref, const revo_pyr *cur, float *R9, float *t3,
               revo_track_result *result)
{
    if (!ctx || !cfg || !ref || !cur || !R9 || !t3) return REVO_ERR_INVALID_ARG;
    revo_track_result res;
    revo_pyr *r = const_cast<revo_pyr *>(ref), *c = const_cast<revo_pyr *>(cur);
    int rc = revo_track_batch(ctx, cfg, 1, &r, &c, R9, t3, &res, nullptr, 0, nullptr);
    if (rc) return rc;
    if (result) *result = res;
    if (res.rc) return res.rc;
    memcpy(R9, res.R, sizeof(float) * 9);
    memcpy(t3, res.t, sizeof(float) * 3);
    return REVO_OK;
}

int revo_track_level(revo_ctx *ctx, const revo_opt_config *cfg, const revo_pyr *ref, const revo_pyr *cur, int lvl, float *R9,
                     float *t3, revo_residual_info *res, float *err, int *n_evals)
{
    if (!ctx || !cfg || !ref || !cur || !R9 || !t3) return REVO_ERR_INVALID_ARG;
    TrackParams prm;
    memset(&prm, 0, sizeof(prm));
    revo_tracker_config_default(&prm.cfg);
    prm.cfg.opt = *cfg; prm.cfg.check_init_values = 0;
    prm.mode = 1; prm.level = lvl; prm.split_world = 1;
    revo_track_result out;
    revo_pyr *r = const_cast<revo_pyr *>(ref), *c = const_cast<revo_pyr *>(cur);
    int rc = run_track(ctx, prm, 1, &r, &c, R9, t3, &out, nullptr, nullptr, 0, nullptr);
    if (rc) return rc;
    if (out.rc) return out.rc;
    memcpy(R9, out.R, sizeof(float) * 9);
    memcpy(t3, out.t, sizeof(float) * 3);
    if (res) *res = out.res;
    if (err) *err = out.error;
    if (n_evals) *n_evals = out.n_evals[lvl];
    return REVO_OK;
}

int revo_eval(revo_ctx *ctx, const revo_opt_config *cfg, const revo_pyr *ref, const revo_pyr *cur, int lvl, const float *R9,
              const float *t3, double *record32)
{
    if (!ctx || !cfg || !ref || !cur || !R9 || !t3 || !record32) return REVO_ERR_INVALID_ARG;
    TrackParams prm;
    memset(&prm, 0, sizeof(prm));
    revo_tracker_config_default(&prm.cfg);
    prm.cfg.opt = *cfg; prm.cfg.check_init_values = 0;
    prm.mode = 2; prm.level = lvl; prm.split_world = 1;
    revo_pyr *r = const_cast<revo_pyr *>(ref), *c = const_cast<revo_pyr *>(cur);
    return run_track(ctx, prm, 1, &r, &c, R9, t3, nullptr, record32, nullptr, 0, nullptr);
}

// ---------------------------------------------------------------------------------------------------
// tracking-quality vote
// ---------------------------------------------------------------------------------------------------
static bool invert4(const double *m /* column-major */, double *inv)
{
    // Gauss-Jordan with partial pivoting on [m | I]
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { a[r][c] = m[c * 4 + r]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        for (int r = k + 1; r < 4; ++r)
            if (fabs(a[r][k]) > fabs(a[piv][k])) piv = r;
        if (fabs(a[piv][k]) < 1e-300) return false;
        if (piv != k)
            for (int c = 0; c < 8; ++c) std::swap(a[piv][c], a[k][c]);
        const double d = 1.0 / a[k][k];
        for (int c = 0; c < 8; ++c) a[k][c] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == k) continue;
            const double f = a[r][k];
            if (f != 0.0)
                for (int c = 0; c < 8; ++c) a[r][c] -= f * a[k][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) inv[c * 4 + r] = a[r][4 + c];
    return true;
}

int revo_track_quality(revo_ctx *ctx, const revo_pyr *cur, int hist_level, int n_past, revo_pyr *const *past,
                       const float *past_world_poses16, const float *estimated_pose16, int n_frames_voting, revo_quality_result *out)
{
    if (!ctx || !cur || !out || n_past < 0 || (n_past > 0 && (!past || !past_world_poses16)) || !estimated_pose16)
        return REVO_ERR_INVALID_ARG;
    if (hist_level < 0 || hist_level >= cur->n_levels) return REVO_ERR_BAD_LEVEL;
    memset(out, 0, sizeof(*out));
    out->status = REVO_TRACKER_STATE_OK;
    int nf = n_past < n_frames_voting ? n_past : n_frames_voting;
    if (nf > 3) nf = 3;                        // histWeights has four entries (tracker.cpp:231-234)
    out->n_frames = nf > 0 ? nf : 0;
    if (nf <= 0) return REVO_OK;               // tracker.cpp:121: nothing to vote with
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    const ImgLevel &L = cur->lv[hist_level];
    QualityArgs a;
    memset(&a, 0, sizeof(a));
    a.n_frames = nf; a.fx = L.fx; a.fy = L.fy; a.cx = L.cx; a.cy = L.cy; a.w = L.w; a.h = L.h;
    double est[16], est_inv[16];
    for (int i = 0; i < 16; ++i) est[i] = estimated_pose16[i];
    if (!invert4(est, est_inv)) return REVO_ERR_INVALID_ARG;
    for (int f = 0; f < nf; ++f) {
        if (!past[f] || hist_level >= past[f]->n_levels) return REVO_ERR_INVALID_ARG;
        const float *pw = past_world_poses16 + 16 * (size_t)f;
        double tr[16];                          // inv(estimatedPose) * pastWorldPose   (tracker.cpp:147)
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 4; ++r) {
                double s = 0;
                for (int k = 0; k < 4; ++k) s += est_inv[k * 4 + r] * (double)pw[c * 4 + k];
                tr[c * 4 + r] = s;
            }
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) a.fr[f].R[c * 3 + r] = (float)tr[c * 4 + r];
        for (int r = 0; r < 3; ++r) a.fr[f].T[r] = (float)tr[12 + r];
        a.fr[f].pts = past[f]->lv[hist_level].pts;
        a.fr[f].n_pts = past[f]->lv[hist_level].n_pts;
        wait_for_build(ctx, past[f]);
    }
    wait_for_build(ctx, cur);
    const size_t words = ((size_t)L.w * L.h + 3) / 4;
    int rc = ensure_scratch(ctx, words * 4 + 256);
    if (rc) return rc;
    unsigned *d_mbits = (unsigned *)ctx->scratch;
    int *d_counters = (int *)((uint8_t *)ctx->scratch + align_up(words * 4, 64));
    // returnOrigEdges(histogramLevel): the Canny output before the fill-in (imgpyramidrgbd.h:69-77)
    const uint8_t *d_edges = (cur->cfg.use_edge_hist && hist_level > 0) ? L.edges_orig : L.edges;
    rc = launch_quality(ctx, a, L.depth, d_edges, cur->cfg.depth_min, cur->cfg.depth_max, d_mbits, d_counters);
    if (rc) return rc;
    int c[16];
    REVO_CUDA(ctx, cudaMemcpyAsync(c, d_counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    REVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    static const float kHistWeights[4] = {0.f, 1.f, 1.25f, 1.5f};
    float measure = 0.f;
    for (int k = 0; k < 4; ++k) { out->histogram[k] = c[k]; out->overlaps[k] = c[4 + k]; }
    for (int k = 1; k <= nf; ++k) measure += (float)c[4 + k] * kHistWeights[k];     // tracker.cpp:176-181
    out->overlap_measure = measure;
    out->out_of_bounds = c[8];
    // tracker.cpp:183: histogram.size() = 1 + frames that took part
    out->status = (measure >= (float)c[4] || nf + 1 < 4) ? REVO_TRACKER_STATE_OK : REVO_TRACKER_STATE_NEW_KF;
    return REVO_OK;
}

// ---------------------------------------------------------------------------------------------------
// multi-GPU split of one pair (one process per GPU; mailboxes exchanged as CUDA IPC handles)
// ---------------------------------------------------------------------------------------------------
struct SplitBlob {
    cudaIpcMemHandle_t handle;   // 64 bytes
    int32_t rank, world;
    int32_t pid_lo, device;
    char pad[REVO_SPLIT_HANDLE_BYTES - 64 - 16];
};
static_assert(sizeof(SplitBlob) == REVO_SPLIT_HANDLE_BYTES, "blob size");

int revo_split_export(revo_ctx *ctx, int rank, int world, void *handle_out)
{
    if (!ctx || !handle_out || world < 1 || world > 16 || rank < 0 || rank >= world) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->split_local) {
        REVO_CUDA(ctx, cudaMalloc(&ctx->split_local, 16384));
        REVO_CUDA(ctx, cudaMemset(ctx->split_local, 0, 16384));
    }
    ctx->split_rank = rank; ctx->split_world = world;
    SplitBlob b;
    memset(&b, 0, sizeof(b));
    REVO_CUDA(ctx, cudaIpcGetMemHandle(&b.handle, ctx->split_local));
    b.rank = rank; b.world = world; b.device = ctx->device;
    memcpy(handle_out, &b, sizeof(b));
    return REVO_OK;
}

int revo_split_open(revo_ctx *ctx, const void *handles)
{
    if (!ctx || !handles || !ctx->split_local) return REVO_ERR_INVALID_ARG;
    REVO_CUDA(ctx, cudaSetDevice(ctx->device));
    const SplitBlob *b = (const SplitBlob *)handles;
    for (int r = 0; r < ctx->split_world; ++r) {
        if (b[r].rank != r || b[r].world != ctx->split_world) return REVO_ERR_COMM;
        if (r == ctx->split_rank) { ctx->split_peers[r] = ctx->split_local; continue; }
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, b[r].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cuda_fail(ctx, e, "cudaIpcOpenMemHandle"); return REVO_ERR_COMM; }
        ctx->split_peers[r] = p;
    }
    ctx->split_seq = 0;
    return REVO_OK;
}

int revo_track_split(revo_ctx *ctx, const revo_tracker_config *cfg, const revo_pyr *ref, const revo_pyr *cur, float *R9, float *t3,
                     revo_track_result *result)
{
    if (!ctx || !cfg || !ref || !cur || !R9 || !t3) return REVO_ERR_INVALID_ARG;
    if (ctx->split_world < 1 || !ctx->split_peers[ctx->split_rank]) return REVO_ERR_COMM;
    TrackParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.cfg = *cfg; prm.mode = 0;
    prm.split_rank = ctx->split_rank; prm.split_world = ctx->split_world;
    prm.split_seq0 = ctx->split_seq;
    ctx->split_seq += (1ull << 24);   // every launch owns a disjoint range of flag values
    for (int r = 0; r < 16; ++r) prm.split_peers[r] = ctx->split_peers[r];
    revo_track_result out;
    revo_pyr *r = const_cast<revo_pyr *>(ref), *c = const_cast<revo_pyr *>(cur);
    int rc = run_track(ctx, prm, 1, &r, &c, R9, t3, &out, nullptr, nullptr, 0, nullptr);
    if (rc) return rc;
    if (result) *result = out;
    if (out.rc) return out.rc;
    memcpy(R9, out.R, sizeof(float) * 9);
    memcpy(t3, out.t, sizeof(float) * 3);
    return REVO_OK;
}

}  // extern "C"
