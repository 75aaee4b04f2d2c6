"""Test infrastructure: the WHOLE library on the CPU.

capi.cu (the C ABI: slab layout, descriptor tables, launch sequencing, batch / keyframe / tracking / quality entry points),
pyramid.cu, the bit-mask Canny of canny.cu and the cluster tracking engine of track.cu are compiled with g++ into
``librevo_b200_emu.so`` with the same exported symbols as the CUDA library: the kernels run on the emulation layer of
``_cuda_emu.py``, the CUDA runtime calls of the host code go to a fake runtime (device memory = host memory, streams and events
are no-ops, cudaLaunchKernelEx starts emulated thread-block clusters).  Tests load it through ``revo_b200.api`` by pointing
``api._LIB_PATH`` at it -- the product itself never does (it has no CPU path).  Not built: the tile / TMA Canny fallback, the
task-queue and ping-pong engines, the multi-GPU split (they return REVO_ERR_UNSUPPORTED / fail here)."""
import ctypes as C
import os
import re
import subprocess

import _cuda_emu as E

ROOT = E.ROOT

FAKE_CUDA = r'''
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
'''

STUBS = r'''
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
int launch_track_lean(revo_ctx *ctx, const PairDesc *d_pairs, int n_pairs, const TrackParams &prm, revo_track_result *d_results,
                      double *d_records, revo_trace_entry *d_trace, int *d_trace_counts, int *d_work_counter);   // scratch/experiments/track_lean.cu
#ifndef EMU_WITH_LEAN
int launch_track_lean(revo_ctx *ctx, const PairDesc *, int, const TrackParams &, revo_track_result *, double *, revo_trace_entry *, int *, int *)
{
    ctx->last_error = "lean engine: not in this build";
    return REVO_ERR_UNSUPPORTED;
}
#endif
bool make_gray_tensor_map(void *, const uint8_t *, int, int, int, size_t) { return false; }
}
'''


def _strip_includes(text):
    return re.sub(r'^#include [<"][^\n]*\n', "", text, flags=re.M).replace("#pragma once", "").replace("#pragma unroll", "")


def build(out_dir, with_lean=False):
    """with_lean: also compile scratch/experiments/track_lean.cu and dispatch tracking engine 4 to it (what
    scratch/experiments/enable_lean_engine.patch does to the library)."""
    rd = lambda *p: open(os.path.join(ROOT, *p)).read()      # noqa: E731
    internal, common = rd("revo_b200", "csrc", "internal.h"), rd("revo_b200", "csrc", "track_common.cuh")
    pyr, canny, track, capi = (rd("revo_b200", "csrc", f) for f in ("pyramid.cu", "canny.cu", "track.cu", "capi.cu"))

    common = E._strip_functions(common, ["ldg_quad", "rcp_approx", "smem_u32", "mbar_init", "mbar_expect_tx", "mbar_wait", "st_async_b64"])
    # canny.cu: the bit-mask pipeline only
    a = canny.index("// counts -> wrapping u8 histogram")
    cbody = E._strip_functions(canny[a:], ["dp4a_us"])
    i = cbody.index("int launch_canny(revo_ctx *ctx")
    j = cbody.index("\n", cbody.index("return launch_canny_bits(", i))
    k = cbody.index("\n}\n", j)
    cbody = cbody[:j] + "\n    ctx->last_error = \"tile / TMA Canny fallback: not in the emulated build\";\n    return REVO_ERR_UNSUPPORTED;" + cbody[k:]
    cbody = "namespace revo {\n" + cbody
    # track.cu: the PTX helpers of the multi-GPU mailboxes are shims
    track = E._strip_functions(track, ["st_release_sys", "ld_acquire_sys"])
    track = track.replace("extern __shared__ float s_pts[];", "float *s_pts = emu::cta->dyn.data();")
    track = re.sub(r'\n[^\n]*asm volatile\("fence\.mbarrier_init[^\n]*\n', "\n", track)
    track = track.replace("__threadfence_system();", "")

    files = [common, pyr, cbody, track]
    if with_lean:
        lean = rd("scratch", "experiments", "track_lean.cu")
        lean = re.sub(r"#define REVO_LDG_QUAD.*?#undef REVO_LDG_QUAD\n", "", lean, flags=re.S)
        lean = E._strip_functions(lean, ["lds3", "sts3", "ffma2", "fmul2", "pin"])
        lean = lean.replace("extern __shared__ float s_pts[];", "float *s_pts = emu::cta->dyn.data();")
        lean = re.sub(r'\n[^\n]*asm volatile\("fence\.mbarrier_init[^\n]*\n', "\n", lean)
        files.append(lean)
        capi = capi.replace("engine < 0 || engine > 3 ||", "engine < 0 || engine > 4 ||")
        capi = capi.replace("    if (engine != 2 && engine != 3) engine = 1;", "    if (engine != 2 && engine != 3 && engine != 4) engine = 1;")
        capi = capi.replace("        else if (engine == 3) rc = launch_track_pp(ctx, d_pairs, n, prm, d_res, d_rec, d_tr, d_tc, d_wc);\n",
                            "        else if (engine == 3) rc = launch_track_pp(ctx, d_pairs, n, prm, d_res, d_rec, d_tr, d_tc, d_wc);\n"
                            "        else if (engine == 4) rc = launch_track_lean(ctx, d_pairs, n, prm, d_res, d_rec, d_tr, d_tc, d_wc);\n")
        assert "launch_track_lean" in capi
    device = "\n".join(_strip_includes(t) for t in files)
    device = E._launches(E._device_text(device))
    assert "asm" not in device and "<<<" not in device, "unexpected PTX / launch syntax left"
    device = re.sub(r"(\.|->)(gridDim|blockDim)\b", r"\1\2_", device)      # members of cudaLaunchConfig_t, not the built-ins
    host = _strip_includes(capi)
    generic = E.RUNNER[E.RUNNER.index("// @GENERIC_BEGIN"):E.RUNNER.index("// @GENERIC_END")]
    dp4a = E.CANNY_SHIMS[E.CANNY_SHIMS.index("namespace revo {"):E.CANNY_SHIMS.index("struct revo_ctx")]
    src_text = (E.PRELUDE + generic + FAKE_CUDA + dp4a + _strip_includes(internal) + STUBS.replace("namespace revo {", "namespace revo {", 1)
                + device + host)
    # the stubs need the declarations of internal.h, and come before capi.cu; the tensor-map stub replaces canny.cu's
    src, lib = os.path.join(out_dir, "revo_b200_emu.cpp"), os.path.join(out_dir, "librevo_b200_emu.so")
    open(src, "w").write(src_text)
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-mfma", "-ffp-contract=fast", "-shared", "-fPIC", "-pthread", "-fvisibility=default",
                    *(["-DEMU_WITH_LEAN"] if with_lean else []), *E.EXTRA_FLAGS, "-I", os.path.join(ROOT, "include"), src, "-o", lib], check=True)
    return lib
