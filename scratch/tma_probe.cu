#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, int bw, int bh, int x0, int y0, unsigned *out)
{
    extern __shared__ __align__(1024) uint8_t g[];
    __shared__ alignas(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bw * bh) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(g)), "l"(&tm), "r"(x0), "r"(y0), "r"(s32(&bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(s32(g)), "l"(&tm), "r"(x0), "r"(y0), "r"(0), "r"(s32(&bar)) : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)), "r"(0) : "memory");
    } while (!ok);
    unsigned s = 0;
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) s += g[i];
    atomicAdd(out, s);
}
int main(int argc, char **argv)
{
    int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), x0 = atoi(argv[4]), y0 = atoi(argv[5]);
    int w = 640, h = 480;
    uint8_t *d; cudaMalloc(&d, w * h); 
    uint8_t *hbuf = (uint8_t *)malloc(w * h); for (int i = 0; i < w * h; ++i) hbuf[i] = (uint8_t)(i % 7 + 1);
    cudaMemcpy(d, hbuf, w * h, cudaMemcpyHostToDevice);
    unsigned *out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    PFN fn = (PFN)p;
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, 1}; cuuint64_t strides[2] = {(cuuint64_t)w, (cuuint64_t)w * h};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}; cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d box %dx%d at (%d,%d): encode=%d ", rank, bw, bh, x0, y0, (int)r);
    if (rank == 2) k<2><<<1, 128, bw * bh>>>(tm, bw, bh, x0, y0, out); else k<3><<<1, 128, bw * bh>>>(tm, bw, bh, x0, y0, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned res = 0; cudaMemcpy(&res, out, 4, cudaMemcpyDeviceToHost);
    unsigned expect = 0;
    for (int yy = 0; yy < bh; ++yy) for (int xx = 0; xx < bw; ++xx) { int gx = x0 + xx, gy = y0 + yy; if (gx >= 0 && gx < w && gy >= 0 && gy < h) expect += hbuf[gy * w + gx]; }
    printf("kernel=%s sum=%u expect=%u\n", cudaGetErrorString(e), res, expect);
    return 0;
}
