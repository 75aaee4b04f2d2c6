// Micro-probe (scratch, not part of the library): what do the pieces of the LM step cost on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I revo_b200/csrc -I include scratch/probes/lm_probe.cu -o scratch/probes/lm_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "internal.h"
#include "track_common.cuh"
using namespace revo;

__global__ void k_chain(double *out, long long *cyc, int lanes, int n)
{
    const int lane = threadIdx.x & 31;
    double a = 1.0 + lane * 1e-9, b = 0.999999, c = 1e-9;
    float fa = 1.0f + lane * 1e-6f, fb = 0.99999f, fc = 1e-6f;
    long long t0, t1, t2, t3;
    __syncwarp();
    t0 = clock64();
    if (lane < lanes)
        for (int i = 0; i < n; ++i) a = fma(a, b, c);
    __syncwarp();
    t1 = clock64();
    if (lane < lanes)
        for (int i = 0; i < n; ++i) fa = fmaf(fa, fb, fc);
    __syncwarp();
    t2 = clock64();
    // 4 independent double chains (ILP 4)
    double a0 = a, a1 = a + 1, a2 = a + 2, a3 = a + 3;
    if (lane < lanes)
        for (int i = 0; i < n; ++i) { a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
    __syncwarp();
    t3 = clock64();
    if (lane == 0 && blockIdx.x == 0 && threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + fa + a0 + a1 + a2 + a3;
}

template <typename T>
__global__ void k_solve(const double *rec, T *out, long long *cyc, int reps)
{
    __shared__ double srec[32];
    __shared__ Trial tr;
    __shared__ lmreal q[4], t[3];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) srec[threadIdx.x] = rec[threadIdx.x];
    if (threadIdx.x == 0) { q[0] = q[1] = q[2] = 0; q[3] = 1; t[0] = t[1] = t[2] = 0; }
    __syncthreads();
    T x[6];
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) solve6_warp<T>(srec + kRecA, srec + kRecB, 1.0 / srec[kRecGood], (T)(1.0 + 0.2 * r), lane, x);
    __syncwarp();
    long long t1 = clock64();
    T y[6];
    if (lane == 0)
        for (int r = 0; r < reps; ++r) solve6<T>(srec + kRecA, srec + kRecB, 1.0 / srec[kRecGood], (T)(1.0 + 0.2 * r), y);
    __syncwarp();
    long long t2 = clock64();
    for (int r = 0; r < reps; ++r) lm_propose_warp(srec, q, t, 0.2f * r, tr, lane);
    __syncwarp();
    long long t3 = clock64();
    if (lane == 0)
        for (int r = 0; r < reps; ++r) lm_propose(srec, q, t, 0.2f * r, tr);
    __syncwarp();
    long long t4 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = (t1 - t0) / reps; cyc[1] = (t2 - t1) / reps; cyc[2] = (t3 - t2) / reps; cyc[3] = (t4 - t3) / reps; }
    if (lane < 6) out[lane] = x[0] + y[0] + (T)tr.R[0];
}

int main()
{
    double *d_out; long long *d_cyc; long long h[4];
    cudaMalloc(&d_out, 1 << 20); cudaMalloc(&d_cyc, 64);
    for (int lanes : {1, 2, 4, 8, 16, 32}) {
        k_chain<<<1, 32>>>(d_out, d_cyc, lanes, 1000);
        cudaMemcpy(h, d_cyc, 32, cudaMemcpyDeviceToHost);
        printf("1 warp, %2d active lanes: DFMA chain %.2f cyc/op, FFMA chain %.2f cyc/op, 4 DFMA chains %.2f cyc/op\n", lanes, h[0] / 1000.0, h[1] / 1000.0, h[2] / 4000.0);
    }
    for (int warps : {4, 16}) {
        k_chain<<<148, 32 * warps>>>(d_out, d_cyc, 1, 1000);
        cudaMemcpy(h, d_cyc, 32, cudaMemcpyDeviceToHost);
        printf("%d warps/SM, lane 0 of each active: DFMA chain %.2f cyc/op, FFMA %.2f, 4 chains %.2f\n", warps, h[0] / 1000.0, h[1] / 1000.0, h[2] / 4000.0);
        k_chain<<<148, 32 * warps>>>(d_out, d_cyc, 32, 1000);
        cudaMemcpy(h, d_cyc, 32, cudaMemcpyDeviceToHost);
        printf("%d warps/SM, all lanes active: DFMA chain %.2f cyc/op, FFMA %.2f, 4 chains %.2f\n", warps, h[0] / 1000.0, h[1] / 1000.0, h[2] / 4000.0);
    }
    // a plausible record: A = J^T J of random J, b
    double rec[32] = {0};
    {
        double A[6][6] = {{0}}, b[6] = {0};
        unsigned s = 12345;
        for (int n = 0; n < 500; ++n) {
            double J[6];
            for (int i = 0; i < 6; ++i) { s = s * 1664525u + 1013904223u; J[i] = ((s >> 8) / 16777216.0 - 0.5) * (i < 3 ? 1 : 3); }
            s = s * 1664525u + 1013904223u;
            double r = ((s >> 8) / 16777216.0 - 0.5);
            for (int i = 0; i < 6; ++i) { b[i] += J[i] * r; for (int j = 0; j < 6; ++j) A[i][j] += J[i] * J[j]; }
        }
        int k = 0;
        for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) rec[k++] = A[i][j];
        for (int i = 0; i < 6; ++i) rec[21 + i] = b[i] * 0.01;
        rec[29] = 500;
    }
    double *d_rec; cudaMalloc(&d_rec, 256); cudaMemcpy(d_rec, rec, 256, cudaMemcpyHostToDevice);
    for (int blocks : {1, 148, 592}) {
        k_solve<float><<<blocks, 32>>>(d_rec, (float *)d_out, d_cyc, 20);
        cudaMemcpy(h, d_cyc, 32, cudaMemcpyDeviceToHost);
        printf("float,  %3d CTAs x 1 warp: solve6_warp %lld, solve6 (1 lane) %lld, lm_propose_warp %lld, lm_propose (1 lane) %lld cycles\n", blocks, h[0], h[1], h[2], h[3]);
        k_solve<double><<<blocks, 32>>>(d_rec, d_out, d_cyc, 20);
        cudaMemcpy(h, d_cyc, 32, cudaMemcpyDeviceToHost);
        printf("double, %3d CTAs x 1 warp: solve6_warp %lld, solve6 (1 lane) %lld, (lm_propose* in lmreal) %lld, %lld cycles\n", blocks, h[0], h[1], h[2], h[3]);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
