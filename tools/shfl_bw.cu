// Micro-benchmark: cost of moving one complex128 per thread to a partner thread, per SM --
// (a) four SHFL.BFLY.32, (b) STS.128 + LDS.128 through shared memory -- with 8 warps per SM.
// Question it answers (kernel A): would exchanging phi between the two levels by warp shuffles
// (levels interleaved in one warp) be cheaper than the shared-memory hand-over?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o shfl_bw tools/shfl_bw.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, long long *cyc, int iters) {
    extern __shared__ double2 sm[];
    const int tid = threadIdx.x;
    double2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_double2(tid + i, tid - i);
    __syncthreads();
    long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                v[i].x = __shfl_xor_sync(0xffffffffu, v[i].x, 16);
                v[i].y = __shfl_xor_sync(0xffffffffu, v[i].y, 16);
            } else {
                sm[i * blockDim.x + tid] = v[i];
            }
        }
        if (MODE == 1) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = sm[i * blockDim.x + (tid ^ 16)];
            __syncwarp();
        }
    }
    long long c1 = clock64();
    __syncthreads();
    if (tid == 0) cyc[blockIdx.x] = c1 - c0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + tid] = s;
}

int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4000;
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 512 * 16);
    for (int mode = 0; mode < 2; ++mode)
        for (int threads : {128, 256, 512}) {
            if (mode == 0) k<0><<<148, threads, 0>>>(out, cyc, iters);
            else k<1><<<148, threads, 8 * threads * 16>>>(out, cyc, iters);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            const double moves = (double)(threads / 32) * 8 * iters;     // warp-wide complex128 exchanges per SM
            printf("%s threads=%d: %s, %.2f clk per warp-wide complex128 exchange per SM (cycles %lld)\n",
                   mode ? "STS.128+LDS.128" : "4 x SHFL.32    ", threads, cudaGetErrorString(e), (double)h[0] / moves, h[0]);
        }
    return 0;
}
