// Does the FP64 tensor-core instruction (DMMA, mma.sync.m8n8k4.f64) run on a pipe of its own on B200, beside the
// FP64 FMA pipe?  Three kernels with W warps per scheduler: DFMA chains only, DMMA chains only, and both interleaved.
// If the mixed kernel takes max(t_dfma, t_dmma) the pipes are separate (and a transform could put part of its
// butterflies on the tensor pipe); if it takes the sum they share the FP64 units.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_probe tools/dmma_probe.cu && tools/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int NF, int NM>
__global__ void mix(double *out, long long *cyc, int iters, double x) {
    double f[NF > 0 ? NF : 1];
    double m0[NM > 0 ? NM : 1], m1[NM > 0 ? NM : 1];
#pragma unroll
    for (int k = 0; k < NF; ++k) f[k] = x + k + threadIdx.x * 1e-3;
#pragma unroll
    for (int k = 0; k < NM; ++k) { m0[k] = x * 1e-3 + k; m1[k] = x * 2e-3 + k; }
    const double mu = 1.0000001, c = 1e-9, a = 1e-3 * (threadIdx.x & 7), b = 1e-3;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < NF; ++k) f[k] = fma(f[k], mu, c);
#pragma unroll
            for (int k = 0; k < NM; ++k) dmma(m0[k], m1[k], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < NF; ++k) s += f[k];
#pragma unroll
    for (int k = 0; k < NM; ++k) s += m0[k] + m1[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NF, int NM>
void run(int wps, double *out, long long *cyc) {
    const int iters = 1000, threads = 128 * wps;
    mix<NF, NM><<<148, threads>>>(out, cyc, iters, 1.0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double per_iter = avg / (iters * 4.0);       // cycles per inner round per scheduler
    printf("warps/sched %d  DFMA/round %d  DMMA/round %d : %.1f cycles per round (%.2f per DFMA-equivalent: a DMMA = 8 FMA per thread)\n",
           wps, NF, NM, per_iter, per_iter / (wps * (NF + 8.0 * NM)));
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int w = 1; w <= 4; w *= 2) {
        run<8, 0>(w, out, cyc);
        run<0, 1>(w, out, cyc);
        run<0, 4>(w, out, cyc);
        run<8, 1>(w, out, cyc);
        run<8, 4>(w, out, cyc);
        run<16, 2>(w, out, cyc);
    }
    return 0;
}
