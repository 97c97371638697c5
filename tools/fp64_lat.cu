// FP64 pipe micro-benchmark for B200: issue cost and dependent latency of DFMA / DADD as seen by a kernel with
// few warps per scheduler (kernel A runs 2).  For W warps per scheduler and K independent chains per thread it
// prints cycles per warp-instruction per scheduler; K = 1, W = 1 gives the dependent latency, large K*W the
// issue interval (2 cycles = 16 lanes/clk/SMSP).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_lat tools/fp64_lat.cu && tools/fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

template <int K, bool ADD>
__global__ void chains(double *out, long long *cyc, int iters, double x) {
    double a[K];
#pragma unroll
    for (int k = 0; k < K; ++k) a[k] = x + k + threadIdx.x * 1e-3;
    const double m = 1.0000001, c = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < K; ++k) a[k] = ADD ? a[k] + c : fma(a[k], m, c);
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K, bool ADD>
void run(int warps_per_sched, double *out, long long *cyc) {
    const int iters = 2000, threads = 128 * warps_per_sched;
    chains<K, ADD><<<148, threads>>>(out, cyc, iters, 1.0);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double instr_per_sched = (double)iters * 8 * K * warps_per_sched;
    printf("%s  warps/sched %d  chains/thread %d : %.2f cycles per warp-instruction per scheduler (%.1f per dependent step)\n",
           ADD ? "DADD" : "DFMA", warps_per_sched, K, avg / instr_per_sched, avg / (iters * 8.0));
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int w = 1; w <= 4; w *= 2) {
        run<1, false>(w, out, cyc);
        run<2, false>(w, out, cyc);
        run<4, false>(w, out, cyc);
        run<8, false>(w, out, cyc);
        run<1, true>(w, out, cyc);
        run<4, true>(w, out, cyc);
    }
    return 0;
}
