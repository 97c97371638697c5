// Micro-benchmark: tcgen05.ld / tcgen05.st throughput per SM (bytes per clock) with 4/8/16 warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tools/tmem_bw.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld16(uint32_t a, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a));
}
__device__ __forceinline__ void st16(uint32_t a, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

template <int MODE>   // 0 = ld, 1 = st
__global__ void k(unsigned *out, long long *cyc, int iters) {
    __shared__ uint32_t base_s;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = base_s;
    const int warp = threadIdx.x >> 5;
    const uint32_t t0 = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    uint32_t r[16];
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x + i;
    for (int c = 0; c < 128; c += 16) st16(t0 + c, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    unsigned acc = 0;
    long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 128; c += 16) {
            if (MODE == 0) {
                ld16(t0 + c, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += r[0] ^ r[15];
            } else {
                r[0] = acc + it;
                st16(t0 + c, r);
            }
        }
        if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    long long c1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = c1 - c0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + r[3];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

int main() {
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int threads : {128, 256, 512}) {
            if (threads == 512) continue;   // 4 warp slots per quadrant would need 512 columns per slot pair; keep 2
            if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters); else k<1><<<148, threads>>>(out, cyc, iters);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            double bytes = (double)threads * 128 * 4 * iters;
            printf("%s threads=%d: %s, %.1f bytes/clk/SM (cycles %lld)\n", mode ? "tcgen05.st" : "tcgen05.ld", threads,
                   cudaGetErrorString(e), bytes / (double)h[0], h[0]);
        }
    return 0;
}
