// Discovers the thread <-> (lane, column) mapping of the tcgen05.ld shapes by storing a tag (lane << 16 | column) with
// tcgen05.st.32x32b (thread i <-> lane i) and reading it back with 16x64b / 16x128b / 16x256b.  The mismatch between
// the store and the load shape is an intra-warp data exchange through tensor memory ("what comes next").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_shapes tools/tmem_shapes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(uint32_t *out) {
    __shared__ uint32_t base_s;
    const int tid = threadIdx.x;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = base_s;
    // thread i of warp 0 owns lane i: columns 0..15 <- tag
    uint32_t r[16];
    for (int c = 0; c < 16; ++c) r[c] = ((uint32_t)tid << 16) | (uint32_t)c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(base),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t a0, a1, a2, a3;
    // 16x256b.x1 at lanes 0-15 and 16-31
    for (int half = 0; half < 2; ++half) {
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(base + ((uint32_t)(16 * half) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t *o = out + (0 * 2 + half) * 32 * 4 + tid * 4;
        o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
    }
    for (int half = 0; half < 2; ++half) {
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(base + ((uint32_t)(16 * half) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t *o = out + (1 * 2 + half) * 32 * 4 + tid * 4;
        o[0] = a0; o[1] = a1; o[2] = 0xffffffffu; o[3] = 0xffffffffu;
    }
    for (int half = 0; half < 2; ++half) {
        asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(a0) : "r"(base + ((uint32_t)(16 * half) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t *o = out + (2 * 2 + half) * 32 * 4 + tid * 4;
        o[0] = a0; o[1] = o[2] = o[3] = 0xffffffffu;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory");
}

int main() {
    uint32_t *d, h[3 * 2 * 32 * 4];
    cudaMalloc(&d, sizeof h);
    cudaMemset(d, 0, sizeof h);
    k<<<1, 32>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const char *names[3] = {"16x256b.x1", "16x128b.x1", "16x64b.x1"};
    for (int s = 0; s < 3; ++s)
        for (int half = 0; half < 2; ++half) {
            printf("%s, lane base %d: thread -> (lane,column) per register\n", names[s], 16 * half);
            for (int t = 0; t < 32; ++t) {
                printf("  t%2d:", t);
                for (int r = 0; r < 4; ++r) {
                    uint32_t v = h[((s * 2 + half) * 32 + t) * 4 + r];
                    if (v != 0xffffffffu) printf(" (%2u,%2u)", v >> 16, v & 0xffff);
                }
                printf("\n");
            }
        }
    return 0;
}
