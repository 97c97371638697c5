// One 1-D kinetic application on a line of N = 2^LOG2N points held 16 per thread by TPL = N/16 threads:
//   a <- ifft( kc * fft(a) )      (kc carries the k-space multiplier, any scale factor and the 1/N of the inverse)
// The same three-pass decimation-in-frequency transform and mirror inverse as kernel A (qc_grid.cuh), factored
// out for kernels whose lines are not the two levels of one run (kernel C, qc_grid2d.cuh).  Thread t of the line
// enters with a[n1] = x[t + TPL*n1] and leaves with the result at the same points; `buf` is the line's private
// transpose buffer of GridCfg<LOG2N>::BUF complex elements; line_sync() synchronises the TPL threads of the line.
#pragma once
#include "qc_grid.cuh"

namespace qc {

template <int LOG2N, class Sync>
__device__ __forceinline__ void kinetic_line(cd *a, cd *buf, int t, const double *kc, uint32_t t_tw1, uint32_t t_tw2,
                                             Sync line_sync) {
    typedef GridCfg<LOG2N> C;
    constexpr int R3 = C::R3, Q = C::Q, S1 = C::S1;
    const int k1p = t / R3, n3p = t % R3;
    {
        TmemTwiddles tw(t_tw1);
        dft16_table_out<false>(a, tw, [&](int k1, cd val) { buf[k1 * S1 + t] = val; });
    }
    line_sync();
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) a[n2] = buf[k1p * S1 + n2 * R3 + n3p];
    if (R3 == 1) dft16<false>(a);
    if (R3 > 1) {
        __syncwarp();
        {
            TmemTwiddles tw(t_tw2);
            dft16_table_out<false>(a, tw, [&](int k2, cd val) { buf[k1p * S1 + pos2<R3, Q>(k2, n3p)] = val; });
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 16; ++e) a[e] = buf[k1p * S1 + pos2<R3, Q>(n3p * Q + e / R3, e % R3)];
#pragma unroll
        for (int q = 0; q < Q; ++q) dft<R3, false>(a + q * R3);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) a[e] = cscale(a[e], kc[e]);
    if (R3 > 1) {
#pragma unroll
        for (int q = 0; q < Q; ++q) dft<R3, true>(a + q * R3);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 16; ++e) buf[k1p * S1 + pos2<R3, Q>(n3p * Q + e / R3, e % R3)] = a[e];
        __syncwarp();
        {
            TmemTwiddles tw(t_tw2);
            dft16_table_in_inv(a, tw, [&](int k2) { return buf[k1p * S1 + pos2<R3, Q>(k2, n3p)]; });
        }
        __syncwarp();
    }
    if (R3 == 1) dft16<true>(a);
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) buf[k1p * S1 + n2 * R3 + n3p] = a[n2];
    line_sync();
    {
        TmemTwiddles tw(t_tw1);
        dft16_table_in_inv(a, tw, [&](int k1) { return buf[k1 * S1 + t]; });
    }
    line_sync();       // the buffer may be rewritten by the next call
}

// index of the frequency that element e of thread t holds between the forward and the inverse transform
template <int LOG2N>
__device__ __forceinline__ int line_freq(int t, int e) {
    typedef GridCfg<LOG2N> C;
    const int k1p = t / C::R3, n3p = t % C::R3;
    const int q = e / C::R3, k3 = e % C::R3;
    return k1p + 16 * (n3p * C::Q + q) + 256 * k3;
}

}  // namespace qc
