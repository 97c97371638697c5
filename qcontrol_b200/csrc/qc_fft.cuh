// Register-resident small DFTs used by the shared-memory FFT of the grid kernel.
//
// Everything here works on arrays whose indices are compile-time constants after
// unrolling, so the data lives in registers.  Sign convention follows numpy.fft
// (the reference's only FFT, phys_base.py:32-34): forward = exp(-2*pi*i*n*k/N),
// inverse = exp(+...), and the 1/N of the inverse is folded into the k-space
// multiplier by the caller.
#pragma once
#include <cuda_runtime.h>

namespace qc {

typedef double2 cd;

__host__ __device__ __forceinline__ cd cmk(double re, double im) { return make_double2(re, im); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return cmk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return cmk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cmul(cd a, cd b) {
    return cmk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ cd cmulc(cd a, cd b) {
    return cmk(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ cd cscale(cd a, double s) { return cmk(a.x * s, a.y * s); }
// acc + a*b
__device__ __forceinline__ cd cfma(cd a, cd b, cd acc) {
    return cmk(fma(a.x, b.x, fma(-a.y, b.y, acc.x)), fma(a.x, b.y, fma(a.y, b.x, acc.y)));
}
// a / b, correctly rounded, from rb = RN(1 / b): q = RN(a rb), r = a - b q (one FMA), RN(q + r rb) is the
// correctly rounded quotient (Markstein 1990; the one exception, a significand of b of all ones, has probability
// 2^-52).  Three dependent FMAs and no branch, against the ~15-instruction division routine with its slow-path
// branch: where one divisor serves many dividends (renormalisation, rotating frame, the field law) the reference's
// `x / s` is reproduced bit for bit at a fraction of the latency.
__device__ __forceinline__ double div_rn_by(double a, double b, double rb) {
    const double q = a * rb;
    const double r = fma(-q, b, a);
    return fma(r, rb, q);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ cd mul_mi(cd a) {
    return INV ? cmk(-a.y, a.x) : cmk(a.y, -a.x);
}
// multiply by the forward twiddle w (given as forward value); conjugate it for the inverse
template <bool INV>
__device__ __forceinline__ cd twmul(cd a, cd w) { return INV ? cmulc(a, w) : cmul(a, w); }

#define QC_SQRT1_2 0.70710678118654752440
#define QC_COS_PI_8 0.92387953251128675613
#define QC_SIN_PI_8 0.38268343236508977173

// multiply by w8^1 = (1 - i)/sqrt2 (forward) or its conjugate
template <bool INV>
__device__ __forceinline__ cd mul_w8_1(cd a) {
    return INV ? cmk((a.x - a.y) * QC_SQRT1_2, (a.x + a.y) * QC_SQRT1_2)
               : cmk((a.x + a.y) * QC_SQRT1_2, (a.y - a.x) * QC_SQRT1_2);
}
// multiply by w8^3 = (-1 - i)/sqrt2 (forward) or its conjugate
template <bool INV>
__device__ __forceinline__ cd mul_w8_3(cd a) {
    return INV ? cmk((-a.x - a.y) * QC_SQRT1_2, (a.x - a.y) * QC_SQRT1_2)
               : cmk((a.y - a.x) * QC_SQRT1_2, (-a.x - a.y) * QC_SQRT1_2);
}

template <bool INV>
__device__ __forceinline__ void dft2(cd &a0, cd &a1) {
    cd t = a0;
    a0 = cadd(t, a1);
    a1 = csub(t, a1);
}

template <bool INV>
__device__ __forceinline__ void dft4(cd &a0, cd &a1, cd &a2, cd &a3) {
    cd t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

// 8-point DFT, natural order in and out (stride-1 register array of 8).
template <bool INV>
__device__ __forceinline__ void dft8(cd *a) {
    // n = 4*na + nb ; k = ka + 2*kb
    dft2<INV>(a[0], a[4]);
    dft2<INV>(a[1], a[5]);
    dft2<INV>(a[2], a[6]);
    dft2<INV>(a[3], a[7]);
    a[5] = mul_w8_1<INV>(a[5]);
    a[6] = mul_mi<INV>(a[6]);
    a[7] = mul_w8_3<INV>(a[7]);
    dft4<INV>(a[0], a[1], a[2], a[3]);  // ka = 0 : X[0], X[2], X[4], X[6]
    dft4<INV>(a[4], a[5], a[6], a[7]);  // ka = 1 : X[1], X[3], X[5], X[7]
    cd x1 = a[4], x2 = a[1], x3 = a[5], x4 = a[2], x5 = a[6], x6 = a[3];
    a[1] = x1;
    a[2] = x2;
    a[3] = x3;
    a[4] = x4;
    a[5] = x5;
    a[6] = x6;
}

// 16-point DFT, natural order in and out.
template <bool INV>
__device__ __forceinline__ void dft16(cd *a) {
    // n = 4*na + nb ; k = ka + 4*kb
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) dft4<INV>(a[nb], a[4 + nb], a[8 + nb], a[12 + nb]);
    // a[4*ka + nb] *= w16^(ka*nb)
    const cd w1 = cmk(QC_COS_PI_8, -QC_SIN_PI_8);   // w16^1
    const cd w3 = cmk(QC_SIN_PI_8, -QC_COS_PI_8);   // w16^3
    const cd w9 = cmk(-QC_COS_PI_8, QC_SIN_PI_8);   // w16^9
    a[5] = twmul<INV>(a[5], w1);
    a[6] = mul_w8_1<INV>(a[6]);                      // w16^2
    a[7] = twmul<INV>(a[7], w3);
    a[9] = mul_w8_1<INV>(a[9]);                      // w16^2
    a[10] = mul_mi<INV>(a[10]);                      // w16^4
    a[11] = mul_w8_3<INV>(a[11]);                    // w16^6
    a[13] = twmul<INV>(a[13], w3);
    a[14] = mul_w8_3<INV>(a[14]);                    // w16^6
    a[15] = twmul<INV>(a[15], w9);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) dft4<INV>(a[4 * ka], a[4 * ka + 1], a[4 * ka + 2], a[4 * ka + 3]);
    // X[ka + 4*kb] sits in a[4*ka + kb]: transpose the 4x4 register tile
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            cd t = a[4 * i + j];
            a[4 * i + j] = a[4 * j + i];
            a[4 * j + i] = t;
        }
}


// constant twiddles w16^(ka*nb) between the two radix-4 stages of the 16-point DFT (element a[4*ka + nb])
template <bool INV>
__device__ __forceinline__ void dft16_internal_twiddles(cd *a) {
    const cd w1 = cmk(QC_COS_PI_8, -QC_SIN_PI_8);   // w16^1
    const cd w3 = cmk(QC_SIN_PI_8, -QC_COS_PI_8);   // w16^3
    const cd w9 = cmk(-QC_COS_PI_8, QC_SIN_PI_8);   // w16^9
    a[5] = twmul<INV>(a[5], w1);
    a[6] = mul_w8_1<INV>(a[6]);
    a[7] = twmul<INV>(a[7], w3);
    a[9] = mul_w8_1<INV>(a[9]);
    a[10] = mul_mi<INV>(a[10]);
    a[11] = mul_w8_3<INV>(a[11]);
    a[13] = twmul<INV>(a[13], w3);
    a[14] = mul_w8_3<INV>(a[14]);
    a[15] = twmul<INV>(a[15], w9);
}

// ---- 16-point DFTs streamed with the shared-memory traffic around them, twiddles w^k from a per-thread table.
// ---- TwSrc::prefetch(g) starts fetching the four twiddles {w^g, w^(g+4), w^(g+8), w^(g+12)} of radix-4 group g,
// ---- TwSrc::get(w4) waits for them (tensor-memory table, see qc_grid.cuh).  Working group by group keeps
// ---- register live ranges short: each group's outputs leave for shared memory as soon as they exist.
template <bool INV, class TwSrc, class Sink>
__device__ __forceinline__ void dft16_table_out(cd *a, TwSrc &tw, Sink sink) {
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) dft4<INV>(a[nb], a[4 + nb], a[8 + nb], a[12 + nb]);
    dft16_internal_twiddles<INV>(a);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) {
        tw.prefetch(ka);
        dft4<INV>(a[4 * ka], a[4 * ka + 1], a[4 * ka + 2], a[4 * ka + 3]);   // X[ka + 4 kb] in a[4 ka + kb]
        cd w4[4];
        tw.get(w4);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
            if (ka == 0 && kb == 0)
                sink(0, a[0]);
            else
                sink(ka + 4 * kb, cmul(a[4 * ka + kb], w4[kb]));
        }
    }
}

// inverse 4-point DFT of y_i * conj(w_i) with the twiddles of inputs 2 and 3 fused into the first butterfly
// stage: t0 = x0 + y2 conj(w2) (four FMAs), t1 = 2 x0 - t0 (two FMAs) instead of a complex multiply and two
// complex additions -- four instructions fewer per group.  (The same fusion inside the constant-twiddle
// stage of dft16 was measured slower: it lengthens the dependency chains of a latency-bound kernel.)
template <bool FIRST_IS_ONE>
__device__ __forceinline__ void dft4_inv_conj_twiddled(cd &y0, cd &y1, cd &y2, cd &y3, const cd *w) {
    const cd x0 = FIRST_IS_ONE ? y0 : cmulc(y0, w[0]);
    const cd x1 = cmulc(y1, w[1]);
    cd t0, t1, t2, u;
    t0.x = fma(y2.x, w[2].x, fma(y2.y, w[2].y, x0.x));
    t0.y = fma(y2.y, w[2].x, fma(-y2.x, w[2].y, x0.y));
    t1.x = fma(2.0, x0.x, -t0.x);
    t1.y = fma(2.0, x0.y, -t0.y);
    t2.x = fma(y3.x, w[3].x, fma(y3.y, w[3].y, x1.x));
    t2.y = fma(y3.y, w[3].x, fma(-y3.x, w[3].y, x1.y));
    u.x = fma(2.0, x1.x, -t2.x);
    u.y = fma(2.0, x1.y, -t2.y);
    const cd t3 = mul_mi<true>(u);
    y0 = cadd(t0, t2);
    y2 = csub(t0, t2);
    y1 = cadd(t1, t3);
    y3 = csub(t1, t3);
}

template <class TwSrc, class Source>
__device__ __forceinline__ void dft16_table_in_inv(cd *a, TwSrc &tw, Source source) {
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
        tw.prefetch(nb);
        cd x0 = source(nb), x1 = source(nb + 4), x2 = source(nb + 8), x3 = source(nb + 12);
        cd w4[4];
        tw.get(w4);
        if (nb == 0)
            dft4_inv_conj_twiddled<true>(x0, x1, x2, x3, w4);
        else
            dft4_inv_conj_twiddled<false>(x0, x1, x2, x3, w4);
        a[nb] = x0;
        a[4 + nb] = x1;
        a[8 + nb] = x2;
        a[12 + nb] = x3;
    }
    dft16_internal_twiddles<true>(a);
#pragma unroll
    for (int ka = 0; ka < 4; ++ka) dft4<true>(a[4 * ka], a[4 * ka + 1], a[4 * ka + 2], a[4 * ka + 3]);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            cd t = a[4 * i + j];
            a[4 * i + j] = a[4 * j + i];
            a[4 * j + i] = t;
        }
}

template <int R, bool INV>
__device__ __forceinline__ void dft(cd *a) {
    if (R == 2) dft2<INV>(a[0], a[1]);
    if (R == 4) dft4<INV>(a[0], a[1], a[2], a[3]);
    if (R == 8) dft8<INV>(a);
    if (R == 16) dft16<INV>(a);
}

}  // namespace qc
