// Kernel A, 512-thread form for np = 2048 ("K8"): the same fused Newton-polynomial time stepper as grid_sweep_kernel
// (qc_grid.cuh; Hamil2DNonTrivial, hamil_2d.py:34-73; prop_cpu, phys_base.py:129-178), laid out for FOUR warps per
// scheduler instead of two.
//
// EXPERIMENTAL (QC_GRID_VAR=2), kept as the measured answer to "why not more warps": parity-green and SLOWER than the
// 256-thread form (0.549 against 0.615 of the FP64 peak).
//
// Why it was built.  The 256-thread forms hold 16 grid points per thread in 255 registers, i.e. two warps per
// scheduler, and keep the FP64 pipe 72 % busy whatever the grid size or the barrier structure (profiles/r02b_*).
// Here a thread owns EIGHT points (124 registers), a level is eight warps, a CTA sixteen.  Eight points per thread
// mean a four-pass transform (8 x 8 x 8 x 4) with three exchanges per direction instead of two; shared memory could
// not carry that (12 accesses per point and order against 8), so two of
// the three exchanges go through TENSOR MEMORY, whose load / store shapes differ in how they map threads to lanes:
//   tcgen05.st.32x32b   thread i  <-> lane i, consecutive columns
//   tcgen05.ld.16x256b  thread 4a + b <-> lanes a and a + 8, column pairs 2b, 2b+1 of every eight columns
// (measured map: tools/tmem_shapes.cu, profiles/r01k_tmem_ld_shapes.log).  Storing with one shape and loading with the
// other moves lane bits 3-4 into the register index and two column bits into lane bits 0-1; and because warps w and
// w + 4 of a level share a lane quadrant, reading the partner's columns moves the third bit.
//
// Index maps (N = 2048, n = 256 n1 + 32 n2 + 4 n3 + n4, k = k1 + 8 k2 + 64 k3 + 512 k4; thread T = 32 W + L of a level,
// W = 4 hb + q: q = lane quadrant, hb = which warp of the quadrant's pair):
//   x layout      registers n1;            thread: n2 = W, 4 n3 + n4 = L                      (x = T + 256 n1)
//   pass 1 over n1, twiddle w_2048^(k1 T);  exchange X1 through shared memory (the only one that crosses quadrants)
//   layout 2      registers n2;            thread: k1 = 2 q + (L & 1), n3 = 4 hb + (L >> 3), n4 = (L >> 1) & 3
//   pass 2 over n2, twiddle w_256^(k2 (4 n3 + n4));  exchange X2 through tensor memory (warp pair of a quadrant)
//   layout 3      registers n3;            thread: k2 = 4 hb + (L & 3), k1 bit 0 = (L >> 2) & 1, n4 = L >> 3
//   pass 3 over n3, twiddle w_32^(k3 n4);   exchange X3 through tensor memory (inside the warp)
//   layout 4      registers (k3 >> 2, n4); thread: k3 & 3 = L & 3, k2 & 3 = (L >> 2) & 3, k1 bit 0 = L >> 4
//   pass 4 over n4 (two 4-point transforms), kinetic multiplier on the permuted layout, and the exact mirror back.
//
// Tensor memory (512 columns per lane, four warp slots (level, hb) per quadrant): [0,128) psi, [128,384) phi of even /
// odd polynomial orders -- thread T of the other level sits on the same lane and reads it in place, as in the TMH form
// of qc_grid.cuh --, [384,512) exchange staging, 64 columns per level and quadrant.
// Shared memory: one X1 buffer per level (33 KB; the inverse writes exactly where the same thread read, so one buffer
// and two level barriers per order suffice), the three twiddle tables (36 KB), d(x) and the kinetic multiplier
// (48 KB), the TMA-staged trajectory slice.
//
// Covers what the bench path needs: prescribed and Krotov field laws, rotating frame, renormalisation, trajectory in /
// out.  Local control, the one-step / operator entry points and small batches keep the 256-thread forms.
//
// What it showed (profiles/r02k_*, r02u_*): more warps cover latency, they do not create issue slots.  A scheduler issues
// one instruction per cycle and does not prefer FP64 work; this form needs 0.54 non-FP64 instructions per FP64
// instruction (table loads from shared memory, register moves into sixteen-register tcgen05 operands, per-thread fixed
// costs over eight points) against 0.34 of the 256-thread form, and its pipe is 65 % busy with ready warps waiting.
#pragma once
#include "qc_grid.cuh"

namespace qc {

struct Grid8Cfg {
    static constexpr int N = 2048, TPL = 256, NTHREADS = 512, PPT = 8;
    static constexpr int KS = 260;                  // k1 stride of the X1 buffer, padded: reads of a quarter warp hit 8 distinct 16-byte banks
    static constexpr int BUF = 8 * KS;
    static constexpr int MAXCH = 128;
    static constexpr size_t OFF_TW1 = sizeof(cd) * 2 * BUF;              // [8][256]  w_2048^(k1 T)
    static constexpr size_t OFF_TW2 = OFF_TW1 + sizeof(cd) * 8 * 256;    // [8][32]   w_256^(k2 m2)
    static constexpr size_t OFF_TW3 = OFF_TW2 + sizeof(cd) * 8 * 32;     // [8][4]    w_32^(k3 n4)
    static constexpr size_t OFF_DV = OFF_TW3 + sizeof(cd) * 8 * 4;
    static constexpr size_t OFF_XP = OFF_DV + sizeof(cd) * MAXCH;
    static constexpr size_t OFF_RED = OFF_XP + sizeof(double) * MAXCH;   // 32 doubles
    static constexpr size_t OFF_KC = OFF_RED + sizeof(double) * 32;      // [8][256]
    static constexpr size_t OFF_DC = OFF_KC + sizeof(double) * 8 * 256;  // [2][8][256]
    static constexpr size_t OFF_SLICE = (OFF_DC + sizeof(double) * 2 * 8 * 256 + 127) / 128 * 128;
    static constexpr size_t SMEM_BYTES = OFF_SLICE + sizeof(cd) * N + 16;
    static constexpr int TW_TOTAL = 8 * 256 + 8 * 32 + 8 * 4;            // complex elements of GridParams::tw8
};

#define QC_R16(r) "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
                  "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
#define QC_W16(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])

__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : QC_W16(r)
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        QC_R16(r)
        : "memory");
}
__device__ __forceinline__ uint32_t dlo(double x) { return (uint32_t)__double2loint(x); }
__device__ __forceinline__ uint32_t dhi(double x) { return (uint32_t)__double2hiint(x); }
__device__ __forceinline__ double mkd(uint32_t lo, uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }

// Stores go out one value at a time: a double is a natural register pair and a complex128 a natural quad, whereas the
// sixteen-register operand of an x16 store is filled by sixteen register moves (ptxas pins it to a window).
__device__ __forceinline__ void tmem_st_d(uint32_t taddr, double x) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(dlo(x)), "r"(dhi(x)) : "memory");
}
__device__ __forceinline__ void tmem_st_c(uint32_t taddr, cd v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(dlo(v.x)), "r"(dhi(v.x)), "r"(dlo(v.y)),
                 "r"(dhi(v.y))
                 : "memory");
}
#ifndef QC_K8_ST
#define QC_K8_ST 3
#endif
// eight complex128 values <-> 32 columns of the thread's own lane, four columns per value
__device__ __forceinline__ void tmem_store_c8(uint32_t taddr, const cd *v) {
    if (QC_K8_ST & 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) tmem_st_c(taddr + 4 * i, v[i]);
    } else {
        tmem_store_c4(taddr, v);
        tmem_store_c4(taddr + 16, v + 4);
    }
}
__device__ __forceinline__ void tmem_store_c4x(uint32_t taddr, const cd *v) {
    if (QC_K8_ST & 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) tmem_st_c(taddr + 4 * i, v[i]);
    } else {
        tmem_store_c4(taddr, v);
    }
}
// four complex values in the split layout of the exchanges: columns [re0 re1 re2 re3 im0 im1 im2 im3], two each
__device__ __forceinline__ void tmem_store_split4(uint32_t taddr, const cd *v) {
    if (QC_K8_ST & 2) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            tmem_st_d(taddr + 2 * c, v[c].x);
            tmem_st_d(taddr + 8 + 2 * c, v[c].y);
        }
    } else {
        uint32_t r[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            r[2 * c] = dlo(v[c].x); r[2 * c + 1] = dhi(v[c].x);
            r[8 + 2 * c] = dlo(v[c].y); r[8 + 2 * c + 1] = dhi(v[c].y);
        }
        tmem_st16(taddr, r);
    }
}

// inverse 8-point DFT of y_i * conj(w_i), w_0 = 1, natural order in and out.  The twiddles of inputs 4..7 are fused
// into the first butterfly stage (t = x_i + y_{i+4} conj(w_{i+4}): four FMAs, x_i - ... = 2 x_i - t: two FMAs), like
// dft4_inv_conj_twiddled of qc_fft.cuh.
__device__ __forceinline__ void idft8_conj_twiddled(cd *y, const cd *w) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const cd x = i == 0 ? y[0] : cmulc(y[i], w[i]);
        const cd u = y[i + 4], wu = w[i + 4];
        cd t;
        t.x = fma(u.x, wu.x, fma(u.y, wu.y, x.x));
        t.y = fma(u.y, wu.x, fma(-u.x, wu.y, x.y));
        y[i] = t;
        y[i + 4] = cmk(fma(2.0, x.x, -t.x), fma(2.0, x.y, -t.y));
    }
    y[5] = mul_w8_1<true>(y[5]);
    y[6] = mul_mi<true>(y[6]);
    y[7] = mul_w8_3<true>(y[7]);
    dft4<true>(y[0], y[1], y[2], y[3]);
    dft4<true>(y[4], y[5], y[6], y[7]);
    const cd x1 = y[4], x2 = y[1], x3 = y[5], x4 = y[2], x5 = y[6], x6 = y[3];
    y[1] = x1; y[2] = x2; y[3] = x3; y[4] = x4; y[5] = x5; y[6] = x6;
}

__global__ void __launch_bounds__(Grid8Cfg::NTHREADS, 1) grid8_sweep_kernel(GridParams p) {
    typedef Grid8Cfg C;
    constexpr int N = C::N, TPL = C::TPL, KS = C::KS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int lev = tid >> 8, T = tid & 255, W = T >> 5, L = tid & 31;
    const int q = W & 3, hb = W >> 2;
    if ((int)blockIdx.x >= p.run_count || p.run_first + (int)blockIdx.x >= p.batch) return;
    const int run = p.run_first + (int)blockIdx.x;

    cd *buf = reinterpret_cast<cd *>(smem_raw) + lev * C::BUF;
    cd *s_tw1 = reinterpret_cast<cd *>(smem_raw + C::OFF_TW1);
    cd *s_tw2 = reinterpret_cast<cd *>(smem_raw + C::OFF_TW2);
    cd *s_tw3 = reinterpret_cast<cd *>(smem_raw + C::OFF_TW3);
    cd *s_dv = reinterpret_cast<cd *>(smem_raw + C::OFF_DV);
    double *s_xp = reinterpret_cast<double *>(smem_raw + C::OFF_XP);
    double *s_red = reinterpret_cast<double *>(smem_raw + C::OFF_RED);
    double *s_kc = reinterpret_cast<double *>(smem_raw + C::OFF_KC);
    double *s_dc = reinterpret_cast<double *>(smem_raw + C::OFF_DC) + lev * 8 * 256;
    cd *s_slice = reinterpret_cast<cd *>(smem_raw + C::OFF_SLICE);
    uint64_t *s_mbar = reinterpret_cast<uint64_t *>(s_slice + N);

    // named barriers: 1,2 level-wide (X1); 3..6 "phi of level n, order parity p is available"; 7..14 warp pairs (X2)
    const int bar_level = 1 + lev, bar_pair = 7 + lev * 4 + q;
    const int BAR_FULL_OWN = 3 + 2 * lev, BAR_FULL_OTH = 3 + 2 * (1 - lev);
    const bool writer = tid == 0;

    const long long prun = (long long)p.poly_stride * run;
    for (int i = tid; i < p.nch; i += C::NTHREADS) {
        s_xp[i] = p.xp[i];
        s_dv[i] = p.dv[prun * p.nch + i];
    }
    for (int i = tid; i < C::TW_TOTAL; i += C::NTHREADS) s_tw1[i] = p.tw8[i];       // the three tables are contiguous
    const double emin = p.sc[prun * 4 + 0], emax = p.sc[prun * 4 + 1], t_sc = p.sc[prun * 4 + 2], eL = p.sc[prun * 4 + 3];
    const double c1 = 4.0 / (emax - emin);                             // phys_base.py:100
    const double c2 = 2.0 * (emax + emin) / (emax - emin);             // phys_base.py:101
    const int nt = (int)p.ctl[(long long)p.ctl_stride * run * 4 + 1];
    // per-run scalars that are only needed between the polynomial loops live in shared memory, not in registers
    // (128 registers per thread: the order loop needs all of them)
    double *s_sc = s_red + 16;                 // [0] dx [1] h_lambda [2] c1 [3,4] final phase
    if (tid == 0) {
        double ph_s, ph_c;
        sincos(-2.0 * t_sc * (emax + emin) / (emax - emin), &ph_s, &ph_c);   // phys_base.py:174
        s_sc[0] = p.dx[(long long)p.dx_stride * run];
        s_sc[1] = p.ctl[(long long)p.ctl_stride * run * 4 + 0];
        s_sc[2] = c1;
        s_sc[3] = ph_c;
        s_sc[4] = ph_s;
    }

    // per-thread constants, parked in shared memory (each thread reads back only what it wrote): the diagonal
    // d(x) = c1*(v(x) +- eL) - c2 in the x layout, the kinetic multiplier c1/N * akx2[k] in the k layout (layout 4)
    {
        const double *vr = p.v + p.v_stride * run + (long long)lev * N;
        const double sgn = lev == 0 ? eL : -eL;                    // hamil_2d.py:50-51, 66-67
#pragma unroll
        for (int i = 0; i < 8; ++i) s_dc[i * 256 + T] = fma(c1, vr[T + TPL * i] + sgn, -c2);
        if (lev == 0) {
            const double kscale = c1 / (double)N;                   // 1/N of numpy's ifft
            const int k1 = 2 * q + ((L >> 4) & 1), k2 = 4 * hb + ((L >> 2) & 3);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k3 = 4 * (e >> 2) + (L & 3), k4 = e & 3;
                s_kc[e * 256 + T] = p.kin[k1 + 8 * k2 + 64 * k3 + 512 * k4] * kscale;
            }
        }
    }

    cd psi[8], a[8];
    {
        const cd *src = p.psi + ((long long)run * 2 + lev) * N;
#pragma unroll
        for (int i = 0; i < 8; ++i) psi[i] = src[T + TPL * i];
    }
    // ---- tensor memory ---------------------------------------------------------------------------------
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_red + 30);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_tmem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_psi = t_lane + (uint32_t)((lev * 2 + hb) * 32);
    const uint32_t t_phi = t_lane + 128 + (uint32_t)((lev * 2 + hb) * 32);            // + 128 * parity
    const uint32_t t_phi_oth = t_lane + 128 + (uint32_t)(((1 - lev) * 2 + hb) * 32);
    const uint32_t t_stg = t_lane + 384 + (uint32_t)(lev * 64);                       // the warp pair's staging columns
    const uint32_t t_blk = t_stg + (uint32_t)(32 * hb);                               // the block this warp reads in X2 and owns in X3

    // addresses of the four layouts
    const int x1_w = T;                                                   // + k1 * KS
    const int x1_r = (2 * q + (L & 1)) * KS + 16 * hb + (L >> 1);        // + n2 * 32
    const cd *tw1 = s_tw1 + T;                                            // + k1 * 256
    const cd *tw2 = s_tw2 + 16 * hb + (L >> 1);                           // + k2 * 32
    const cd *tw3 = s_tw3 + (L >> 3);                                     // + k3 * 4
    const double *kcp = s_kc + T;                                         // + e * 256
    const double *dcp = s_dc + T;                                         // + n1 * 256

    const cd *traj_in = p.traj_in ? p.traj_in + (long long)run * (p.nt_max + 1) * N : nullptr;
    const bool staged = traj_in && (p.field_mode == 1 || p.field_mode == 2);
    uint32_t slice_phase = 0;
    cd *traj_out = p.traj_out ? p.traj_out + (long long)run * (p.nt_max + 1) * N : nullptr;
    const long long erun = (long long)p.E_stride * run * (p.nt_max + 1);
    const long long xrun = (long long)p.expL_stride * run * (p.nt_max + 1);
    cd *E_out = p.E_out + (long long)run * (p.nt_max + 1);
    const int ov_lev = p.field_mode == 2 ? 1 : 0;          // fitter.py:700-706, 791-796
    const int last_j = (p.orders ? p.orders[prun] : p.nch) - 2;
    if (staged && tid == 0) mbar_init(s_mbar, 1);
    __syncthreads();
    if (staged && tid == 0 && p.l_first <= nt)
        tma_load_1d(s_slice, traj_in + (long long)(nt - p.l_first) * N, (uint32_t)(N * sizeof(cd)), s_mbar);

    for (int l = p.l_first; l < p.l_first + p.nsteps && l <= nt; ++l) {
        const bool start_only = (l == 0);   // start(): only the initial field is evaluated (propagation.py:360-361)
        // ---- rotating frame in (propagation.py:383-390) ---------------------
        if (p.hf_hide && !start_only) {
            const cd eLx = p.expL[xrun + l];
            const double den = eLx.x * eLx.x + eLx.y * eLx.y, rden = 1.0 / den;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (lev == 0) {
                    cd n = cmulc(psi[i], eLx);
                    psi[i] = cmk(div_rn_by(n.x, den, rden), div_rn_by(n.y, den, rden));
                } else {
                    psi[i] = cmul(psi[i], eLx);
                }
            }
        }
        // ---- field of this step (fitter.py:650-811) --------------------------
        double E;
        if (p.field_mode == 0) {
            const int le = p.reversed_E ? (l == 0 ? nt : nt - l + 1) : l;
            E = p.E_base[erun + le].x;
        } else {
            double part = 0.0;
            if (lev == ov_lev) {
                mbar_wait(s_mbar, slice_phase);      // the bulk copy of this step's slice has landed
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const cd c = s_slice[T + TPL * i];
                    // forward: Im(conj(psi) * chi) ; backward: Im(conj(psi_old) * chi)
                    part += p.field_mode == 1 ? (psi[i].x * c.y - psi[i].y * c.x) : (c.x * psi[i].y - c.y * psi[i].x);
                }
            }
            double t0s, t1s;
            level_sums<C::NTHREADS, TPL>(part, s_red, tid, t0s, t1s);
            const double tot = t0s + t1s;
            E = -2.0 * (tot * s_sc[0]) / s_sc[1];
            // every thread is past the CTA-wide barriers of the sums, so the buffer has been read: fetch the slice of
            // the next step behind the polynomial orders of this one
            slice_phase ^= 1u;
            const int ln = l + 1;
            if (tid == 0 && ln <= nt && ln < p.l_first + p.nsteps)
                tma_load_1d(s_slice, traj_in + (long long)(nt - ln) * N, (uint32_t)(N * sizeof(cd)), s_mbar);
        }
        if (writer) {
            E_out[l] = cmk(E, 0.0);
            if (p.write_E_base) p.E_base[erun + l] = cmk(E, 0.0);
        }
        if (start_only) continue;
        const double cE = s_sc[2] * E;

        // ---- Newton recurrence (phys_base.py:158-172): phi = psi, psi <- dv0 * psi (parked in TMEM) -----
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            a[i] = psi[i];
            psi[i] = cmul(psi[i], s_dv[0]);
        }
        tmem_store_c8(t_psi, psi);
        for (int j = 0; j <= last_j; ++j) {
            const int par = j & 1;
            const uint32_t t_phi_j = t_phi + (uint32_t)(par * 128);
            tmem_store_c8(t_phi_j, a);                // own copy for the pointwise phase = hand-over to the other level
            // ---- pass 1 over n1, twiddle, exchange X1 (shared memory) ----
            {
                cd w[8];
#pragma unroll
                for (int k1 = 1; k1 < 8; ++k1) w[k1] = tw1[k1 * 256];       // in flight behind the butterflies
                dft8<false>(a);
                buf[x1_w] = a[0];
#pragma unroll
                for (int k1 = 1; k1 < 8; ++k1) buf[x1_w + k1 * KS] = cmul(a[k1], w[k1]);
            }
            tmem_wait_st();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(BAR_FULL_OWN + par, C::NTHREADS);          // "phi of this parity is there"
            bar_sync(bar_level, TPL);
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) a[n2] = buf[x1_r + n2 * 32];
            // ---- pass 2 over n2, twiddle, exchange X2 (tensor memory, warp pair) ----
            {
                cd w[8];
#pragma unroll
                for (int k2 = 1; k2 < 8; ++k2) w[k2] = tw2[k2 * 32];
                dft8<false>(a);
#pragma unroll
                for (int k2 = 1; k2 < 8; ++k2) a[k2] = cmul(a[k2], w[k2]);
            }
#pragma unroll
            for (int d = 0; d < 2; ++d) tmem_store_split4(t_stg + (uint32_t)(32 * d + 16 * hb), a + 4 * d);
            tmem_wait_st();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_sync(bar_pair, 64);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                uint32_t r0[16], r1[16];
                tmem_ld_16x256b_x4(t_blk, r0);
                tmem_ld_16x256b_x4(t_blk + (16u << 16), r1);
                tmem_wait_ld();
#pragma unroll
                for (int hp = 0; hp < 2; ++hp)
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
                        a[4 * hp + sl] = cmk(mkd(r0[8 * hp + 2 * sl], r0[8 * hp + 2 * sl + 1]),
                                             mkd(r0[8 * hp + 4 + 2 * sl], r0[8 * hp + 4 + 2 * sl + 1]));
                        a[4 * hp + 2 + sl] = cmk(mkd(r1[8 * hp + 2 * sl], r1[8 * hp + 2 * sl + 1]),
                                                 mkd(r1[8 * hp + 4 + 2 * sl], r1[8 * hp + 4 + 2 * sl + 1]));
                    }
            }
            // ---- pass 3 over n3, twiddle, exchange X3 (tensor memory, inside the warp) ----
            {
                cd w[8];
#pragma unroll
                for (int k3 = 1; k3 < 8; ++k3) w[k3] = tw3[k3 * 4];
                dft8<false>(a);
#pragma unroll
                for (int k3 = 1; k3 < 8; ++k3) a[k3] = cmul(a[k3], w[k3]);
            }
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) tmem_store_split4(t_blk + (uint32_t)(16 * pp), a + 4 * pp);
            tmem_wait_st();
            {
                uint32_t r0[16], r1[16];
                tmem_ld_16x256b_x4(t_blk, r0);
                tmem_ld_16x256b_x4(t_blk + (16u << 16), r1);
                tmem_wait_ld();
#pragma unroll
                for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
                        a[4 * pp + sl] = cmk(mkd(r0[8 * pp + 2 * sl], r0[8 * pp + 2 * sl + 1]),
                                             mkd(r0[8 * pp + 4 + 2 * sl], r0[8 * pp + 4 + 2 * sl + 1]));
                        a[4 * pp + 2 + sl] = cmk(mkd(r1[8 * pp + 2 * sl], r1[8 * pp + 2 * sl + 1]),
                                                 mkd(r1[8 * pp + 4 + 2 * sl], r1[8 * pp + 4 + 2 * sl + 1]));
                    }
            }
            // ---- pass 4 over n4, kinetic multiplier (c1 and the 1/N of the inverse folded in), inverse pass 4 ----
            dft4<false>(a[0], a[1], a[2], a[3]);
            dft4<false>(a[4], a[5], a[6], a[7]);
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = cscale(a[e], kcp[e * 256]);
            dft4<true>(a[0], a[1], a[2], a[3]);
            dft4<true>(a[4], a[5], a[6], a[7]);
            // ---- X3 back: store in the load's shape, load in the store's ----
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[16];
#pragma unroll
                for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
                        const cd v = a[4 * pp + 2 * half + sl];
                        r[8 * pp + 2 * sl] = dlo(v.x); r[8 * pp + 2 * sl + 1] = dhi(v.x);
                        r[8 * pp + 4 + 2 * sl] = dlo(v.y); r[8 * pp + 4 + 2 * sl + 1] = dhi(v.y);
                    }
                tmem_st_16x256b_x4(t_blk + ((uint32_t)(16 * half) << 16), r);
            }
            tmem_wait_st();
            {
                cd w[8];
#pragma unroll
                for (int k3 = 1; k3 < 8; ++k3) w[k3] = tw3[k3 * 4];
                {
                    uint32_t r0[16], r1[16];
                    tmem_ld16(t_blk, r0);
                    tmem_ld16(t_blk + 16, r1);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        a[c] = cmk(mkd(r0[2 * c], r0[2 * c + 1]), mkd(r0[8 + 2 * c], r0[8 + 2 * c + 1]));
                        a[4 + c] = cmk(mkd(r1[2 * c], r1[2 * c + 1]), mkd(r1[8 + 2 * c], r1[8 + 2 * c + 1]));
                    }
                }
                idft8_conj_twiddled(a, w);            // inverse pass 3: k3 -> n3
            }
            // ---- X2 back ----
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t r[16];
#pragma unroll
                for (int hp = 0; hp < 2; ++hp)
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
                        const cd v = a[4 * hp + 2 * half + sl];
                        r[8 * hp + 2 * sl] = dlo(v.x); r[8 * hp + 2 * sl + 1] = dhi(v.x);
                        r[8 * hp + 4 + 2 * sl] = dlo(v.y); r[8 * hp + 4 + 2 * sl + 1] = dhi(v.y);
                    }
                tmem_st_16x256b_x4(t_blk + ((uint32_t)(16 * half) << 16), r);
            }
            tmem_wait_st();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_sync(bar_pair, 64);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            {
                cd w[8];
#pragma unroll
                for (int k2 = 1; k2 < 8; ++k2) w[k2] = tw2[k2 * 32];
                {
                    uint32_t r0[16], r1[16];
                    tmem_ld16(t_stg + (uint32_t)(16 * hb), r0);
                    tmem_ld16(t_stg + (uint32_t)(32 + 16 * hb), r1);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        a[c] = cmk(mkd(r0[2 * c], r0[2 * c + 1]), mkd(r0[8 + 2 * c], r0[8 + 2 * c + 1]));
                        a[4 + c] = cmk(mkd(r1[2 * c], r1[2 * c + 1]), mkd(r1[8 + 2 * c], r1[8 + 2 * c + 1]));
                    }
                }
                idft8_conj_twiddled(a, w);            // inverse pass 2: k2 -> n2
            }
            // ---- X1 back: every thread writes exactly where it read ----
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) buf[x1_r + n2 * 32] = a[n2];
            bar_sync(bar_level, TPL);
            {
                cd w[8];
#pragma unroll
                for (int k1 = 1; k1 < 8; ++k1) w[k1] = tw1[k1 * 256];
#pragma unroll
                for (int k1 = 0; k1 < 8; ++k1) a[k1] = buf[x1_w + k1 * KS];
                idft8_conj_twiddled(a, w);            // inverse pass 1 -> kinetic term (scaled by c1) at x = T + 256 n1
            }
            // ---- pointwise part (hamil_2d.py:60-71, phys_base.py:98-111) and accumulation (phys_base.py:170-172) ----
            bar_sync(BAR_FULL_OTH + par, C::NTHREADS);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const double xpj = s_xp[j];
            const cd dvj = s_dv[j + 1];
#pragma unroll
            for (int c4 = 0; c4 < 8; c4 += 4) {
                uint32_t rp[16], rs[16], ro[16];
                tmem_ld16(t_psi + c4 * 4, rp);
                tmem_ld16(t_phi_j + c4 * 4, rs);
                tmem_ld16(t_phi_oth + (uint32_t)(par * 128) + c4 * 4, ro);
                cd oth4[4], own4[4], acc4[4];
                tmem_wait_ld();
                tmem_unpack_c4(rp, acc4);
                tmem_unpack_c4(rs, own4);
                tmem_unpack_c4(ro, oth4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n1 = c4 + i;
                    const double dd = dcp[n1 * 256] - xpj;
                    cd r;
                    r.x = fma(dd, own4[i].x, fma(-cE, oth4[i].x, a[n1].x));
                    r.y = fma(dd, own4[i].y, fma(-cE, oth4[i].y, a[n1].y));
                    a[n1] = r;
                    acc4[i] = cfma(dvj, r, acc4[i]);
                }
                tmem_store_c4x(t_psi + c4 * 4, acc4);
            }
        }
        tmem_wait_st();
#pragma unroll
        for (int c4 = 0; c4 < 8; c4 += 4) {
            uint32_t rp[16];
            tmem_ld16(t_psi + c4 * 4, rp);
            tmem_wait_ld();
            tmem_unpack_c4(rp, psi + c4);
        }
        // ---- final phase (phys_base.py:174-176), norms, renormalisation --------
        double nrm = 0.0;
        const cd phase = cmk(s_sc[3], s_sc[4]);
        const double dx = s_sc[0];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            psi[i] = cmul(psi[i], phase);
            nrm = fma(psi[i].x, psi[i].x, fma(psi[i].y, psi[i].y, nrm));
        }
        double n0, n1s;
        level_sums<C::NTHREADS, TPL>(nrm, s_red, tid, n0, n1s);
        n0 *= dx;
        n1s *= dx;
        if (writer) {
            p.norms[run * 2 + 0] = n0;
            p.norms[run * 2 + 1] = n1s;
        }
        const double tot = fabs(n0 + n1s);
        if (p.renorm && tot > 0.0) {                     // propagation.py:430-433
            const double s = sqrt(tot), rs = 1.0 / s;       // x / s bit for bit, see div_rn_by
#pragma unroll
            for (int i = 0; i < 8; ++i) psi[i] = cmk(div_rn_by(psi[i].x, s, rs), div_rn_by(psi[i].y, s, rs));
        }
        // ---- record the rotating-frame state (fitter.py:301-309, 322-329) -----
        if (traj_out && lev == ov_lev) {
            cd *dst = traj_out + (long long)l * N;
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[T + TPL * i] = psi[i];
        }
        // ---- rotating frame out (propagation.py:440-441) ---------------------
        if (p.hf_hide) {
            const cd eLx = p.expL[xrun + l];      // read again rather than kept across the polynomial loop
            const double den = eLx.x * eLx.x + eLx.y * eLx.y, rden = 1.0 / den;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (lev == 0) {
                    psi[i] = cmul(psi[i], eLx);
                } else {
                    cd n = cmulc(psi[i], eLx);
                    psi[i] = cmk(div_rn_by(n.x, den, rden), div_rn_by(n.y, den, rden));
                }
            }
        }
    }
    {
        cd *dst = p.psi + ((long long)run * 2 + lev) * N;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[T + TPL * i] = psi[i];
    }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
}

}  // namespace qc
