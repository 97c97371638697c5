// Kernel A, 512-thread form for np = 1024 with TWO runs per CTA ("K8x2"): the layout of qc_grid8.cuh -- eight points per
// thread, 124 registers, transform exchanges through tensor memory -- at the size where it gives a scheduler FOUR
// INDEPENDENT phases instead of two.  A level of 1024 points is four warps, one per scheduler (= one per tensor-memory
// lane quadrant); two runs = four levels = sixteen warps, and the four warps of a scheduler belong to four different
// levels that meet nobody but their own level (two level barriers per order) and their partner level (the phi
// hand-over, one order of slack).  EXPERIMENTAL (QC_GRID_VAR=2 at np = 1024): it was built to test whether the lock
// step of the eight warps of a level is what holds qc_grid8.cuh back at np = 2048.  It is not: parity-green, 0.562
// against 0.621 of the default two-runs-per-CTA form -- the issue port is.
//
// Transform: N = 1024 = 8 x 8 x 4 x 4, n = 128 n1 + 16 n2 + 4 n3 + n4, k = k1 + 8 k2 + 64 k3 + 256 k4; thread T = 32 q + L.
//   x layout   registers n1;          thread: n2 = T >> 4, n3 = (L >> 2) & 3, n4 = L & 3            (x = T + 128 n1)
//   pass 1 over n1, twiddle w_1024^(k1 T); exchange X1 through shared memory (crosses the quadrants)
//   layout 2   registers n2;          thread: k1 = 2 q + (L & 1), n3 = L >> 3, n4 = (L >> 1) & 3
//   pass 2 over n2, twiddle w_128^(k2 (4 n3 + n4)); exchange X2 through tensor memory, inside the warp
//   layout 3   registers (k2 >> 2, n3); thread: k2 & 3 = L & 3, k1 bit 0 = (L >> 2) & 1, n4 = L >> 3
//   pass 3 over n3 (two 4-point transforms), twiddle w_16^(k3 n4); exchange X3 through tensor memory, inside the warp
//   layout 4   registers (k2 >> 2, n4); thread: k3 = L & 3, k2 & 3 = (L >> 2) & 3, k1 bit 0 = L >> 4
//   pass 4 over n4, kinetic multiplier, and the mirror image back.
// Both tensor-memory exchanges are the intra-warp one of qc_grid8.cuh (st.32x32b + ld.16x256b and its mirror): lane
// bits 3-4 trade places with two register bits, no barrier wider than a warp.
#pragma once
#include "qc_grid8.cuh"

namespace qc {

struct Grid8x2Cfg {
    static constexpr int N = 1024, TPL = 128, RUNS = 2, NTHREADS = 512;
    static constexpr int KS = 132;                  // k1 stride of an X1 buffer (padded: conflict-free reads)
    static constexpr int BUF = 8 * KS;
    static constexpr int MAXCH = 128;
    static constexpr size_t OFF_TW1 = sizeof(cd) * 4 * BUF;                      // [8][128] w_1024^(k1 T)
    static constexpr size_t OFF_TW2 = OFF_TW1 + sizeof(cd) * 8 * 128;            // [8][16]  w_128^(k2 m2)
    static constexpr size_t OFF_TW3 = OFF_TW2 + sizeof(cd) * 8 * 16;             // [4][4]   w_16^(k3 n4)
    static constexpr size_t OFF_DV = OFF_TW3 + sizeof(cd) * 4 * 4;               // [2][MAXCH]
    static constexpr size_t OFF_XP = OFF_DV + sizeof(cd) * 2 * MAXCH;            // [2][MAXCH]
    static constexpr size_t OFF_RED = OFF_XP + sizeof(double) * 2 * MAXCH;       // [2][16] partial sums, [2][8] scalars, tmem slot
    static constexpr size_t OFF_KC = OFF_RED + sizeof(double) * 64;              // [4][8][128]  (c1 belongs to the run)
    static constexpr size_t OFF_DC = OFF_KC + sizeof(double) * 4 * 8 * 128;      // [4][8][128]
    static constexpr size_t OFF_SLICE = (OFF_DC + sizeof(double) * 4 * 8 * 128 + 127) / 128 * 128;   // [2][N] + 2 mbarriers
    static constexpr size_t SMEM_BYTES = OFF_SLICE + sizeof(cd) * 2 * N + 32;
    static constexpr int TW_TOTAL = 8 * 128 + 8 * 16 + 4 * 4;
};

__global__ void __launch_bounds__(Grid8x2Cfg::NTHREADS, 1) grid8x2_sweep_kernel(GridParams p) {
    typedef Grid8x2Cfg C;
    constexpr int N = C::N, TPL = C::TPL, KS = C::KS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, L = tid & 31, q = warp & 3, slot4 = warp >> 2, r = slot4 >> 1, lev = slot4 & 1;
    const int T = 32 * q + L;
    const int slot = (int)blockIdx.x * 2 + r;
    const bool valid = slot < p.run_count && p.run_first + slot < p.batch;
    const int run = valid ? p.run_first + slot : p.run_first;

    cd *buf = reinterpret_cast<cd *>(smem_raw) + slot4 * C::BUF;
    cd *s_tw1 = reinterpret_cast<cd *>(smem_raw + C::OFF_TW1);
    cd *s_tw2 = reinterpret_cast<cd *>(smem_raw + C::OFF_TW2);
    cd *s_tw3 = reinterpret_cast<cd *>(smem_raw + C::OFF_TW3);
    cd *s_dv = reinterpret_cast<cd *>(smem_raw + C::OFF_DV) + r * C::MAXCH;
    double *s_xp = reinterpret_cast<double *>(smem_raw + C::OFF_XP) + r * C::MAXCH;
    double *s_redall = reinterpret_cast<double *>(smem_raw + C::OFF_RED);
    double *s_red = s_redall + r * 16;              // per-warp partial sums of this run: [lev][q]
    double *s_sc = s_redall + 32 + r * 8;           // [0] dx [1] h_lambda [2] c1 [3,4] final phase
    double *s_kc = reinterpret_cast<double *>(smem_raw + C::OFF_KC) + slot4 * 8 * 128;
    double *s_dc = reinterpret_cast<double *>(smem_raw + C::OFF_DC) + slot4 * 8 * 128;
    cd *s_slice = reinterpret_cast<cd *>(smem_raw + C::OFF_SLICE) + r * N;
    uint64_t *s_mbar = reinterpret_cast<uint64_t *>(reinterpret_cast<cd *>(smem_raw + C::OFF_SLICE) + 2 * N) + r;

    // named barriers: 1..4 level-wide (X1), 5..12 "phi of (run, level, parity) is available", 13..14 both levels of a run
    const int bar_level = 1 + slot4, bar_run = 13 + r;
    const int BAR_FULL_OWN = 5 + 4 * r + 2 * lev, BAR_FULL_OTH = 5 + 4 * r + 2 * (1 - lev);
    const bool writer = valid && lev == 0 && T == 0;
    constexpr int RT = 2 * TPL;                     // threads of a run
    const int rtid = lev * TPL + T;

    const long long prun = (long long)p.poly_stride * run;
    for (int i = rtid; i < p.nch; i += RT) {
        s_xp[i] = p.xp[i];
        s_dv[i] = p.dv[prun * p.nch + i];
    }
    for (int i = tid; i < C::TW_TOTAL; i += C::NTHREADS) s_tw1[i] = p.tw8[i];       // the three tables are contiguous
    const double emin = p.sc[prun * 4 + 0], emax = p.sc[prun * 4 + 1], t_sc = p.sc[prun * 4 + 2], eL = p.sc[prun * 4 + 3];
    const double c1 = 4.0 / (emax - emin);                             // phys_base.py:100
    const double c2 = 2.0 * (emax + emin) / (emax - emin);             // phys_base.py:101
    const int nt = (int)p.ctl[(long long)p.ctl_stride * run * 4 + 1];
    if (rtid == 0) {
        double ph_s, ph_c;
        sincos(-2.0 * t_sc * (emax + emin) / (emax - emin), &ph_s, &ph_c);   // phys_base.py:174
        s_sc[0] = p.dx[(long long)p.dx_stride * run];
        s_sc[1] = p.ctl[(long long)p.ctl_stride * run * 4 + 0];
        s_sc[2] = c1;
        s_sc[3] = ph_c;
        s_sc[4] = ph_s;
    }
    {
        const double *vr = p.v + p.v_stride * run + (long long)lev * N;
        const double sgn = lev == 0 ? eL : -eL;                    // hamil_2d.py:50-51, 66-67
#pragma unroll
        for (int i = 0; i < 8; ++i) s_dc[i * 128 + T] = fma(c1, vr[T + TPL * i] + sgn, -c2);
    }
    // the kinetic multiplier c1 / N * akx2[k] at the eight frequencies this thread holds in layout 4
    {
        const double kscale = c1 / (double)N;                       // 1/N of numpy's ifft
        const int k1 = 2 * q + ((L >> 4) & 1);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k2 = 4 * (e >> 2) + ((L >> 2) & 3), k3 = L & 3, k4 = e & 3;
            s_kc[e * 128 + T] = p.kin[k1 + 8 * k2 + 64 * k3 + 256 * k4] * kscale;
        }
    }
    cd psi[8], a[8];
    {
        const cd *src = p.psi + ((long long)run * 2 + lev) * N;
#pragma unroll
        for (int i = 0; i < 8; ++i) psi[i] = src[T + TPL * i];
    }
    // ---- tensor memory ---------------------------------------------------------------------------------
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_redall + 62);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_tmem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0 || tid == 256) mbar_init(s_mbar, 1);      // thread 0 of each run (r = tid >> 8)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t t_psi = t_lane + (uint32_t)(slot4 * 32);
    const uint32_t t_phi = t_lane + 128 + (uint32_t)(slot4 * 32);                     // + 128 * parity
    const uint32_t t_phi_oth = t_lane + 128 + (uint32_t)((slot4 ^ 1) * 32);
    const uint32_t t_blk = t_lane + 384 + (uint32_t)(slot4 * 32);                     // this warp's staging columns

    const int x1_w = T;                                                   // + k1 * KS
    const int x1_r = (2 * q + (L & 1)) * KS + 4 * (L >> 3) + ((L >> 1) & 3);   // + n2 * 16
    const cd *tw1 = s_tw1 + T;                                            // + k1 * 128
    const cd *tw2 = s_tw2 + 4 * (L >> 3) + ((L >> 1) & 3);               // + k2 * 16
    const cd *tw3 = s_tw3 + (L >> 3);                                     // + k3 * 4
    const double *dcp = s_dc + T;                                         // + n1 * 128
    const double *kcp = s_kc + T;                                         // + e * 128

    const cd *traj_in = p.traj_in ? p.traj_in + (long long)run * (p.nt_max + 1) * N : nullptr;
    const bool staged = traj_in && (p.field_mode == 1 || p.field_mode == 2);
    uint32_t slice_phase = 0;
    cd *traj_out = p.traj_out ? p.traj_out + (long long)run * (p.nt_max + 1) * N : nullptr;
    const long long erun = (long long)p.E_stride * run * (p.nt_max + 1);
    const long long xrun = (long long)p.expL_stride * run * (p.nt_max + 1);
    cd *E_out = p.E_out + (long long)run * (p.nt_max + 1);
    const int ov_lev = p.field_mode == 2 ? 1 : 0;          // fitter.py:700-706, 791-796
    const int last_j = (p.orders ? p.orders[prun] : p.nch) - 2;
    if (valid && staged && rtid == 0 && p.l_first <= nt)
        tma_load_1d(s_slice, traj_in + (long long)(nt - p.l_first) * N, (uint32_t)(N * sizeof(cd)), s_mbar);

    // sums over the threads of each level of THIS run; every thread of the run gets both
    auto run_sums = [&](double val, double &s0, double &s1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        if (L == 0) s_red[lev * 4 + q] = val;
        bar_sync(bar_run, RT);
        s0 = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
        s1 = (s_red[4] + s_red[5]) + (s_red[6] + s_red[7]);
        bar_sync(bar_run, RT);
    };

    for (int l = p.l_first; l < p.l_first + p.nsteps && l <= nt && valid; ++l) {
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
            run_sums(part, t0s, t1s);
            const double tot = t0s + t1s;
            E = -2.0 * (tot * s_sc[0]) / s_sc[1];
            slice_phase ^= 1u;
            const int ln = l + 1;
            if (rtid == 0 && ln <= nt && ln < p.l_first + p.nsteps)
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
                for (int k1 = 1; k1 < 8; ++k1) w[k1] = tw1[k1 * 128];       // in flight behind the butterflies
                dft8<false>(a);
                buf[x1_w] = a[0];
#pragma unroll
                for (int k1 = 1; k1 < 8; ++k1) buf[x1_w + k1 * KS] = cmul(a[k1], w[k1]);
            }
            tmem_wait_st();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(BAR_FULL_OWN + par, RT);                   // "phi of this parity is there"
            bar_sync(bar_level, TPL);
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) a[n2] = buf[x1_r + n2 * 16];
            // ---- pass 2 over n2, twiddle, exchange X2 (tensor memory, inside the warp) ----
            {
                cd w[8];
#pragma unroll
                for (int k2 = 1; k2 < 8; ++k2) w[k2] = tw2[k2 * 16];
                dft8<false>(a);
#pragma unroll
                for (int k2 = 1; k2 < 8; ++k2) a[k2] = cmul(a[k2], w[k2]);
            }
            auto exchange_fwd = [&]() {               // a[4 pp + c] at lane (s, a') -> a[4 pp + s] at lane (a', c)
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) tmem_store_split4(t_blk + (uint32_t)(16 * pp), a + 4 * pp);
                tmem_wait_st();
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
            };
            auto exchange_back = [&]() {              // the mirror: store in the load's shape, load in the store's
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t rr[16];
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp)
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) {
                            const cd v = a[4 * pp + 2 * half + sl];
                            rr[8 * pp + 2 * sl] = dlo(v.x); rr[8 * pp + 2 * sl + 1] = dhi(v.x);
                            rr[8 * pp + 4 + 2 * sl] = dlo(v.y); rr[8 * pp + 4 + 2 * sl + 1] = dhi(v.y);
                        }
                    tmem_st_16x256b_x4(t_blk + ((uint32_t)(16 * half) << 16), rr);
                }
                tmem_wait_st();
                uint32_t r0[16], r1[16];
                tmem_ld16(t_blk, r0);
                tmem_ld16(t_blk + 16, r1);
                tmem_wait_ld();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    a[c] = cmk(mkd(r0[2 * c], r0[2 * c + 1]), mkd(r0[8 + 2 * c], r0[8 + 2 * c + 1]));
                    a[4 + c] = cmk(mkd(r1[2 * c], r1[2 * c + 1]), mkd(r1[8 + 2 * c], r1[8 + 2 * c + 1]));
                }
            };
            exchange_fwd();                           // registers (k2 >> 2, n3)
            // ---- pass 3 over n3, twiddle w_16^(k3 n4), exchange X3 ----
            {
                cd w[4];
#pragma unroll
                for (int k3 = 1; k3 < 4; ++k3) w[k3] = tw3[k3 * 4];
                dft4<false>(a[0], a[1], a[2], a[3]);
                dft4<false>(a[4], a[5], a[6], a[7]);
#pragma unroll
                for (int k3 = 1; k3 < 4; ++k3) {
                    a[k3] = cmul(a[k3], w[k3]);
                    a[4 + k3] = cmul(a[4 + k3], w[k3]);
                }
            }
            exchange_fwd();                           // registers (k2 >> 2, n4)
            // ---- pass 4 over n4, kinetic multiplier (c1 and the 1/N of the inverse folded in), inverse pass 4 ----
            dft4<false>(a[0], a[1], a[2], a[3]);
            dft4<false>(a[4], a[5], a[6], a[7]);
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = cscale(a[e], kcp[e * 128]);
            dft4<true>(a[0], a[1], a[2], a[3]);
            dft4<true>(a[4], a[5], a[6], a[7]);
            exchange_back();                          // registers (k2 >> 2, k3), layout 3
            {
                cd w[4];
#pragma unroll
                for (int k3 = 1; k3 < 4; ++k3) w[k3] = tw3[k3 * 4];
                dft4_inv_conj_twiddled<true>(a[0], a[1], a[2], a[3], w);      // inverse pass 3: k3 -> n3
                dft4_inv_conj_twiddled<true>(a[4], a[5], a[6], a[7], w);
            }
            exchange_back();                          // registers k2, layout 2
            {
                cd w[8];
#pragma unroll
                for (int k2 = 1; k2 < 8; ++k2) w[k2] = tw2[k2 * 16];
                idft8_conj_twiddled(a, w);            // inverse pass 2: k2 -> n2
            }
            // ---- X1 back: every thread writes exactly where it read ----
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) buf[x1_r + n2 * 16] = a[n2];
            bar_sync(bar_level, TPL);
            {
                cd w[8];
#pragma unroll
                for (int k1 = 1; k1 < 8; ++k1) w[k1] = tw1[k1 * 128];
#pragma unroll
                for (int k1 = 0; k1 < 8; ++k1) a[k1] = buf[x1_w + k1 * KS];
                idft8_conj_twiddled(a, w);            // inverse pass 1 -> kinetic term (scaled by c1) at x = T + 128 n1
            }
            // ---- pointwise part (hamil_2d.py:60-71, phys_base.py:98-111) and accumulation (phys_base.py:170-172) ----
            bar_sync(BAR_FULL_OTH + par, RT);
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
                    const double dd = dcp[n1 * 128] - xpj;
                    cd rr;
                    rr.x = fma(dd, own4[i].x, fma(-cE, oth4[i].x, a[n1].x));
                    rr.y = fma(dd, own4[i].y, fma(-cE, oth4[i].y, a[n1].y));
                    a[n1] = rr;
                    acc4[i] = cfma(dvj, rr, acc4[i]);
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
        run_sums(nrm, n0, n1s);
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
    if (valid) {
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
