// Local control with a fed-back carrier-frequency multiplier (task type local_control_projection):
// the per-step scalar work of PropagationSolver.step / FittingSolver that kernel A needs on the device
// because it depends on the state of the previous step.
//
//   FreqMultiplier      fitter.py:814-836   Euler step of the multiplier ODE towards freq_mult_patched
//   step()              propagation.py:379-407   eL, exp_L, energy window, t_sc from the new multiplier
//   math_base.points    math_base.py:153-188     Newton table of exp(-i z t_sc), rebuilt every step
//   do_the_thing        fitter.py:598-623   dA/dt from <psi_e|psi_g>, <psi_e|dV|psi_g>; patched multiplier
//
// The reference evaluates all of this in Python / numpy scalars; the expressions below keep its operand
// order and use the non-contracting intrinsics, so that the only differences left are the last-ulp ones of
// sincos / pow / hypot and of the order of the two inner-product reductions.
#pragma once
#include "qc_fft.cuh"

namespace qc {

// per-run block of doubles behind GridParams::lc
enum LocalSlot {
    LC_FM = 0,        // freq_mult of the last step
    LC_VEL,           // freq_mult_vel
    LC_PATCHED,       // freq_mult_patched
    LC_HAPPY,         // dAdt_happy
    LC_ERR,           // 0 ok; 1 "freq_mult_patched has to be positive or zero"; 2 "Frequency multiplier has to be ..."
    LC_SRE, LC_SIM,   // psigc_psie of the last step
    LC_DRE, LC_DIM,   // psigc_dv_psie of the last step
    LC_DADT,          // dA/dt of the last step
    LC_ERR_STEP,      // step at which LC_ERR was raised
    LC_EXTR0 = 12,    // energy extremes (propagation.py:371-375), 4 values
    LC_E0 = 16, LC_NUL, LC_NU, LC_KE, LC_LAMB, LC_POW, LC_EPS, LC_DT,
    LC_STRIDE = 24,
    LC_NSTATE = 11    // slots written back by the kernel
};
// per-step scalars handed from warp 0 to the CTA
enum LocalStep { LS_EL = 0, LS_EMIN, LS_EMAX, LS_TSC, LS_XR, LS_XI, LS_PHR, LS_PHI, LS_STOP, LS_COUNT = 12 };

__device__ __forceinline__ double pmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double padd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double psub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double pdiv(double a, double b) { return __ddiv_rn(a, b); }

constexpr double LC_HZ_TO_CM = 3.33564095e-11;      // phys_base.py:15
constexpr double LC_CM_TO_ERG = 1.98644568e-16;     // phys_base.py:12
constexpr double LC_RED_PLANCK = 1.054572e-27;      // phys_base.py:14
constexpr double LC_PI = 3.141592653589793;

// principal square root of a complex number, CPython's cmath.sqrt algorithm (normal range)
__device__ __forceinline__ cd csqrt_py(cd z) {
    if (z.x == 0.0 && z.y == 0.0) return cmk(0.0, z.y);
    double ax = fabs(z.x) / 8.0;
    const double ay = fabs(z.y);
    const double s = 2.0 * sqrt(padd(ax, hypot(ax, ay / 8.0)));
    const double d = ay / (2.0 * s);
    return z.x >= 0.0 ? cmk(s, copysign(d, z.y)) : cmk(d, copysign(s, z.y));
}

// Executed by warp 0 at the beginning of step l (l >= 1).  s_lc: the run's LocalSlot block (shared memory);
// s_ls: LocalStep output; s_dv: Newton coefficients output; s_xp: nodes; rt: transposed partial products
// rt[j*nch + i] = prod_{m<=j} (xp_i - xp_m); rdiag[k] = 1 / rt[(k-1)*nch + k].
__device__ __noinline__ void local_step_setup(double *s_lc, double *s_ls, cd *s_dv, const double *s_xp,
                                               const double *__restrict__ rt, const double *__restrict__ rdiag, int nch,
                                               double t, int l, int lane) {
    if (lane == 0) {
        const double fm_prev = s_lc[LC_FM], patched = s_lc[LC_PATCHED];
        double vel = s_lc[LC_VEL];
        int err = 0;
        if (patched < 0.0) err = 1;                                // fitter.py:817-821
        else if (fm_prev < 0.0) err = 2;
        const double dt = s_lc[LC_DT], adt = fabs(dt);
        const double eL = pdiv(pmul(pmul(s_lc[LC_NUL], fm_prev), LC_HZ_TO_CM), 2.0);      // propagation.py:379
        // multiplier ODE, fitter.py:823-832
        const double first = psub(fm_prev, patched);
        const double second = pmul(vel, pow(pdiv(patched, fm_prev), s_lc[LC_POW]));
        const double acc = psub(pmul(-s_lc[LC_KE], first), pmul(s_lc[LC_LAMB], second));
        vel = padd(vel, pmul(acc, adt));
        const double fm = padd(fm_prev, pmul(vel, adt));
        // exp_L = sqrt(exp(i 2 pi nu fm t)), task_manager.py:84, propagation.py:386
        const double theta = pmul(pmul(pmul(pmul(2.0, LC_PI), s_lc[LC_NU]), fm), t);
        double sn, cs;
        sincos(theta, &sn, &cs);
        const cd xl = csqrt_py(cmk(cs, sn));
        // energy window, propagation.py:392-407
        const double E0 = s_lc[LC_E0];
        const double e0 = padd(padd(s_lc[LC_EXTR0 + 0], E0), eL), e1 = padd(psub(s_lc[LC_EXTR0 + 1], E0), eL);
        const double e2 = psub(padd(s_lc[LC_EXTR0 + 2], E0), eL), e3 = psub(psub(s_lc[LC_EXTR0 + 3], E0), eL);
        const double emax = fmax(fmax(e0, e1), fmax(e2, e3)), emin = fmin(fmin(e0, e1), fmin(e2, e3));
        const double t_sc = pdiv(pdiv(pmul(pmul(dt, psub(emax, emin)), LC_CM_TO_ERG), 4.0), LC_RED_PLANCK);
        double ps, pc;
        sincos(pdiv(pmul(pmul(-2.0, t_sc), padd(emax, emin)), psub(emax, emin)), &ps, &pc);   // phys_base.py:174
        s_ls[LS_EL] = eL; s_ls[LS_EMIN] = emin; s_ls[LS_EMAX] = emax; s_ls[LS_TSC] = t_sc;
        s_ls[LS_XR] = xl.x; s_ls[LS_XI] = xl.y; s_ls[LS_PHR] = pc; s_ls[LS_PHI] = ps;
        s_ls[LS_STOP] = err ? 1.0 : 0.0;
        if (err) {
            s_lc[LC_ERR] = (double)err;
            s_lc[LC_ERR_STEP] = (double)l;
        } else {
            s_lc[LC_FM] = fm;
            s_lc[LC_VEL] = vel;
        }
    }
    __syncwarp();
    const double t_sc = s_ls[LS_TSC];
    // Newton table: dv_i = (f(xp_i) - dv_0 - sum_{j<i-1} dv_{j+1} R[i][j]) / R[i][i-1]  (math_base.py:172-186),
    // rows i = lane + 32 r, the sums kept per row and advanced as soon as a coefficient is known
    constexpr int RPL = 4;
    cd base[RPL], acc[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        acc[r] = cmk(0.0, 0.0);
        base[r] = cmk(0.0, 0.0);
        if (i < nch) {
            double sn, cs;
            sincos(-pmul(s_xp[i], t_sc), &sn, &cs);                 // phys_base.py:116-126
            base[r] = cmk(cs, sn);
        }
    }
    if (lane == 0) s_dv[0] = base[0];
    __syncwarp();
    const cd dv0 = s_dv[0];
#pragma unroll
    for (int r = 0; r < RPL; ++r) base[r] = cmk(psub(base[r].x, dv0.x), psub(base[r].y, dv0.y));
    double rnext[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const int i = lane + 32 * r;
        rnext[r] = (i < nch && nch > 1) ? __ldg(rt + i) : 0.0;      // column j = 0
    }
    for (int k = 1; k < nch; ++k) {
        double rcur[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            rcur[r] = rnext[r];
            const int i = lane + 32 * r;
            rnext[r] = (i < nch && k + 1 < nch) ? __ldg(rt + (long long)k * nch + i) : 0.0;
        }
        if (lane == (k & 31)) {
            cd num = cmk(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < RPL; ++r)
                if (r == (k >> 5)) num = cmk(psub(base[r].x, acc[r].x), psub(base[r].y, acc[r].y));
            const double scl = rdiag[k];                            // numpy complex / real: multiply by 1/res
            s_dv[k] = cmk(pmul(num.x, scl), pmul(num.y, scl));
        }
        __syncwarp();
        const cd dvk = s_dv[k];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const int i = lane + 32 * r;
            if (i > k && i < nch) {
                acc[r].x = padd(acc[r].x, pmul(dvk.x, rcur[r]));
                acc[r].y = padd(acc[r].y, pmul(dvk.y, rcur[r]));
            }
        }
    }
    __syncwarp();
}

// Executed by thread 0 after the step: the feedback of fitter.py:598-623.  S = <psi_e|psi_g> and
// D = <psi_e|(v_g - v_e)|psi_g> already carry dx.
__device__ __forceinline__ void local_feedback(double *s_lc, cd S, cd D) {
    const double coef2 = pdiv(pmul(-4.0, LC_CM_TO_ERG), LC_RED_PLANCK);
    const cd Sge2 = cmk(psub(pmul(S.x, S.x), pmul(S.y, S.y)), padd(pmul(S.x, S.y), pmul(S.y, S.x)));
    const cd Sdvge = cmk(psub(pmul(S.x, D.x), pmul(S.y, D.y)), padd(pmul(S.x, D.y), pmul(S.y, D.x)));
    const double freq_cm = pmul(LC_HZ_TO_CM, s_lc[LC_NUL]);
    const double f = pmul(freq_cm, s_lc[LC_FM]);
    const double body_im = padd(Sdvge.y, pmul(f, Sge2.y));
    const double dAdt = pmul(body_im, coef2);
    const double eps = s_lc[LC_EPS], happy = s_lc[LC_HAPPY];
    if (dAdt >= 0.0) {
        s_lc[LC_HAPPY] = dAdt;
    } else {
        double patched;
        if (Sge2.y > eps)
            patched = pdiv(psub(happy, pmul(coef2, Sdvge.y)), pmul(pmul(Sge2.y, freq_cm), coef2));
        else if (Sge2.y > 0.0)
            patched = pdiv(psub(happy, pmul(coef2, Sdvge.y)), pmul(pmul(eps, freq_cm), coef2));
        else if (Sge2.y < -eps)
            patched = pdiv(-Sdvge.y, pmul(Sge2.y, freq_cm));
        else
            patched = pdiv(Sdvge.y, pmul(eps, freq_cm));
        if (patched < 0.0) patched = 0.0;
        s_lc[LC_PATCHED] = patched;
    }
    s_lc[LC_SRE] = S.x; s_lc[LC_SIM] = S.y; s_lc[LC_DRE] = D.x; s_lc[LC_DIM] = D.y; s_lc[LC_DADT] = dAdt;
}

}  // namespace qc
