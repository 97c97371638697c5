// Kernel B: N-level systems without a spatial grid (np == 1): the tridiagonal
// angular-momentum / Bose-Hubbard Hamiltonians BH_X, BH_Z, the qubit and the
// trivial two-level model (hamil_2d.py:76-187).
//
// One thread owns one basis vector of one run for a whole sweep: state, Newton
// recurrence and renormalisation are register-only.  The nb basis vectors of a
// run sit in adjacent lanes so the unitary-transformation field update
// (fitter.py:739-778), which couples them at every step, is a few shuffles.
// Trajectory slices are stored [slice][level][thread] so that a warp reads and
// writes contiguous 512-byte rows.
#pragma once
#include "qc_fft.cuh"

namespace qc {

struct NLevelParams {
    int batch, nb, nbp, nch, nt_max, hamil;
    int renorm, field_mode, reversed_E, write_E_base;
    int l_first, nsteps, single_step;
    int no_retire;               // diagnostics (QC_NLEVEL_RETIRE=0): every warp runs the whole launch like in round 1
    long long nthreads;          // batch * nbp
    cd *psi;                     // [batch][nb][NL]
    const double *vdiag;         // [vb][NL]  the constant arrays v[n][1] (two_levels only)
    int v_stride;
    const double *uwd;           // [ub][3]
    int uwd_stride;
    const double *xp;            // [nch]
    const cd *dv;                // [pb][nch]
    const double *sc;            // [pb][4]
    const int *orders;           // [pb] polynomial terms summed per run (opt-in truncation) or null
    int poly_stride;
    const double *dx;
    int dx_stride;
    cd *E_base;                  // [eb][nt_max+1]
    int E_stride;
    const cd *s_env;             // [sb][nt_max+1]
    int s_stride;
    const double *ctl;           // [cb][4] h_lambda, nt
    int ctl_stride;
    const cd *a0;                // [batch][nb]
    const cd *E_step;            // [batch]
    cd *E_out;                   // [batch][nt_max+1]
    const cd *traj_in;           // [nt_max+1][NL][nthreads]
    cd *traj_out;
    double *norms;               // [batch][nb][NL]
};

// phi_out = H(E) phi for the tridiagonal level matrix (diag real, off-diagonals complex)
template <int NL>
__device__ __forceinline__ void tri_apply(const double *dg, const cd *up, const cd *lo, const cd *in, cd *out) {
#pragma unroll
    for (int g = 0; g < NL; ++g) {
        cd acc = cmk(0.0, 0.0);
        if (g > 0) acc = cmul(lo[g], in[g - 1]);            // H[g, g-1]
        acc = cadd(acc, cscale(in[g], dg[g]));              // H[g, g]
        if (g < NL - 1) acc = cadd(acc, cmul(up[g + 1], in[g + 1]));  // H[g, g+1]
        out[g] = acc;
    }
}

// NCH > 0: the interpolation nodes and coefficients (nch <= NCH) are cached in registers and the polynomial loop
// is unrolled, so the time loop has no loads besides the one-step-ahead prefetch; NCH == 0: any nch, read
// through L1 every order.
template <int NL, int NCH>
__global__ void __launch_bounds__(128) nlevel_sweep_kernel(NLevelParams p) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = gt < p.nthreads;
    const int run = in_range ? (int)(gt / p.nbp) : 0;
    const int vec = in_range ? (int)(gt % p.nbp) : 0;
    const bool active = in_range && vec < p.nb;

    const long long prun = (long long)p.poly_stride * run;
    const double emin = p.sc[prun * 4 + 0], emax = p.sc[prun * 4 + 1], t_sc = p.sc[prun * 4 + 2];
    const double c1 = 4.0 / (emax - emin);
    const double c2 = 2.0 * (emax + emin) / (emax - emin);
    const double dx = p.dx[(long long)p.dx_stride * run];
    const double h_lambda = p.ctl[(long long)p.ctl_stride * run * 4 + 0];
    // ctl[run] = (h_lambda, nt, retired, -): a retired run (converged, diverged or stopped in an earlier iteration of
    // the control loop) keeps its slot in the plan but does no step and no store
    const bool retired = p.ctl[(long long)p.ctl_stride * run * 4 + 2] != 0.0;
    const int nt = retired ? -1 : (int)p.ctl[(long long)p.ctl_stride * run * 4 + 1];
    const double inv_h_lambda = 1.0 / h_lambda;
    // A warp leaves the time loop with its longest live run (the basis vectors of a run never straddle warps, so the
    // shuffles below stay convergent).  With the runs of a plan ordered by length (control.NLevelBatch) a launch
    // keeps only the warps of the runs that are still propagating, and its tail costs nothing.
    int warp_last = (in_range && !retired) ? nt : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) warp_last = max(warp_last, __shfl_xor_sync(0xffffffffu, warp_last, o));
    const int l_end = (p.single_step || p.no_retire) ? p.l_first + p.nsteps : min(p.l_first + p.nsteps, warp_last + 1);
    const double U = p.uwd ? p.uwd[(long long)p.uwd_stride * run * 3 + 0] : 0.0;
    const double W = p.uwd ? p.uwd[(long long)p.uwd_stride * run * 3 + 1] : 0.0;
    const double delta = p.uwd ? p.uwd[(long long)p.uwd_stride * run * 3 + 2] : 0.0;
    double ph_s, ph_c;
    sincos(-2.0 * t_sc * (emax + emin) / (emax - emin), &ph_s, &ph_c);
    const cd phase = cmk(ph_c, ph_s);
    const cd *dv = p.dv + prun * p.nch;
    const int nterms = p.orders ? p.orders[prun] : p.nch;
    constexpr int NCR = NCH > 0 ? NCH : 1;
    double xp_r[NCR];
    cd dv_r[NCR];
    if (NCH > 0) {
#pragma unroll
        for (int j = 0; j < NCR; ++j) {
            xp_r[j] = j < p.nch ? p.xp[j] : 0.0;
            dv_r[j] = j < p.nch ? dv[j] : cmk(0.0, 0.0);
        }
    }
    const cd a0 = (p.a0 && active) ? p.a0[(long long)run * p.nb + vec] : cmk(0.0, 0.0);

    // E-independent parts of the level matrix
    const double lq = (NL - 1) / 2.0;
    double jc[NL];          // sqrt(l(l+1) - (l-vi+1)(l-vi)),  hamil_2d.py:110
    double dg0[NL];         // field-free diagonal
#pragma unroll
    for (int vi = 0; vi < NL; ++vi) {
        jc[vi] = vi ? sqrt(lq * (lq + 1) - (lq - vi + 1) * (lq - vi)) : 0.0;
        const double m = lq - vi;
        if (p.hamil == 2) dg0[vi] = 2.0 * m * U + 2.0 * m * m * W;            // BH_X  hamil_2d.py:108-109
        else if (p.hamil == 3) dg0[vi] = 2.0 * m * m * U;                      // BH_Z  hamil_2d.py:138-139
        else if (p.hamil == 4) dg0[vi] = -2.0 * m * U + 2.0 * m * m * W;       // qubit hamil_2d.py:168-169
        else dg0[vi] = p.vdiag[(long long)p.v_stride * run * NL + vi];         // two_levels hamil_2d.py:90-91
    }

    cd psi[NL], phi[NL], tmp[NL];
    if (active) {
#pragma unroll
        for (int n = 0; n < NL; ++n) psi[n] = p.psi[((long long)run * p.nb + vec) * NL + n];
    } else {
#pragma unroll
        for (int n = 0; n < NL; ++n) psi[n] = cmk(0.0, 0.0);
    }
    const long long erun = (long long)p.E_stride * run * (p.nt_max + 1);
    const long long srun = (long long)p.s_stride * run * (p.nt_max + 1);
    cd *E_out = p.E_out + (long long)run * (p.nt_max + 1);
    const long long slice = (long long)NL * p.nthreads;

    // The per-step inputs (field table entry, envelope, the stored slice of the opposite sweep) depend only on
    // the step number: they are fetched one step ahead, so their HBM latency (every slice is touched once) is
    // hidden behind the polynomial recurrence of the current step instead of heading its dependency chain.
    struct StepIn {
        cd base, chi[NL];
        double s;
    };
    const bool ut = !p.single_step && p.field_mode != 0;
    auto fetch = [&](int l) {
        StepIn in;
        in.base = cmk(0.0, 0.0);
        in.s = 0.0;
#pragma unroll
        for (int n = 0; n < NL; ++n) in.chi[n] = cmk(0.0, 0.0);
        if (p.single_step || l > nt || l >= p.l_first + p.nsteps) return in;
        if (!ut) {
            const int le = p.reversed_E ? (l == 0 ? nt : nt - l + 1) : l;
            in.base = p.E_base[erun + le];
            return in;
        }
        in.base = p.E_base[erun + l];
        in.s = p.s_env[srun + l].x;
        if (active && l > 0) {
            // UT forward (fitter.py:739-778): chi_tlist[-l] is slice nt+1-l of the backward record
            const cd *tr = p.traj_in + (long long)(nt + 1 - l) * slice + gt;
#pragma unroll
            for (int n = 0; n < NL; ++n) in.chi[n] = tr[(long long)n * p.nthreads];
        }
        return in;
    };
    StepIn nxt = fetch(p.l_first);
    const cd E_single = p.single_step ? p.E_step[run] : cmk(0.0, 0.0);

    for (int l = p.l_first; l < l_end; ++l) {
        const bool live = active && l <= nt;
        const StepIn cur = nxt;
        // ---- field -------------------------------------------------------------
        cd E;
        if (p.single_step) {
            E = E_single;
        } else if (!ut) {
            E = cur.base;
        } else {
            cd part = cmk(0.0, 0.0);
            if (live && l > 0) {
#pragma unroll
                for (int n = 0; n < NL; ++n) {
                    // a0[v] * cprod(chi, psi),  cprod(chi, psi) = dx * conj(psi) * chi  (math_base.py:21)
                    part = cadd(part, cmul(a0, cscale(cmulc(cur.chi[n], psi[n]), dx)));
                }
            }
            for (int o = 1; o < p.nbp; o <<= 1) {
                part.x += __shfl_xor_sync(0xffffffffu, part.x, o);
                part.y += __shfl_xor_sync(0xffffffffu, part.y, o);
            }
            if (l <= nt) {
                E = l > 0 ? cmk(cur.base.x - div_rn_by(cur.s * part.y, h_lambda, inv_h_lambda), cur.base.y) : cur.base;
            } else {
                E = cmk(0.0, 0.0);
            }
        }
        if (live && vec == 0 && !p.single_step) {
            E_out[l] = E;
            if (p.write_E_base) p.E_base[erun + l] = E;
        }
        // the only loads of the time loop; issued here, consumed one iteration later (no other load may share
        // their scoreboard slot between here and the stores above, or the stores wait for HBM)
        nxt = fetch(l + 1);
        if (l == 0 || !live) continue;

        // ---- level matrix for this field (rebuilt per step like hamil_2d.py:100-118) -------
        double dg[NL];
        cd up[NL], lo[NL];       // up[vi] = H[vi-1, vi], lo[vi] = H[vi, vi-1]
#pragma unroll
        for (int vi = 0; vi < NL; ++vi) {
            dg[vi] = dg0[vi];
            up[vi] = lo[vi] = cmk(0.0, 0.0);
            if (p.hamil == 2) {                                            // BH_X: P = R = -delta*E*jc
                up[vi] = lo[vi] = cscale(cscale(E, -delta), jc[vi]);
            } else if (p.hamil == 3) {                                     // BH_Z: field on the diagonal
                dg[vi] = dg0[vi] + 2.0 * (lq - vi) * E.x;
                up[vi] = lo[vi] = cmk(-delta * jc[vi], 0.0);
            } else if (p.hamil == 4) {                                     // qubit: P = -delta*E*jc, R = conj
                up[vi] = cscale(cscale(E, -delta), jc[vi]);
                lo[vi] = cmk(up[vi].x, -up[vi].y);
            } else if (vi == 1) {                                          // two_levels: -E_full, -conj(E_full)
                up[vi] = cmk(-E.x, -E.y);
                lo[vi] = cmk(-E.x, E.y);
            }
        }
        // ---- Newton recurrence ---------------------------------------------------------
#pragma unroll
        for (int n = 0; n < NL; ++n) {
            phi[n] = psi[n];
            psi[n] = cmul(psi[n], NCH > 0 ? dv_r[0] : dv[0]);
        }
        // residum (phys_base.py:98-111) in the reference's operand order: folding c1, c2, xp_j into the level
        // matrix shortens the chain by 8 % but moves the ill-conditioned Jz configuration beyond 1e-9
        auto order = [&](double xpj, cd dvj) {
            tri_apply<NL>(dg, up, lo, phi, tmp);
#pragma unroll
            for (int n = 0; n < NL; ++n) {
                cd r = cscale(tmp[n], c1);
                r = csub(r, cscale(phi[n], c2));
                r = cadd(cscale(phi[n], -xpj), r);
                phi[n] = r;
                psi[n] = cadd(psi[n], cmul(r, dvj));
            }
        };
        if (NCH > 0) {
#pragma unroll
            for (int j = 0; j < NCR - 1; ++j)
                if (j < nterms - 1) order(xp_r[j], dv_r[j + 1]);
        } else {
            for (int j = 0; j < nterms - 1; ++j) order(p.xp[j], dv[j + 1]);
        }
        double nsum = 0.0;
        double nl[NL];
#pragma unroll
        for (int n = 0; n < NL; ++n) {
            psi[n] = cmul(psi[n], phase);
            nl[n] = (psi[n].x * psi[n].x + psi[n].y * psi[n].y) * dx;
            nsum += nl[n];
        }
        if (!p.single_step) {
            if (p.renorm && fabs(nsum) > 0.0) {
                const double s = sqrt(fabs(nsum));
                const double rs = 1.0 / s;
#pragma unroll
                for (int n = 0; n < NL; ++n) psi[n] = cmk(div_rn_by(psi[n].x, s, rs), div_rn_by(psi[n].y, s, rs));
            }
            if (l == nt || l == p.l_first + p.nsteps - 1) {
#pragma unroll
                for (int n = 0; n < NL; ++n) p.norms[((long long)run * p.nb + vec) * NL + n] = nl[n];
            }
            if (p.traj_out) {
                cd *dst = p.traj_out + (long long)l * slice + gt;
#pragma unroll
                for (int n = 0; n < NL; ++n) dst[(long long)n * p.nthreads] = psi[n];
            }
        }
    }
    if (active && !retired) {
#pragma unroll
        for (int n = 0; n < NL; ++n) p.psi[((long long)run * p.nb + vec) * NL + n] = psi[n];
    }
}

}  // namespace qc
