// Kernel C: Newton-polynomial time stepper for ONE-level wavefunctions on a genuine two-dimensional grid,
//   H = T_x + T_y + V(x, y),   T_x = ifft_x(akx2 * fft_x .),  T_y likewise
// -- the "hamil_2d 512x512 grid" item of BASELINE config 5.  The reference has no such
// operator; this is a synthetic extension whose CPU restatement (the 2-D test checker under oracle/) is the 2-D analogue of
// phys_base.hamil_cpu / residum_cpu / prop_cpu with numpy.fft.fft2, and which is additionally tied to the
// reference's 1-D path through separable potentials (tests/test_gpu_grid2d.py).
//
// A 512 x 512 wavefunction (4 MB) does not fit one SM, so the kinetic term is applied as two 1-D passes: every
// polynomial order has a ROW phase (T_y phi + V phi, lines contiguous in memory) and a COLUMN phase
// (T_x phi, the affine map and the Newton accumulation), separated by grid-wide barriers of a cooperative
// launch; phi and the row-phase result travel between the phases through L2 (12 MB working set), while the
// accumulator psi stays in tensor memory in the column mapping for the whole time step.  64 CTAs of 8 warps per
// wavefunction, one 512-point line per warp (16 points per thread, the transform of kernel A at N = 512).
#pragma once
#include <cooperative_groups.h>

#include "qc_line.cuh"

namespace qc {

struct Grid2DParams {
    int n;                       // points per dimension (512)
    int nch, nsteps, renorm;
    int wave_first, wave_count;  // wavefunctions [wave_first, wave_first + wave_count) of the batch in this launch
    cd *psi;                     // [batch][n][n]  state (in / out)
    cd *phi;                     // [wave_count][n][n]  recurrence vector
    cd *tmp;                     // [wave_count][n][n]  row-phase result
    const double *v;             // [vb][n][n]
    long long v_stride;          // 0 or n*n
    const double *kinx, *kiny;   // [n] akx2, aky2 (real parts)
    const cd *tw1, *tw2;         // twiddle roots of the 512-point transform (same tables as kernel A)
    const double *xp;            // [nch]
    const cd *dv;                // [nch]
    double emin, emax, t_sc;
    double dxdy;                 // area element of the norm
    double *norm_acc;            // [nsteps][wave_count][TILES]  per-CTA partial sums (summed in a fixed order)
    double *norms;               // [batch] last norm before renormalisation
};

constexpr int G2_LOG2N = 9;
constexpr int G2_LINES = 8;      // lines (warps) per CTA

struct Grid2DCfg {
    typedef GridCfg<G2_LOG2N> L;
    static constexpr int N = L::N;
    static constexpr int NTHREADS = 32 * G2_LINES;
    static constexpr int TILES = N / G2_LINES;         // CTAs per wavefunction
    static constexpr int TS = N + 1;                   // row stride of the column-phase staging tile (bank skew)
    static constexpr size_t SMEM_BYTES =
        sizeof(cd) * (size_t)(G2_LINES * L::BUF + G2_LINES * TS) + sizeof(double) * (3 * 128 + 32);
};

__device__ __forceinline__ int tile_first_line(unsigned block) { return (int)(block % Grid2DCfg::TILES) * G2_LINES; }

__global__ void __launch_bounds__(Grid2DCfg::NTHREADS, 1) grid2d_sweep_kernel(Grid2DParams p) {
    typedef Grid2DCfg C;
    typedef GridCfg<G2_LOG2N> L;
    constexpr int N = C::N, TPL = L::TPL;
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd *sm = reinterpret_cast<cd *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    cd *buf = sm + warp * L::BUF;
    // column phase staging: the CTA's 8 adjacent columns are 128 contiguous bytes in every grid row, so the tile is
    // moved with whole 128-byte lines (8 threads per row) and transposed through shared memory -- a warp reading
    // its column straight from global memory would touch 32 different lines per load instruction
    cd *tile = sm + G2_LINES * L::BUF;
    cd *s_dv = tile + G2_LINES * C::TS;
    double *s_xp = reinterpret_cast<double *>(s_dv + 128);
    double *s_red = s_xp + 128;
    const int wslot = blockIdx.x / C::TILES;            // wavefunction slot of this launch
    const int tix = blockIdx.x % C::TILES;
    const int line = tix * G2_LINES + warp;            // the row (row phase) / column (column phase) of this warp
    const long long wave = p.wave_first + wslot;
    const long long NN = (long long)N * N;
    cd *psi_g = p.psi + wave * NN;
    cd *phi_g = p.phi + (long long)wslot * NN;
    cd *tmp_g = p.tmp + (long long)wslot * NN;
    const double *v_g = p.v + p.v_stride * wave;

    for (int i = tid; i < p.nch; i += C::NTHREADS) {
        s_xp[i] = p.xp[i];
        s_dv[i] = p.dv[i];
    }
    const double c1 = 4.0 / (p.emax - p.emin), c2 = 2.0 * (p.emax + p.emin) / (p.emax - p.emin);
    double ph_s, ph_c;
    sincos(-2.0 * p.t_sc * (p.emax + p.emin) / (p.emax - p.emin), &ph_s, &ph_c);
    const cd phase = cmk(ph_c, ph_s);
    double kcx[16], kcy[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int k = line_freq<G2_LOG2N>(lane, e);
        kcx[e] = p.kinx[k] * (c1 / (double)N);
        kcy[e] = p.kiny[k] * (c1 / (double)N);
    }
    // tensor memory: accumulator psi (column mapping) and the two twiddle tables, as in kernel A
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(s_red + 30);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_tmem)),
                     "n"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    const uint32_t t_psi = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 256);
    const uint32_t t_phi = t_psi + 64;      // phi in the column mapping: the output of a column phase is its next input
    const uint32_t t_tw1 = t_psi + 128, t_tw2 = t_psi + 192;
    tmem_fill_twiddles(t_tw1, p.tw1[L::S1 + lane]);
    tmem_fill_twiddles(t_tw2, p.tw2[L::R3 + (lane % L::R3)]);
    tmem_wait_st();
    __syncthreads();
    auto line_sync = [&]() { __syncwarp(); };
    // grid-wide barrier of the cooperative launch (a hand-written per-wavefunction barrier -- one atomic arrival per
    // CTA on a growing counter -- measured 886 us per step against 856 us with this one)
    auto grid_barrier = [&]() {
        __syncthreads();
        grid.sync();
    };
    double cv[16];                                      // c1 * V on this warp's row (row mapping never changes)
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) cv[n1] = c1 * v_g[(long long)line * N + lane + TPL * n1];
    const int y0 = tile_first_line(blockIdx.x);
    auto tile_load = [&](const cd *g) {                 // tile[c][x] <- g[x][y0 + c]
        const int c = tid & 7, r = tid >> 3;
#pragma unroll 4
        for (int it = 0; it < N / 32; ++it) tile[c * C::TS + it * 32 + r] = g[(long long)(it * 32 + r) * N + y0 + c];
    };
    auto tile_store = [&](cd *g) {                      // g[x][y0 + c] <- tile[c][x]
        const int c = tid & 7, r = tid >> 3;
#pragma unroll 4
        for (int it = 0; it < N / 32; ++it) g[(long long)(it * 32 + r) * N + y0 + c] = tile[c * C::TS + it * 32 + r];
    };
    cd *col = tile + warp * C::TS;                      // this warp's column inside the tile
    const int last_j = p.nch - 2;

    for (int step = 0; step < p.nsteps; ++step) {
        cd a[16], own[16];
        // ---- psi accumulator <- dv0 * psi (column mapping: x = lane + 32 n1, y = line) ---------------------
        tile_load(psi_g);
        __syncthreads();
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) own[n1] = col[lane + TPL * n1];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) a[n1] = cmul(own[n1], s_dv[0]);
#pragma unroll
        for (int c4 = 0; c4 < 16; c4 += 4) {
            tmem_store_c4(t_psi + c4 * 4, a + c4);
            tmem_store_c4(t_phi + c4 * 4, own + c4);       // phi_0 = psi
        }
        __syncthreads();
        for (int j = 0; j <= last_j; ++j) {
            const cd *src = j == 0 ? psi_g : phi_g;           // phi_0 = psi
            // ---- row phase: tmp = c1 * (T_y phi + V phi) on row x = line, points y = lane + 32 n1 ----------
            {
                const cd *row = src + (long long)line * N;
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) own[n1] = a[n1] = row[lane + TPL * n1];
                kinetic_line<G2_LOG2N>(a, buf, lane, kcy, t_tw1, t_tw2, line_sync);
                cd *out = tmp_g + (long long)line * N;
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1)
                    out[lane + TPL * n1] = cmk(fma(cv[n1], own[n1].x, a[n1].x), fma(cv[n1], own[n1].y, a[n1].y));
            }
            grid_barrier();
            // ---- column phase: phi' = tmp + c1 T_x phi - (c2 + xp_j) phi ; psi += dv_{j+1} phi' ------------
            {
                const double shift = c2 + s_xp[j];
                const cd dvj = s_dv[j + 1];
                tile_load(tmp_g);                             // stays in the tile until the pointwise stage
                tmem_wait_st();
#pragma unroll
                for (int c4 = 0; c4 < 16; c4 += 4) {
                    uint32_t rs[16];
                    tmem_ld16(t_phi + c4 * 4, rs);
                    tmem_wait_ld();
                    tmem_unpack_c4(rs, own + c4);
                }
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) a[n1] = own[n1];
                kinetic_line<G2_LOG2N>(a, buf, lane, kcx, t_tw1, t_tw2, line_sync);
                __syncthreads();
#pragma unroll
                for (int c4 = 0; c4 < 16; c4 += 4) {
                    uint32_t rp[16];
                    tmem_ld16(t_psi + c4 * 4, rp);
                    cd acc4[4];
                    tmem_wait_ld();
                    tmem_unpack_c4(rp, acc4);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int n1 = c4 + i;
                        const cd r1 = col[lane + TPL * n1];
                        cd r;
                        r.x = fma(-shift, own[n1].x, a[n1].x + r1.x);
                        r.y = fma(-shift, own[n1].y, a[n1].y + r1.y);
                        col[lane + TPL * n1] = r;
                        own[n1] = r;
                        acc4[i] = cfma(dvj, r, acc4[i]);
                    }
                    tmem_store_c4(t_psi + c4 * 4, acc4);
                    tmem_store_c4(t_phi + c4 * 4, own + c4);
                }
                __syncthreads();
                tile_store(phi_g);
            }
            grid_barrier();
        }
        // ---- final phase, norm over the whole grid, renormalisation, state out (column mapping) -------------
        tmem_wait_st();
        double nrm = 0.0;
#pragma unroll
        for (int c4 = 0; c4 < 16; c4 += 4) {
            uint32_t rp[16];
            tmem_ld16(t_psi + c4 * 4, rp);
            tmem_wait_ld();
            tmem_unpack_c4(rp, a + c4);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            a[i] = cmul(a[i], phase);
            nrm = fma(a[i].x, a[i].x, fma(a[i].y, a[i].y, nrm));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        if (lane == 0) s_red[warp] = nrm;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < G2_LINES; ++w) s += s_red[w];
            p.norm_acc[((long long)step * p.wave_count + wslot) * C::TILES + tix] = s;
        }
        grid_barrier();
        double whole = 0.0;
        {
            const double *part = p.norm_acc + ((long long)step * p.wave_count + wslot) * C::TILES;
            for (int i = 0; i < C::TILES; ++i) whole += part[i];
        }
        const double tot = fabs(whole * p.dxdy);
        if (p.renorm && tot > 0.0) {
            const double s = sqrt(tot), rs = 1.0 / s;
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = cmk(div_rn_by(a[i].x, s, rs), div_rn_by(a[i].y, s, rs));
        }
        if (tix == 0 && tid == 0) p.norms[wave] = tot;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) col[lane + TPL * n1] = a[n1];
        __syncthreads();
        tile_store(psi_g);
        grid_barrier();
    }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

}  // namespace qc
