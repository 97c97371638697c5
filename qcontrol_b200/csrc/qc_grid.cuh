// Kernel A: fused Newton-polynomial time stepper for the two-level FFT-grid
// Hamiltonian (Hamil2DNonTrivial, hamil_2d.py:34-73).
//
// One CTA owns one run (one two-level wavefunction) for a whole launch of `nsteps` time steps.
// 2*TPL threads, TPL = N/16: thread (level, t) owns the 16 grid points x = t + TPL*n1 of its level.
//
// Where the data lives
//   registers      the recurrence vector phi (16 complex), the diagonal d(x) and the k-space multiplier
//   tensor memory  per thread 256 columns: the accumulator psi, the thread's own copy of phi, and the two
//                  twiddle tables w_N^(k t), w_{16 R3}^(k n3) (tcgen05.st once, tcgen05.ld per use).  TMEM
//                  has its own datapath (measured 790 B/clk/SM read, 930 B/clk/SM write on B200,
//                  tools/tmem_bw.cu), so none of this competes with shared memory
//   shared memory  only what has to cross threads: the four FFT transposes per polynomial order and the
//                  phi hand-over between the two levels (coupling term -E*phi_other)
//   HBM            once per time step: trajectory slice in / out (Krotov), the state at launch begin / end
//
// FFT: N = 16*16*R3 (R3 = N/256), decimation in frequency forward, the exact mirror inverse, so no
// reordering pass is needed and the k-space multiply runs on the permuted layout.
//
// Synchronisation (N >= 512): the two levels are separate groups of warps that only meet at the coupling
// term.  Each level synchronises its own transposes on a private named barrier; phi is handed over through
// two buffers (order parity) with one "available" named barrier each (bar.arrive by the producer, bar.sync by
// the consumer), so a level waits for the other one only when it is more than a whole polynomial order ahead
// (+5 % over a single hand-over buffer; forcing a half-order skew between the levels on top of that changed
// nothing).  No "has been read" handshake is needed: a level passes its wait for the partner's order j+1 only
// after the partner has finished order j, i.e. has read this level's buffer of order j, and that buffer is not
// rewritten before order j+2 (dropping the handshake: +1.6 % at np = 2048).
//
// Packed form (RPC > 0, np <= 512): RPC runs per CTA, so that an SM holds eight warps instead of two or four.
// np = 256 (ILV): four runs per CTA, the two levels of a run interleaved in one warp (lanes L and L^16), the coupling
// term a warp shuffle in k space, no barrier wider than a warp (0.26 -> 0.60 of the FP64 peak).  np = 512: two runs
// per CTA, a level per warp and the hand-over of the plain form with five named barriers per run (0.42 -> 0.52;
// the interleaved layout measured 0.51 here, 0.51 against 0.56 at np = 1024 / 2048).
//
// History of this kernel with ncu evidence: profiles/r01a (twiddle tables in shared memory, 0.41 of the
// FP64 peak), r01c (register power chains, 0.48), r01d/r01e (TMEM accumulator and tables, 0.52), r01g (one
// transpose buffer + double-buffered hand-over, 0.555).
#pragma once
#include <cooperative_groups.h>

#include <cstdint>

#include "qc_fft.cuh"
#include "qc_local.cuh"

#ifndef QC_RDV_MBAR
#define QC_RDV_MBAR 1
#endif
#ifndef QC_TMH_MBAR
#define QC_TMH_MBAR 0
#endif
#ifndef QC_CL_SPLIT_PUSH
#define QC_CL_SPLIT_PUSH 1
#endif
#ifndef QC_CL_LATE_ARRIVE
#define QC_CL_LATE_ARRIVE 2
#endif

namespace qc {

struct GridParams {
    int batch, nch, nt_max;
    int run_first, run_count;    // this launch covers runs [run_first, run_first + run_count)
    int hf_hide, renorm, field_mode, reversed_E, write_E_base;
    int l_first, nsteps;
    int single_step;             // 1: qc_prop_step semantics (explicit E, no transform / renorm)
    int cplx_coupling;           // 1 (single_step only): complex E_full, level 0 gets -E_full*psi_1, level 1 -conj(E_full)*psi_0
    cd *psi;                     // [batch][2][N]
    const double *v;             // [vb][2][N]
    long long v_stride;
    const double *kin;           // [N]
    const cd *tw1;               // [16][S1]  (row 1 = w_N^t is the only row the kernel reads)
    const cd *tw2;               // [16][R3]  (row 1 = w_{16 R3}^{n3})
    const double *xp;            // [nch]
    const cd *dv;                // [pb][nch]
    const double *sc;            // [pb][4] emin emax t_sc eL
    const int *orders;           // [pb] polynomial terms summed per run (truncated series, opt-in), or null = nch
    int poly_stride;             // 0 or 1 (in runs)
    const double *dx;            // [db]
    int dx_stride;
    cd *E_base;                  // [eb][nt_max+1]
    int E_stride;                // 0 or 1 (in runs)
    const cd *expL;              // [xb][nt_max+1]
    int expL_stride;
    const double *ctl;           // [cb][4] h_lambda, nt
    int ctl_stride;
    const cd *E_step;            // [batch] explicit field for single_step
    cd *E_out;                   // [batch][nt_max+1]
    const cd *traj_in;           // [batch][nt_max+1][N]  or null
    cd *traj_out;                // [batch][nt_max+1][N]  or null
    double *norms;               // [batch][2]
    // field_mode 4 (local control with a fed-back carrier-frequency multiplier, qc_local.cuh)
    double *lc;                  // [batch][LC_STRIDE] state and constants of the multiplier ODE
    const double *lc_t;          // [batch][nt_max+1] step times t_l
    const double *lc_rt;         // [nch][nch] transposed partial products of the Newton table
    const double *lc_rdiag;      // [nch] reciprocal pivots
    const cd *tw8;               // np = 2048, 512-thread form (qc_grid8.cuh): [8][256] w_2048^(k1 T), [8][32] w_256^(k2 m), [8][4] w_32^(k3 n)
};

// CL = false: one CTA holds both levels of a run.  CL = true: a cluster of two CTAs, one level each, on two SMs --
// the only form for np = 4096 (a level alone fills the registers and shared memory of an SM) and the
// low-latency form for batches too small to fill the GPU; the levels talk through distributed shared memory.
// RPC > 0: the PACKED form for small grids -- RPC runs per CTA so that a CTA has four warps (all four tensor-memory
// lane quadrants in use) and an SM holds eight.  ILV: the two levels of a run are interleaved inside one warp (lanes
// L and L^16; np = 256, where a level is half a warp); otherwise a level is one warp (np = 512), see the kernel.
// VAR (plain one-run-per-CTA form at np = 2048 only) selects kernel variants that are compared on the GPU by
// tools/variant_compare.py:  bit 0 (TMH) = the phi hand-over between the two levels goes through TENSOR MEMORY instead
// of shared memory.  Thread t of level 0 (warp w) and thread t of level 1 (warp w + 4) sit on the same tensor-memory
// lane, so (a) they can share ONE copy of each twiddle table, which frees 128 of the pair's 512 columns, and (b) the
// freed columns hold the second phi buffer of each level, so the partner reads this level's phi of either order parity
// straight from the lane with tcgen05.ld.  Shared memory then carries nothing but the four FFT transposes per order:
// 10 -> 8 complex128 accesses per point and order (-20 % of the busiest pipe), 200 KB -> 68 KB per CTA.
// With RPC = 128 / TPL runs per CTA (np = 1024: two runs) the same pairing holds for smaller grids: run r owns lane
// quadrants r * WPL .. (r + 1) * WPL - 1 (WPL = TPL / 32 warps per level), level 0 in warps 0-3, level 1 in warps 4-7.
template <int LOG2N, bool CL = false, int RPC = 0, bool ILV = false, int VAR = 0>
struct GridCfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int TPL = N / 16;          // threads per level
    static constexpr int R3 = N / 256;          // third radix (1,2,4,8,16)
    static constexpr int Q = 16 / R3;           // R3-point DFTs per thread in pass 3
    static constexpr int S1 = TPL + (R3 < 8 ? R3 : 0);   // padded row stride of the transpose buffers
    static constexpr int BUF = 16 * S1;         // complex elements per transpose buffer
    static constexpr int NLEV_CTA = CL ? 1 : 2; // levels per CTA
    static constexpr bool PK = RPC > 0;
    static constexpr int RT = 2 * TPL;          // threads per run (packed form)
    static constexpr int NTHREADS = PK ? RPC * RT : NLEV_CTA * TPL;
    static constexpr int MAXCH = 128;
    static constexpr bool NAMED = !CL && !PK && TPL >= 32;   // a level is made of whole warps: per-level named barriers
    static constexpr bool TMH = (VAR & 1) != 0;              // phi hand-over through tensor memory
    static_assert(!TMH || (!CL && !ILV && (PK ? RPC * TPL == 128 : TPL == 128)),
                  "tensor-memory hand-over: thread t of the two levels must sit in warps w and w + 4");
    static constexpr int TM_COLS = NTHREADS > 128 ? 512 : 256;   // 256 TMEM columns per warp slot
    // shared memory (complex elements): per level ONE transpose buffer (forward and inverse passes use it in turn:
    // every element is either private to an R3-lane group or read by the thread that wrote it, except across the two
    // level barriers) and TWO phi hand-over buffers, so that the levels may run up to one polynomial order apart;
    // then dv, xp and 32 doubles of reduction scratch (the last one carries the TMEM base address)
    // packed form: per run two transpose buffers (level-per-warp layout: and the four hand-over buffers), dv, xp and
    // 16 doubles of reduction scratch; then 32 doubles per CTA
    static constexpr int LEVEL_ELEMS = (ILV || TMH) ? BUF : BUF + 2 * N;     // complex elements of shared memory per level
    static constexpr size_t RUN_BYTES = sizeof(cd) * (size_t)(2 * LEVEL_ELEMS + MAXCH) + sizeof(double) * (size_t)(MAXCH + 16);
    // TMA staging of the trajectory stream (plain TMH form): the slice of the previous opposite sweep that the field
    // law of the NEXT time step reads is fetched by one cp.async.bulk (the TMA engine's 1-D bulk copy, SASS UBLKCP)
    // into a shared-memory buffer while the current step computes; an mbarrier with a byte count says when it landed.
    // The shared memory for it is what the tensor-memory hand-over freed.
    static constexpr bool TMA_SLICE = TMH && !PK;
    static constexpr size_t PLAIN_BYTES =
        sizeof(cd) * (size_t)(NLEV_CTA * LEVEL_ELEMS) + sizeof(double) * (size_t)(3 * MAXCH + 32 + LC_STRIDE + LS_COUNT);
    static constexpr size_t SLICE_OFFSET = (PLAIN_BYTES + 127) / 128 * 128;
    static constexpr size_t SMEM_BYTES =
        PK ? RPC * RUN_BYTES + sizeof(double) * 32 : TMA_SLICE ? SLICE_OFFSET + sizeof(cd) * (size_t)N + 16 : PLAIN_BYTES;
};

// skewed position of element (k2, n3) inside a lane group's private 16*R3 region
template <int R3, int Q>
__device__ __forceinline__ int pos2(int k2, int n3) {
    return k2 * R3 + ((n3 + k2 / Q) & (R3 - 1));
}

__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- TMA 1-D bulk copy global -> shared memory, completion counted in bytes on an mbarrier -------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(b), "r"(parity)
            : "memory");
    } while (!done);
}

// ---- tensor memory as a per-thread scratchpad -----------------------------------------------------------
// Each thread owns one TMEM lane (warp w -> lanes 32*(w%4)..+31; warps w and w+4 use different columns).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// four complex128 values <-> 16 TMEM columns
__device__ __forceinline__ void tmem_store_c4(uint32_t taddr, const cd *v) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r[4 * i + 0] = (uint32_t)__double2loint(v[i].x);
        r[4 * i + 1] = (uint32_t)__double2hiint(v[i].x);
        r[4 * i + 2] = (uint32_t)__double2loint(v[i].y);
        r[4 * i + 3] = (uint32_t)__double2hiint(v[i].y);
    }
    tmem_st16(taddr, r);
}
__device__ __forceinline__ void tmem_unpack_c4(const uint32_t (&r)[16], cd *v) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        v[i] = cmk(__hiloint2double((int)r[4 * i + 1], (int)r[4 * i + 0]), __hiloint2double((int)r[4 * i + 3], (int)r[4 * i + 2]));
}
// per-thread twiddle table in TMEM: 16 complex values in radix-4 group order, T[4 g + h] = w^(g + 4 h)
struct TmemTwiddles {
    uint32_t taddr;
    uint32_t r[16];
    __device__ __forceinline__ explicit TmemTwiddles(uint32_t a) : taddr(a) {}
    __device__ __forceinline__ void prefetch(int g) { tmem_ld16(taddr + 16 * g, r); }
    __device__ __forceinline__ void get(cd *w4) {
        tmem_wait_ld();
        tmem_unpack_c4(r, w4);
    }
};
// fill the table of one root: powers by pairing (depth log2 k, ~4 roundings), stored group-major
__device__ __forceinline__ void tmem_fill_twiddles(uint32_t taddr, const cd w) {
    cd pw[16];
    pw[0] = cmk(1.0, 0.0);
    pw[1] = w;
#pragma unroll
    for (int k = 2; k < 16; ++k) pw[k] = cmul(pw[k / 2], pw[k - k / 2]);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        cd grp[4] = {pw[g], pw[g + 4], pw[g + 8], pw[g + 12]};
        tmem_store_c4(taddr + 16 * g, grp);
    }
}

// Sum `val` over the threads of each level; every thread gets both sums.
template <int NTHREADS, int TPL>
__device__ __forceinline__ void level_sums(double val, double *s_red, int tid, double &sum0, double &sum1) {
    constexpr int G = TPL < 32 ? TPL : 32;        // lanes that share a level inside a warp
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if ((tid & (G - 1)) == 0) s_red[tid / G] = val;
    __syncthreads();
    constexpr int NG = NTHREADS / G;
    sum0 = 0.0;
    sum1 = 0.0;
#pragma unroll
    for (int g = 0; g < NG / 2; ++g) sum0 += s_red[g];
#pragma unroll
    for (int g = NG / 2; g < NG; ++g) sum1 += s_red[g];
    __syncthreads();
}

// Packed form: lanes 0-15 of every warp belong to level 0, lanes 16-31 to level 1; a run is W warps that meet on
// the run's own named barrier.
template <int W>
__device__ __forceinline__ void packed_sums(double val, double *s_red, int rtid, int bar_id, double &sum0, double &sum1) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    const double oth = __shfl_xor_sync(0xffffffffu, val, 16);
    const bool hi = (rtid >> 4) & 1;
    const double a0 = hi ? oth : val, a1 = hi ? val : oth;      // this warp's share of the two level sums
    if (W == 1) {
        sum0 = a0;
        sum1 = a1;
        return;
    }
    if ((rtid & 31) == 0) {
        s_red[2 * (rtid >> 5)] = a0;
        s_red[2 * (rtid >> 5) + 1] = a1;
    }
    bar_sync(bar_id, 32 * W);
    sum0 = 0.0;
    sum1 = 0.0;
#pragma unroll
    for (int w = 0; w < W; ++w) {
        sum0 += s_red[2 * w];
        sum1 += s_red[2 * w + 1];
    }
    bar_sync(bar_id, 32 * W);
}

// Packed form, level-per-warp layout: the first W/2 warps of a run are level 0, the others level 1.
template <int W>
__device__ __forceinline__ void packed_level_sums(double val, double *s_red, int rtid, int bar_id, double &sum0, double &sum1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if ((rtid & 31) == 0) s_red[rtid >> 5] = val;
    bar_sync(bar_id, 32 * W);
    sum0 = 0.0;
    sum1 = 0.0;
#pragma unroll
    for (int w = 0; w < W / 2; ++w) {
        sum0 += s_red[w];
        sum1 += s_red[W / 2 + w];
    }
    bar_sync(bar_id, 32 * W);
}

// ---- cluster form: split-phase cluster barrier and sums over the two CTAs ----------------------------------
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// Sum `val` over the threads of this CTA (= one level) and fetch the partner CTA's sum through distributed
// shared memory; every thread gets (sum of level 0, sum of level 1).  Two full cluster barriers.
template <int NTHREADS>
__device__ __forceinline__ void cluster_sums(double val, double *s_red, int tid, int lev, double &sum0, double &sum1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = val;
    __syncthreads();
    double own = 0.0;
#pragma unroll
    for (int g = 0; g < NTHREADS / 32; ++g) own += s_red[g];
    if (tid == 0) s_red[16] = own;
    cl_arrive();
    cl_wait();
    const double *remote = cooperative_groups::this_cluster().map_shared_rank(s_red + 16, 1 - lev);
    const double oth = *remote;
    cl_arrive();
    cl_wait();
    sum0 = lev == 0 ? own : oth;
    sum1 = lev == 0 ? oth : own;
}

template <int LOG2N, bool LOCAL = false, bool CL = false, int RPC = 0, bool ILV = false, int VAR = 0>
__global__ void __launch_bounds__(GridCfg<LOG2N, CL, RPC, ILV, VAR>::NTHREADS, 1) grid_sweep_kernel(GridParams p) {
    static_assert(!(LOCAL && CL), "local control runs in the single-CTA form");
    static_assert(!(RPC > 0 && (LOCAL || CL)), "the packed form exists for the plain single-CTA kernel only");
    static_assert(!ILV || RPC > 0, "interleaved levels are a layout of the packed form");
    typedef GridCfg<LOG2N, CL, RPC, ILV, VAR> C;
    static_assert(ILV || RPC == 0 || C::TMH || GridCfg<LOG2N>::TPL == 32, "packed level-per-warp layout: a level is one warp");
    static_assert(!(C::TMH && LOCAL), "local control keeps the shared-memory hand-over (it reuses the buffers)");
    constexpr bool TMH = C::TMH;
    constexpr int N = C::N, TPL = C::TPL, R3 = C::R3, Q = C::Q, S1 = C::S1, BUF = C::BUF;
    constexpr bool PK = C::PK;
    constexpr bool PKL = PK && !ILV;              // packed, a level per warp: shared-memory hand-over like the plain form
    constexpr bool NAMED = C::NAMED || PKL;       // hand-over through full barriers (bar.arrive / bar.sync)
    constexpr int HN = PK ? C::RT : C::NTHREADS;  // threads that meet on a hand-over barrier
    constexpr int RW = C::RT / 32;                // warps per run (packed form)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    // packed form: run slot rslot of this CTA, thread rtid of the run.  ILV: the two levels of grid point set t sit
    // in lanes L and L^16 of one warp, so the coupling term is a warp shuffle (in k space, where both levels hold
    // the same frequencies) instead of a shared-memory hand-over.  A run synchronises on named barriers of its own
    // TMP (tensor-memory hand-over with several runs per CTA): level 0 of every run in warps 0-3, level 1 in warps 4-7,
    // run r in the lane quadrants r * WPL ..; rtid is the thread's index inside its run, level 0 first, as elsewhere
    constexpr bool TMP = TMH && PK;
    constexpr int WPL = TPL >= 32 ? TPL / 32 : 1;      // warps per level
    const int rslot = !PK ? 0 : TMP ? ((tid >> 5) & 3) / WPL : tid / C::RT;
    const int rtid = !PK ? tid : TMP ? (tid >> 7) * TPL + (((tid >> 5) & 3) % WPL) * 32 + (tid & 31) : tid - rslot * C::RT;
    cd *sm = reinterpret_cast<cd *>(smem_raw + (PK ? rslot * C::RUN_BYTES : 0));
    const int lev = CL ? (int)cooperative_groups::this_cluster().block_rank() : ILV ? (rtid >> 4) & 1 : rtid / TPL;
    const int t = CL ? tid : ILV ? (rtid & 15) | ((rtid >> 5) << 4) : rtid - lev * TPL;
    cd *bufA = sm + (CL ? 0 : lev) * C::LEVEL_ELEMS;
    cd *bufB = bufA;
    // hand-over buffers: in the cluster form a level PUSHES its phi into the partner CTA's shared memory (remote
    // stores are fire-and-forget), so that the reads in the pointwise phase stay local
    cd *stash0 = CL ? cooperative_groups::this_cluster().map_shared_rank(bufA + BUF, 1 - lev) : bufA + BUF;
    const cd *stash0_other = CL ? bufA + BUF : sm + (1 - lev) * C::LEVEL_ELEMS + BUF;
    cd *s_dv = sm + (ILV ? 2 * BUF : C::NLEV_CTA * C::LEVEL_ELEMS);
    double *s_xp = reinterpret_cast<double *>(s_dv + C::MAXCH);
    double *s_red = s_xp + C::MAXCH;              // 32 doubles of reduction scratch
    double *s_lc = s_red + 32;                    // local control: the run's state / constants ...
    double *s_ls = s_lc + LC_STRIDE;              // ... and the scalars of the current step

    const int slot = CL ? (int)blockIdx.x / 2 : PK ? (int)blockIdx.x * RPC + rslot : (int)blockIdx.x;     // cluster-uniform
    // packed form: an empty run slot of the last CTA keeps its threads (they take part in the CTA-wide barriers
    // around the tensor-memory allocation), reads the first run and does no step and no store
    const bool valid = slot < p.run_count && p.run_first + slot < p.batch;
    if (!PK && !valid) return;
    const int run = valid ? p.run_first + slot : p.run_first;
    auto level_sync = [&]() {                     // all threads of this thread's level (packed form: of the run)
        if (TMP && WPL > 1) bar_sync(1 + 7 * rslot + lev, TPL);
        else if (PKL) __syncwarp();
        else if (PK) {
            if (RW == 1) __syncwarp(); else bar_sync(1 + rslot, C::RT);
        } else if (NAMED) bar_sync(1 + lev, TPL); else __syncthreads();
    };
    // named barrier ids.  Plain form: 1,2 level-local transposes; 3..6 "phi of level n, order parity p is available".
    // Packed form: interleaved layout one id per run (1 + rslot); level-per-warp layout five per run (four "available"
    // ids and one for sums).  A "phi has been read" handshake is not needed: the buffers alternate with the order
    // parity, and a level passes its wait for the partner's order j+1 only after the partner has finished order j.
    // TMP: seven per run (two level barriers, four "available" ids, one for sums).  RDV (TMP with a level per warp,
    // np = 512, four runs per CTA): a run is two warps, and 4 x 5 ids would not fit the 15 a CTA has, so the four
    // "available" signals of a run are MBARRIERS in shared memory (32 arrivals each: every lane of the producer warp
    // arrives once its own tensor-memory stores have completed; the consumer warp spins on the phase parity) -- the
    // same one-order slack as the named barriers of the other forms.  (QC_RDV_MBAR = 0: the symmetric rendezvous of
    // the two warps once per order on one named barrier, the first version of this form.)
    constexpr bool RDV = TMP && WPL == 1;
    // QC_TMH_MBAR = 1 (diagnostic): the same mbarriers (TPL arrivals each) in the other tensor-memory hand-over forms
    // instead of the named barriers, whose bar.sync also makes the consumer warps of a level wait for each other --
    // measured neutral at np = 2048 (prescribed 0.616 / 0.615, Krotov forward 0.602 / 0.611) and worse at np = 1024
    // (0.591 / 0.621), so the named barriers stay
    constexpr bool RDVM = TMH && (RDV ? QC_RDV_MBAR : QC_TMH_MBAR);
    uint64_t *s_hb = PK ? reinterpret_cast<uint64_t *>(smem_raw + RPC * C::RUN_BYTES + 64) + 4 * rslot
                        : reinterpret_cast<uint64_t *>(s_red + 24);                                 // [lev][parity]
    uint32_t hb_use0 = 0, hb_use1 = 0;            // completed waits on the partner's barrier of parity 0 / 1
    const int bar_run = RDV ? 1 + rslot : TMP ? 1 + 7 * rslot + 6 : PKL ? 1 + 5 * rslot + 4 : 1 + rslot;
    const int bar_full = TMP ? 1 + 7 * rslot + 2 : PKL ? 1 + 5 * rslot : 3;
    const int bar_rdv = 1 + RPC + rslot;
    auto both_sums = [&](double val, double &s0, double &s1) {
        if (CL) cluster_sums<C::NTHREADS>(val, s_red, tid, lev, s0, s1);
        else if (PKL) packed_level_sums<RW>(val, s_red, rtid, bar_run, s0, s1);
        else if (PK) packed_sums<RW>(val, s_red, rtid, bar_run, s0, s1);
        else level_sums<C::NTHREADS, TPL>(val, s_red, tid, s0, s1);
    };
    const bool writer = (PK ? rtid == 0 && valid : tid == 0) && (!CL || lev == 0);   // the one thread that writes per-run scalars
    const int k1p = t / R3;       // k1' : which pass-1 output row this thread transforms in pass 2
    const int n3p = t % R3;       // n3' (also the lane's index c inside its group in pass 3)
    const int BAR_FULL_OWN = bar_full + 2 * lev, BAR_FULL_OTH = bar_full + 2 * (1 - lev);

    const long long prun = (long long)p.poly_stride * run;
    for (int i = rtid; i < p.nch; i += (PK ? C::RT : C::NTHREADS)) {
        s_xp[i] = p.xp[i];
        s_dv[i] = p.dv[prun * p.nch + i];
    }
    if (LOCAL && tid < LC_STRIDE) s_lc[tid] = p.lc[(long long)run * LC_STRIDE + tid];
    double emin = p.sc[prun * 4 + 0], emax = p.sc[prun * 4 + 1], t_sc = p.sc[prun * 4 + 2], eL = p.sc[prun * 4 + 3];
    double c1 = 4.0 / (emax - emin);                             // phys_base.py:100
    double c2 = 2.0 * (emax + emin) / (emax - emin);             // phys_base.py:101
    const double dx = p.dx[(long long)p.dx_stride * run];
    const double h_lambda = p.ctl[(long long)p.ctl_stride * run * 4 + 0];
    const int nt = (int)p.ctl[(long long)p.ctl_stride * run * 4 + 1];
    double ph_s, ph_c;
    sincos(-2.0 * t_sc * (emax + emin) / (emax - emin), &ph_s, &ph_c);   // phys_base.py:174
    cd phase = cmk(ph_c, ph_s);

    // per-thread constants in registers: diagonal d(x) = c1*(v(x) +- eL) - c2 and the k-space multiplier
    // c1/N * akx2[k] at the (permuted) frequencies this thread holds after the forward transform
    double dcoef[16], kc[16];
    auto load_coefs = [&]() {
        const double *vr = p.v + p.v_stride * run + (long long)lev * N;
        const double sgn = lev == 0 ? eL : -eL;                  // hamil_2d.py:50-51, 66-67
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) dcoef[n1] = fma(c1, vr[t + TPL * n1] + sgn, -c2);
        const double kscale = c1 / (double)N;                     // 1/N of numpy's ifft
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int q = e / R3, k3 = e % R3;
            const int k = k1p + 16 * (n3p * Q + q) + 256 * k3;
            kc[e] = p.kin[k] * kscale;
        }
    };
    load_coefs();

    cd psi[16], a[16];
    {
        const cd *src = p.psi + ((long long)run * 2 + lev) * N;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) psi[n1] = src[t + TPL * n1];
    }
    // ---- tensor memory: one warp allocates, everybody learns the base address ------------------------
    uint32_t *s_tmem = PK ? reinterpret_cast<uint32_t *>(smem_raw + RPC * C::RUN_BYTES) : reinterpret_cast<uint32_t *>(s_red + 30);
    if (RDVM && tid == 32) {
        uint64_t *hb = PK ? reinterpret_cast<uint64_t *>(smem_raw + RPC * C::RUN_BYTES + 64) : reinterpret_cast<uint64_t *>(s_red + 24);
        for (int i = 0; i < 4 * (PK ? RPC : 1); ++i) mbar_init(hb + i, TPL);
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(s_tmem)),
                     "n"(C::TM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *s_tmem;
    const int warp = tid >> 5;
    // plain layout: 256 private columns per warp slot: psi, own phi, the two twiddle tables.
    // TMH layout: the 512 columns of a lane are shared by thread t of level 0 (warp w) and of level 1 (warp w + 4):
    // [0,128) psi of level 0 / 1, [128,256) ONE copy of the two twiddle tables (same t, same tables), [256,512) phi of
    // level 0 / 1 for even / odd polynomial orders -- the partner reads them in place
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_psi = TMH ? t_lane + (uint32_t)(lev * 64) : t_lane + (uint32_t)((warp >> 2) * 256);
    const uint32_t t_phi = TMH ? t_lane + 256 + (uint32_t)(lev * 128) : t_psi + 64;
    const uint32_t t_phi_oth = t_lane + 256 + (uint32_t)((1 - lev) * 128);          // TMH only
    const uint32_t t_tw1 = TMH ? t_lane + 128 : t_psi + 128, t_tw2 = t_tw1 + 64;
    if (!TMH || lev == 0) {
        tmem_fill_twiddles(t_tw1, p.tw1[S1 + t]);        // w_N^t          (host-computed root, forward value)
        tmem_fill_twiddles(t_tw2, p.tw2[R3 + n3p]);      // w_{16 R3}^{n3'}
        tmem_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const cd *traj_in = p.traj_in ? p.traj_in + (long long)run * (p.nt_max + 1) * N : nullptr;
    // TMA staging of the stored trajectory (C::TMA_SLICE): one slice ahead of the time loop
    cd *s_slice = reinterpret_cast<cd *>(smem_raw + (C::TMA_SLICE ? C::SLICE_OFFSET : 0));
    uint64_t *s_mbar = reinterpret_cast<uint64_t *>(s_slice + N);
    const bool staged = C::TMA_SLICE && traj_in && !p.single_step && (p.field_mode == 1 || p.field_mode == 2);
    uint32_t slice_phase = 0;
    cd *traj_out = p.traj_out ? p.traj_out + (long long)run * (p.nt_max + 1) * N : nullptr;
    const long long erun = (long long)p.E_stride * run * (p.nt_max + 1);
    const long long xrun = (long long)p.expL_stride * run * (p.nt_max + 1);
    cd *E_out = p.E_out + (long long)run * (p.nt_max + 1);
    // which level records / reads trajectories: forward sweeps record level 0 and read the stored
    // level 1 of chi; backward sweeps record level 1 and read level 0 of psi (fitter.py:700-706, 791-796)
    const int ov_lev = p.field_mode == 2 ? 1 : 0;
    const int last_j = (p.orders ? p.orders[prun] : p.nch) - 2;
    if (staged) {
        if (tid == 0) mbar_init(s_mbar, 1);
        __syncthreads();
        if (tid == 0 && p.l_first <= nt) tma_load_1d(s_slice, traj_in + (long long)(nt - p.l_first) * N, (uint32_t)(N * sizeof(cd)), s_mbar);
    }

    for (int l = p.l_first; l < p.l_first + p.nsteps && l <= nt && (!PK || valid); ++l) {
        const bool start_only = (l == 0);   // start(): only the initial field is evaluated (propagation.py:360-361)
        // ---- local control: this step's multiplier, energy window and Newton table (qc_local.cuh) ----
        if (LOCAL && !start_only) {
            if (tid < 32)
                local_step_setup(s_lc, s_ls, s_dv, s_xp, p.lc_rt, p.lc_rdiag, p.nch,
                                 p.lc_t[(long long)run * (p.nt_max + 1) + l], l, tid);
            __syncthreads();
            if (s_ls[LS_STOP] != 0.0) break;             // the reference raises here (fitter.py:817-821)
            eL = s_ls[LS_EL]; emin = s_ls[LS_EMIN]; emax = s_ls[LS_EMAX]; t_sc = s_ls[LS_TSC];
            c1 = 4.0 / (emax - emin);
            c2 = 2.0 * (emax + emin) / (emax - emin);
            phase = cmk(s_ls[LS_PHR], s_ls[LS_PHI]);
            load_coefs();
        }
        // ---- rotating frame in (propagation.py:383-390) ---------------------
        cd eLx = cmk(1.0, 0.0);
        if (p.hf_hide && !start_only) {
            eLx = LOCAL ? cmk(s_ls[LS_XR], s_ls[LS_XI]) : p.expL[xrun + l];
            const double den = eLx.x * eLx.x + eLx.y * eLx.y, rden = 1.0 / den;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
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
        double Ei = 0.0;
        if (p.single_step) {
            E = p.E_step[run].x;
            if (p.cplx_coupling) Ei = lev == 0 ? p.E_step[run].y : -p.E_step[run].y;
        } else if (p.field_mode == 0 || LOCAL) {
            const int le = p.reversed_E ? (l == 0 ? nt : nt - l + 1) : l;
            E = p.E_base[erun + le].x;
        } else {
            double part = 0.0;
            if (lev == ov_lev) {
                const cd *tr = staged ? s_slice : traj_in + (long long)(nt - l) * N;
                if (staged) mbar_wait(s_mbar, slice_phase);      // the bulk copy of this step's slice has landed
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const cd c = tr[t + TPL * i];
                    // forward: Im(conj(psi) * chi) ; backward: Im(conj(psi_old) * chi)
                    part += p.field_mode == 1 ? (psi[i].x * c.y - psi[i].y * c.x) : (c.x * psi[i].y - c.y * psi[i].x);
                }
            }
            double t0s, t1s;
            both_sums(part, t0s, t1s);
            const double tot = t0s + t1s;
            E = -2.0 * (tot * dx) / h_lambda;
            if (staged) {
                // every thread is past the CTA-wide barriers of the sums, so the buffer has been read: fetch the slice
                // of the next step behind the 63 polynomial orders of this one
                slice_phase ^= 1u;
                const int ln = l + 1;
                if (tid == 0 && ln <= nt && ln < p.l_first + p.nsteps)
                    tma_load_1d(s_slice, traj_in + (long long)(nt - ln) * N, (uint32_t)(N * sizeof(cd)), s_mbar);
            }
        }
        if (writer && !p.single_step) {
            E_out[l] = cmk(E, LOCAL ? s_lc[LC_FM] : 0.0);      // local control: the multiplier rides in the imaginary slot
            if (p.write_E_base) p.E_base[erun + l] = cmk(E, 0.0);
        }
        if (start_only) continue;
        const double cE = c1 * E;
        const double cEi = c1 * Ei;
        const double cEk = cE / (double)N, cEik = cEi / (double)N;     // packed form: coupling applied before the inverse transform

        // ---- Newton recurrence (phys_base.py:158-172): phi = psi, psi <- dv0 * psi (parked in TMEM) -----
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            a[i] = psi[i];
            psi[i] = cmul(psi[i], s_dv[0]);
        }
#pragma unroll
        for (int c4 = 0; c4 < 16; c4 += 4) tmem_store_c4(t_psi + c4 * 4, psi + c4);
        for (int j = 0; j <= last_j; ++j) {
            const double xpj = s_xp[j];
            const cd dvj = s_dv[j + 1];
            // hand phi over to the other level (shared memory) once it has read the previous one, and keep
            // an own copy in TMEM for the pointwise phase
            const int par = j & 1;
            cd *stash = stash0 + par * N;
            const cd *stash_other = stash0_other + par * N;
            // cluster form, QC_CL_SPLIT_PUSH: only the first four of the sixteen remote stores leave here; the others
            // follow in three groups further down, re-read from the own copy in tensor memory, so that the slow
            // distributed-shared-memory path never holds a queue of sixteen stores per thread (their source registers
            // stay locked until the store has been taken: 12 % of the warp time in the long-scoreboard stall)
            constexpr bool SPLIT = CL && QC_CL_SPLIT_PUSH;
            if (!ILV && !TMH) {
#pragma unroll
                for (int n1 = 0; n1 < (SPLIT ? 4 : 16); ++n1) stash[t + TPL * n1] = a[n1];
            }
            auto push_group = [&](int g) {
                uint32_t rg[16];
                cd v4[4];
                tmem_ld16(t_phi + 16 * g, rg);
                tmem_wait_ld();
                tmem_unpack_c4(rg, v4);
#pragma unroll
                for (int i = 0; i < 4; ++i) stash[t + TPL * (4 * g + i)] = v4[i];
            };
            const uint32_t t_phi_j = TMH ? t_phi + (uint32_t)(par * 64) : t_phi;
#pragma unroll
            for (int c4 = 0; c4 < 16; c4 += 4) tmem_store_c4(t_phi_j + c4 * 4, a + c4);
            if (NAMED && !TMH) bar_arrive(BAR_FULL_OWN + par, HN);   // producer side of a named barrier (no fence needed)
            // cluster form: one split-phase cluster barrier per order is the whole protocol -- the partner waits for
            // it before reading, and since the buffers alternate, this level's wait of order j+1 already implies
            // that the partner has finished reading buffer `par` of order j.  The arrive has RELEASE semantics (SASS:
            // MEMBAR.ALL.GPU): it waits until the remote stores above have landed in the partner's shared memory, so it
            // is issued behind the forward transform, when they have (right after the stores it cost 20 % of the warp
            // time in the memory-barrier stall: np = 4096 0.39 -> 0.42 of the FP64 peak, profiles/r02z_*, r02ab_*); the
            // partner needs it an order later.
            if (CL && !QC_CL_LATE_ARRIVE) cl_arrive();
            // pass 1 : DFT over n1, twiddle w_N^(k1*t), transpose
            {
                TmemTwiddles tw(t_tw1);
                dft16_table_out<false>(a, tw, [&](int k1, cd val) { bufA[k1 * S1 + t] = val; });
            }
            if (CL && QC_CL_LATE_ARRIVE == 1) cl_arrive();
            if (SPLIT) {
                tmem_wait_st();
                push_group(1);
            }
            if (TMH && (!RDV || RDVM)) {
                // hand-over through tensor memory: the asynchronous stores of phi have long landed behind the
                // butterflies of pass 1; make them visible to the partner warp and signal "phi of this parity is there"
                tmem_wait_st();
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (RDVM) mbar_arrive(s_hb + 2 * lev + par);
                else bar_arrive(BAR_FULL_OWN + par, HN);
            }
            level_sync();                                                           // (#1)
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) a[n2] = bufA[k1p * S1 + n2 * R3 + n3p];
            // pass 2 : DFT over n2, twiddle w_{16 R3}^(k2*n3), transpose inside the R3-lane group
            if (R3 == 1) dft16<false>(a);
            if (R3 > 1) {
                __syncwarp();
                {
                    TmemTwiddles tw(t_tw2);
                    dft16_table_out<false>(a, tw, [&](int k2, cd val) { bufA[k1p * S1 + pos2<R3, Q>(k2, n3p)] = val; });
                }
                if (SPLIT) push_group(2);
                __syncwarp();
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    a[e] = bufA[k1p * S1 + pos2<R3, Q>(n3p * Q + e / R3, e % R3)];
                // pass 3 : Q DFTs of size R3 over n3
#pragma unroll
                for (int q = 0; q < Q; ++q) dft<R3, false>(a + q * R3);
            }
            if (SPLIT && R3 > 1) push_group(3);
            if (CL && QC_CL_LATE_ARRIVE == 2 && !SPLIT) cl_arrive();
            // k-space: kinetic multiplier (with c1 and the 1/N of the inverse transform folded in)
            if (!ILV) {
#pragma unroll
                for (int e = 0; e < 16; ++e) a[e] = cscale(a[e], kc[e]);
            } else {
                // ... and the coupling -E * phi_other (hamil_2d.py:70-71): the partner lane holds the other level at
                // the same frequencies; the transform is linear, so the term may join before the inverse passes
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    cd o, r;
                    o.x = __shfl_xor_sync(0xffffffffu, a[e].x, 16);
                    o.y = __shfl_xor_sync(0xffffffffu, a[e].y, 16);
                    r.x = fma(kc[e], a[e].x, -cEk * o.x);
                    r.y = fma(kc[e], a[e].y, -cEk * o.y);
                    if (p.cplx_coupling) {           // orig=True form of hamil_2d.py:40-47
                        r.x = fma(cEik, o.y, r.x);
                        r.y = fma(-cEik, o.x, r.y);
                    }
                    a[e] = r;
                }
            }
            if (R3 > 1) {
#pragma unroll
                for (int q = 0; q < Q; ++q) dft<R3, true>(a + q * R3);
                __syncwarp();                 // the group's forward reads of this row are done
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    bufB[k1p * S1 + pos2<R3, Q>(n3p * Q + e / R3, e % R3)] = a[e];
                __syncwarp();
                {
                    TmemTwiddles tw(t_tw2);       // inverse pass 2 with its conjugate twiddles on the load side
                    dft16_table_in_inv(a, tw, [&](int k2) { return bufB[k1p * S1 + pos2<R3, Q>(k2, n3p)]; });
                }
                __syncwarp();
            }
            if (R3 == 1) dft16<true>(a);
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) bufB[k1p * S1 + n2 * R3 + n3p] = a[n2];
            if (SPLIT) cl_arrive();       // every group has long landed: the release fence finds nothing to wait for
            level_sync();                                                           // (#2)
            // inverse pass 1 -> kinetic term (already scaled by c1) at x = t + TPL*n1
            {
                TmemTwiddles tw(t_tw1);
                dft16_table_in_inv(a, tw, [&](int k1) { return bufB[k1 * S1 + t]; });
            }
            // pointwise part (hamil_2d.py:60-71, phys_base.py:98-111) and accumulation (phys_base.py:170-172),
            // in chunks of four grid points: psi and own phi from TMEM, the other level's phi from shared memory
            if (NAMED && !RDV && !RDVM) bar_sync(BAR_FULL_OTH + par, HN);
            if (RDVM) {
                if (par == 0) mbar_wait(s_hb + 2 * (1 - lev), hb_use0++ & 1u);
                else mbar_wait(s_hb + 2 * (1 - lev) + 1, hb_use1++ & 1u);
            } else if (RDV) {
                tmem_wait_st();
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                bar_sync(bar_rdv, HN);
            }
            if (TMH) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (CL) cl_wait();
            tmem_wait_st();
#pragma unroll
            for (int c4 = 0; c4 < 16; c4 += 4) {
                uint32_t rp[16], rs[16], ro[16];
                tmem_ld16(t_psi + c4 * 4, rp);
                tmem_ld16(t_phi_j + c4 * 4, rs);
                if (TMH) tmem_ld16(t_phi_oth + (uint32_t)(par * 64) + c4 * 4, ro);
                cd oth4[4], own4[4], acc4[4];
                if (!ILV && !TMH) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) oth4[i] = stash_other[t + TPL * (c4 + i)];
                }
                tmem_wait_ld();
                tmem_unpack_c4(rp, acc4);
                tmem_unpack_c4(rs, own4);
                if (TMH) tmem_unpack_c4(ro, oth4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n1 = c4 + i;
                    const double dd = dcoef[n1] - xpj;
                    cd r;
                    if (ILV) {                       // the coupling term is already inside a[]
                        r.x = fma(dd, own4[i].x, a[n1].x);
                        r.y = fma(dd, own4[i].y, a[n1].y);
                    } else {
                        r.x = fma(dd, own4[i].x, fma(-cE, oth4[i].x, a[n1].x));
                        r.y = fma(dd, own4[i].y, fma(-cE, oth4[i].y, a[n1].y));
                        if (p.cplx_coupling) {       // orig=True form of hamil_2d.py:40-47 (uniform branch, off in sweeps)
                            r.x = fma(cEi, oth4[i].y, r.x);
                            r.y = fma(-cEi, oth4[i].x, r.y);
                        }
                    }
                    a[n1] = r;
                    acc4[i] = cfma(dvj, r, acc4[i]);
                }
                tmem_store_c4(t_psi + c4 * 4, acc4);
            }
            if (!NAMED && !CL && !PK) {
                __syncthreads();                                                   // (#3) hand-over buffer reuse
            }
        }
        tmem_wait_st();
#pragma unroll
        for (int c4 = 0; c4 < 16; c4 += 4) {
            uint32_t rp[16];
            tmem_ld16(t_psi + c4 * 4, rp);
            tmem_wait_ld();
            tmem_unpack_c4(rp, psi + c4);
        }
        // ---- final phase (phys_base.py:174-176), norms, renormalisation --------
        double nrm = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            psi[i] = cmul(psi[i], phase);
            nrm = fma(psi[i].x, psi[i].x, fma(psi[i].y, psi[i].y, nrm));
        }
        if (!p.single_step) {
            double n0, n1s;
            both_sums(nrm, n0, n1s);
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
                for (int i = 0; i < 16; ++i) psi[i] = cmk(div_rn_by(psi[i].x, s, rs), div_rn_by(psi[i].y, s, rs));
            }
            // ---- record the rotating-frame state (fitter.py:301-309, 322-329) -----
            if (traj_out && lev == ov_lev) {
                cd *dst = traj_out + (long long)l * N;
#pragma unroll
                for (int i = 0; i < 16; ++i) dst[t + TPL * i] = psi[i];
            }
            // ---- local control: <psi_e|psi_g>, <psi_e|v_g - v_e|psi_g> and the feedback (propagation.py:435-437,
            //      fitter.py:598-623); the two levels meet through the hand-over buffer -----------------------
            if (LOCAL) {
#pragma unroll
                for (int n1 = 0; n1 < 16; ++n1) stash0[t + TPL * n1] = psi[n1];
                __syncthreads();
                double q[4] = {0.0, 0.0, 0.0, 0.0};
                if (lev == 0) {
                    const double *v0 = p.v + p.v_stride * run, *v1 = v0 + N;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const cd u = stash0_other[t + TPL * i];
                        const double qr = psi[i].x * u.x + psi[i].y * u.y, qi = psi[i].x * u.y - psi[i].y * u.x;
                        const double w = v0[t + TPL * i] - v1[t + TPL * i];
                        q[0] += qr; q[1] += qi; q[2] += qr * w; q[3] += qi * w;
                    }
                }
                double tot[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    double a0, a1;
                    level_sums<C::NTHREADS, TPL>(q[c], s_red, tid, a0, a1);
                    tot[c] = a0 + a1;
                }
                if (tid == 0) {
                    local_feedback(s_lc, cmk(tot[0] * dx, tot[1] * dx), cmk(tot[2] * dx, tot[3] * dx));
                }
            }
            // ---- rotating frame out (propagation.py:440-441) ---------------------
            if (p.hf_hide) {
                const double den = eLx.x * eLx.x + eLx.y * eLx.y, rden = 1.0 / den;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (lev == 0) {
                        psi[i] = cmul(psi[i], eLx);
                    } else {
                        cd n = cmulc(psi[i], eLx);
                        psi[i] = cmk(div_rn_by(n.x, den, rden), div_rn_by(n.y, den, rden));
                    }
                }
            }
        } else if (!PK) {
            __syncthreads();
        } else if (PKL) {
            bar_sync(bar_run, C::RT);
        } else {
            level_sync();
        }
    }
    if (!PK || valid) {
        cd *dst = p.psi + ((long long)run * 2 + lev) * N;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) dst[t + TPL * n1] = psi[n1];
    }
    if (LOCAL) {
        __syncthreads();
        if (tid < LC_NSTATE) p.lc[(long long)run * LC_STRIDE + tid] = s_lc[tid];
    }
    if (CL) {                 // the partner may still be reading this CTA's hand-over buffer
        cl_arrive();
        cl_wait();
    }
    tmem_wait_st();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TM_COLS) : "memory");
}

}  // namespace qc
