// C ABI of qcontrol_b200 (see include/qcontrol_b200.h).  Host-side plumbing only:
// device buffers, uploads, kernel launches.  The numerics live in qc_grid.cuh
// (FFT grid, kernel A) and qc_nlevel.cuh (N-level systems, kernel B).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <iterator>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/qcontrol_b200.h"
#include "qc_grid.cuh"
#include "qc_grid8.cuh"
#include "qc_grid8x2.cuh"
#include "qc_grid2d.cuh"
#include "qc_nlevel.cuh"

using qc::cd;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

static int32_t fail(int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define QC_CUDA(call)                                                                       \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess)                                                             \
            return fail(QC_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                \
    } while (0)

// ---------------------------------------------------------------------------
// objects
// ---------------------------------------------------------------------------
struct qc_context {
    int device;
    cudaStream_t stream;
    bool owns_stream;
    cudaEvent_t ev0, ev1;
    cudaStream_t aux[2];     // extra streams for the chunked host pipeline of qc_propagate_host (lazy)
    cudaEvent_t aux_ev[3];
    bool aux_ready;
    long long launches;
    int sm_count;
    int live_plans;      // plans that still point at this context
    bool released;       // qc_destroy was called while plans were alive: freed with the last plan
};

static void pool_context_gone(int dev);
static void context_free(qc_context *ctx) {
    cudaSetDevice(ctx->device);
    pool_context_gone(ctx->device);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->aux_ready) {
        for (auto &a : ctx->aux) cudaStreamDestroy(a);
        for (auto &e : ctx->aux_ev) cudaEventDestroy(e);
    }
    delete ctx;
}

// ---- device memory pool ------------------------------------------------------------------------------------------
// cudaFree waits for the whole device and cudaMalloc of a multi-GB trajectory buffer costs tens of milliseconds per GB;
// a batch entry point that sets the next sub-batch up on a second thread while the GPU propagates the current one
// (newcheb.run_variants) would stall on both.  Released buffers are therefore kept per device and handed out again to
// requests of exactly the same size (the sub-batches of a shape group are alike); everything cached is given back to
// the driver when an allocation fails, when qc_mem_info is asked for the free memory, and when the last context of the
// device is destroyed.  Reused memory is as uninitialised as fresh memory: no caller may expect zeros.
// An ARENA on top of that (qc_pool_reserve): one block reserved while the device is idle, from which every later
// request is cut (first fit, coalescing free list).  A fresh cudaMalloc beside a running sweep -- every SM held by one
// CTA for the whole sweep -- was measured to return only when the sweep ends (0.3-0.9 s per plan set-up).
struct Arena {
    char *base = nullptr;
    size_t size = 0, live = 0;
    bool released = false;
    std::map<size_t, size_t> holes;            // offset -> length
    void *cut(size_t bytes) {
        bytes = (bytes + 511) / 512 * 512;
        for (auto it = holes.begin(); it != holes.end(); ++it) {
            if (it->second < bytes) continue;
            const size_t off = it->first, len = it->second;
            holes.erase(it);
            if (len > bytes) holes.emplace(off + bytes, len - bytes);
            live += bytes;
            return base + off;
        }
        return nullptr;
    }
    bool owns(const void *q) const { return base && (const char *)q >= base && (const char *)q < base + size; }
    void put(void *q, size_t bytes) {
        bytes = (bytes + 511) / 512 * 512;
        size_t off = (size_t)((char *)q - base);
        live -= bytes;
        auto nxt = holes.lower_bound(off);
        if (nxt != holes.begin()) {
            auto prv = std::prev(nxt);
            if (prv->first + prv->second == off) {
                off = prv->first;
                bytes += prv->second;
                holes.erase(prv);
            }
        }
        if (nxt != holes.end() && off + bytes == nxt->first) {
            bytes += nxt->second;
            holes.erase(nxt);
        }
        holes.emplace(off, bytes);
    }
};

struct DevPool {
    std::mutex mu;
    std::multimap<size_t, void *> blocks[64];
    Arena arena[64];
    int contexts[64] = {0};
    // smallest cached block of at least `bytes` and at most four times that (the last sub-batch of a group is
    // shorter than the others and may reuse their blocks; a small request must not swallow a trajectory buffer)
    void *take(int dev, size_t bytes, size_t *cap) {
        std::lock_guard<std::mutex> g(mu);
        Arena &a = arena[dev & 63];
        if (a.base && !a.released) {
            if (void *q = a.cut(bytes)) {
                *cap = bytes;
                return q;
            }
        }
        auto &m = blocks[dev & 63];
        auto it = m.lower_bound(bytes);
        if (it == m.end() || it->first > 4 * bytes) return nullptr;
        void *q = it->second;
        *cap = it->first;
        m.erase(it);
        return q;
    }
    void give(int dev, void *q, size_t bytes) {
        void *dead = nullptr;
        {
            std::lock_guard<std::mutex> g(mu);
            Arena &a = arena[dev & 63];
            if (a.owns(q)) {
                a.put(q, bytes);
                if (a.released && a.live == 0) {       // the last block of a released arena came back
                    dead = a.base;
                    a = Arena();
                }
            } else {
                blocks[dev & 63].emplace(bytes, q);
            }
        }
        if (dead) cudaFree(dead);
    }
    cudaError_t reserve(int dev, size_t bytes) {
        std::lock_guard<std::mutex> g(mu);
        Arena &a = arena[dev & 63];
        if (a.base) return cudaSuccess;                // one arena per device at a time
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) return e;
        a = Arena();
        a.base = (char *)q;
        a.size = bytes;
        a.holes.emplace(0, bytes);
        return cudaSuccess;
    }
    void unreserve(int dev) {
        void *dead = nullptr;
        {
            std::lock_guard<std::mutex> g(mu);
            Arena &a = arena[dev & 63];
            if (!a.base) return;
            if (a.live == 0) {
                dead = a.base;
                a = Arena();
            } else {
                a.released = true;                     // freed when its last block is returned
            }
        }
        if (dead) cudaFree(dead);
    }
    void drain(int dev) {
        std::multimap<size_t, void *> m;
        {
            std::lock_guard<std::mutex> g(mu);
            m.swap(blocks[dev & 63]);
        }
        for (auto &kv : m) cudaFree(kv.second);
    }
};
static DevPool g_pool;

static void pool_context_gone(int dev) {
    bool last;
    {
        std::lock_guard<std::mutex> g(g_pool.mu);
        last = --g_pool.contexts[dev & 63] <= 0;
    }
    if (last) {
        g_pool.drain(dev);
        g_pool.unreserve(dev);
    }
}

static cudaError_t pool_alloc(void **out, size_t bytes, size_t *cap) {
    int dev = 0;
    cudaGetDevice(&dev);
    if ((*out = g_pool.take(dev, bytes, cap)) != nullptr) return cudaSuccess;
    *cap = bytes;
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        g_pool.drain(dev);
        e = cudaMalloc(out, bytes);
    }
    return e;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    size_t cap = 0;                            // bytes of the block behind p (>= n * sizeof(T))
    int dev = 0;
    cudaError_t alloc(size_t count) {
        if (p) cudaDeviceSynchronize();        // re-allocation: work in flight may still read the old block
        release();
        n = count;
        if (!count) return cudaSuccess;
        cudaGetDevice(&dev);
        return pool_alloc(reinterpret_cast<void **>(&p), count * sizeof(T), &cap);
    }
    void release() {                           // the caller guarantees that no work in flight uses the block
        if (p) g_pool.give(dev, p, cap);
        p = nullptr;
        n = 0;
    }
};

struct PolySlot {
    DevBuf<double> xp, sc;
    DevBuf<cd> dv;
    DevBuf<int> orders;      // optional: polynomial terms actually summed per run (qc_set_poly_orders), else nch
    int batched = 0;
    bool set = false, truncated = false;
};

struct qc_plan {
    qc_context *ctx;
    qc_plan_desc d;
    bool grid;
    int log2n;
    int nbp;                 // nb padded to a power of two (N-level thread groups)
    long long nthreads;      // batch * nbp
    size_t state_elems;      // batch*nb*nlevs*np
    size_t traj_elems;
    DevBuf<cd> psi, E_base, expL[2], s_env, E_out, a0, E_step, tw1, tw2, tw8, goal, gc, tmp_row;
    DevBuf<cd> traj[2];
    DevBuf<cd> snap[4];
    DevBuf<double> v, kin, uwd, dx, ctl, norms;
    int v_batched = 0, uwd_batched = 0, dx_batched = 0, expL_batched[2] = {0, 0}, s_batched = 0, ctl_batched = 0;
    bool static_set = false, ctl_set = false;
    PolySlot poly[3];                 // slot 2: internal identity polynomial (one bare H application)
    DevBuf<cd> work[4];               // psi backup, H psi, T psi, P psi (instrumentation)
    DevBuf<double> zero_v, akx, xgrid, instr;
    DevBuf<double> lc, lc_t, lc_rt, lc_rdiag;     // local control (qc_set_local_control)
    bool lc_set = false;
};

// ---------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------
__global__ void broadcast_rows_kernel(cd *dst, const cd *src, long long row, long long rows) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < row * rows) dst[i] = src[i % row];
}

// gc[run][v] = sum_n dx * sum_x conj(psi[x]) * goal[x]      (cprod(goal, psi), math_base.py:21)
__global__ void goal_overlap_kernel(const cd *psi, const cd *goal, long long goal_stride, const double *dx,
                                    int dx_stride, int per_vec, cd *gc, int nb) {
    const int item = blockIdx.x;          // run*nb + v
    const int run = item / nb, v = item % nb;
    const cd *a = psi + (long long)item * per_vec;
    const cd *g = goal + goal_stride * run + (long long)v * per_vec;
    double re = 0.0, im = 0.0;
    for (int i = threadIdx.x; i < per_vec; i += blockDim.x) {
        const cd p = a[i], q = g[i];
        re += p.x * q.x + p.y * q.y;
        im += p.x * q.y - p.y * q.x;
    }
    __shared__ double sr[32], si[32];
    for (int o = 16; o > 0; o >>= 1) {
        re += __shfl_xor_sync(0xffffffffu, re, o);
        im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if ((threadIdx.x & 31) == 0) {
        sr[threadIdx.x >> 5] = re;
        si[threadIdx.x >> 5] = im;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tr = 0.0, ti = 0.0;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) {
            tr += sr[w];
            ti += si[w];
        }
        const double h = dx[(long long)dx_stride * run];
        gc[item] = qc::cmk(tr * h, ti * h);
    }
}

// Krotov terminal condition, fitter.py:376-392 (grid plans, nb == 1, two levels)
__global__ void krotov_terminal_kernel(cd *psi, const cd *goal, long long goal_stride, const cd *gc,
                                       const cd *sqrt_hf, int hf_stride, const double *dx, int dx_stride,
                                       cd *traj0, long long traj_run_stride, int np) {
    const int run = blockIdx.x;
    cd *st = psi + (long long)run * 2 * np;
    const cd *gu = goal + goal_stride * run + np;        // goal level 1
    const cd g = gc[run];
    double acc = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const cd c = qc::cmul(g, gu[i]);
        acc += c.x * c.x + c.y * c.y;
    }
    __shared__ double sr[32];
    __shared__ double s_norm;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sr[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) t += sr[w];
        s_norm = fabs(t * dx[(long long)dx_stride * run]);
    }
    __syncthreads();
    const double nrm = s_norm;
    const double s = nrm > 0.0 ? sqrt(nrm) : 1.0;
    const cd hf = sqrt_hf[(long long)hf_stride * run];
    cd *tr = traj0 ? traj0 + traj_run_stride * run : nullptr;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        cd c = qc::cmul(g, gu[i]);
        c = qc::cmk(c.x / s, c.y / s);
        st[i] = qc::cmk(0.0, 0.0);
        st[np + i] = c;
        if (tr) tr[i] = qc::cmul(c, hf);
    }
}

// copy the current state into a trajectory slice
__global__ void record_grid_kernel(const cd *psi, cd *traj, long long traj_run_stride, long long slice_off,
                                   int level, int np) {
    const int run = blockIdx.x;
    const cd *src = psi + ((long long)run * 2 + level) * np;
    cd *dst = traj + traj_run_stride * run + slice_off;
    for (int i = threadIdx.x; i < np; i += blockDim.x) dst[i] = src[i];
}

__global__ void record_nlevel_kernel(const cd *psi, cd *traj_slice, int nb, int nbp, int nl, long long nthreads) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gt >= nthreads) return;
    const int run = (int)(gt / nbp), v = (int)(gt % nbp);
    for (int n = 0; n < nl; ++n)
        traj_slice[(long long)n * nthreads + gt] = v < nb ? psi[((long long)run * nb + v) * nl + n] : qc::cmk(0.0, 0.0);
}

// out[run][x] = traj[run][index][x]: one recorded slice of a grid trajectory, contiguous for a single copy to the
// host (the run stride (nt_max+1)*np*16 B of the buffer exceeds the 2 GiB pitch limit of a strided copy at the
// shipped sizes: nt = 230000, np = 1024 -> 3.77 GB)
__global__ void gather_grid_slice_kernel(const cd *traj, cd *out, long long run_stride, int np_) {
    const cd *src = traj + run_stride * blockIdx.x;
    cd *dst = out + (long long)np_ * blockIdx.x;
    for (int i = threadIdx.x; i < np_; i += blockDim.x) dst[i] = src[i];
}

__global__ void gather_nlevel_kernel(const cd *traj_slice, cd *out, int nb, int nbp, int nl, long long nthreads) {
    const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gt >= nthreads) return;
    const int run = (int)(gt / nbp), v = (int)(gt % nbp);
    if (v >= nb) return;
    for (int n = 0; n < nl; ++n) out[((long long)run * nb + v) * nl + n] = traj_slice[(long long)n * nthreads + gt];
}


// Per (run, vector, level) reductions of PropagationSolver.step's instrumentation block
// (propagation.py:462-470, 260-299; phys_base.py:196-288).  out[item][20]:
//  0,1 cnorm   2,3 cener   4,5 overlp0   6,7 overlpf   8,9 <x>   10,11 <x^2>   12,13 <p>   14,15 <p^2>
//  16 max|psi|   17 Re psi at argmax|Re psi|   18,19 cprod(psi_0, psi_1) (sigma moments, written for level 0)
__global__ void instrument_kernel(const cd *psi, const cd *hpsi, const cd *tpsi, const cd *ppsi, const cd *psi0,
                                  const cd *psif, const double *x, const double *dx, int dx_stride, double two_m,
                                  int nb, int nl, int np, double *out) {
    const int item = blockIdx.x;                 // (run*nb + v)*nl + n
    const int n = item % nl, run = item / (nl * nb);
    const long long off = (long long)item * np;
    const cd *a = psi + off;
    double acc[16];
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    double mabs = 0.0, mre_abs = -1.0, mre = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) {
        const cd p = a[i];
        const double w = p.x * p.x + p.y * p.y;
        acc[0] += w;
        if (hpsi) { const cd h = hpsi[off + i]; acc[2] += h.x * p.x + h.y * p.y; acc[3] += h.x * p.y - h.y * p.x; }
        if (psi0) { const cd g = psi0[off + i]; acc[4] += p.x * g.x + p.y * g.y; acc[5] += p.x * g.y - p.y * g.x; }
        if (psif) { const cd g = psif[off + i]; acc[6] += p.x * g.x + p.y * g.y; acc[7] += p.x * g.y - p.y * g.x; }
        const double xi = x ? x[i] : 0.0;
        acc[8] += w * xi;
        acc[10] += w * xi * xi;
        if (ppsi) { const cd h = ppsi[off + i]; acc[12] += h.x * p.x + h.y * p.y; acc[13] += h.x * p.y - h.y * p.x; }
        if (tpsi) { const cd h = qc::cmk(tpsi[off + i].x * two_m, tpsi[off + i].y * two_m);
                    acc[14] += h.x * p.x + h.y * p.y; acc[15] += h.x * p.y - h.y * p.x; }
        if (n == 0 && nl == 2) { const cd q = a[np + i]; acc[1] += q.x * p.x + q.y * p.y; acc[9] += q.x * p.y - q.y * p.x; }
        const double ab = sqrt(w);
        if (ab > mabs) mabs = ab;
        if (fabs(p.x) > mre_abs) { mre_abs = fabs(p.x); mre = p.x; }
    }
    __shared__ double red[16][32];
    __shared__ double rmax[3][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) / 32;
    for (int k = 0; k < 16; ++k) {
        double v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][wid] = v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double oa = __shfl_xor_sync(0xffffffffu, mabs, o);
        const double ora = __shfl_xor_sync(0xffffffffu, mre_abs, o);
        const double orv = __shfl_xor_sync(0xffffffffu, mre, o);
        if (oa > mabs) mabs = oa;
        if (ora > mre_abs) { mre_abs = ora; mre = orv; }
    }
    if (lane == 0) { rmax[0][wid] = mabs; rmax[1][wid] = mre_abs; rmax[2][wid] = mre; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double h = dx[(long long)dx_stride * run];
        double t[16];
        for (int k = 0; k < 16; ++k) { t[k] = 0.0; for (int w = 0; w < nw; ++w) t[k] += red[k][w]; }
        double ma = 0.0, mra = -1.0, mr = 0.0;
        for (int w = 0; w < nw; ++w) {
            if (rmax[0][w] > ma) ma = rmax[0][w];
            if (rmax[1][w] > mra) { mra = rmax[1][w]; mr = rmax[2][w]; }
        }
        double *o = out + (long long)item * 20;
        o[0] = t[0] * h;  o[1] = 0.0;
        o[2] = t[2] * h;  o[3] = t[3] * h;
        o[4] = t[4] * h;  o[5] = t[5] * h;
        o[6] = t[6] * h;  o[7] = t[7] * h;
        o[8] = t[8] * h;  o[9] = 0.0;
        o[10] = t[10] * h; o[11] = 0.0;
        o[12] = t[12] * h; o[13] = t[13] * h;
        o[14] = t[14] * h; o[15] = t[15] * h;
        o[16] = ma; o[17] = mr;
        o[18] = t[1] * h; o[19] = t[9] * h;
    }
}

// out[run] = sum_{l=1..nt_run} |E_out[run][l]|^2   (E_int without the dt factor, fitter.py:338-341)
__global__ void field_integral_kernel(const cd *E_out, const double *ctl, int ctl_stride, int nt_max, double *out) {
    const int run = blockIdx.x;
    const int nt = (int)ctl[(long long)ctl_stride * run * 4 + 1];
    const cd *row = E_out + (long long)run * (nt_max + 1);
    double acc = 0.0;
    for (int l = 1 + threadIdx.x; l <= nt; l += blockDim.x) acc += row[l].x * row[l].x + row[l].y * row[l].y;
    __shared__ double sr[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sr[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) t += sr[w];
        out[run] = t;
    }
}

// FP64 FMA throughput: 8 independent chains per thread
__global__ void fp64_peak_kernel(double *out, int iters, double seed) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c);
            a1 = fma(a1, m, c);
            a2 = fma(a2, m, c);
            a3 = fma(a3, m, c);
            a4 = fma(a4, m, c);
            a5 = fma(a5, m, c);
            a6 = fma(a6, m, c);
            a7 = fma(a7, m, c);
        }
    }
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------
static bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

static int ilog2(int x) {
    int r = 0;
    while ((1 << r) < x) ++r;
    return r;
}

static void fill_twiddles(int log2n, std::vector<cd> &t1, std::vector<cd> &t2) {
    const int N = 1 << log2n, TPL = N / 16, R3 = N / 256;
    const int S1 = TPL + (R3 < 8 ? R3 : 0);
    const long double two_pi = 6.283185307179586476925286766559005768L;
    t1.assign((size_t)16 * S1, qc::cmk(1.0, 0.0));
    for (int k1 = 0; k1 < 16; ++k1)
        for (int t = 0; t < TPL; ++t) {
            const int m = (k1 * t) % N;
            const long double ang = -two_pi * (long double)m / (long double)N;
            t1[(size_t)k1 * S1 + t] = qc::cmk((double)cosl(ang), (double)sinl(ang));
        }
    const int M = 16 * R3;
    t2.assign((size_t)16 * R3, qc::cmk(1.0, 0.0));
    for (int k2 = 0; k2 < 16; ++k2)
        for (int n3 = 0; n3 < R3; ++n3) {
            const int m = (k2 * n3) % M;
            const long double ang = -two_pi * (long double)m / (long double)M;
            t2[(size_t)k2 * R3 + n3] = qc::cmk((double)cosl(ang), (double)sinl(ang));
        }
}

// tables of the 512-thread form at np = 2048 (qc_grid8.cuh): forward values exp(-2 pi i m / M)
static void fill_twiddles8(std::vector<cd> &t) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    auto root = [&](int m, int M) {
        const long double ang = -two_pi * (long double)(m % M) / (long double)M;
        return qc::cmk((double)cosl(ang), (double)sinl(ang));
    };
    t.clear();
    for (int k1 = 0; k1 < 8; ++k1)
        for (int T = 0; T < 256; ++T) t.push_back(root(k1 * T, 2048));
    for (int k2 = 0; k2 < 8; ++k2)
        for (int m = 0; m < 32; ++m) t.push_back(root(k2 * m, 256));
    for (int k3 = 0; k3 < 8; ++k3)
        for (int n = 0; n < 4; ++n) t.push_back(root(k3 * n, 32));
}

// ... and of its np = 1024 sibling (qc_grid8x2.cuh)
static void fill_twiddles8x2(std::vector<cd> &t) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    auto root = [&](int m, int M) {
        const long double ang = -two_pi * (long double)(m % M) / (long double)M;
        return qc::cmk((double)cosl(ang), (double)sinl(ang));
    };
    t.clear();
    for (int k1 = 0; k1 < 8; ++k1)
        for (int T = 0; T < 128; ++T) t.push_back(root(k1 * T, 1024));
    for (int k2 = 0; k2 < 8; ++k2)
        for (int m = 0; m < 16; ++m) t.push_back(root(k2 * m, 128));
    for (int k3 = 0; k3 < 4; ++k3)
        for (int n = 0; n < 4; ++n) t.push_back(root(k3 * n, 16));
}

template <typename T>
static int32_t upload(qc_plan *pl, DevBuf<T> &buf, const void *host, size_t count) {
    if (buf.n != count) QC_CUDA(buf.alloc(count));
    QC_CUDA(cudaMemcpyAsync(buf.p, host, count * sizeof(T), cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

// upload rows[ (batched ? batch : 1) ][row] into a buffer that is always [batch][row]
static int32_t upload_rows_expanded(qc_plan *pl, DevBuf<cd> &buf, const double *host, int batched, size_t row) {
    const size_t batch = pl->d.batch;
    if (buf.n != batch * row) QC_CUDA(buf.alloc(batch * row));
    cudaStream_t st = pl->ctx->stream;
    if (batched) {
        QC_CUDA(cudaMemcpyAsync(buf.p, host, batch * row * sizeof(cd), cudaMemcpyHostToDevice, st));
    } else if (batch * row <= (size_t)1 << 26) {
        // up to 1 GiB: replicate on the host and copy (no kernel: see the note on zeros in qc_plan_create)
        std::vector<cd> full(batch * row);
        const cd *src = reinterpret_cast<const cd *>(host);
        for (size_t b = 0; b < batch; ++b) memcpy(full.data() + b * row, src, row * sizeof(cd));
        QC_CUDA(cudaMemcpyAsync(buf.p, full.data(), batch * row * sizeof(cd), cudaMemcpyHostToDevice, st));
        QC_CUDA(cudaStreamSynchronize(st));
        return QC_OK;
    } else {
        if (pl->tmp_row.n < row) QC_CUDA(pl->tmp_row.alloc(row));
        QC_CUDA(cudaMemcpyAsync(pl->tmp_row.p, host, row * sizeof(cd), cudaMemcpyHostToDevice, st));
        const long long total = (long long)batch * row;
        broadcast_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(buf.p, pl->tmp_row.p, (long long)row,
                                                                              (long long)batch);
        pl->ctx->launches++;
        QC_CUDA(cudaGetLastError());
    }
    QC_CUDA(cudaStreamSynchronize(st));
    return QC_OK;
}

// ---------------------------------------------------------------------------
// library / context
// ---------------------------------------------------------------------------
extern "C" int32_t qc_abi_version(void) { return QC_ABI_VERSION; }
extern "C" const char *qc_last_error(void) { return g_err.c_str(); }

extern "C" int32_t qc_create(int32_t device, void *stream, qc_context **out) {
    if (!out) return fail(QC_ERR_ARG, "qc_create: out is NULL");
    int ndev = 0;
    QC_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(QC_ERR_ARG, "qc_create: device %d of %d", device, ndev);
    QC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    QC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(QC_ERR_ARG, "qc_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    qc_context *c = new qc_context();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->launches = 0;
    c->live_plans = 0;
    c->aux_ready = false;
    c->released = false;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->owns_stream = false;
    } else {
        QC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->owns_stream = true;
    }
    QC_CUDA(cudaEventCreate(&c->ev0));
    QC_CUDA(cudaEventCreate(&c->ev1));
    {
        std::lock_guard<std::mutex> g(g_pool.mu);
        g_pool.contexts[device & 63]++;
    }
    *out = c;
    return QC_OK;
}

extern "C" int32_t qc_destroy(qc_context *ctx) {
    if (!ctx) return QC_OK;
    if (ctx->live_plans > 0) {       // plans outlive the handle: defer, never leave them a dangling pointer
        ctx->released = true;
        return QC_OK;
    }
    context_free(ctx);
    return QC_OK;
}

extern "C" int32_t qc_synchronize(qc_context *ctx) {
    if (!ctx) return fail(QC_ERR_ARG, "qc_synchronize: NULL context");
    QC_CUDA(cudaSetDevice(ctx->device));
    QC_CUDA(cudaStreamSynchronize(ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_timer_start(qc_context *ctx) {
    if (!ctx) return fail(QC_ERR_ARG, "qc_timer_start: NULL context");
    QC_CUDA(cudaSetDevice(ctx->device));
    QC_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_timer_stop(qc_context *ctx, double *elapsed_ms) {
    if (!ctx || !elapsed_ms) return fail(QC_ERR_ARG, "qc_timer_stop: NULL argument");
    QC_CUDA(cudaSetDevice(ctx->device));
    QC_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    QC_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    QC_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *elapsed_ms = ms;
    return QC_OK;
}

extern "C" int64_t qc_launch_count(qc_context *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int32_t qc_mem_info(qc_context *ctx, int64_t *free_bytes, int64_t *total_bytes) {
    if (!ctx || !free_bytes || !total_bytes) return fail(QC_ERR_ARG, "qc_mem_info: NULL argument");
    QC_CUDA(cudaSetDevice(ctx->device));
    g_pool.drain(ctx->device);                 // cached blocks are free memory
    size_t f = 0, t = 0;
    QC_CUDA(cudaMemGetInfo(&f, &t));
    *free_bytes = (int64_t)f;
    *total_bytes = (int64_t)t;
    return QC_OK;
}

extern "C" int32_t qc_pool_reserve(qc_context *ctx, int64_t bytes) {
    if (!ctx || bytes < 0) return fail(QC_ERR_ARG, "qc_pool_reserve: bad argument");
    QC_CUDA(cudaSetDevice(ctx->device));
    if (bytes == 0) {
        g_pool.unreserve(ctx->device);
        return QC_OK;
    }
    g_pool.drain(ctx->device);
    const cudaError_t e = g_pool.reserve(ctx->device, (size_t)bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(QC_ERR_NOMEM, "qc_pool_reserve: %lld bytes: %s", (long long)bytes, cudaGetErrorString(e));
    }
    return QC_OK;
}

extern "C" int32_t qc_sm_count(qc_context *ctx, int32_t *sm_count) {
    if (!ctx || !sm_count) return fail(QC_ERR_ARG, "qc_sm_count: NULL argument");
    *sm_count = ctx->sm_count;
    return QC_OK;
}

extern "C" int32_t qc_measure_fp64_peak(qc_context *ctx, double *tflops) {
    if (!ctx || !tflops) return fail(QC_ERR_ARG, "qc_measure_fp64_peak: NULL argument");
    QC_CUDA(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    QC_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double *out = nullptr;
    QC_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        QC_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 1.0 + rep);
        ctx->launches++;
        QC_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        QC_CUDA(cudaEventSynchronize(ctx->ev1));
        float ms = 0.f;
        QC_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(out);
    *tflops = best;
    return QC_OK;
}

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
extern "C" int32_t qc_plan_create(qc_context *ctx, const qc_plan_desc *desc, qc_plan **out) {
    if (!ctx || !desc || !out) return fail(QC_ERR_ARG, "qc_plan_create: NULL argument");
    const qc_plan_desc &d = *desc;
    if (d.batch < 1 || d.nt_max < 1) return fail(QC_ERR_ARG, "qc_plan_create: batch=%d nt_max=%d", d.batch, d.nt_max);
    if (!is_pow2(d.nch) || d.nch < 2 || d.nch > 128)
        return fail(QC_ERR_ARG, "qc_plan_create: nch=%d must be a power of two in [2,128]", d.nch);
    const bool grid = d.hamil == QC_HAMIL_NONTRIV;
    if (grid) {
        if (!is_pow2(d.np) || d.np < 256 || d.np > 4096)
            return fail(QC_ERR_ARG, "qc_plan_create: grid plans need np in {256,512,1024,2048,4096}, got %d", d.np);
        if (d.nlevs != 2 || d.nb != 1)
            return fail(QC_ERR_ARG, "qc_plan_create: the FFT-grid Hamiltonian has nlevs=2, nb=1 (task_manager.py:925-932)");
    } else {
        if (d.hamil < QC_HAMIL_TWO_LEVELS || d.hamil > QC_HAMIL_QUBIT)
            return fail(QC_ERR_ARG, "qc_plan_create: unknown hamil %d", d.hamil);
        if (d.np != 1) return fail(QC_ERR_ARG, "qc_plan_create: N-level plans need np=1 (newcheb.py:887-889), got %d", d.np);
        if (d.nlevs < 2 || d.nlevs > 8) return fail(QC_ERR_ARG, "qc_plan_create: nlevs=%d outside [2,8]", d.nlevs);
        if (d.hamil == QC_HAMIL_TWO_LEVELS && d.nlevs != 2) return fail(QC_ERR_ARG, "qc_plan_create: two_levels needs nlevs=2");
        if (d.nb < 1 || d.nb > 32) return fail(QC_ERR_ARG, "qc_plan_create: nb=%d outside [1,32]", d.nb);
        if (d.hf_hide) return fail(QC_ERR_ARG, "qc_plan_create: hf_hide is not defined for N-level systems (task_manager.py:916-917)");
    }
    QC_CUDA(cudaSetDevice(ctx->device));
    qc_plan *pl = new qc_plan();
    pl->ctx = ctx;
    ctx->live_plans++;
    pl->d = d;
    pl->grid = grid;
    pl->log2n = grid ? ilog2(d.np) : 0;
    pl->nbp = 1;
    while (pl->nbp < d.nb) pl->nbp <<= 1;
    pl->nthreads = (long long)d.batch * pl->nbp;
    pl->state_elems = (size_t)d.batch * d.nb * d.nlevs * d.np;
    const size_t nt1 = (size_t)d.nt_max + 1;
    cudaError_t e = cudaSuccess;
    auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    chk(pl->psi.alloc(pl->state_elems));
    chk(pl->E_base.alloc((size_t)d.batch * nt1));
    chk(pl->E_out.alloc((size_t)d.batch * nt1));
    chk(pl->E_step.alloc(d.batch));
    chk(pl->norms.alloc((size_t)d.batch * d.nb * d.nlevs));
    chk(pl->gc.alloc((size_t)d.batch * d.nb));
    chk(pl->a0.alloc((size_t)d.batch * d.nb));
    if (d.store_traj) {
        pl->traj_elems = grid ? (size_t)d.batch * nt1 * d.np : nt1 * (size_t)d.nlevs * (size_t)pl->nthreads;
        chk(pl->traj[0].alloc(pl->traj_elems));
        chk(pl->traj[1].alloc(pl->traj_elems));
    }
    if (e != cudaSuccess) {
        const int32_t rc = fail(QC_ERR_NOMEM, "qc_plan_create: device allocation failed: %s", cudaGetErrorString(e));
        qc_plan_destroy(pl);
        cudaGetLastError();
        return rc;
    }
    {
        // zeros by copy, not by cudaMemset: a memset is a kernel, and while another plan's sweep holds every SM (one CTA
        // per SM for a whole sweep) a kernel of this stream -- and every copy queued behind it -- waits for a CTA to
        // retire; plan set-up runs beside the propagation of the previous sub-batch (newcheb.run_variants)
        const size_t nz = std::max(pl->E_base.n, std::max(pl->E_out.n, pl->a0.n));
        std::vector<cd> zeros(nz, qc::cmk(0.0, 0.0));
        QC_CUDA(cudaMemcpyAsync(pl->E_base.p, zeros.data(), pl->E_base.n * sizeof(cd), cudaMemcpyHostToDevice, ctx->stream));
        QC_CUDA(cudaMemcpyAsync(pl->E_out.p, zeros.data(), pl->E_out.n * sizeof(cd), cudaMemcpyHostToDevice, ctx->stream));
        if (pl->a0.n) QC_CUDA(cudaMemcpyAsync(pl->a0.p, zeros.data(), pl->a0.n * sizeof(cd), cudaMemcpyHostToDevice, ctx->stream));
        QC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (grid) {
        std::vector<cd> t1, t2;
        fill_twiddles(pl->log2n, t1, t2);
        int32_t rc = upload(pl, pl->tw1, t1.data(), t1.size());
        if (rc == QC_OK) rc = upload(pl, pl->tw2, t2.data(), t2.size());
        if (rc == QC_OK && (pl->log2n == 11 || pl->log2n == 10)) {
            std::vector<cd> t8;
            if (pl->log2n == 11) fill_twiddles8(t8); else fill_twiddles8x2(t8);
            rc = upload(pl, pl->tw8, t8.data(), t8.size());
        }
        if (rc != QC_OK) {
            qc_plan_destroy(pl);
            return rc;
        }
    }
    // default control row: h_lambda = 1, nt = nt_max
    const double ctl[4] = {1.0, (double)d.nt_max, 0.0, 0.0};
    int32_t rc = upload(pl, pl->ctl, ctl, 4);
    if (rc != QC_OK) {
        qc_plan_destroy(pl);
        return rc;
    }
    pl->ctl_batched = 0;
    *out = pl;
    return QC_OK;
}

// Upper bound of the device memory a plan of this shape allocates over its life: state + four snapshots + four
// instrumentation work buffers + one staging buffer, the per-step scalar tables, the two trajectory buffers, and
// the per-run constants.  newcheb.run_variants sizes its sub-batches with it (qc_mem_info gives the free bytes).
extern "C" int32_t qc_plan_footprint(const qc_plan_desc *desc, int64_t *bytes) {
    if (!desc || !bytes) return fail(QC_ERR_ARG, "qc_plan_footprint: NULL argument");
    const qc_plan_desc &d = *desc;
    if (d.batch < 1 || d.nt_max < 1 || d.np < 1 || d.nlevs < 1 || d.nb < 1 || d.nch < 1)
        return fail(QC_ERR_ARG, "qc_plan_footprint: non-positive size");
    const bool grid = d.hamil == QC_HAMIL_NONTRIV;
    int nbp = 1;
    while (nbp < d.nb) nbp <<= 1;
    const double state = (double)d.batch * d.nb * d.nlevs * d.np * sizeof(cd);
    const double nt1 = (double)d.nt_max + 1.0;
    double total = 10.0 * state;                                           // psi, snap[4], work[4], goal
    total += (double)d.batch * nt1 * (5.0 * sizeof(cd) + sizeof(double));   // E_base, E_out, exp_L x 2, s_env, step times
    if (d.store_traj)
        total += 2.0 * sizeof(cd) * (grid ? (double)d.batch * nt1 * d.np : nt1 * d.nlevs * (double)d.batch * nbp);
    total += (double)d.batch * ((double)d.nlevs * d.np * sizeof(double) + (double)d.nch * sizeof(cd) + 1024.0);
    total += 1 << 20;
    *bytes = (int64_t)total;
    return QC_OK;
}

extern "C" int32_t qc_plan_destroy(qc_plan *pl) {
    if (!pl) return QC_OK;
    cudaSetDevice(pl->ctx->device);
    cudaStreamSynchronize(pl->ctx->stream);
    pl->psi.release(); pl->E_base.release(); pl->expL[0].release(); pl->expL[1].release(); pl->s_env.release(); pl->E_out.release();
    pl->a0.release(); pl->E_step.release(); pl->tw1.release(); pl->tw2.release(); pl->tw8.release(); pl->goal.release();
    pl->gc.release(); pl->tmp_row.release(); for (auto &b : pl->snap) b.release(); pl->traj[0].release(); pl->traj[1].release();
    pl->v.release(); pl->kin.release(); pl->uwd.release(); pl->dx.release(); pl->ctl.release(); pl->norms.release();
    for (auto &s : pl->poly) { s.xp.release(); s.sc.release(); s.dv.release(); s.orders.release(); }
    for (auto &b : pl->work) b.release();
    pl->zero_v.release(); pl->akx.release(); pl->xgrid.release(); pl->instr.release();
    pl->lc.release(); pl->lc_t.release(); pl->lc_rt.release(); pl->lc_rdiag.release();
    qc_context *ctx = pl->ctx;
    delete pl;
    if (--ctx->live_plans == 0 && ctx->released) context_free(ctx);
    return QC_OK;
}

extern "C" int32_t qc_set_static(qc_plan *pl, const double *v, int32_t v_batched, const double *akx2_re,
                                 const double *uwd, int32_t uwd_batched, const double *dx, int32_t dx_batched) {
    if (!pl || !v || !dx) return fail(QC_ERR_ARG, "qc_set_static: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const qc_plan_desc &d = pl->d;
    int32_t rc = upload(pl, pl->v, v, (size_t)(v_batched ? d.batch : 1) * d.nlevs * d.np);
    if (rc) return rc;
    pl->v_batched = v_batched ? 1 : 0;
    if (pl->grid) {
        if (!akx2_re) return fail(QC_ERR_ARG, "qc_set_static: akx2_re is required for grid plans");
        if ((rc = upload(pl, pl->kin, akx2_re, d.np))) return rc;
    } else if (d.hamil != QC_HAMIL_TWO_LEVELS) {
        if (!uwd) return fail(QC_ERR_ARG, "qc_set_static: (U, W, delta) required for BH_X/BH_Z/QUBIT");
    }
    if (uwd) {
        if ((rc = upload(pl, pl->uwd, uwd, (size_t)(uwd_batched ? d.batch : 1) * 3))) return rc;
        pl->uwd_batched = uwd_batched ? 1 : 0;
    }
    if ((rc = upload(pl, pl->dx, dx, dx_batched ? d.batch : 1))) return rc;
    pl->dx_batched = dx_batched ? 1 : 0;
    pl->static_set = true;
    return QC_OK;
}

extern "C" int32_t qc_set_poly(qc_plan *pl, int32_t slot, const double *xp, const double *dv, const double *sc,
                               int32_t batched) {
    if (!pl || !xp || !dv || !sc) return fail(QC_ERR_ARG, "qc_set_poly: NULL argument");
    if (slot < 0 || slot > 1) return fail(QC_ERR_ARG, "qc_set_poly: slot %d", slot);
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const size_t rows = batched ? pl->d.batch : 1;
    for (size_t r = 0; r < rows; ++r)
        if (!(sc[r * 4 + 1] > sc[r * 4 + 0]))
            return fail(QC_ERR_ARG, "qc_set_poly: emax (%g) must exceed emin (%g) in row %zu", sc[r * 4 + 1], sc[r * 4], r);
    PolySlot &s = pl->poly[slot];
    int32_t rc = upload(pl, s.xp, xp, pl->d.nch);
    if (!rc) rc = upload(pl, s.dv, dv, rows * pl->d.nch);
    if (!rc) rc = upload(pl, s.sc, sc, rows * 4);
    if (rc) return rc;
    s.batched = batched ? 1 : 0;
    s.set = true;
    return QC_OK;
}

extern "C" int32_t qc_set_poly_orders(qc_plan *pl, int32_t slot, const int32_t *orders, int32_t batched) {
    if (!pl) return fail(QC_ERR_ARG, "qc_set_poly_orders: NULL plan");
    if (slot < 0 || slot > 1 || !pl->poly[slot].set) return fail(QC_ERR_STATE, "qc_set_poly_orders: slot %d has no polynomial yet", slot);
    PolySlot &s = pl->poly[slot];
    if (!orders) {
        s.truncated = false;
        return QC_OK;
    }
    if ((batched != 0) != (s.batched != 0)) return fail(QC_ERR_ARG, "qc_set_poly_orders: batched must match qc_set_poly");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const size_t rows = batched ? pl->d.batch : 1;
    for (size_t r = 0; r < rows; ++r)
        if (orders[r] < 2 || orders[r] > pl->d.nch)
            return fail(QC_ERR_ARG, "qc_set_poly_orders: order %d of row %zu outside [2, nch=%d]", orders[r], r, pl->d.nch);
    if (s.orders.n != rows) QC_CUDA(s.orders.alloc(rows));
    QC_CUDA(cudaMemcpyAsync(s.orders.p, orders, rows * sizeof(int), cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    s.truncated = true;
    return QC_OK;
}

extern "C" int32_t qc_set_state(qc_plan *pl, const double *psi) {
    if (!pl || !psi) return fail(QC_ERR_ARG, "qc_set_state: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(pl->psi.p, psi, pl->state_elems * sizeof(cd), cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_get_state(qc_plan *pl, double *psi) {
    if (!pl || !psi) return fail(QC_ERR_ARG, "qc_get_state: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(psi, pl->psi.p, pl->state_elems * sizeof(cd), cudaMemcpyDeviceToHost, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_set_step_scalars(qc_plan *pl, int32_t which, const double *vals, int32_t batched) {
    if (!pl || !vals) return fail(QC_ERR_ARG, "qc_set_step_scalars: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const size_t row = (size_t)pl->d.nt_max + 1;
    if (which == 0) return upload_rows_expanded(pl, pl->E_base, vals, batched, row);
    if (which == 1 || which == 3) {
        const int slot = which == 1 ? 0 : 1;
        pl->expL_batched[slot] = batched ? 1 : 0;
        return upload(pl, pl->expL[slot], vals, (batched ? pl->d.batch : 1) * row);
    }
    if (which == 2) {
        pl->s_batched = batched ? 1 : 0;
        return upload(pl, pl->s_env, vals, (batched ? pl->d.batch : 1) * row);
    }
    return fail(QC_ERR_ARG, "qc_set_step_scalars: which=%d", which);
}

extern "C" int32_t qc_set_control(qc_plan *pl, const double *ctl, int32_t batched, const double *a0) {
    if (!pl || !ctl) return fail(QC_ERR_ARG, "qc_set_control: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const size_t rows = batched ? pl->d.batch : 1;
    for (size_t r = 0; r < rows; ++r) {
        const double nt = ctl[r * 4 + 1];
        if (!(nt >= 1.0) || nt > pl->d.nt_max)
            return fail(QC_ERR_ARG, "qc_set_control: nt=%g of row %zu outside [1, nt_max=%d]", nt, r, pl->d.nt_max);
        if (ctl[r * 4 + 0] == 0.0) return fail(QC_ERR_ARG, "qc_set_control: h_lambda of row %zu is zero", r);
    }
    int32_t rc = upload(pl, pl->ctl, ctl, rows * 4);
    if (rc) return rc;
    pl->ctl_batched = batched ? 1 : 0;
    if (a0) {
        QC_CUDA(cudaMemcpyAsync(pl->a0.p, a0, pl->a0.n * sizeof(cd), cudaMemcpyHostToDevice, pl->ctx->stream));
        QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    }
    return QC_OK;
}

// ---------------------------------------------------------------------------
// sweeps
// ---------------------------------------------------------------------------
#ifndef QC_GRID_VAR_DEFAULT
#define QC_GRID_VAR_DEFAULT 1
#endif
template <int LOG2N, bool LOCAL = false, bool CL = false, int RPC = 0, bool ILV = false, int VAR = 0>
static int32_t launch_grid(qc_plan *pl, const qc::GridParams &gp, cudaStream_t stream) {
    typedef qc::GridCfg<LOG2N, CL, RPC, ILV, VAR> C;
    static bool attr_done[64] = {false};
    const int dev = pl->ctx->device;
    if (!attr_done[dev & 63]) {
        QC_CUDA(cudaFuncSetAttribute(qc::grid_sweep_kernel<LOG2N, LOCAL, CL, RPC, ILV, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)C::SMEM_BYTES));
        attr_done[dev & 63] = true;
    }
    if constexpr (CL) {                              // one cluster of two CTAs (one level each) per run
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2u * (unsigned)gp.run_count);
        cfg.blockDim = dim3(C::NTHREADS);
        cfg.dynamicSmemBytes = C::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        QC_CUDA(cudaLaunchKernelEx(&cfg, qc::grid_sweep_kernel<LOG2N, LOCAL, CL>, gp));
    } else {
        const int ctas = RPC > 0 ? (gp.run_count + RPC - 1) / RPC : gp.run_count;
        qc::grid_sweep_kernel<LOG2N, LOCAL, CL, RPC, ILV, VAR><<<ctas, C::NTHREADS, C::SMEM_BYTES, stream>>>(gp);
    }
    pl->ctx->launches++;
    QC_CUDA(cudaGetLastError());
    return QC_OK;
}

static int32_t launch_grid8(qc_plan *pl, const qc::GridParams &gp, cudaStream_t stream) {
    typedef qc::Grid8Cfg C;
    static bool attr_done[64] = {false};
    const int dev = pl->ctx->device;
    if (!attr_done[dev & 63]) {
        QC_CUDA(cudaFuncSetAttribute(qc::grid8_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        attr_done[dev & 63] = true;
    }
    qc::grid8_sweep_kernel<<<gp.run_count, C::NTHREADS, C::SMEM_BYTES, stream>>>(gp);
    pl->ctx->launches++;
    QC_CUDA(cudaGetLastError());
    return QC_OK;
}

static int32_t launch_grid8x2(qc_plan *pl, const qc::GridParams &gp, cudaStream_t stream) {
    typedef qc::Grid8x2Cfg C;
    static bool attr_done[64] = {false};
    const int dev = pl->ctx->device;
    if (!attr_done[dev & 63]) {
        QC_CUDA(cudaFuncSetAttribute(qc::grid8x2_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        attr_done[dev & 63] = true;
    }
    qc::grid8x2_sweep_kernel<<<(gp.run_count + 1) / 2, C::NTHREADS, C::SMEM_BYTES, stream>>>(gp);
    pl->ctx->launches++;
    QC_CUDA(cudaGetLastError());
    return QC_OK;
}

template <int NL>
static int32_t launch_nlevel(qc_plan *pl, const qc::NLevelParams &np_) {
    const int threads = 128;
    const unsigned blocks = (unsigned)((np_.nthreads + threads - 1) / threads);
    if (np_.nch <= 8)
        qc::nlevel_sweep_kernel<NL, 8><<<blocks, threads, 0, pl->ctx->stream>>>(np_);
    else if (np_.nch <= 16)
        qc::nlevel_sweep_kernel<NL, 16><<<blocks, threads, 0, pl->ctx->stream>>>(np_);
    else
        qc::nlevel_sweep_kernel<NL, 0><<<blocks, threads, 0, pl->ctx->stream>>>(np_);
    pl->ctx->launches++;
    QC_CUDA(cudaGetLastError());
    return QC_OK;
}

struct ApplyOpts {
    int run_first = 0, run_count = -1;   // sub-range of the batch (grid plans), -1 = all
    cudaStream_t stream = nullptr;       // launch stream override
    bool active = false;        // bare operator application through the identity polynomial (slot 2)
    bool cplx = false;          // complex coupling (orig=True form)
    const double *v = nullptr;  // potential override (device) or null
    const double *kin = nullptr;
};

static int32_t run_sweep(qc_plan *pl, const qc_sweep_args *a, bool single_step, const ApplyOpts &ap = ApplyOpts()) {
    if (!pl || !a) return fail(QC_ERR_ARG, "qc_sweep: NULL argument");
    const qc_plan_desc &d = pl->d;
    if (!pl->static_set) return fail(QC_ERR_STATE, "qc_sweep: qc_set_static has not been called");
    if (a->poly_slot < 0 || a->poly_slot > (ap.active ? 2 : 1) || !pl->poly[a->poly_slot].set)
        return fail(QC_ERR_STATE, "qc_sweep: polynomial slot %d has not been set", a->poly_slot);
    if (a->nsteps < 1 || a->l_first < 0 || a->l_first > d.nt_max)
        return fail(QC_ERR_ARG, "qc_sweep: l_first=%d nsteps=%d (nt_max=%d)", a->l_first, a->nsteps, d.nt_max);
    const int mode = a->field_mode;
    if (mode < 0 || mode > 4) return fail(QC_ERR_ARG, "qc_sweep: field_mode %d", mode);
    const bool local = mode == QC_FIELD_LOCAL_PROJ;
    if (local && (!pl->grid || !d.hf_hide || single_step || !pl->lc_set))
        return fail(QC_ERR_STATE, "qc_sweep: QC_FIELD_LOCAL_PROJ needs a grid plan with hf_hide and qc_set_local_control");
    if (pl->grid && mode == QC_FIELD_UT_FWD) return fail(QC_ERR_ARG, "qc_sweep: UT field update needs an N-level plan (newcheb.py:887-889)");
    if (!pl->grid && (mode == QC_FIELD_KROTOV_FWD || mode == QC_FIELD_KROTOV_BWD))
        return fail(QC_ERR_ARG, "qc_sweep: Krotov field update needs a grid plan");
    if (a->traj_read > 1 || a->traj_write > 1 || (a->traj_read >= 0 && a->traj_read == a->traj_write))
        return fail(QC_ERR_ARG, "qc_sweep: traj_read=%d traj_write=%d", a->traj_read, a->traj_write);
    if ((a->traj_read >= 0 || a->traj_write >= 0) && !d.store_traj)
        return fail(QC_ERR_STATE, "qc_sweep: plan was created without trajectory buffers");
    if (mode != QC_FIELD_PRESCRIBED && !local && a->traj_read < 0)
        return fail(QC_ERR_ARG, "qc_sweep: field mode %d reads the previous sweep; traj_read must be 0 or 1", mode);
    if (d.hf_hide && !single_step && !local && !pl->expL[a->poly_slot].p) return fail(QC_ERR_STATE, "qc_sweep: exp_L table (qc_set_step_scalars which=1) missing");
    if (mode == QC_FIELD_UT_FWD && !pl->s_env.p) return fail(QC_ERR_STATE, "qc_sweep: s_env table (qc_set_step_scalars which=2) missing");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const PolySlot &ps = pl->poly[a->poly_slot];
    if (pl->grid) {
        qc::GridParams g;
        memset(&g, 0, sizeof g);
        g.run_first = ap.run_first; g.run_count = ap.run_count < 0 ? d.batch : ap.run_count;
        cudaStream_t lst = ap.stream ? ap.stream : pl->ctx->stream;
        g.batch = d.batch; g.nch = ap.active ? 2 : d.nch; g.nt_max = d.nt_max; g.cplx_coupling = ap.cplx ? 1 : 0;
        g.hf_hide = single_step ? 0 : d.hf_hide; g.renorm = a->renorm; g.field_mode = mode;
        g.reversed_E = a->reversed_E; g.write_E_base = a->write_E_base;
        g.l_first = a->l_first; g.nsteps = a->nsteps; g.single_step = single_step ? 1 : 0;
        g.psi = pl->psi.p; g.v = ap.v ? ap.v : pl->v.p; g.v_stride = (!ap.v && pl->v_batched) ? 2LL * d.np : 0;
        g.kin = ap.kin ? ap.kin : pl->kin.p;
        g.tw1 = pl->tw1.p; g.tw2 = pl->tw2.p; g.tw8 = pl->tw8.p; g.xp = ps.xp.p; g.dv = ps.dv.p; g.sc = ps.sc.p;
        g.orders = (ps.truncated && !ap.active && !local) ? ps.orders.p : nullptr;
        g.poly_stride = ps.batched; g.dx = pl->dx.p; g.dx_stride = pl->dx_batched;
        g.E_base = pl->E_base.p; g.E_stride = 1; g.expL = pl->expL[a->poly_slot].p; g.expL_stride = pl->expL_batched[a->poly_slot];
        g.ctl = pl->ctl.p; g.ctl_stride = pl->ctl_batched; g.E_step = pl->E_step.p; g.E_out = pl->E_out.p;
        g.traj_in = a->traj_read >= 0 ? pl->traj[a->traj_read].p : nullptr;
        g.traj_out = a->traj_write >= 0 ? pl->traj[a->traj_write].p : nullptr;
        g.norms = pl->norms.p;
        g.lc = pl->lc.p; g.lc_t = pl->lc_t.p; g.lc_rt = pl->lc_rt.p; g.lc_rdiag = pl->lc_rdiag.p;
        if (local) {
            switch (pl->log2n) {
                case 8: return launch_grid<8, true>(pl, g, lst);
                case 9: return launch_grid<9, true>(pl, g, lst);
                case 10: return launch_grid<10, true>(pl, g, lst);
                case 11: return launch_grid<11, true>(pl, g, lst);
            }
            return fail(QC_ERR_ARG, "qc_sweep: unsupported np");
        }
        // cluster form (two SMs per run): the only form for np = 4096; at np = 2048 it shortens a step by 11 % when
        // the batch cannot fill the GPU anyway (237 -> 209 us); below that the cluster barrier costs more than the
        // second SM brings (np = 1024: 133 -> 160 us).  QC_GRID_CLUSTER=0/1 overrides the heuristic (tests).
        bool cluster = pl->log2n == 12 || (pl->log2n == 11 && g.run_count * 2 <= pl->ctx->sm_count);
        if (const char *e = getenv("QC_GRID_CLUSTER")) cluster = pl->log2n == 12 || (pl->log2n >= 9 && atoi(e) != 0);
        if (cluster) {
            switch (pl->log2n) {
                case 9: return launch_grid<9, false, true>(pl, g, lst);
                case 10: return launch_grid<10, false, true>(pl, g, lst);
                case 11: return launch_grid<11, false, true>(pl, g, lst);
                case 12: return launch_grid<12, false, true>(pl, g, lst);
            }
        }
        // packed form (np <= 512): several runs per CTA so that every SM holds eight warps -- np = 256: four runs, the
        // levels interleaved in every warp (0.26 -> 0.60 of the FP64 peak); np = 512: two runs, a level per warp
        // (0.42 -> 0.52).  Batches that are resident all at once in the one-run-per-CTA form (two CTAs per SM) keep
        // it: its single step is 4-12 % shorter.
        // QC_GRID_PACK=0/1 overrides the heuristic (tests run both forms).
        bool packed = pl->log2n <= 9 && g.run_count > 2 * pl->ctx->sm_count;
        if (const char *e = getenv("QC_GRID_PACK")) packed = pl->log2n <= 9 && atoi(e) != 0;
        if (packed) {
            bool four = QC_GRID_VAR_DEFAULT == 1;       // np = 512: four runs per CTA with the tensor-memory hand-over
            if (const char *e = getenv("QC_GRID_TMP")) four = atoi(e) != 0;
            switch (pl->log2n) {
                case 8: return launch_grid<8, false, false, 4, true>(pl, g, lst);
                case 9:
                    if (four) return launch_grid<9, false, false, 4, false, 1>(pl, g, lst);
                    return launch_grid<9, false, false, 2, false>(pl, g, lst);
            }
        }
        // kernel variants of the plain form (qc_grid.cuh, GridCfg::VAR); QC_GRID_VAR overrides the default (A/B runs)
        int var = QC_GRID_VAR_DEFAULT;
        if (const char *e = getenv("QC_GRID_VAR")) var = atoi(e);
        switch (pl->log2n) {
            case 8: return launch_grid<8>(pl, g, lst);
            case 9: return launch_grid<9>(pl, g, lst);
            case 10:
                // 512-thread form, two runs per CTA (qc_grid8x2.cuh): sweeps of the prescribed / Krotov field laws
                if (var == 2 && !single_step && !ap.cplx && !ap.active && mode <= QC_FIELD_KROTOV_BWD)
                    return launch_grid8x2(pl, g, lst);
                // two runs per CTA with the tensor-memory hand-over once the batch fills every SM anyway (the plain
                // form holds two one-run CTAs per SM: same residency, 20 % more shared-memory traffic)
                {
                    bool two = var == 1 && g.run_count > pl->ctx->sm_count;
                    if (const char *e = getenv("QC_GRID_TMP")) two = atoi(e) != 0;      // tests force either form
                    if (two) return launch_grid<10, false, false, 2, false, 1>(pl, g, lst);
                }
                return launch_grid<10>(pl, g, lst);
            case 11:
                // 512-thread form (qc_grid8.cuh): sweeps of the prescribed / Krotov field laws
                if (var == 2 && !single_step && !ap.cplx && !ap.active && mode <= QC_FIELD_KROTOV_BWD)
                    return launch_grid8(pl, g, lst);
                if (var >= 1) return launch_grid<11, false, false, 0, false, 1>(pl, g, lst);
                return launch_grid<11>(pl, g, lst);
        }
        return fail(QC_ERR_ARG, "qc_sweep: unsupported np");
    }
    qc::NLevelParams n;
    memset(&n, 0, sizeof n);
    n.batch = d.batch; n.nb = d.nb; n.nbp = pl->nbp; n.nch = ap.active ? 2 : d.nch; n.nt_max = d.nt_max; n.hamil = d.hamil;
    n.renorm = a->renorm; n.field_mode = mode; n.reversed_E = a->reversed_E; n.write_E_base = a->write_E_base;
    n.l_first = a->l_first; n.nsteps = a->nsteps; n.single_step = single_step ? 1 : 0; n.nthreads = pl->nthreads;
    n.psi = pl->psi.p; n.vdiag = pl->v.p; n.v_stride = pl->v_batched; n.uwd = pl->uwd.p; n.uwd_stride = pl->uwd_batched;
    n.xp = ps.xp.p; n.dv = ps.dv.p; n.sc = ps.sc.p; n.poly_stride = ps.batched;
    n.orders = (ps.truncated && !ap.active) ? ps.orders.p : nullptr;
    if (const char *e = getenv("QC_NLEVEL_RETIRE")) n.no_retire = atoi(e) == 0; n.dx = pl->dx.p; n.dx_stride = pl->dx_batched;
    n.E_base = pl->E_base.p; n.E_stride = 1; n.s_env = pl->s_env.p; n.s_stride = pl->s_batched;
    n.ctl = pl->ctl.p; n.ctl_stride = pl->ctl_batched; n.a0 = pl->a0.p; n.E_step = pl->E_step.p; n.E_out = pl->E_out.p;
    n.traj_in = a->traj_read >= 0 ? pl->traj[a->traj_read].p : nullptr;
    n.traj_out = a->traj_write >= 0 ? pl->traj[a->traj_write].p : nullptr;
    n.norms = pl->norms.p;
    switch (d.nlevs) {
        case 2: return launch_nlevel<2>(pl, n);
        case 3: return launch_nlevel<3>(pl, n);
        case 4: return launch_nlevel<4>(pl, n);
        case 5: return launch_nlevel<5>(pl, n);
        case 6: return launch_nlevel<6>(pl, n);
        case 7: return launch_nlevel<7>(pl, n);
        case 8: return launch_nlevel<8>(pl, n);
    }
    return fail(QC_ERR_ARG, "qc_sweep: unsupported nlevs");
}

extern "C" int32_t qc_sweep(qc_plan *pl, const qc_sweep_args *a) { return run_sweep(pl, a, false); }

extern "C" int32_t qc_prop_step(qc_plan *pl, int32_t poly_slot, const double *E) {
    if (!pl || !E) return fail(QC_ERR_ARG, "qc_prop_step: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(pl->E_step.p, E, pl->d.batch * sizeof(cd), cudaMemcpyHostToDevice, pl->ctx->stream));
    qc_sweep_args a;
    memset(&a, 0, sizeof a);
    a.poly_slot = poly_slot; a.field_mode = QC_FIELD_PRESCRIBED; a.l_first = 1; a.nsteps = 1; a.renorm = 0;
    a.traj_read = -1; a.traj_write = -1;
    int32_t rc = run_sweep(pl, &a, true);
    if (rc) return rc;
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}


// ---------------------------------------------------------------------------
// bare operator applications and instrumentation
// ---------------------------------------------------------------------------
static int32_t ensure_identity_poly(qc_plan *pl, double eL) {
    PolySlot &s = pl->poly[2];
    const int nch = pl->d.nch;
    std::vector<double> xp(nch, 0.0), dv(2 * (size_t)nch, 0.0);
    dv[2] = 1.0;                                   // dv = (0, 1, 0, ...): psi_out = (Hn - 0) psi
    const double sc[4] = {-2.0, 2.0, 0.0, eL};     // c1 = 1, c2 = 0, phase = 1
    int32_t rc = upload(pl, s.xp, xp.data(), nch);
    if (!rc) rc = upload(pl, s.dv, dv.data(), nch);
    if (!rc) rc = upload(pl, s.sc, sc, 4);
    s.batched = 0;
    s.set = rc == QC_OK;
    return rc;
}

// psi <- Op psi on the device (the caller has saved psi if it needs it)
static int32_t apply_operator(qc_plan *pl, int32_t which, int32_t orig, const double *E_host, double eL) {
    const qc_plan_desc &d = pl->d;
    cudaStream_t st = pl->ctx->stream;
    int32_t rc = ensure_identity_poly(pl, (which == 0 && !orig) ? eL : 0.0);
    if (rc) return rc;
    if (which == 0 && E_host) {
        QC_CUDA(cudaMemcpyAsync(pl->E_step.p, E_host, d.batch * sizeof(cd), cudaMemcpyHostToDevice, st));
    } else {
        QC_CUDA(cudaMemsetAsync(pl->E_step.p, 0, d.batch * sizeof(cd), st));
    }
    ApplyOpts ap;
    ap.active = true;
    ap.cplx = pl->grid && which == 0 && orig;
    if (which != 0) {
        if (!pl->grid) return fail(QC_ERR_ARG, "qc_apply_hamil: kinetic / momentum applications need a grid plan");
        if (pl->zero_v.n != (size_t)2 * d.np) {
            QC_CUDA(pl->zero_v.alloc((size_t)2 * d.np));
            QC_CUDA(cudaMemsetAsync(pl->zero_v.p, 0, pl->zero_v.n * sizeof(double), st));
        }
        ap.v = pl->zero_v.p;
        if (which == 2) {
            if (!pl->akx.p) return fail(QC_ERR_STATE, "qc_apply_hamil: momentum multiplier not set (qc_set_instr_grid)");
            ap.kin = pl->akx.p;
        }
    }
    qc_sweep_args a;
    memset(&a, 0, sizeof a);
    a.poly_slot = 2; a.field_mode = QC_FIELD_PRESCRIBED; a.l_first = 1; a.nsteps = 1; a.renorm = 0;
    a.traj_read = -1; a.traj_write = -1;
    return run_sweep(pl, &a, true, ap);
}

extern "C" int32_t qc_apply_hamil(qc_plan *pl, int32_t which, int32_t orig, const double *E, double eL, double *out) {
    if (!pl || !out) return fail(QC_ERR_ARG, "qc_apply_hamil: NULL argument");
    if (which < 0 || which > 2) return fail(QC_ERR_ARG, "qc_apply_hamil: which=%d", which);
    if (!pl->static_set) return fail(QC_ERR_STATE, "qc_apply_hamil: qc_set_static has not been called");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    cudaStream_t st = pl->ctx->stream;
    if (pl->work[0].n != pl->state_elems) QC_CUDA(pl->work[0].alloc(pl->state_elems));
    const size_t bytes = pl->state_elems * sizeof(cd);
    QC_CUDA(cudaMemcpyAsync(pl->work[0].p, pl->psi.p, bytes, cudaMemcpyDeviceToDevice, st));
    int32_t rc = apply_operator(pl, which, orig, E, eL);
    if (rc) return rc;
    QC_CUDA(cudaMemcpyAsync(out, pl->psi.p, bytes, cudaMemcpyDeviceToHost, st));
    QC_CUDA(cudaMemcpyAsync(pl->psi.p, pl->work[0].p, bytes, cudaMemcpyDeviceToDevice, st));
    QC_CUDA(cudaStreamSynchronize(st));
    return QC_OK;
}

extern "C" int32_t qc_set_instr_grid(qc_plan *pl, const double *x, const double *akx_re) {
    if (!pl || !x) return fail(QC_ERR_ARG, "qc_set_instr_grid: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    int32_t rc = upload(pl, pl->xgrid, x, pl->d.np);
    if (!rc && akx_re) rc = upload(pl, pl->akx, akx_re, pl->d.np);
    return rc;
}

extern "C" int32_t qc_instrument(qc_plan *pl, const double *E_full, double two_m, int32_t psi0_slot, int32_t psif_slot,
                                 double *out) {
    if (!pl || !out) return fail(QC_ERR_ARG, "qc_instrument: NULL argument");
    if (!pl->static_set) return fail(QC_ERR_STATE, "qc_instrument: qc_set_static has not been called");
    if (psi0_slot > 3 || psif_slot > 3) return fail(QC_ERR_ARG, "qc_instrument: snapshot slots");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const qc_plan_desc &d = pl->d;
    cudaStream_t st = pl->ctx->stream;
    const size_t bytes = pl->state_elems * sizeof(cd);
    for (int k = 0; k < 4; ++k)
        if (pl->work[k].n != pl->state_elems) QC_CUDA(pl->work[k].alloc(pl->state_elems));
    QC_CUDA(cudaMemcpyAsync(pl->work[0].p, pl->psi.p, bytes, cudaMemcpyDeviceToDevice, st));
    const bool moments = pl->grid && pl->akx.p;
    for (int which = 0; which < (moments ? 3 : 1); ++which) {
        int32_t rc = apply_operator(pl, which, 1, E_full, 0.0);
        if (rc) return rc;
        QC_CUDA(cudaMemcpyAsync(pl->work[1 + which].p, pl->psi.p, bytes, cudaMemcpyDeviceToDevice, st));
        QC_CUDA(cudaMemcpyAsync(pl->psi.p, pl->work[0].p, bytes, cudaMemcpyDeviceToDevice, st));
    }
    const size_t items = (size_t)d.batch * d.nb * d.nlevs;
    if (pl->instr.n != items * 20) QC_CUDA(pl->instr.alloc(items * 20));
    instrument_kernel<<<(unsigned)items, d.np >= 256 ? 256 : 32, 0, st>>>(
        pl->psi.p, pl->work[1].p, moments ? pl->work[2].p : nullptr, moments ? pl->work[3].p : nullptr,
        psi0_slot >= 0 ? pl->snap[psi0_slot].p : nullptr, psif_slot >= 0 ? pl->snap[psif_slot].p : nullptr,
        pl->xgrid.p, pl->dx.p, pl->dx_batched, two_m, d.nb, d.nlevs, d.np, pl->instr.p);
    pl->ctx->launches++;
    QC_CUDA(cudaGetLastError());
    QC_CUDA(cudaMemcpyAsync(out, pl->instr.p, items * 20 * sizeof(double), cudaMemcpyDeviceToHost, st));
    QC_CUDA(cudaStreamSynchronize(st));
    return QC_OK;
}

extern "C" int32_t qc_set_local_control(qc_plan *pl, const double *params, const double *t_tab, const double *xp) {
    if (!pl || !params || !t_tab || !xp) return fail(QC_ERR_ARG, "qc_set_local_control: NULL argument");
    if (!pl->grid || !pl->d.hf_hide) return fail(QC_ERR_STATE, "qc_set_local_control: needs a grid plan with hf_hide");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const int nch = pl->d.nch, B = pl->d.batch;
    std::vector<double> lc((size_t)B * qc::LC_STRIDE, 0.0);
    for (int b = 0; b < B; ++b) {
        double *r = &lc[(size_t)b * qc::LC_STRIDE];
        r[qc::LC_FM] = 1.0;                                 // FitterDynamicState, fitter.py:22-31
        r[qc::LC_PATCHED] = 1.0;
        const double *q = params + (size_t)b * 12;
        for (int k = 0; k < 4; ++k) r[qc::LC_EXTR0 + k] = q[k];
        r[qc::LC_E0] = q[4]; r[qc::LC_NUL] = q[5]; r[qc::LC_NU] = q[6]; r[qc::LC_KE] = q[7]; r[qc::LC_LAMB] = q[8];
        r[qc::LC_POW] = q[9]; r[qc::LC_EPS] = q[10]; r[qc::LC_DT] = q[11];
        if (q[11] == 0.0) return fail(QC_ERR_ARG, "qc_set_local_control: dt of row %d is zero", b);
    }
    // partial products of the Newton table (math_base.py:176-186), transposed: rt[j][i] = prod_{m<=j} (xp_i - xp_m)
    std::vector<double> rt((size_t)nch * nch, 0.0), rdiag(nch, 0.0);
    for (int i = 1; i < nch; ++i) {
        double res = 1.0;
        for (int j = 0; j < i; ++j) {
            res *= (xp[i] - xp[j]);
            rt[(size_t)j * nch + i] = res;
        }
        rdiag[i] = 1.0 / res;
    }
    int32_t rc = upload(pl, pl->lc, lc.data(), lc.size());
    if (!rc) rc = upload(pl, pl->lc_t, t_tab, (size_t)B * (pl->d.nt_max + 1));
    if (!rc) rc = upload(pl, pl->lc_rt, rt.data(), rt.size());
    if (!rc) rc = upload(pl, pl->lc_rdiag, rdiag.data(), rdiag.size());
    if (rc) return rc;
    pl->lc_set = true;
    return QC_OK;
}

extern "C" int32_t qc_get_local_state(qc_plan *pl, double *out) {
    if (!pl || !out) return fail(QC_ERR_ARG, "qc_get_local_state: NULL argument");
    if (!pl->lc_set) return fail(QC_ERR_STATE, "qc_get_local_state: qc_set_local_control has not been called");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpy2DAsync(out, qc::LC_NSTATE * sizeof(double), pl->lc.p, qc::LC_STRIDE * sizeof(double),
                              qc::LC_NSTATE * sizeof(double), pl->d.batch, cudaMemcpyDeviceToHost, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_set_local_state(qc_plan *pl, const double *state) {
    if (!pl || !state) return fail(QC_ERR_ARG, "qc_set_local_state: NULL argument");
    if (!pl->lc_set) return fail(QC_ERR_STATE, "qc_set_local_state: qc_set_local_control has not been called");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpy2DAsync(pl->lc.p, qc::LC_STRIDE * sizeof(double), state, 4 * sizeof(double), 4 * sizeof(double),
                              pl->d.batch, cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_get_fields(qc_plan *pl, double *E_out) {
    if (!pl || !E_out) return fail(QC_ERR_ARG, "qc_get_fields: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(E_out, pl->E_out.p, pl->E_out.n * sizeof(cd), cudaMemcpyDeviceToHost, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_field_integral(qc_plan *pl, double *out) {
    if (!pl || !out) return fail(QC_ERR_ARG, "qc_field_integral: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const qc_plan_desc &d = pl->d;
    if (pl->instr.n < (size_t)d.batch) QC_CUDA(pl->instr.alloc((size_t)d.batch * 20 * d.nb * d.nlevs));
    field_integral_kernel<<<d.batch, 128, 0, pl->ctx->stream>>>(pl->E_out.p, pl->ctl.p, pl->ctl_batched, d.nt_max, pl->instr.p);
    pl->ctx->launches++;
    QC_CUDA(cudaGetLastError());
    QC_CUDA(cudaMemcpyAsync(out, pl->instr.p, d.batch * sizeof(double), cudaMemcpyDeviceToHost, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_get_norms(qc_plan *pl, double *norms) {
    if (!pl || !norms) return fail(QC_ERR_ARG, "qc_get_norms: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(norms, pl->norms.p, pl->norms.n * sizeof(double), cudaMemcpyDeviceToHost, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_snapshot_set(qc_plan *pl, int32_t slot, const double *psi_host) {
    if (!pl || slot < 0 || slot > 3) return fail(QC_ERR_ARG, "qc_snapshot_set: slot=%d", slot);
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    if (pl->snap[slot].n != pl->state_elems) QC_CUDA(pl->snap[slot].alloc(pl->state_elems));
    if (psi_host) {
        QC_CUDA(cudaMemcpyAsync(pl->snap[slot].p, psi_host, pl->state_elems * sizeof(cd), cudaMemcpyHostToDevice,
                                pl->ctx->stream));
        QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    } else {
        QC_CUDA(cudaMemcpyAsync(pl->snap[slot].p, pl->psi.p, pl->state_elems * sizeof(cd), cudaMemcpyDeviceToDevice,
                                pl->ctx->stream));
    }
    return QC_OK;
}

extern "C" int32_t qc_snapshot_load(qc_plan *pl, int32_t slot) {
    if (!pl || slot < 0 || slot > 3 || !pl->snap[slot].p) return fail(QC_ERR_ARG, "qc_snapshot_load: slot=%d is empty", slot);
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(pl->psi.p, pl->snap[slot].p, pl->state_elems * sizeof(cd), cudaMemcpyDeviceToDevice,
                            pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_goal_overlap(qc_plan *pl, int32_t goal_slot, double *gc) {
    if (!pl || goal_slot < 0 || goal_slot > 3 || !pl->snap[goal_slot].p)
        return fail(QC_ERR_ARG, "qc_goal_overlap: snapshot slot %d is empty", goal_slot);
    if (!pl->dx.p) return fail(QC_ERR_STATE, "qc_goal_overlap: qc_set_static has not been called");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const qc_plan_desc &d = pl->d;
    const int per_vec = d.nlevs * d.np;
    goal_overlap_kernel<<<d.batch * d.nb, per_vec >= 256 ? 256 : 32, 0, pl->ctx->stream>>>(
        pl->psi.p, pl->snap[goal_slot].p, (long long)d.nb * per_vec, pl->dx.p, pl->dx_batched, per_vec, pl->gc.p, d.nb);
    pl->ctx->launches++;
    QC_CUDA(cudaGetLastError());
    if (gc) {
        QC_CUDA(cudaMemcpyAsync(gc, pl->gc.p, pl->gc.n * sizeof(cd), cudaMemcpyDeviceToHost, pl->ctx->stream));
        QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    }
    return QC_OK;
}

extern "C" int32_t qc_krotov_terminal(qc_plan *pl, int32_t goal_slot, const double *gc, const double *sqrt_hf_T,
                                      int32_t hf_batched, int32_t traj_write) {
    if (!pl || !sqrt_hf_T) return fail(QC_ERR_ARG, "qc_krotov_terminal: NULL argument");
    if (!pl->grid) return fail(QC_ERR_ARG, "qc_krotov_terminal: grid plans only (task_manager.py:925-932)");
    if (goal_slot < 0 || goal_slot > 3 || !pl->snap[goal_slot].p)
        return fail(QC_ERR_ARG, "qc_krotov_terminal: snapshot slot %d is empty", goal_slot);
    if (traj_write > 1 || (traj_write >= 0 && !pl->d.store_traj)) return fail(QC_ERR_ARG, "qc_krotov_terminal: traj_write=%d", traj_write);
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const qc_plan_desc &d = pl->d;
    if (gc) QC_CUDA(cudaMemcpyAsync(pl->gc.p, gc, pl->gc.n * sizeof(cd), cudaMemcpyHostToDevice, pl->ctx->stream));
    const size_t hrows = hf_batched ? d.batch : 1;
    if (pl->tmp_row.n < hrows) QC_CUDA(pl->tmp_row.alloc(hrows));
    QC_CUDA(cudaMemcpyAsync(pl->tmp_row.p, sqrt_hf_T, hrows * sizeof(cd), cudaMemcpyHostToDevice, pl->ctx->stream));
    krotov_terminal_kernel<<<d.batch, 256, 0, pl->ctx->stream>>>(
        pl->psi.p, pl->snap[goal_slot].p, 2LL * d.np, pl->gc.p, pl->tmp_row.p, hf_batched ? 1 : 0, pl->dx.p,
        pl->dx_batched, traj_write >= 0 ? pl->traj[traj_write].p : nullptr, (long long)(d.nt_max + 1) * d.np, d.np);
    pl->ctx->launches++;
    QC_CUDA(cudaGetLastError());
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_get_traj(qc_plan *pl, int32_t buffer, int32_t index, double *out) {
    if (!pl || !out) return fail(QC_ERR_ARG, "qc_get_traj: NULL argument");
    if (buffer < 0 || buffer > 1 || !pl->d.store_traj || index < 0 || index > pl->d.nt_max)
        return fail(QC_ERR_ARG, "qc_get_traj: buffer=%d index=%d", buffer, index);
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const qc_plan_desc &d = pl->d;
    cudaStream_t st = pl->ctx->stream;
    if (pl->grid) {
        const size_t n = (size_t)d.batch * d.np;
        if (pl->goal.n < n) QC_CUDA(pl->goal.alloc(n));
        gather_grid_slice_kernel<<<d.batch, 256, 0, st>>>(pl->traj[buffer].p + (size_t)index * d.np, pl->goal.p,
                                                          (long long)(d.nt_max + 1) * d.np, d.np);
        pl->ctx->launches++;
        QC_CUDA(cudaGetLastError());
        QC_CUDA(cudaMemcpyAsync(out, pl->goal.p, n * sizeof(cd), cudaMemcpyDeviceToHost, st));
    } else {
        const size_t n = (size_t)d.batch * d.nb * d.nlevs;
        if (pl->goal.n < n) QC_CUDA(pl->goal.alloc(n));
        const cd *slice = pl->traj[buffer].p + (size_t)index * d.nlevs * pl->nthreads;
        gather_nlevel_kernel<<<(unsigned)((pl->nthreads + 127) / 128), 128, 0, st>>>(slice, pl->goal.p, d.nb, pl->nbp,
                                                                                   d.nlevs, pl->nthreads);
        pl->ctx->launches++;
        QC_CUDA(cudaGetLastError());
        QC_CUDA(cudaMemcpyAsync(out, pl->goal.p, n * sizeof(cd), cudaMemcpyDeviceToHost, st));
    }
    QC_CUDA(cudaStreamSynchronize(st));
    return QC_OK;
}

extern "C" int32_t qc_record_state(qc_plan *pl, int32_t buffer, int32_t index, int32_t level) {
    if (!pl) return fail(QC_ERR_ARG, "qc_record_state: NULL plan");
    if (buffer < 0 || buffer > 1 || !pl->d.store_traj || index < 0 || index > pl->d.nt_max)
        return fail(QC_ERR_ARG, "qc_record_state: buffer=%d index=%d", buffer, index);
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const qc_plan_desc &d = pl->d;
    if (pl->grid) {
        if (level < 0 || level > 1) return fail(QC_ERR_ARG, "qc_record_state: level=%d", level);
        record_grid_kernel<<<d.batch, 256, 0, pl->ctx->stream>>>(pl->psi.p, pl->traj[buffer].p,
                                                                (long long)(d.nt_max + 1) * d.np,
                                                                (long long)index * d.np, level, d.np);
    } else {
        cd *slice = pl->traj[buffer].p + (size_t)index * d.nlevs * pl->nthreads;
        record_nlevel_kernel<<<(unsigned)((pl->nthreads + 127) / 128), 128, 0, pl->ctx->stream>>>(
            pl->psi.p, slice, d.nb, pl->nbp, d.nlevs, pl->nthreads);
    }
    pl->ctx->launches++;
    QC_CUDA(cudaGetLastError());
    return QC_OK;
}

extern "C" int32_t qc_propagate_host(qc_plan *pl, const qc_sweep_args *args, const double *psi_in, double *psi_out) {
    if (!pl || !args || !psi_in || !psi_out) return fail(QC_ERR_ARG, "qc_propagate_host: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    qc_context *ctx = pl->ctx;
    cudaStream_t st = ctx->stream;
    const qc_plan_desc &d = pl->d;
    // np = 2048: 4 waves of 148 CTAs; np = 256 in the packed form: one wave is 148 SMs x 2 CTAs x 4 runs
    int chunk = pl->grid && pl->log2n == 8 ? 1184 : 592;
    if (const char *e = getenv("QC_E2E_CHUNK")) chunk = atoi(e) > 0 ? atoi(e) : chunk;
    if (!pl->grid || d.batch <= chunk || args->field_mode != QC_FIELD_PRESCRIBED) {
        QC_CUDA(cudaMemcpyAsync(pl->psi.p, psi_in, pl->state_elems * sizeof(cd), cudaMemcpyHostToDevice, st));
        int32_t rc = run_sweep(pl, args, false);
        if (rc) return rc;
        QC_CUDA(cudaMemcpyAsync(psi_out, pl->psi.p, pl->state_elems * sizeof(cd), cudaMemcpyDeviceToHost, st));
        QC_CUDA(cudaStreamSynchronize(st));
        return QC_OK;
    }
    // Runs are independent: cut the batch into chunks and pipeline H2D(i+1) / kernel(i) / D2H(i-1) over
    // three streams so that the PCIe copies hide behind the propagation.
    if (!ctx->aux_ready) {
        for (auto &a : ctx->aux) QC_CUDA(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        for (auto &e : ctx->aux_ev) QC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->aux_ready = true;
    }
    cudaStream_t streams[3] = {st, ctx->aux[0], ctx->aux[1]};
    QC_CUDA(cudaEventRecord(ctx->aux_ev[0], st));
    QC_CUDA(cudaStreamWaitEvent(ctx->aux[0], ctx->aux_ev[0], 0));
    QC_CUDA(cudaStreamWaitEvent(ctx->aux[1], ctx->aux_ev[0], 0));
    const size_t per_run = (size_t)d.nb * d.nlevs * d.np;
    int k = 0;
    for (int first = 0; first < d.batch; first += chunk, ++k) {
        const int count = d.batch - first < chunk ? d.batch - first : chunk;
        cudaStream_t s = streams[k % 3];
        const size_t off = (size_t)first * per_run;
        QC_CUDA(cudaMemcpyAsync(pl->psi.p + off, reinterpret_cast<const cd *>(psi_in) + off, count * per_run * sizeof(cd),
                                cudaMemcpyHostToDevice, s));
        ApplyOpts ap;
        ap.run_first = first;
        ap.run_count = count;
        ap.stream = s;
        int32_t rc = run_sweep(pl, args, false, ap);
        if (rc) return rc;
        QC_CUDA(cudaMemcpyAsync(reinterpret_cast<cd *>(psi_out) + off, pl->psi.p + off, count * per_run * sizeof(cd),
                                cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < 2; ++i) {
        QC_CUDA(cudaEventRecord(ctx->aux_ev[1 + i], ctx->aux[i]));
        QC_CUDA(cudaStreamWaitEvent(st, ctx->aux_ev[1 + i], 0));
    }
    QC_CUDA(cudaStreamSynchronize(st));
    return QC_OK;
}


// ---------------------------------------------------------------------------
// two-dimensional grids (kernel C, qc_grid2d.cuh)
// ---------------------------------------------------------------------------
struct qc_plan2d {
    qc_context *ctx;
    int n, nch, batch;
    int v_batched = 0;
    bool static_set = false, poly_set = false;
    double emin = 0, emax = 0, t_sc = 0, dxdy = 0;
    DevBuf<cd> psi, phi, tmp, dv, tw1, tw2;
    DevBuf<double> v, kinx, kiny, xp, norm_acc, norms;
};
static const int kWavesPerLaunch = 2;        // 2 x 64 CTAs of a cooperative launch on 148 SMs

extern "C" int32_t qc_plan2d_create(qc_context *ctx, int32_t n, int32_t nch, int32_t batch, qc_plan2d **out) {
    if (!ctx || !out) return fail(QC_ERR_ARG, "qc_plan2d_create: NULL argument");
    if (n != 512) return fail(QC_ERR_ARG, "qc_plan2d_create: the 2-D stepper is built for 512 x 512 grids, got n=%d", n);
    if (!is_pow2(nch) || nch < 2 || nch > 128) return fail(QC_ERR_ARG, "qc_plan2d_create: nch=%d must be a power of two in [2,128]", nch);
    if (batch < 1) return fail(QC_ERR_ARG, "qc_plan2d_create: batch=%d", batch);
    QC_CUDA(cudaSetDevice(ctx->device));
    int coop = 0;
    QC_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!coop) return fail(QC_ERR_CUDA, "qc_plan2d_create: the device does not support cooperative launches");
    qc_plan2d *pl = new qc_plan2d();
    pl->ctx = ctx;
    pl->n = n; pl->nch = nch; pl->batch = batch;
    const size_t nn = (size_t)n * n;
    const int waves = batch < kWavesPerLaunch ? batch : kWavesPerLaunch;
    cudaError_t e = cudaSuccess;
    auto chk = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    chk(pl->psi.alloc(nn * batch));
    chk(pl->phi.alloc(nn * waves));
    chk(pl->tmp.alloc(nn * waves));
    chk(pl->dv.alloc(nch));
    chk(pl->xp.alloc(nch));
    chk(pl->kinx.alloc(n));
    chk(pl->kiny.alloc(n));
    chk(pl->norms.alloc(batch));
    std::vector<cd> t1, t2;
    fill_twiddles(qc::G2_LOG2N, t1, t2);
    chk(pl->tw1.alloc(t1.size()));
    chk(pl->tw2.alloc(t2.size()));
    if (e == cudaSuccess) chk(cudaMemcpy(pl->tw1.p, t1.data(), t1.size() * sizeof(cd), cudaMemcpyHostToDevice));
    if (e == cudaSuccess) chk(cudaMemcpy(pl->tw2.p, t2.data(), t2.size() * sizeof(cd), cudaMemcpyHostToDevice));
    if (e == cudaSuccess) chk(cudaMemset(pl->norms.p, 0, batch * sizeof(double)));
    if (e != cudaSuccess) {
        delete pl;
        return fail(QC_ERR_CUDA, "qc_plan2d_create: %s", cudaGetErrorString(e));
    }
    ctx->live_plans++;
    *out = pl;
    return QC_OK;
}

extern "C" int32_t qc_plan2d_destroy(qc_plan2d *pl) {
    if (!pl) return QC_OK;
    cudaSetDevice(pl->ctx->device);
    cudaStreamSynchronize(pl->ctx->stream);
    pl->psi.release(); pl->phi.release(); pl->tmp.release(); pl->dv.release(); pl->tw1.release(); pl->tw2.release();
    pl->v.release(); pl->kinx.release(); pl->kiny.release(); pl->xp.release(); pl->norm_acc.release(); pl->norms.release();
    qc_context *ctx = pl->ctx;
    delete pl;
    if (--ctx->live_plans == 0 && ctx->released) context_free(ctx);
    return QC_OK;
}

extern "C" int32_t qc_plan2d_set_static(qc_plan2d *pl, const double *v, int32_t v_batched, const double *akx2_re,
                                        const double *aky2_re, double dxdy) {
    if (!pl || !v || !akx2_re || !aky2_re) return fail(QC_ERR_ARG, "qc_plan2d_set_static: NULL argument");
    if (!(dxdy > 0.0)) return fail(QC_ERR_ARG, "qc_plan2d_set_static: dxdy=%g", dxdy);
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    const size_t nn = (size_t)pl->n * pl->n;
    const size_t count = nn * (v_batched ? pl->batch : 1);
    if (pl->v.n != count) QC_CUDA(pl->v.alloc(count));
    QC_CUDA(cudaMemcpyAsync(pl->v.p, v, count * sizeof(double), cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaMemcpyAsync(pl->kinx.p, akx2_re, pl->n * sizeof(double), cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaMemcpyAsync(pl->kiny.p, aky2_re, pl->n * sizeof(double), cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    pl->v_batched = v_batched ? 1 : 0;
    pl->dxdy = dxdy;
    pl->static_set = true;
    return QC_OK;
}

extern "C" int32_t qc_plan2d_set_poly(qc_plan2d *pl, const double *xp, const double *dv, const double *sc) {
    if (!pl || !xp || !dv || !sc) return fail(QC_ERR_ARG, "qc_plan2d_set_poly: NULL argument");
    if (!(sc[1] > sc[0])) return fail(QC_ERR_ARG, "qc_plan2d_set_poly: emax (%g) must exceed emin (%g)", sc[1], sc[0]);
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(pl->xp.p, xp, pl->nch * sizeof(double), cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaMemcpyAsync(pl->dv.p, dv, pl->nch * sizeof(cd), cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    pl->emin = sc[0]; pl->emax = sc[1]; pl->t_sc = sc[2];
    pl->poly_set = true;
    return QC_OK;
}

extern "C" int32_t qc_plan2d_set_state(qc_plan2d *pl, const double *psi) {
    if (!pl || !psi) return fail(QC_ERR_ARG, "qc_plan2d_set_state: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(pl->psi.p, psi, pl->psi.n * sizeof(cd), cudaMemcpyHostToDevice, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_plan2d_get_state(qc_plan2d *pl, double *psi) {
    if (!pl || !psi) return fail(QC_ERR_ARG, "qc_plan2d_get_state: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(psi, pl->psi.p, pl->psi.n * sizeof(cd), cudaMemcpyDeviceToHost, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_plan2d_get_norms(qc_plan2d *pl, double *norms) {
    if (!pl || !norms) return fail(QC_ERR_ARG, "qc_plan2d_get_norms: NULL argument");
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    QC_CUDA(cudaMemcpyAsync(norms, pl->norms.p, pl->batch * sizeof(double), cudaMemcpyDeviceToHost, pl->ctx->stream));
    QC_CUDA(cudaStreamSynchronize(pl->ctx->stream));
    return QC_OK;
}

extern "C" int32_t qc_plan2d_sweep(qc_plan2d *pl, int32_t nsteps, int32_t renorm) {
    if (!pl) return fail(QC_ERR_ARG, "qc_plan2d_sweep: NULL argument");
    if (!pl->static_set || !pl->poly_set) return fail(QC_ERR_STATE, "qc_plan2d_sweep: qc_plan2d_set_static / qc_plan2d_set_poly first");
    if (nsteps < 1) return fail(QC_ERR_ARG, "qc_plan2d_sweep: nsteps=%d", nsteps);
    QC_CUDA(cudaSetDevice(pl->ctx->device));
    typedef qc::Grid2DCfg C;
    static bool attr_done[64] = {false};
    if (!attr_done[pl->ctx->device & 63]) {
        QC_CUDA(cudaFuncSetAttribute(qc::grid2d_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        attr_done[pl->ctx->device & 63] = true;
    }
    const size_t acc = (size_t)nsteps * kWavesPerLaunch * C::TILES;
    if (pl->norm_acc.n < acc) QC_CUDA(pl->norm_acc.alloc(acc));
    for (int first = 0; first < pl->batch; first += kWavesPerLaunch) {
        const int count = pl->batch - first < kWavesPerLaunch ? pl->batch - first : kWavesPerLaunch;
        qc::Grid2DParams g;
        memset(&g, 0, sizeof g);
        g.n = pl->n; g.nch = pl->nch; g.nsteps = nsteps; g.renorm = renorm;
        g.wave_first = first; g.wave_count = count;
        g.psi = pl->psi.p; g.phi = pl->phi.p; g.tmp = pl->tmp.p;
        g.v = pl->v.p; g.v_stride = pl->v_batched ? (long long)pl->n * pl->n : 0;
        g.kinx = pl->kinx.p; g.kiny = pl->kiny.p; g.tw1 = pl->tw1.p; g.tw2 = pl->tw2.p;
        g.xp = pl->xp.p; g.dv = pl->dv.p; g.emin = pl->emin; g.emax = pl->emax; g.t_sc = pl->t_sc; g.dxdy = pl->dxdy;
        g.norm_acc = pl->norm_acc.p; g.norms = pl->norms.p;
        void *args[] = {&g};
        QC_CUDA(cudaLaunchCooperativeKernel((const void *)qc::grid2d_sweep_kernel, dim3(C::TILES * count), dim3(C::NTHREADS),
                                            args, C::SMEM_BYTES, pl->ctx->stream));
        pl->ctx->launches++;
    }
    QC_CUDA(cudaGetLastError());
    return QC_OK;
}
