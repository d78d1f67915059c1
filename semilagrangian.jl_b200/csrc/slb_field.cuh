// slb_field.cuh -- K4 (charge density), K5 (Fourier Poisson solve) and the deterministic
// reductions behind compute_ee / compute_ke.
//
// Reference: src/util_poisson.jl:68-79 (compute_charge!), src/poisson.jl:7-15,139-144
// (compute_elfield!), src/util_poisson.jl:41-53,156-162 (compute_ke / compute_ee).
// All reductions are atomics-free and run in a fixed order, so electric-energy histories
// are reproducible run to run ("hard parts").
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

// ------------------------------------------------------------------------------------------
// K4: rho[a] = sum_b f[a + nsp*b].  f is [nsp, nv]; lanes run along the contiguous space
// index a (coalesced 256 B rows), the 8 warps of a block and the grid's y dimension split
// the velocity range; each thread keeps 4 independent accumulators, the 8 warps combine
// through shared memory in a fixed order, and chunk partials are summed by k_charge_final.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_charge_partial(const double* __restrict__ f, long long nsp, long long nv, long long chunk,
                 double* __restrict__ partial)
{
    __shared__ double sm[8][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    long long a = (long long)blockIdx.x * 32 + tx;
    long long b0 = (long long)blockIdx.y * chunk;
    long long b1 = b0 + chunk < nv ? b0 + chunk : nv;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    if (a < nsp) {
        const double* p = f + a;
        long long b = b0 + ty;
        for (; b + 24 < b1; b += 32) {
            acc0 += __ldg(p + nsp * b);
            acc1 += __ldg(p + nsp * (b + 8));
            acc2 += __ldg(p + nsp * (b + 16));
            acc3 += __ldg(p + nsp * (b + 24));
        }
        for (; b < b1; b += 8) acc0 += __ldg(p + nsp * b);
    }
    sm[ty][tx] = (acc0 + acc1) + (acc2 + acc3);
    __syncthreads();
    if (ty == 0 && a < nsp) {
        double s = sm[0][tx];
#pragma unroll
        for (int i = 1; i < 8; ++i) s += sm[i][tx];
        partial[(long long)blockIdx.y * nsp + a] = s;
    }
}

__global__ void k_charge_final(const double* __restrict__ partial, long long nsp, int nchunk, double dv,
                               double* __restrict__ rho)
{
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nsp) return;
    double s = 0.0;
    for (int c = 0; c < nchunk; ++c) s += partial[(long long)c * nsp + a];
    rho[a] = dv * s;
}

// ------------------------------------------------------------------------------------------
// deterministic reductions: MODE 0 sum(x), 1 sum(x^2)
// ------------------------------------------------------------------------------------------
#include "slb_devfn.cuh"

template <int MODE>
__global__ void __launch_bounds__(256) k_reduce_partial(const double* __restrict__ x, long long n,
                                                        double* __restrict__ partial)
{
    __shared__ double sm[32];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        double v = __ldg(x + i);
        acc += (MODE == 1) ? v * v : v;
    }
    double r = slb_block_reduce(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// out[0] = scale * sum(partial[0..m))
__global__ void __launch_bounds__(256) k_reduce_final(const double* __restrict__ partial, int m, double scale,
                                                      double* __restrict__ out)
{
    __shared__ double sm[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += 256) acc += partial[i];
    double r = slb_block_reduce(acc, sm);
    if (threadIdx.x == 0) out[0] = scale * r;
}

__global__ void k_sub_scalar(double* __restrict__ x, long long n, const double* __restrict__ s)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] -= s[0];
}

// ------------------------------------------------------------------------------------------
// K5: direct DFT along one dim of the (small) space grid: <= 256^2 / 1024 points per dim.
//   out(a,k,b) = sum_j in(a,j,b) * tw[(j*k) mod n],   tw[m] = exp(-+2 pi i m/n)
// Convention of src/fftbig.jl:162-176 (FFTW): forward exp(-i..) unnormalised; inverse 1/n.
// One block per line: the line and the twiddle table are staged in shared memory, thread k
// accumulates output k.  Fusions: MULT applies fctv_k = i*m (src/poisson.jl:8,14) to the input
// on the fly; REAL_OUT stores only the real part (the E field, src/poisson.jl:142).
// The first version (one thread per output reading global memory) took 16-52 us per pass on
// a 128^2 grid under ncu: 8% of a Strang step.
// ------------------------------------------------------------------------------------------
template <bool REAL_IN, bool INVERSE, bool MULT, bool REAL_OUT>
__global__ void __launch_bounds__(256)
k_dft_line(const void* __restrict__ in_, void* __restrict__ out_, long long inner, int n,
           const double2* __restrict__ tw, const double* __restrict__ mult)
{
    extern __shared__ double2 dsm[];
    double2* line = dsm;
    double2* twd = dsm + n;
    const long long ln = blockIdx.x;
    const long long a = ln % inner, b = ln / inner;
    const long long base = a + inner * (long long)n * b;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double2 t = __ldg(tw + j);
        if (INVERSE) t.y = -t.y;
        twd[j] = t;
        double2 v;
        if (REAL_IN)
            v = make_double2(__ldg(reinterpret_cast<const double*>(in_) + base + inner * j), 0.0);
        else
            v = __ldg(reinterpret_cast<const double2*>(in_) + base + inner * j);
        if (MULT) {
            double mm = __ldg(mult + base + inner * j);
            v = make_double2(-v.y * mm, v.x * mm);
        }
        line[j] = v;
    }
    __syncthreads();
    const double scale = INVERSE ? 1.0 / (double)n : 1.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        double sr = 0.0, si = 0.0;
        int m = 0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            double2 t = twd[m];
            double2 v = line[j];
            sr = fma(v.x, t.x, sr);
            sr = fma(-v.y, t.y, sr);
            si = fma(v.x, t.y, si);
            si = fma(v.y, t.x, si);
            m += k;
            if (m >= n) m -= n;
        }
        long long o = base + inner * k;
        if (REAL_OUT)
            reinterpret_cast<double*>(out_)[o] = sr * scale;
        else
            reinterpret_cast<double2*>(out_)[o] = make_double2(sr * scale, si * scale);
    }
}

// ------------------------------------------------------------------------------------------
// compute_ke: partial[v] = vsq[v] * sum_{a < nsp} f[a + nsp*v]   (one block per v)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ke_partial(const double* __restrict__ f, long long nsp,
                                                    const double* __restrict__ vsq, double* __restrict__ partial)
{
    __shared__ double sm[32];
    const double* p = f + nsp * (long long)blockIdx.x;
    double acc = 0.0;
    for (long long a = threadIdx.x; a < nsp; a += 256) acc += __ldg(p + a);
    double r = slb_block_reduce(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = vsq[blockIdx.x] * r;
}

// ------------------------------------------------------------------------------------------
// K5f: the whole field solve of one initcoef! (src/poisson.jl:171-176) in ONE cooperative
// kernel: rho = scale * sum_c partial[c] (the per-chunk sums of K4, or line sums, or a gathered
// rho), mean subtraction (src/util_poisson.jl:77), forward DFTs over every space dim, and for
// each component the multiplier i*k_x/|k|^2 + inverse DFTs + real part (src/poisson.jl:139-144).
// The 11 small launches this replaces cost ~85 us per solve at 128^2 (launch gaps included),
// 6 % of a fused Strang step on one GPU and 15 % on eight.  Phases are separated by grid-wide
// barriers; every reduction runs in a fixed order, so histories stay reproducible.
// Lines are distributed block-cyclically; line + twiddles sit in shared memory (2 n double2).
// ------------------------------------------------------------------------------------------
#define SLB_FIELD_MAXDIM 3
struct FieldArgs {
    const double* partial;   // [nchunk][ntot]
    int nchunk;
    double scale;
    int subtract_mean;
    int nsp;
    int ext[SLB_FIELD_MAXDIM];
    long long ntot;
    const double2* tw[SLB_FIELD_MAXDIM];    // forward twiddles exp(-2 pi i m / n) per dim
    const double* mult[SLB_FIELD_MAXDIM];   // imag part of fctv_k per component
    double* rho;                            // out: charge density (after mean subtraction)
    double* E[SLB_FIELD_MAXDIM];            // out: field components
    double2* wa;                            // work: spectrum ping
    double2* wb;                            // work: spectrum pong
    double2* wc[SLB_FIELD_MAXDIM];          // work: one buffer per component (inverse passes)
    double2* wd[SLB_FIELD_MAXDIM];
    double* red;                            // work: gridDim.x block sums
};

// one DFT line: out(k) = sum_j in(j) * tw^(jk) (conjugated when INVERSE, scaled by 1/n)
template <bool REAL_IN, bool INVERSE, bool MULT, bool REAL_OUT>
__device__ __forceinline__ void field_dft_line(const void* in_, void* out_, long long base, long long inner, int n,
                                               const double2* __restrict__ tw, const double* __restrict__ mult, double shift,
                                               double2* line, double2* twd)
{
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double2 t = tw[j];
        if (INVERSE) t.y = -t.y;
        twd[j] = t;
        double2 v;
        if (REAL_IN)
            v = make_double2(reinterpret_cast<const double*>(in_)[base + inner * j] - shift, 0.0);
        else
            v = reinterpret_cast<const double2*>(in_)[base + inner * j];
        if (MULT) {
            const double mm = mult[base + inner * j];
            v = make_double2(-v.y * mm, v.x * mm);
        }
        line[j] = v;
    }
    __syncthreads();
    const double scale = INVERSE ? 1.0 / (double)n : 1.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        // four interleaved partial sums (j mod 4): the dependent FMA chain is 4x shorter; the order of
        // the additions is fixed, so results are reproducible
        double ar[4] = {0.0, 0.0, 0.0, 0.0}, ai[4] = {0.0, 0.0, 0.0, 0.0};
        int mq[4];
        const int k4 = (int)(((long long)4 * k) % n);
#pragma unroll
        for (int q = 0; q < 4; ++q) mq[q] = (int)(((long long)q * k) % n);
        int j = 0;
        for (; j + 4 <= n; j += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 t = twd[mq[q]];
                const double2 v = line[j + q];
                ar[q] = fma(v.x, t.x, ar[q]);
                ar[q] = fma(-v.y, t.y, ar[q]);
                ai[q] = fma(v.x, t.y, ai[q]);
                ai[q] = fma(v.y, t.x, ai[q]);
                mq[q] += k4;
                if (mq[q] >= n) mq[q] -= n;
            }
        }
        for (int q = 0; j < n; ++j, ++q) {  // n % 4 leftover elements
            const double2 t = twd[mq[q]];
            const double2 v = line[j];
            ar[q] = fma(v.x, t.x, ar[q]);
            ar[q] = fma(-v.y, t.y, ar[q]);
            ai[q] = fma(v.x, t.y, ai[q]);
            ai[q] = fma(v.y, t.x, ai[q]);
        }
        const double sr = (ar[0] + ar[1]) + (ar[2] + ar[3]);
        const double si = (ai[0] + ai[1]) + (ai[2] + ai[3]);
        const long long o = base + inner * k;
        if (REAL_OUT)
            reinterpret_cast<double*>(out_)[o] = sr * scale;
        else
            reinterpret_cast<double2*>(out_)[o] = make_double2(sr * scale, si * scale);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128) k_field_solve(const __grid_constant__ FieldArgs fa)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double2 fsm2[];
    __shared__ double redsm[32];
    __shared__ double mean_sm;
    const long long ntot = fa.ntot;
    // ---- phase A: rho = scale * sum of the partial sums; block sums for the mean --------------
    {
        double acc = 0.0;
        for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < ntot; a += (long long)gridDim.x * blockDim.x) {
            double s = 0.0;
            for (int c = 0; c < fa.nchunk; ++c) s += fa.partial[(long long)c * ntot + a];
            s *= fa.scale;
            fa.rho[a] = s;
            acc += s;
        }
        const double r = slb_block_reduce(acc, redsm);
        if (threadIdx.x == 0) fa.red[blockIdx.x] = r;
    }
    grid.sync();
    // ---- mean (every block adds the block sums in the same order) ------------------------------
    if (threadIdx.x == 0) {
        double m = 0.0;
        if (fa.subtract_mean) {
            for (unsigned b = 0; b < gridDim.x; ++b) m += fa.red[b];
            m /= (double)ntot;
        }
        mean_sm = m;
    }
    __syncthreads();
    const double mean = mean_sm;
    // ---- forward transforms ------------------------------------------------------------------------
    const int nsp = fa.nsp;
    double2* cur = fa.wa;
    double2* nxt = fa.wb;
    {
        const int n = fa.ext[0];
        const long long nlines = ntot / n;
        for (long long ln = blockIdx.x; ln < nlines; ln += gridDim.x)  // reads the raw rho, subtracts the mean on the fly
            field_dft_line<true, false, false, false>(fa.rho, cur, (long long)n * ln, 1, n, fa.tw[0], nullptr, mean, fsm2, fsm2 + n);
    }
    grid.sync();
    if (fa.subtract_mean)  // nobody reads the raw rho any more: store the mean-free charge density
        for (long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x; a < ntot; a += (long long)gridDim.x * blockDim.x)
            fa.rho[a] -= mean;
    long long inner = fa.ext[0];
    for (int d = 1; d < nsp; ++d) {
        const int n = fa.ext[d];
        const long long nlines = ntot / n;
        for (long long ln = blockIdx.x; ln < nlines; ln += gridDim.x) {
            const long long a = ln % inner, b = ln / inner;
            field_dft_line<false, false, false, false>(cur, nxt, a + inner * (long long)n * b, inner, n, fa.tw[d], nullptr, 0.0, fsm2, fsm2 + n);
        }
        grid.sync();
        double2* t = cur;
        cur = nxt;
        nxt = t;
        inner *= n;
    }
    // ---- inverse transforms, all components in the same phases ---------------------------------------
    // pass over dim 0 applies the multiplier; the last pass keeps the real part
    {
        const int n = fa.ext[0];
        const long long nlines = ntot / n;
        for (long long w = blockIdx.x; w < nlines * nsp; w += gridDim.x) {
            const int x = (int)(w / nlines);
            const long long ln = w % nlines;
            if (nsp == 1)
                field_dft_line<false, true, true, true>(cur, fa.E[x], (long long)n * ln, 1, n, fa.tw[0], fa.mult[x], 0.0, fsm2, fsm2 + n);
            else
                field_dft_line<false, true, true, false>(cur, fa.wc[x], (long long)n * ln, 1, n, fa.tw[0], fa.mult[x], 0.0, fsm2, fsm2 + n);
        }
    }
    inner = fa.ext[0];
    for (int d = 1; d < nsp; ++d) {
        grid.sync();
        const int n = fa.ext[d];
        const long long nlines = ntot / n;
        const bool last = d == nsp - 1;
        for (long long w = blockIdx.x; w < nlines * nsp; w += gridDim.x) {
            const int x = (int)(w / nlines);
            const long long ln = w % nlines;
            const long long a = ln % inner, b = ln / inner;
            const long long base = a + inner * (long long)n * b;
            const double2* src = (d % 2 == 1) ? fa.wc[x] : fa.wd[x];
            if (last)
                field_dft_line<false, true, false, true>(src, fa.E[x], base, inner, n, fa.tw[d], nullptr, 0.0, fsm2, fsm2 + n);
            else
                field_dft_line<false, true, false, false>(src, (d % 2 == 1) ? fa.wd[x] : fa.wc[x], base, inner, n, fa.tw[d], nullptr, 0.0, fsm2,
                                                          fsm2 + n);
        }
        inner *= n;
    }
}


// ------------------------------------------------------------------------------------------
// K5c: the same field solve for power-of-two space grids up to 256 points per dim, as radix-2 FFTs: a small
// cooperative grid (32 blocks x 8 warps, one line per warp in shared memory) with two grid barriers instead of five
// around O(n^2) line DFTs.  (A first version ran inside one 8-block cluster: 96 transforms per SM made it
// shared-memory-bandwidth bound, 35-41 us; the work is tiny but wants to be spread.)  Two solves per Strang step:
// 4 % of the step on one GPU, 15 % when the grid is sharded over eight.
//   phase 0: rho = scale * sum_c partial[c]  (stored raw), forward FFT of every x1 line          -> wa
//   phase 1: per x1 wavenumber: forward FFT along x2 (the spectrum, in shared memory; its (0,0) entry / N is the
//            mean of src/util_poisson.jl:77), then per component the multiplier i k_x/|k|^2 (fctv_k,
//            src/poisson.jl:7-15; its zero mode is 0, so the mean never reaches E) and the inverse FFT along x2 -> wc[x]
//   phase 2: per component the inverse FFT of every x1 line, real part -> E_x; rho -= mean
// One space dim (1D1V): a single warp does forward FFT, multiplier and inverse FFT of the one line.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SLB_FFT_THREADS) k_field_fft(const __grid_constant__ FieldFftArgs fa)
{
    namespace cg = cooperative_groups;
    extern __shared__ double2 fsm2[];
    const int n1 = fa.n1, n2 = fa.n2, l1 = fa.l1, l2 = fa.l2;
    const int nmax = n1 > n2 ? n1 : n2;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double2* tw1 = fsm2;
    double2* tw2 = fsm2 + n1;
    double2* xa = fsm2 + n1 + n2 + (size_t)(2 * w) * nmax;  // two lines per warp
    double2* xb = xa + nmax;
    field_fill_tws(tw1, fa.tw1, n1, threadIdx.x, blockDim.x);
    if (fa.nsp == 2) field_fill_tws(tw2, fa.tw2, n2, threadIdx.x, blockDim.x);
    __syncthreads();
    const int gw = blockIdx.x * nw + w, W = gridDim.x * nw;
    const long long ntot = (long long)n1 * n2;
    if (fa.nsp != 2) {  // one space dim: the whole solve by one warp (shared with the step-program interpreter)
        if (gw == 0) field_fft_1d_warp(fa, tw1, xa, xb, lane);
        return;
    }
    // ---- phase 0: rho, forward transform of the x1 lines (spectrum index in bit-reversed position) -----------------
    for (int j = gw; j < n2; j += W) {
        const long long base = (long long)n1 * j;
        for (int a = lane; a < n1; a += 32) {
            double s = 0.0;
            for (int c = 0; c < fa.nchunk; ++c) s += fa.partial[(long long)c * ntot + base + a];
            s *= fa.scale;
            fa.rho[base + a] = s;
            xa[a] = make_double2(s, 0.0);
        }
        field_warp_fft<false>(xa, n1, l1, tw1, lane);
        for (int p = lane; p < n1; p += 32) fa.wa[base + p] = xa[p];
    }
    cg::this_grid().sync();
    // ---- phase 1: columns (position p1 holds wavenumber k1 = brev(p1)) -----------------------------------------------
    for (int p1 = gw; p1 < n1; p1 += W) {
        const int k1 = field_brev(p1, l1);
        for (int j = lane; j < n2; j += 32) xa[j] = fa.wa[p1 + (long long)n1 * j];
        field_warp_fft<false>(xa, n2, l2, tw2, lane);
        if (p1 == 0 && lane == 0) *fa.mean = fa.subtract_mean ? xa[0].x / (double)ntot : 0.0;
        const double sc = 1.0 / (double)n2;
        for (int x = 0; x < 2; ++x) {
            __syncwarp();
            for (int p2 = lane; p2 < n2; p2 += 32) {
                const double2 v = xa[p2];
                const double mm = fa.mult[x][k1 + (long long)n1 * field_brev(p2, l2)];
                xb[p2] = make_double2(-v.y * mm, v.x * mm);
            }
            field_warp_fft<true>(xb, n2, l2, tw2, lane);
            for (int j = lane; j < n2; j += 32) {
                const double2 v = xb[j];
                fa.wc[x][p1 + (long long)n1 * j] = make_double2(v.x * sc, v.y * sc);
            }
        }
    }
    cg::this_grid().sync();
    // ---- phase 2: inverse x1 lines of both components (input already in bit-reversed order); mean removal ------------
    const double mean = *fa.mean;
    const double sc1 = 1.0 / (double)n1;
    for (int t = gw; t < 2 * n2; t += W) {
        const int x = t / n2, j = t - x * n2;
        const long long base = (long long)n1 * j;
        __syncwarp();
        for (int p = lane; p < n1; p += 32) xa[p] = fa.wc[x][base + p];
        field_warp_fft<true>(xa, n1, l1, tw1, lane);
        for (int a = lane; a < n1; a += 32) {
            fa.E[x][base + a] = xa[a].x * sc1;
            if (x == 0 && fa.subtract_mean) fa.rho[base + a] -= mean;
        }
    }
}
