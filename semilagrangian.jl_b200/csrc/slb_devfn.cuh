// slb_devfn.cuh -- device functions shared by several translation units (all __forceinline__: no external symbols):
// the fixed-order block reduction, the warp-level radix-2 FFTs of the field solve (see slb_field.cuh, K5c) and the
// one-space-dim field solve done by a single warp (used by k_field_fft and by the step-program interpreter,
// slb_program.cuh).
#pragma once
#include <cuda_runtime.h>

__device__ __forceinline__ double slb_block_reduce(double v, double* sm)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        int nw = (blockDim.x + 31) >> 5;
        r = lane < nw ? sm[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    return r;  // valid in thread 0
}

#define SLB_FFT_BLOCKS 32     // cooperative grid: 32 blocks x 8 warps = one line per warp in the widest phase at 128^2
#define SLB_FFT_THREADS 256
#define SLB_FFT_NMAX 256
struct FieldFftArgs {
    const double* partial;   // [nchunk][n1*n2]
    int nchunk;
    double scale;
    int subtract_mean;
    int nsp, n1, n2, l1, l2;  // n2 = 1, l2 = 0 for one space dim
    const double2* tw1;       // exp(-2 pi i m / n1)
    const double2* tw2;
    const double* mult[2];
    double* rho;
    double* E[2];
    double2* wa;
    double2* wc[2];
    double* mean;             // one double of work space
};

// In-place radix-2 FFTs of a line held in shared memory, by one warp, without any bit-reversal pass:
//   forward  = decimation in frequency : natural order in  -> BIT-REVERSED order out   (stages n/2 ... 1)
//   inverse  = decimation in time      : bit-reversed in   -> natural order out        (stages 1 ... n/2)
// so the spectrum simply stays in bit-reversed order between the two (the multiplier is looked up at the reversed
// index).  Twiddles come from a per-stage table tws[half + k] = exp(-2 pi i k / (2 half)), k < half, so that the
// lanes of a stage read consecutive entries (a stride-n/m walk through one table is an 8-way bank conflict; that
// and the bit-reversed scatter were half of all shared-memory wavefronts of the first version of this kernel).
// LOGN > 5: line length known at compile time, the n / 64 butterflies of a lane unrolled.
template <bool INVERSE, int LOGN>
__device__ __forceinline__ void field_warp_fft_t(double2* x, int n_rt, int logn_rt, const double2* tws, int lane)
{
    const int logn = LOGN > 0 ? LOGN : logn_rt;
    const int n = LOGN > 0 ? (1 << LOGN) : n_rt;
    constexpr int NBF = LOGN > 5 ? (1 << (LOGN - 6)) : 1;  // butterflies per lane and stage (compile-time form)
#pragma unroll
    for (int st = 1; st <= (LOGN > 0 ? LOGN : 30); ++st) {
        if (LOGN == 0 && st > logn) break;
        const int s = INVERSE ? st : logn + 1 - st;
        const int half = 1 << (s - 1);
        __syncwarp();
        if (LOGN > 5) {
            double2 u[NBF], v[NBF], w[NBF];
            int i0[NBF];
#pragma unroll
            for (int q = 0; q < NBF; ++q) {
                const int b = lane + 32 * q;
                const int k = b & (half - 1);
                i0[q] = ((b >> (s - 1)) << s) + k;
                w[q] = tws[half + k];
                u[q] = x[i0[q]];
                v[q] = x[i0[q] + half];
            }
#pragma unroll
            for (int q = 0; q < NBF; ++q) {
                if (INVERSE) {
                    const double tr = fma(v[q].x, w[q].x, v[q].y * w[q].y), ti = fma(v[q].y, w[q].x, -v[q].x * w[q].y);  // conj(w) v
                    x[i0[q]] = make_double2(u[q].x + tr, u[q].y + ti);
                    x[i0[q] + half] = make_double2(u[q].x - tr, u[q].y - ti);
                } else {
                    const double dr = u[q].x - v[q].x, di = u[q].y - v[q].y;
                    x[i0[q]] = make_double2(u[q].x + v[q].x, u[q].y + v[q].y);
                    x[i0[q] + half] = make_double2(fma(dr, w[q].x, -di * w[q].y), fma(dr, w[q].y, di * w[q].x));
                }
            }
        } else {
            for (int b = lane; b < (n >> 1); b += 32) {
                const int k = b & (half - 1);
                const int i0 = ((b >> (s - 1)) << s) + k;
                const int i1 = i0 + half;
                const double2 w = tws[half + k];
                const double2 u = x[i0], v = x[i1];
                if (INVERSE) {
                    const double tr = fma(v.x, w.x, v.y * w.y), ti = fma(v.y, w.x, -v.x * w.y);
                    x[i0] = make_double2(u.x + tr, u.y + ti);
                    x[i1] = make_double2(u.x - tr, u.y - ti);
                } else {
                    const double dr = u.x - v.x, di = u.y - v.y;
                    x[i0] = make_double2(u.x + v.x, u.y + v.y);
                    x[i1] = make_double2(fma(dr, w.x, -di * w.y), fma(dr, w.y, di * w.x));
                }
            }
        }
    }
    __syncwarp();
}

template <bool INVERSE>
__device__ __forceinline__ void field_warp_fft(double2* x, int n, int logn, const double2* tws, int lane)
{
    switch (logn) {  // block-uniform
    case 6: field_warp_fft_t<INVERSE, 6>(x, n, logn, tws, lane); break;
    case 7: field_warp_fft_t<INVERSE, 7>(x, n, logn, tws, lane); break;
    case 8: field_warp_fft_t<INVERSE, 8>(x, n, logn, tws, lane); break;
    default: field_warp_fft_t<INVERSE, 0>(x, n, logn, tws, lane); break;
    }
}

__device__ __forceinline__ int field_brev(int i, int logn) { return logn ? (int)(__brev((unsigned)i) >> (32 - logn)) : 0; }

// per-stage twiddle table of a dim from its forward twiddles tw[m] = exp(-2 pi i m / n)
__device__ __forceinline__ void field_fill_tws(double2* tws, const double2* tw, int n, int tid, int nthreads)
{
    for (int q = tid; q < n; q += nthreads) {
        if (q == 0) {
            tws[0] = make_double2(1.0, 0.0);
        } else {
            const int half = 1 << (31 - __clz(q)), k = q - half;  // q = half + k
            tws[q] = tw[k * (n / (2 * half))];
        }
    }
}

// One space dim (1D1V), ONE warp: rho = scale * sum_c partial[c] (stored raw, then mean-free), forward FFT, multiplier
// i k/|k|^2 (src/poisson.jl:7-15, :139-144), inverse FFT, real part -> E.  tw1 = per-stage twiddle table in shared
// memory (field_fill_tws), xa / xb = two lines of n1 double2 in shared memory.  E_local (optional): a second copy of E
// (shared memory of the calling block; it also parks the raw rho until the mean is known, so that rho - mean is stored
// once instead of read back and modified); write_global = false: rho and E are not stored to global memory;
// mult_brev (optional): the multiplier already in bit-reversed order (shared memory).
__device__ __forceinline__ void field_fft_1d_warp(const FieldFftArgs& fa, const double2* tw1, double2* xa, double2* xb, int lane,
                                                  double* E_local = nullptr, bool write_global = true, const double* mult_brev = nullptr)
{
    const int n1 = fa.n1, l1 = fa.l1;
    // four points per lane at a time: their loads are independent of each other and of the stores below (a loop that
    // stores rho between the loads serialises one L2 round trip per element: 2.5 us of a 4 us solve); the additions per
    // point keep their order (chunk 0, 1, ...)
    for (int a0 = lane; a0 < n1; a0 += 128) {
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        for (int c = 0; c < fa.nchunk; ++c) {
            const double* pc = fa.partial + (long long)c * n1 + a0;
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = (a0 + 32 * q < n1) ? pc[32 * q] : 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] += v[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int a = a0 + 32 * q;
            if (a < n1) {
                const double sq = s[q] * fa.scale;
                if (E_local)
                    E_local[a] = sq;
                else if (write_global)
                    fa.rho[a] = sq;
                xa[a] = make_double2(sq, 0.0);
            }
        }
    }
    field_warp_fft<false>(xa, n1, l1, tw1, lane);
    const double mean = fa.subtract_mean ? xa[0].x / (double)n1 : 0.0;
    __syncwarp();
    for (int p = lane; p < n1; p += 32) {
        const double2 v = xa[p];
        const double mm = mult_brev ? mult_brev[p] : fa.mult[0][field_brev(p, l1)];
        xb[p] = make_double2(-v.y * mm, v.x * mm);
    }
    field_warp_fft<true>(xb, n1, l1, tw1, lane);
    const double sc = 1.0 / (double)n1;
    for (int a = lane; a < n1; a += 32) {
        const double e = xb[a].x * sc;
        if (E_local) {
            if (write_global) {
                fa.rho[a] = E_local[a] - mean;   // == the stored raw value minus the mean, as the two-step form computes it
                fa.E[0][a] = e;
            }
            E_local[a] = e;
        } else if (write_global) {
            fa.E[0][a] = e;
            fa.rho[a] -= mean;
        }
    }
}
