// slb_field.cuh -- K4 (charge density), K5 (Fourier Poisson solve) and the deterministic
// reductions behind compute_ee / compute_ke.
//
// Reference: src/util_poisson.jl:68-79 (compute_charge!), src/poisson.jl:7-15,139-144
// (compute_elfield!), src/util_poisson.jl:41-53,156-162 (compute_ke / compute_ee).
// All reductions are atomics-free and run in a fixed order, so electric-energy histories
// are reproducible run to run ("hard parts").
#pragma once
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
