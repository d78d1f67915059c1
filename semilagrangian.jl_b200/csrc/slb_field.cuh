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
// K5: direct DFT along one dim of a small N-D array (the space grid: <= 256^2 / 1024 points
// per dim).  out(a,k,b) = sum_j in(a,j,b) * tw[(j*k) mod n], tw[m] = exp(-+2 pi i m/n).
// Convention of src/fftbig.jl:162-176 (FFTW): forward exp(-i..) unnormalised; inverse 1/n.
// ------------------------------------------------------------------------------------------
template <bool REAL_IN, bool INVERSE>
__global__ void __launch_bounds__(128)
k_dft_dim(const void* __restrict__ in_, double2* __restrict__ out, long long inner, int n, long long total,
          const double2* __restrict__ tw)
{
    long long id = (long long)blockIdx.x * 128 + threadIdx.x;
    if (id >= total) return;
    long long a = id % inner;
    long long r = id / inner;
    int k = (int)(r % n);
    long long b = r / n;
    long long base = a + inner * (long long)n * b;
    double sr = 0.0, si = 0.0;
    int m = 0;
    for (int j = 0; j < n; ++j) {
        double2 t = __ldg(tw + m);
        double ti = INVERSE ? -t.y : t.y;
        if (REAL_IN) {
            double v = __ldg(reinterpret_cast<const double*>(in_) + base + inner * j);
            sr = fma(v, t.x, sr);
            si = fma(v, ti, si);
        } else {
            double2 v = __ldg(reinterpret_cast<const double2*>(in_) + base + inner * j);
            sr += v.x * t.x - v.y * ti;
            si += v.x * ti + v.y * t.x;
        }
        m += k;
        if (m >= n) m -= n;
    }
    if (INVERSE) {
        double s = 1.0 / (double)n;
        sr *= s;
        si *= s;
    }
    out[id] = make_double2(sr, si);
}

// D = (i * m) .* C   (fctv_k is purely imaginary: src/poisson.jl:8,14)
__global__ void k_mult_imag(const double2* __restrict__ c, const double* __restrict__ m, double2* __restrict__ d,
                            long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double2 v = c[i];
    double mm = m[i];
    d[i] = make_double2(-v.y * mm, v.x * mm);
}

__global__ void k_real_part(const double2* __restrict__ c, double* __restrict__ e, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) e[i] = c[i].x;
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
