// slb_points.cuh -- the N-D PER-POINT interpolation of the unsplit 2-D solvers.
//
// Reference: interpolate!(fp, fi, bufdec::Array{OpTuple}, interp_t) (src/interpolation.jl:561-621)
// and its closure twin interpolate!(fp, fi, dec::Function, interp_t) (:401-429), N = 2.  Every
// grid point (i, j) has its own displacement (a_1, a_2) in grid units:
//     d_x = floor(a_x), t_x = a_x - d_x, w^x = tabfct_x(t_x)              (getprecal, :345-353)
//     tab[a, b] = w^1_a * w^2_b                                            (dotprod, :112-118)
//     fp[i, j] = sum_{a, b} res[(i + d_1 - p_1/2 + a) mod n_1, (j + d_2 - p_2/2 + b) mod n_2] * tab[a, b]
// with res = sol(interp_t, fi), the 1-D B-spline solves along each dim (:48-94) -- done by the
// caller with the K2 pre-solve kernel before this one.  fi may be a field of OpTuple{2}
// (displacement fields of the Adams-Bashforth time algorithms, src/advection.jl:391-580): stored
// as `ncomp` planes [n1, n2, ncomp], all sharing the point's weights.
//
// Mapping: one thread per point, lanes along the contiguous index i, so for a smooth displacement
// field a warp's gathers for one stencil tap fall into one or two 256-byte row segments (served
// from L1/L2: these grids are a few MB).  Weights by FMA Horner (== Base.evalpoly) from the
// constant bank.  EXACT reproduces the reference's operation order (rounded tab products, rounded
// res .* tab products, summed in column-major order); the default contracts each stencil row into
// an FMA chain and applies w^2_b once per row ((p+1)^2 + (p+1) FMAs instead of 2 (p+1)^2 flops).
//
// The per-point body is __host__ __device__ so that tests can run the same code on the CPU
// (csrc/bspline_hosttest.cu) where no GPU is present.
#pragma once
#include "slb_sweep.cuh"
#include <math.h>

#define SLB_POINTS_MAXP1 64

struct PointsArgs {
    const double* res;  // [n1, n2, ncomp], pre-solved
    const double* dec;  // [n1, n2, 2]
    double* out;        // [n1, n2, ncomp]
    int n1, n2, ncomp;
    int pA, pB;    // order + 1 per dim
    int ncA, ncB;  // coefficients per weight polynomial
};

#ifdef __CUDA_ARCH__
#define SLB_PT_MUL(a, b) __dmul_rn(a, b)
#define SLB_PT_ADD(a, b) __dadd_rn(a, b)
#else
#define SLB_PT_MUL(a, b) ((a) * (b))
#define SLB_PT_ADD(a, b) ((a) + (b))
#endif

__host__ __device__ __forceinline__ void slb_pt_split(double alpha, int n, int half, double& t, int& s0)
{
    double fl = floor(alpha);
    t = alpha - fl;
    long long d = (fabs(fl) < 9.0e18) ? (long long)fl - half : 0;
    long long r = d % n;
    s0 = (int)(r < 0 ? r + n : r);
}

// P1 > 0: both dims have order + 1 == P1 (tap loop unrolled, dim-1 weights in registers); P1 == 0:
// run-time orders up to 63 each (dim-1 weights in local memory).  cA / cB: weight polynomial rows, row
// stride sA / sB.
// Register budget decides this kernel: unrolling all (p+1)^2 taps lets the compiler hoist a hundred
// gathers and costs 250 registers (8 warps per SM, every gather at full L2 latency).  Only the taps of
// ONE stencil row are unrolled; the row loop stays rolled and evaluates its dim-2 weight on the fly
// (each w^2_b is used by exactly one row, so nothing is computed twice), two component planes share a
// pass.  Windows that do not wrap around dim 1 use immediate offsets.
template <int P1, bool EXACT>
__host__ __device__ __forceinline__ void slb_point_eval(const PointsArgs& pa, const double* cA, int sA, const double* cB,
                                                        int sB, int i, int j)
{
    constexpr int W = P1 > 0 ? P1 : SLB_POINTS_MAXP1;
    const int pA = P1 > 0 ? P1 : pa.pA;
    const int pB = P1 > 0 ? P1 : pa.pB;
    const int n1 = pa.n1, n2 = pa.n2;
    const long long plane = (long long)n1 * n2;
    const long long idx = (long long)j * n1 + i;
    double tA, tB;
    int sa0, sb0;
    slb_pt_split(pa.dec[idx], n1, (pA - 1) / 2, tA, sa0);
    slb_pt_split(pa.dec[plane + idx], n2, (pB - 1) / 2, tB, sb0);
    double wA[W];
#pragma unroll
    for (int a = 0; a < pA; ++a) {
        double ex = cA[a * sA + pa.ncA - 1];
        for (int k = pa.ncA - 2; k >= 0; --k) ex = fma(tA, ex, cA[a * sA + k]);
        wA[a] = ex;
    }
    // start of the periodic windows: (i + d - p/2) mod n
    int ia0 = i + sa0;
    if (ia0 >= n1) ia0 -= n1;
    int jb0 = j + sb0;
    if (jb0 >= n2) jb0 -= n2;
    const bool nowrap = ia0 + pA <= n1;
    for (int c = 0; c < pa.ncomp; c += 2) {
        const bool two = c + 1 < pa.ncomp;
        const double* r0 = pa.res + (long long)c * plane;
        const double* r1 = two ? r0 + plane : r0;
        double acc0 = 0.0, acc1 = 0.0;
        int jb = jb0;
#pragma unroll 1
        for (int b = 0; b < pB; ++b) {
            double wb = cB[b * sB + pa.ncB - 1];
            for (int k = pa.ncB - 2; k >= 0; --k) wb = fma(tB, wb, cB[b * sB + k]);
            const double* row0 = r0 + (long long)jb * n1;
            const double* row1 = r1 + (long long)jb * n1;
            double in0 = 0.0, in1 = 0.0;
            if (nowrap) {
#pragma unroll
                for (int a = 0; a < pA; ++a) {
                    const double v0 = row0[ia0 + a];
                    const double v1 = two ? row1[ia0 + a] : 0.0;
                    if (EXACT) {
                        const double tab = SLB_PT_MUL(wA[a], wb);
                        const double t0 = SLB_PT_MUL(v0, tab), t1 = SLB_PT_MUL(v1, tab);
                        acc0 = (a == 0 && b == 0) ? t0 : SLB_PT_ADD(acc0, t0);
                        acc1 = (a == 0 && b == 0) ? t1 : SLB_PT_ADD(acc1, t1);
                    } else {
                        in0 = (a == 0) ? v0 * wA[a] : fma(v0, wA[a], in0);
                        in1 = (a == 0) ? v1 * wA[a] : fma(v1, wA[a], in1);
                    }
                }
            } else {
                int ia = ia0;
#pragma unroll
                for (int a = 0; a < pA; ++a) {
                    const double v0 = row0[ia];
                    const double v1 = two ? row1[ia] : 0.0;
                    if (EXACT) {
                        const double tab = SLB_PT_MUL(wA[a], wb);
                        const double t0 = SLB_PT_MUL(v0, tab), t1 = SLB_PT_MUL(v1, tab);
                        acc0 = (a == 0 && b == 0) ? t0 : SLB_PT_ADD(acc0, t0);
                        acc1 = (a == 0 && b == 0) ? t1 : SLB_PT_ADD(acc1, t1);
                    } else {
                        in0 = (a == 0) ? v0 * wA[a] : fma(v0, wA[a], in0);
                        in1 = (a == 0) ? v1 * wA[a] : fma(v1, wA[a], in1);
                    }
                    if (++ia == n1) ia = 0;
                }
            }
            if (!EXACT) {
                acc0 = (b == 0) ? in0 * wb : fma(in0, wb, acc0);
                acc1 = (b == 0) ? in1 * wb : fma(in1, wb, acc1);
            }
            if (++jb == n2) jb = 0;
        }
        pa.out[(long long)c * plane + idx] = acc0;
        if (two) pa.out[(long long)(c + 1) * plane + idx] = acc1;
    }
}

#ifdef __CUDACC__
template <int P1, bool EXACT>
__global__ void __launch_bounds__(128) k_interp2d_points(PointsArgs pa, CoefTab ctA, CoefTab ctB)
{
    int i = blockIdx.x * 128 + threadIdx.x;
    int j = blockIdx.y;
    if (i >= pa.n1) return;
    slb_point_eval<P1, EXACT>(pa, ctA.c, SLB_NCMAX, ctB.c, SLB_NCMAX, i, j);
}

template <bool EXACT>
__global__ void __launch_bounds__(128) k_interp2d_points_generic(PointsArgs pa, const double* __restrict__ coefA,
                                                                 const double* __restrict__ coefB)
{
    int i = blockIdx.x * 128 + threadIdx.x;
    int j = blockIdx.y;
    if (i >= pa.n1) return;
    slb_point_eval<0, EXACT>(pa, coefA, pa.ncA, coefB, pa.ncB, i, j);
}

// dec[i, j, 0] = scale_j * tab_j[j],  dec[i, j, 1] = scale_i * tab_i[i]: the displacement field of
// the unsplit 1D1V solvers (src/poisson.jl:229-247 StdPoisson2d, src/rotation.jl:36-54).
__global__ void __launch_bounds__(128)
k_fill_dec2d(double* __restrict__ dec, int n1, int n2, const double* __restrict__ tab_j, double scale_j,
             const double* __restrict__ tab_i, double scale_i)
{
    int i = blockIdx.x * 128 + threadIdx.x;
    int j = blockIdx.y;
    if (i >= n1) return;
    long long idx = (long long)j * n1 + i;
    dec[idx] = scale_j * __ldg(tab_j + j);
    dec[(long long)n1 * n2 + idx] = scale_i * __ldg(tab_i + i);
}

// out = sum_k coef[k] * x[k], products rounded, summed left to right: the reference's
// sum(map(k -> c(abcoef, k, ord) * t_bufc[k], 1:ord)) on arrays (src/advection.jl:431, :479, :506, :540)
#define SLB_LINCOMB_MAX 8
struct LincombArgs {
    int nterms;
    double coef[SLB_LINCOMB_MAX];
    const double* x[SLB_LINCOMB_MAX];
};
__global__ void __launch_bounds__(256) k_lincomb(double* out, long long n, LincombArgs la)
{
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double acc = __dmul_rn(la.coef[0], la.x[0][i]);
    for (int k = 1; k < la.nterms; ++k) acc = __dadd_rn(acc, __dmul_rn(la.coef[k], la.x[k][i]));
    out[i] = acc;
}
#endif
