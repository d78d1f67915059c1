// slb_sweep.cuh -- K1: the 1-D periodic interpolation sweep (Lagrange / Hermite / B-spline
// stencil part) for sm_100a.
//
// Semantics (reference: src/interpolation.jl:175-193 + :381-396, called from
// src/advection.jl:627-631): for each line u[0..n) along the swept dimension,
//     d = floor(alpha), t = alpha - d, w_j = tabfct[j](t) (Horner, FMA == Base.evalpoly/muladd)
//     out[i] = sum_{j=0..p} u[(i + d - p/2 + j) mod n] * w_j
// The grid is column-major; a sweep along dim `dim` sees it as [inner, n, outer] with
// element (a, k, b) at a + inner*(k + n*b).  Sweeps are out-of-place (in != out), which
// removes the reference's two permutedims! passes and its per-line copy buffer.
//
//  * k_sweep_strided  (dim > 0): one thread per line, lanes along the contiguous `inner`
//    index, so every load/store of a warp is one coalesced 256 B row segment.  The thread
//    marches along k holding the p+1 stencil inputs in a rotating register window: each
//    element is loaded from HBM once and feeds p+1 FMAs from registers.  Loads are issued
//    one group (p+1 elements) ahead of their use.
//  * k_sweep_contig   (dim == 0): one warp per line segment of 32*R outputs.  The warp
//    stages the segment (+p halo, periodic wrap applied on the global index) in shared
//    memory in an [e mod R][e div R] layout, so that lane l can read its R+p window inputs
//    conflict-free, computes R consecutive outputs from registers and writes them with one
//    vector store.
//  * k_sweep_generic: any order <= 63, any dim; thread per line, weights in local memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SLB_P1MAX 14   // fast path: order+1 <= 14
#define SLB_NCMAX 14   // fast path: polynomial coefficients per weight <= 14
#define SLB_MAXD 6

struct CoefTab {  // weight polynomials, passed by value -> constant bank, uniform LDC reads
    double c[SLB_P1MAX * SLB_NCMAX];
};

// alpha = scale * tab[ sum_d idx_d * stride_d ] over the dims below (lo) / above (hi) the swept one
struct AlphaMap {
    const double* tab;
    double scale;
    int nlo, nhi;
    unsigned ext_lo[SLB_MAXD], ext_hi[SLB_MAXD];
    long long str_lo[SLB_MAXD], str_hi[SLB_MAXD];
};

// Generalised addressing of the swept index k on the OUTPUT side (multi-GPU re-shard fused
// into the sweep epilogue): k -> (k % kc) * inner + (k / kc) * kblk.  kc == n: plain layout.
struct OutMap {
    int kc;
    long long kblk;
    long long bstride;  // offset between consecutive outer indices b (plain: n * inner)
};

__device__ __forceinline__ long long slb_alpha_off(const AlphaMap& m, unsigned long long a, unsigned long long b)
{
    long long off = 0;
    // the host trims zero-stride dims: nlo == 0 when alpha does not depend on the inner index
#pragma unroll 1
    for (int d = 0; d < m.nlo; ++d) {
        unsigned long long q = a / m.ext_lo[d];
        off += (long long)(a - q * m.ext_lo[d]) * m.str_lo[d];
        a = q;
    }
#pragma unroll 1
    for (int d = 0; d < m.nhi; ++d) {
        unsigned long long q = b / m.ext_hi[d];
        off += (long long)(b - q * m.ext_hi[d]) * m.str_hi[d];
        b = q;
    }
    return off;
}

// floor/frac split of src/interpolation.jl:381-389 and the start index of the periodic window
__device__ __forceinline__ void slb_split(double alpha, int n, int half, double& t, int& s0)
{
    double fl = floor(alpha);
    t = alpha - fl;
    long long d = (long long)fl - half;
    long long r = d % n;
    s0 = (int)(r < 0 ? r + n : r);
}

template <int P1, bool EXACT>
__device__ __forceinline__ double slb_dot(const double (&x)[P1], const double (&w)[P1], int rot)
{
    // x is a rotating window: logical element j lives in x[(rot + j) % P1]
    double acc;
    if (EXACT) {
        acc = __dmul_rn(x[rot % P1], w[0]);
#pragma unroll
        for (int j = 1; j < P1; ++j) acc = __dadd_rn(acc, __dmul_rn(x[(rot + j) % P1], w[j]));
    } else {
        acc = x[rot % P1] * w[0];
#pragma unroll
        for (int j = 1; j < P1; ++j) acc = fma(x[(rot + j) % P1], w[j], acc);
    }
    return acc;
}

// ------------------------------------------------------------------------------------------
// K1 strided: thread per line
// ------------------------------------------------------------------------------------------
template <int P1, bool EXACT>
__global__ void __launch_bounds__(128)
k_sweep_strided(const double* __restrict__ in, double* __restrict__ out, long long inner, int n, long long nlines,
                AlphaMap am, CoefTab ct, int nc, OutMap om)
{
    long long gid = (long long)blockIdx.x * 128 + threadIdx.x;
    if (gid >= nlines) return;
    long long b = gid / inner;
    long long a = gid - b * inner;

    double alpha = am.scale * __ldg(am.tab + slb_alpha_off(am, (unsigned long long)a, (unsigned long long)b));
    double t;
    int s0;
    slb_split(alpha, n, (P1 - 1) / 2, t, s0);

    double w[P1];
#pragma unroll
    for (int j = 0; j < P1; ++j) w[j] = ct.c[j * SLB_NCMAX + nc - 1];
    for (int k = nc - 2; k >= 0; --k) {
#pragma unroll
        for (int j = 0; j < P1; ++j) w[j] = fma(t, w[j], ct.c[j * SLB_NCMAX + k]);
    }

    const double* pin = in + (b * n) * inner + a;
    double* pout = out + b * om.bstride + a;
    int kk = s0;
    const double* pl = pin + (long long)kk * inner;

#define SLB_LOAD_NEXT(dst)            \
    {                                 \
        dst = __ldg(pl);              \
        ++kk;                         \
        pl += inner;                  \
        if (kk == n) {                \
            kk = 0;                   \
            pl = pin;                 \
        }                             \
    }

    double win[P1];
#pragma unroll
    for (int j = 0; j < P1 - 1; ++j) SLB_LOAD_NEXT(win[j]);

    double nxa[P1], nxb[P1];
#pragma unroll
    for (int r = 0; r < P1; ++r) SLB_LOAD_NEXT(nxa[r]);

    const bool plain = (om.kc >= n);
    double* po = pout;
    int ko = 0;  // position inside the current output k-block

#define SLB_STORE(val)                                  \
    {                                                   \
        *po = (val);                                    \
        po += inner;                                    \
        if (!plain) {                                   \
            if (++ko == om.kc) {                        \
                ko = 0;                                 \
                po += om.kblk - (long long)om.kc * inner; \
            }                                           \
        }                                               \
    }

#define SLB_COMPUTE_GROUP(buf, tb)                          \
    _Pragma("unroll") for (int r = 0; r < P1; ++r)          \
    {                                                       \
        if ((tb) + r < n) {                                 \
            win[(r + P1 - 1) % P1] = buf[r];                \
            double acc = slb_dot<P1, EXACT>(win, w, r);     \
            SLB_STORE(acc);                                 \
        }                                                   \
    }

    for (int t0 = 0; t0 < n; t0 += 2 * P1) {
        if (t0 + P1 < n) {
#pragma unroll
            for (int r = 0; r < P1; ++r) SLB_LOAD_NEXT(nxb[r]);
        }
        SLB_COMPUTE_GROUP(nxa, t0);
        if (t0 + 2 * P1 < n) {
#pragma unroll
            for (int r = 0; r < P1; ++r) SLB_LOAD_NEXT(nxa[r]);
        }
        if (t0 + P1 < n) {
            SLB_COMPUTE_GROUP(nxb, t0 + P1);
        }
    }
#undef SLB_COMPUTE_GROUP
#undef SLB_STORE
#undef SLB_LOAD_NEXT
}

// ------------------------------------------------------------------------------------------
// K1 contiguous: warp per line segment
// ------------------------------------------------------------------------------------------
template <int R>
struct ContigCfg {
    // row pitch (doubles) of the [e mod R][e div R] staging tile; chosen so that both the
    // staging stores (lanes = consecutive e) and the window loads (lanes = consecutive
    // columns) touch every bank pair exactly twice per 32 x 8 B request (the minimum).
    static constexpr int PITCH = (R == 4) ? 40 : (R == 2 ? 48 : 32 + SLB_P1MAX);
};

template <int P1, int R, bool EXACT>
__global__ void __launch_bounds__(256)
k_sweep_contig(const double* __restrict__ in, double* __restrict__ out, int n, long long nlines, AlphaMap am,
               const double* __restrict__ coef, int nc)
{
    constexpr int WARPS = 8;
    constexpr int SEG = 32 * R;           // outputs per warp pass
    constexpr int W = SEG + P1 - 1;       // inputs staged per pass
    constexpr int NLD = (W + 31) / 32;
    constexpr int PITCH = ContigCfg<R>::PITCH;
    constexpr int NX = R + P1 - 1;
    __shared__ double sm[WARPS][R * PITCH];

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    double* tile = sm[wid];

    for (long long line = (long long)blockIdx.x * WARPS + wid; line < nlines; line += (long long)gridDim.x * WARPS) {
        double alpha = am.scale * __ldg(am.tab + slb_alpha_off(am, 0ull, (unsigned long long)line));
        double t;
        int s0;
        slb_split(alpha, n, (P1 - 1) / 2, t, s0);

        // lane j evaluates weight polynomial j, then the warp broadcasts
        double wl = 0.0;
        if (lane < P1) {
            const double* c = coef + lane * nc;
            wl = __ldg(c + nc - 1);
            for (int k = nc - 2; k >= 0; --k) wl = fma(t, wl, __ldg(c + k));
        }
        double w[P1];
#pragma unroll
        for (int j = 0; j < P1; ++j) w[j] = __shfl_sync(0xffffffffu, wl, j);

        const double* lin = in + line * n;
        double* lout = out + line * n;

        for (int seg0 = 0; seg0 < n; seg0 += SEG) {
            int g0 = s0 + seg0;
            if (g0 >= n) g0 -= n;
#pragma unroll
            for (int c = 0; c < NLD; ++c) {
                int e = lane + 32 * c;
                if (e < W) {
                    int g = g0 + e;
                    if (g >= n) {
                        g -= n;
                        if (g >= n) g %= n;
                    }
                    tile[(e % R) * PITCH + e / R] = __ldg(lin + g);
                }
            }
            __syncwarp();
            double x[NX];
#pragma unroll
            for (int jj = 0; jj < NX; ++jj) x[jj] = tile[(jj % R) * PITCH + lane + jj / R];
            double o[R];
#pragma unroll
            for (int m = 0; m < R; ++m) {
                double acc;
                if (EXACT) {
                    acc = __dmul_rn(x[m], w[0]);
#pragma unroll
                    for (int j = 1; j < P1; ++j) acc = __dadd_rn(acc, __dmul_rn(x[m + j], w[j]));
                } else {
                    acc = x[m] * w[0];
#pragma unroll
                    for (int j = 1; j < P1; ++j) acc = fma(x[m + j], w[j], acc);
                }
                o[m] = acc;
            }
            int i0 = seg0 + R * lane;
            double* dst = lout + i0;
            if (i0 + R <= n && ((reinterpret_cast<uintptr_t>(dst) & (R * 8 - 1)) == 0)) {
                if (R == 4) {
                    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(dst), "d"(o[0]), "d"(o[1 % R]),
                                 "d"(o[2 % R]), "d"(o[3 % R])
                                 : "memory");
                } else if (R == 2) {
                    *reinterpret_cast<double2*>(dst) = make_double2(o[0], o[1 % R]);
                } else {
                    dst[0] = o[0];
                }
            } else {
#pragma unroll
                for (int m = 0; m < R; ++m)
                    if (i0 + m < n) dst[m] = o[m];
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------
// generic fallback: thread per line, any order <= 63
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_sweep_generic(const double* __restrict__ in, double* __restrict__ out, long long inner, int n, long long nlines,
                AlphaMap am, const double* __restrict__ coef, int np, int nc, int exact)
{
    long long gid = (long long)blockIdx.x * 128 + threadIdx.x;
    if (gid >= nlines) return;
    long long b = gid / inner;
    long long a = gid - b * inner;
    double alpha = am.scale * __ldg(am.tab + slb_alpha_off(am, (unsigned long long)a, (unsigned long long)b));
    double t;
    int s0;
    slb_split(alpha, n, (np - 1) / 2, t, s0);
    double w[64];
    for (int j = 0; j < np; ++j) {
        const double* c = coef + j * nc;
        double ex = __ldg(c + nc - 1);
        for (int k = nc - 2; k >= 0; --k) ex = fma(t, ex, __ldg(c + k));
        w[j] = ex;
    }
    const double* pin = in + (b * n) * inner + a;
    double* pout = out + (b * n) * inner + a;
    int k0 = s0;
    for (int i = 0; i < n; ++i) {
        int k = k0;
        double acc = 0.0;
        for (int j = 0; j < np; ++j) {
            double v = __ldg(pin + (long long)k * inner);
            if (exact) {
                double pr = __dmul_rn(v, w[j]);
                acc = (j == 0) ? pr : __dadd_rn(acc, pr);
            } else {
                acc = (j == 0) ? v * w[j] : fma(v, w[j], acc);
            }
            if (++k == n) k = 0;
        }
        pout[(long long)i * inner] = acc;
        if (++k0 == n) k0 = 0;
    }
}
