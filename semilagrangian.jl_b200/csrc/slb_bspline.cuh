// slb_bspline.cuh -- K2: periodic B-spline pre-solve  c = A^{-1} u  batched over lines.
//
// Reference: sol(::BSplineLU) src/bsplinelu.jl:179-220,275-284 (cyclic banded LU stored as
// band + lastrows + lastcols) and sol(::BSplineFFT) src/bsplinefft.jl:49-51 (the same
// circulant system solved by FFT).  A is the n x n circulant collocation matrix with
// A[i][(i+m) mod n] = a_|m|, a_m = B((order+1)/2 + m), |m| <= h = (order-1)/2
// (layout pinned by test/test_splinelu.jl:10-29).  Odd orders only (the reference's
// BSplineLU throws for even orders, src/bsplinelu.jl:257-261; even-order circulants are
// singular at the Nyquist mode).
//
// GPU formulation (not the reference's storage): bordered banded LU.  With N = n - h,
//     A = [ T  C ]   T: N x N banded Toeplitz (no wrap), C: N x h, R: h x N, D: h x h
//         [ R  D ]
// T = L U without pivoting (A is symmetric positive definite).  The host precomputes, in
// extended precision, per-row tables  L (N x h),  U (N x h) + 1/diag,  Ri = R U^{-1} (N x h),
// G = T^{-1} C (N x h)  and  Sinv = (D - R T^{-1} C)^{-1} (h x h).  One thread solves one line:
//     forward   y_i = u_i - sum_j L[i][j] y_{i-j} ;  acc_r += Ri[i][r] * y_i
//     border    x2  = Sinv (u2 - acc)
//     backward  w_i = (y_i - sum_j U[i][j] w_{i+j}) / diag_i ;  x_i = w_i - sum_r G[i][r] x2_r
// i.e. 4h FMAs per cell with register windows of h values; every table read is
// warp-uniform.  A warp owns a tile of 32 lines staged in shared memory as [k][33] (pitch
// 33 makes both the transposed staging of contiguous lines and the thread-per-line accesses
// conflict-free); lines too long for shared memory fall back to streaming through global
// memory.  Both B-spline kinds use this solver: for odd orders BSplineFFT and BSplineLU
// solve the same system, and the banded solve costs 4h FMAs per cell
// against ~10 log2(n) for an FFT pair.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/slb200.h"

#define SLB_BSP_HMAX 13  // order <= 27

struct BsplineDev {
    int h, n, N;
    double *L, *U, *invd, *Ri, *G, *Sinv;
};

static void bspline_free(BsplineDev* b)
{
    if (b->L) cudaFree(b->L);
    if (b->U) cudaFree(b->U);
    if (b->invd) cudaFree(b->invd);
    if (b->Ri) cudaFree(b->Ri);
    if (b->G) cudaFree(b->G);
    if (b->Sinv) cudaFree(b->Sinv);
    memset(b, 0, sizeof(*b));
}

// host-side tables (double), produced by bspline_factor
struct BsplineHost {
    int h, n, N;
    std::vector<double> L, U, invd, Ri, G, Sinv;
};

static int bspline_factor(int order, int64_t n64, const double* node_vals, BsplineHost* hostout, std::string& msg)
{
    typedef long double ld;
    const int h = (order - 1) / 2;
    const int n = (int)n64;
    if (h < 1 || h > SLB_BSP_HMAX) { msg = "B-spline order must be odd and in [3,27]"; return SLB_E_UNSUPPORTED; }
    if (n < 2 * h + 2) { msg = "B-spline: line too short for the stencil (n >= order + 1 required)"; return SLB_E_ARG; }
    const int N = n - h;
    // a_m, m = 0..h : node_vals[i] = B(i+1), centre B((order+1)/2) = node_vals[h]
    std::vector<ld> a(h + 1);
    for (int m = 0; m <= h; ++m) a[m] = (ld)node_vals[h + m];
    auto Aent = [&](int i, int j) -> ld {  // cyclic entry A[i][j] = a_{|i-j| cyclic}; n >= 2h+2: no overlap
        int m = ((j - i) % n + n) % n;
        if (n - m < m) m = n - m;
        return m <= h ? a[m] : 0.0L;
    };
    // banded LU of T (rows 0..N-1): Lm[i][j-1] multiplies row i-j, Um[i][j] = U(i, i+j)
    std::vector<ld> Lm((size_t)N * h, 0.0L), Um((size_t)N * (h + 1), 0.0L);
    {
        // work row: entries of row i in columns i-h..i+h
        for (int i = 0; i < N; ++i) {
            std::vector<ld> row(2 * h + 1, 0.0L);
            for (int m = -h; m <= h; ++m) {
                int j = i + m;
                if (j >= 0 && j < N) row[m + h] = a[m < 0 ? -m : m];
            }
            for (int j = h; j >= 1; --j) {  // eliminate column i-j using row i-j
                int pr = i - j;
                if (pr < 0) continue;
                ld mult = row[h - j] / Um[(size_t)pr * (h + 1)];
                Lm[(size_t)i * h + (j - 1)] = mult;
                for (int q = 0; q <= h; ++q) {
                    int col = pr + q;       // absolute column
                    int off = col - i + h;  // position in row window
                    if (off >= 0 && off <= 2 * h) row[off] -= mult * Um[(size_t)pr * (h + 1) + q];
                }
            }
            for (int q = 0; q <= h; ++q) Um[(size_t)i * (h + 1) + q] = (i + q < N) ? row[h + q] : 0.0L;
            if (Um[(size_t)i * (h + 1)] == 0.0L) { msg = "B-spline: singular collocation matrix"; return SLB_E_ARG; }
        }
    }
    // Ri = R U^{-1}: for each border row r solve Ri[r][.] U = R[r][.]
    std::vector<ld> Ri((size_t)N * h, 0.0L), Lc((size_t)N * h, 0.0L), G((size_t)N * h, 0.0L);
    for (int r = 0; r < h; ++r) {
        for (int c = 0; c < N; ++c) {
            ld s = Aent(N + r, c);
            for (int j = 1; j <= h; ++j)
                if (c - j >= 0) s -= Ri[(size_t)(c - j) * h + r] * Um[(size_t)(c - j) * (h + 1) + j];
            Ri[(size_t)c * h + r] = s / Um[(size_t)c * (h + 1)];
        }
    }
    // Lc = L^{-1} C, G = U^{-1} Lc
    for (int q = 0; q < h; ++q) {
        for (int i = 0; i < N; ++i) {
            ld s = Aent(i, N + q);
            for (int j = 1; j <= h; ++j)
                if (i - j >= 0) s -= Lm[(size_t)i * h + (j - 1)] * Lc[(size_t)(i - j) * h + q];
            Lc[(size_t)i * h + q] = s;
        }
        for (int i = N - 1; i >= 0; --i) {
            ld s = Lc[(size_t)i * h + q];
            for (int j = 1; j <= h; ++j)
                if (i + j < N) s -= Um[(size_t)i * (h + 1) + j] * G[(size_t)(i + j) * h + q];
            G[(size_t)i * h + q] = s / Um[(size_t)i * (h + 1)];
        }
    }
    // S = D - Ri^T Lc  (h x h), Sinv by Gauss-Jordan with partial pivoting
    std::vector<ld> S((size_t)h * h), Si((size_t)h * h, 0.0L);
    for (int r = 0; r < h; ++r)
        for (int q = 0; q < h; ++q) {
            ld s = Aent(N + r, N + q);
            for (int i = 0; i < N; ++i) s -= Ri[(size_t)i * h + r] * Lc[(size_t)i * h + q];
            S[(size_t)r * h + q] = s;
        }
    for (int r = 0; r < h; ++r) Si[(size_t)r * h + r] = 1.0L;
    for (int c = 0; c < h; ++c) {
        int piv = c;
        for (int r = c + 1; r < h; ++r)
            if (fabsl(S[(size_t)r * h + c]) > fabsl(S[(size_t)piv * h + c])) piv = r;
        if (S[(size_t)piv * h + c] == 0.0L) { msg = "B-spline: singular Schur complement"; return SLB_E_ARG; }
        if (piv != c)
            for (int q = 0; q < h; ++q) {
                std::swap(S[(size_t)c * h + q], S[(size_t)piv * h + q]);
                std::swap(Si[(size_t)c * h + q], Si[(size_t)piv * h + q]);
            }
        ld inv = 1.0L / S[(size_t)c * h + c];
        for (int q = 0; q < h; ++q) { S[(size_t)c * h + q] *= inv; Si[(size_t)c * h + q] *= inv; }
        for (int r = 0; r < h; ++r)
            if (r != c) {
                ld f = S[(size_t)r * h + c];
                if (f != 0.0L)
                    for (int q = 0; q < h; ++q) { S[(size_t)r * h + q] -= f * S[(size_t)c * h + q]; Si[(size_t)r * h + q] -= f * Si[(size_t)c * h + q]; }
            }
    }
    // round to double and upload
    std::vector<double> dL((size_t)N * h), dU((size_t)N * h), dinv(N), dRi((size_t)N * h), dG((size_t)N * h), dS((size_t)h * h);
    for (int i = 0; i < N; ++i) {
        dinv[i] = (double)(1.0L / Um[(size_t)i * (h + 1)]);
        for (int j = 0; j < h; ++j) {
            dL[(size_t)i * h + j] = (double)Lm[(size_t)i * h + j];
            dU[(size_t)i * h + j] = (double)Um[(size_t)i * (h + 1) + j + 1];
            dRi[(size_t)i * h + j] = (double)Ri[(size_t)i * h + j];
            dG[(size_t)i * h + j] = (double)G[(size_t)i * h + j];
        }
    }
    for (int i = 0; i < h * h; ++i) dS[i] = (double)Si[i];
    hostout->h = h; hostout->n = n; hostout->N = N;
    hostout->L.swap(dL); hostout->U.swap(dU); hostout->invd.swap(dinv); hostout->Ri.swap(dRi); hostout->G.swap(dG); hostout->Sinv.swap(dS);
    return SLB_OK;
}

static int bspline_build(int kind, int order, int64_t n64, const double* node_vals, BsplineDev* out, std::string& msg,
                         BsplineHost* host_out = nullptr)
{
    (void)kind;
    BsplineHost hb_local;
    BsplineHost& hb = host_out ? *host_out : hb_local;
    int rc = bspline_factor(order, n64, node_vals, &hb, msg);
    if (rc) return rc;
    const std::vector<double>&dL = hb.L, &dU = hb.U, &dinv = hb.invd, &dRi = hb.Ri, &dG = hb.G, &dS = hb.Sinv;
    memset(out, 0, sizeof(*out));
    out->h = hb.h; out->n = hb.n; out->N = hb.N;
    cudaError_t e = cudaSuccess;
    auto up = [&](double** dst, const std::vector<double>& src) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(dst, src.size() * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(*dst, src.data(), src.size() * sizeof(double), cudaMemcpyHostToDevice);
    };
    up(&out->L, dL); up(&out->U, dU); up(&out->invd, dinv); up(&out->Ri, dRi); up(&out->G, dG); up(&out->Sinv, dS);
    if (e != cudaSuccess) {
        bspline_free(out);
        msg = std::string("B-spline table upload: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return SLB_E_CUDA;
    }
    return SLB_OK;
}

// Thread-per-line solve.  LD(k) / ST(k, v) access element k of this thread's line (in place).
#ifdef __CUDA_ARCH__
#define SLB_LDG(p) __ldg(p)
#else
#define SLB_LDG(p) (*(p))
#endif

template <int H, class LDF, class STF>
__host__ __device__ __forceinline__ void bspline_solve_line(const BsplineDev& f, LDF LD, STF ST)
{
    const int N = f.N;
    double yw[H], acc[H];
#pragma unroll
    for (int s = 0; s < H; ++s) { yw[s] = 0.0; acc[s] = 0.0; }
    // forward
    for (int i0 = 0; i0 < N; i0 += H) {
#pragma unroll
        for (int r = 0; r < H; ++r) {
            int i = i0 + r;
            if (i < N) {
                double y = LD(i);
                const double* Lr = f.L + (size_t)i * H;
                const double* Rr = f.Ri + (size_t)i * H;
#pragma unroll
                for (int j = 1; j <= H; ++j) y = fma(-SLB_LDG(Lr + j - 1), yw[(r - j + 2 * H) % H], y);
                yw[r] = y;
#pragma unroll
                for (int q = 0; q < H; ++q) acc[q] = fma(SLB_LDG(Rr + q), y, acc[q]);
                ST(i, y);
            }
        }
    }
    // border unknowns
    double x2[H];
    {
        double rhs[H];
#pragma unroll
        for (int r = 0; r < H; ++r) rhs[r] = LD(N + r) - acc[r];
#pragma unroll
        for (int q = 0; q < H; ++q) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < H; ++r) s = fma(SLB_LDG(f.Sinv + q * H + r), rhs[r], s);
            x2[q] = s;
        }
#pragma unroll
        for (int q = 0; q < H; ++q) ST(N + q, x2[q]);
    }
    // backward
    double ww[H];
#pragma unroll
    for (int s = 0; s < H; ++s) ww[s] = 0.0;
    for (int i0 = ((N - 1) / H) * H; i0 >= 0; i0 -= H) {
#pragma unroll
        for (int r = H - 1; r >= 0; --r) {
            int i = i0 + r;
            if (i < N) {
                double w = LD(i);
                const double* Ur = f.U + (size_t)i * H;
                const double* Gr = f.G + (size_t)i * H;
#pragma unroll
                for (int j = 1; j <= H; ++j) w = fma(-SLB_LDG(Ur + j - 1), ww[(r + j) % H], w);
                w *= SLB_LDG(f.invd + i);
                ww[r] = w;
                double x = w;
#pragma unroll
                for (int q = 0; q < H; ++q) x = fma(-SLB_LDG(Gr + q), x2[q], x);
                ST(i, x);
            }
        }
    }
}

#define SLB_BSP_PITCH 33

// tile kernel: each warp owns 32 lines staged in shared memory as [k][33]
template <int H>
__global__ void __launch_bounds__(64)
k_bspline_tile(const double* __restrict__ in, double* __restrict__ out, long long inner, int n, long long nlines,
               BsplineDev f, int contig)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double* tile = smem + (size_t)wid * n * SLB_BSP_PITCH;
    long long tile0 = ((long long)blockIdx.x * (blockDim.x >> 5) + wid) * 32;
    if (tile0 >= nlines) return;
    long long line = tile0 + lane;
    bool active = line < nlines;
    if (contig) {
        for (int l = 0; l < 32; ++l) {
            long long ln = tile0 + l;
            if (ln < nlines) {
                const double* p = in + ln * n;
                for (int k = lane; k < n; k += 32) tile[k * SLB_BSP_PITCH + l] = __ldg(p + k);
            }
        }
    } else if (active) {
        long long b = line / inner, a = line - b * inner;
        const double* p = in + (b * n) * inner + a;
#pragma unroll 8
        for (int k = 0; k < n; ++k) tile[k * SLB_BSP_PITCH + lane] = __ldg(p + (long long)k * inner);
    }
    __syncwarp();
    if (active) {
        double* col = tile + lane;
        bspline_solve_line<H>(
            f, [&](int k) { return col[k * SLB_BSP_PITCH]; }, [&](int k, double v) { col[k * SLB_BSP_PITCH] = v; });
    }
    __syncwarp();
    if (contig) {
        for (int l = 0; l < 32; ++l) {
            long long ln = tile0 + l;
            if (ln < nlines) {
                double* p = out + ln * n;
                for (int k = lane; k < n; k += 32) p[k] = tile[k * SLB_BSP_PITCH + l];
            }
        }
    } else if (active) {
        long long b = line / inner, a = line - b * inner;
        double* p = out + (b * n) * inner + a;
#pragma unroll 8
        for (int k = 0; k < n; ++k) p[(long long)k * inner] = tile[k * SLB_BSP_PITCH + lane];
    }
}

// streaming fallback for lines that do not fit in shared memory: y and x go through `out`
template <int H>
__global__ void __launch_bounds__(128)
k_bspline_stream(const double* __restrict__ in, double* __restrict__ out, long long inner, int n, long long nlines,
                 BsplineDev f)
{
    long long line = (long long)blockIdx.x * 128 + threadIdx.x;
    if (line >= nlines) return;
    long long b = line / inner, a = line - b * inner;
    const double* pi = in + (b * n) * inner + a;
    double* po = out + (b * n) * inner + a;
    const int N = f.N;
    // copy the border entries first: the border step reads u2 through LD(N + r)
    for (int r = 0; r < H; ++r) po[(long long)(N + r) * inner] = pi[(long long)(N + r) * inner];
    // the forward pass reads `in` (phase 0); everything after its last row works in place on `out`
    struct Acc {
        const double* pi;
        double* po;
        long long inner;
        int N;
        int phase;
    } st = {pi, po, inner, N, 0};
    auto LD = [&](int k) -> double {
        if (st.phase == 0 && k < st.N) return __ldg(st.pi + (long long)k * st.inner);
        return st.po[(long long)k * st.inner];
    };
    auto ST = [&](int k, double v) {
        st.po[(long long)k * st.inner] = v;
        if (k == st.N - 1 && st.phase == 0) st.phase = 1;  // forward pass done after its last row
    };
    bspline_solve_line<H>(f, LD, ST);
}

template <int H>
static int bspline_launch(cudaStream_t s, const BsplineDev* f, const double* in, double* out, long long inner, int n,
                          long long outer, int64_t* launches)
{
    long long nlines = inner * outer;
    size_t per_warp = (size_t)n * SLB_BSP_PITCH * sizeof(double);
    static int max_smem = -1;
    if (max_smem < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    if (per_warp <= (size_t)max_smem) {
        int warps = (2 * per_warp <= (size_t)max_smem && 2 * per_warp <= 72 * 1024) ? 2 : 1;
        size_t smem = per_warp * warps;
        cudaFuncSetAttribute(k_bspline_tile<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        long long tiles = (nlines + 31) / 32;
        unsigned blocks = (unsigned)((tiles + warps - 1) / warps);
        k_bspline_tile<H><<<blocks, 32 * warps, smem, s>>>(in, out, inner, n, nlines, *f, inner == 1 ? 1 : 0);
    } else {
        unsigned blocks = (unsigned)((nlines + 127) / 128);
        k_bspline_stream<H><<<blocks, 128, 0, s>>>(in, out, inner, n, nlines, *f);
    }
    (*launches)++;
    return cudaGetLastError() == cudaSuccess ? SLB_OK : SLB_E_CUDA;
}

static int bspline_presolve(cudaStream_t s, const BsplineDev* f, const double* in, double* out, long long inner, int n,
                            long long outer, int64_t* launches)
{
    switch (f->h) {
#define X(H) case H: return bspline_launch<H>(s, f, in, out, inner, n, outer, launches);
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)
#undef X
    }
    return SLB_E_UNSUPPORTED;
}
