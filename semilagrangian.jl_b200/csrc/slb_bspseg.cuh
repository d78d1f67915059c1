// slb_bspseg.cuh -- K2+K1 fused, SEGMENTED: one periodic B-spline sweep (pre-solve c = A^{-1} u AND the order+1
// point stencil) in one pass over HBM, with every line worked on by S threads at once and the line data in
// REGISTERS, not in shared memory.
//
// Reference: sol(interp, line) (src/bsplinelu.jl:179-220, :275-284 | src/bsplinefft.jl:49-58) + the periodic
// stencil (src/interpolation.jl:175-193), per line (src/advection.jl:627-631).
//
// Why: the thread-per-line kernels (slb_bspfused.cuh) must park a whole line in shared memory between the forward
// and the backward substitution -- 1 KB per 128-point line, so 7 warps per SM and a latency-bound march (ncu: issue
// slots 40 % busy, FP64 27 %, DRAM 34 %; 1.55 ms per 128^4 sweep against a 0.66 ms HBM floor; a 1024-point line
// does not fit at all and ran at CPU speed).  The pre-solve is a cascade of first-order recursive filters
// (slb_bsprf.cuh): y[i] = x[i] + z y[i-1].  A first-order recurrence splits exactly: cut the periodic line into S
// segments of M rows; every segment runs the recurrence from a ZERO state (M FMAs, registers), publishes its last
// value L_s, and the true state entering segment s is the geometric sum
//       E_{s-1} = ( sum_{m >= 0} z^(M m) L_{s-1-m} ) / (1 - z^n)          (indices around the ring)
// -- for M = 16 and the poles of order 11 (|z| <= 0.66) the terms fall below 1e-19 after 7 | 3 | 2 | 2 | 1
// segments -- after which element j of the segment is corrected by z^(j+1) E (one FMA with a constant-bank
// operand).  2 FMAs per row and stage instead of 1, but S threads per line, no shared-memory tile and no per-row
// tables: 16 to 32 warps per SM instead of 7.
//
//   k_bspline_seg<H, M, CONTIG> : block = S warps; lane = one of 32 neighbouring lines, warp = one segment of M
//       rows (n = S M, S <= 8).  Strided dims load rows straight into registers (coalesced 256 B per warp and
//       row); dim 0 stages the 32 lines through a transposed shared-memory tile for coalescing.  Carries and the
//       stencil's order-point overlap with the next segment go through small shared-memory exchanges (one block
//       barrier per filter stage).  The stencil is evaluated UNSHIFTED on the segment and the periodic shift
//       floor(alpha) is applied to the row index of the store (dim 0: inside the tile), so no thread ever needs
//       rows of another segment beyond the stencil overlap.
//   k_bspline_wline<H, M>       : long lines (n = 32 M, 512 <= n <= 2048, e.g. the 1024 x 1024 rotation of
//       BASELINE config 2): one WARP per line, lane = segment, carries and overlaps by warp shuffles.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slb_sweep.cuh"
#include "slb_bsprf.cuh"

#define SLB_SEG_HMAX 5      // orders 3 .. 11
#define SLB_SEG_MMAX 64
#define SLB_SEG_SMAX 32

struct BspSegTab {
    int h, n, M, S;
    double invC;
    double z[SLB_SEG_HMAX];
    int nm[SLB_SEG_HMAX];                    // carry terms kept per stage (<= S)
    double zp[SLB_SEG_HMAX][SLB_SEG_MMAX];   // z_k^(j+1)
    double cz[SLB_SEG_HMAX][SLB_SEG_SMAX];   // z_k^(M m) / (1 - z_k^n)
};

struct BspSegArgs {
    const double* in;
    double* out;
    long long inner;   // element stride of the swept index (1: contiguous variant)
    long long nlines;
    int n;
    int nc;            // polynomial coefficients per stencil weight
    AlphaMap am;
    OutMap om;         // strided variant: re-shard fused into the stores (plain: kc >= n)
    InMap im;          // contiguous variant: block-major input lines (plain: c == 0)
    double* linesum;   // strided variant, optional: per-line sums of the outputs
    BspSegTab tab;
};

// Rows of a segment whose correction z_k^(j+1) E is applied: the poles of the order-(2H+1) B-spline symbol are universal
// constants (stage k = k-th smallest |z|), and |z_k|^(j+1) < 1e-19 -- the threshold the carry sums use as well -- beyond
// these (even) counts; slb_bspseg_plan verifies them against the actual poles.  Order 11, M = 32: 6 + 12 + 20 + 32 + 32
// of 5 x 32 corrections per direction remain.
__host__ __device__ constexpr int slb_seg_cut(int H, int k, int M)
{
    const int c = H == 1 ? 34
                : H == 2 ? (k == 0 ? 14 : 52)
                : H == 3 ? (k == 0 ? 10 : k == 1 ? 22 : 70)
                : H == 4 ? (k == 0 ? 8 : k == 1 ? 14 : k == 2 ? 28 : 88)
                         : (k == 0 ? 6 : k == 1 ? 12 : k == 2 ? 20 : k == 3 ? 34 : 106);
    return c < M ? c : M;
}

// host: table from the recursive-filter factorisation; returns false when (order, n) is not on this path
bool slb_bspseg_plan(const BspRfHost& hr, bool wline, bool contig, BspSegTab* tab);
// returns 0 on success, -1 when no kernel is instantiated for the combination, else cudaGetLastError()
int slb_bspseg_launch(const BspSegArgs& a, const CoefTab& ct, bool contig, cudaStream_t stream);
int slb_bspwline_launch(const BspSegArgs& a, const CoefTab& ct, cudaStream_t stream);

#ifdef SLB_BSPSEG_IMPL
// stencil weight j of this lane's line (Horner over the nc coefficients; j is a run-time index)
__device__ __forceinline__ double bspseg_weight(const CoefTab& ct, int nc, int j, double t)
{
    double wj = ct.c[j * SLB_NCMAX + nc - 1];
    for (int k = nc - 2; k >= 0; --k) wj = fma(t, wj, ct.c[j * SLB_NCMAX + k]);
    return wj;
}

// address of output row i of the line that starts at `base` (strided variant; plain or block-major / peer layout)
__device__ __forceinline__ double* bspseg_out_ptr(const OutMap& om, double* out, long long base, long long inner, int n, int i)
{
    if (om.kc >= n) return out + base + (long long)i * inner;
    const int q = i / om.kc, r = i - q * om.kc;
    double* blk = om.npeer > 0 ? om.blk[q] : out + (long long)q * om.kblk;
    return blk + base + (long long)r * inner;
}

// MINB: resident blocks per SM the register allocation is capped for (1: no cap)
template <int H, int M, int S, bool CONTIG, int MINB>
__global__ void __launch_bounds__(32 * S, MINB) k_bspline_seg(const __grid_constant__ BspSegArgs fa, const __grid_constant__ CoefTab ct)
{
    constexpr int P1 = 2 * H + 2, HALO = P1 - 1, HM = HALO < M ? HALO : M;  // HM: values a segment lends to its predecessors
    constexpr int TP = 33;                                                    // tile pitch (dim 0)
    static_assert(M % 2 == 0, "constants are fetched in pairs");
    extern __shared__ __align__(16) double ssm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int n = M * S;
    double* czs = ssm;                       // [H][S]   z_k^(M m) / (1 - z_k^n), zero beyond the kept terms
    double* zpl = czs + H * S;               // [H][M]   z_k^(j+1)
    double* tts = zpl + H * M;               // [32]      fractional shift of each line
    int* s0s = reinterpret_cast<int*>(tts + 32);  // [32] start index of each line's stencil window (64 ints reserved)
    double* Lbuf = tts + 64;                 // [2][S][32]
    double* wsm = Lbuf + 2 * S * 32;         // [P1][32]
    double* hal = wsm + P1 * 32;             // [S][HM][32]
    double* lsb = hal + S * HM * 32;         // [S][32]   line-sum partials
    // dim 0: the transposed tile [n][33] SHARES its memory with the exchange buffers above (it is idle between the moment
    // every thread has its rows in registers and the moment the outputs go back into it; two extra block barriers) --
    // 36 KB instead of 53 KB per block: the shared memory no longer caps an SM at 4 blocks
    double* tile = Lbuf;
    for (int q = threadIdx.x; q < H * S; q += 32 * S) czs[q] = (q % S) < fa.tab.nm[q / S] ? fa.tab.cz[q / S][q % S] : 0.0;
    for (int q = threadIdx.x; q < H * M; q += 32 * S) zpl[q] = fa.tab.zp[q / M][q % M];
    const long long line0 = (long long)blockIdx.x * 32;
    const long long line = line0 + lane;
    const bool active = line < fa.nlines;
    const long long lc = active ? line : fa.nlines - 1;
    long long a = 0, b = 0;
    if (!CONTIG) {
        if (fa.nlines < 0x7fffffffLL) {  // the usual case: 32-bit division
            const unsigned bq = (unsigned)lc / (unsigned)fa.inner;
            b = bq;
            a = (unsigned)lc - bq * (unsigned)fa.inner;
        } else {
            b = lc / fa.inner;
            a = lc - b * fa.inner;
        }
    }
    // ---- load: M rows of this lane's line into registers ------------------------------------------------------
    double v[M];
    if (CONTIG) {
        // the 32 lines of the tile are contiguous runs of n doubles: coalesced reads, transposed into the tile.
        // (Reading M consecutive rows per thread straight from global memory -- 32 lines per warp request, half a
        // sector each, the other half left to L1 -- was not faster: order 11 1.48 -> 1.50 ms, order 5 1.04 -> 1.21.)
        for (int j = w; j < 32; j += S) {
            const long long lj = line0 + j;
            if (lj < fa.nlines) {
                const double* src = fa.in + slb_in_line(fa.im, lj) * n;
#pragma unroll
                for (int k = lane; k < n; k += 32) tile[k * TP + j] = __ldg(src + k);
            }
        }
    } else {
        const double* src = fa.in + (b * n) * fa.inner + a + (long long)(w * M) * fa.inner;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            v[j] = active ? __ldg(src) : 0.0;
            src += fa.inner;
        }
    }
    // ---- shift of every line: warp 0 evaluates it for the block ------------------------------------------------
    if (w == 0) {
        const double alpha = fa.am.scale * __ldg(fa.am.tab + (CONTIG ? slb_alpha_off(fa.am, 0u, (unsigned)lc)
                                                                      : slb_alpha_off(fa.am, (unsigned)a, (unsigned)b)));
        double tt;
        int s0;
        slb_split(alpha, n, (P1 - 1) / 2, tt, s0);
        tts[lane] = tt;
        s0s[lane] = s0;
    }
    __syncthreads();  // the constants (and the tile) are in shared memory
    if (CONTIG) {
#pragma unroll
        for (int j = 0; j < M; ++j) v[j] = tile[(w * M + j) * TP + lane];
        __syncthreads();  // the tile's memory now serves the exchanges
    }
    // ---- recursive-filter cascade ------------------------------------------------------------------------------
    // Every stage: the segment runs the recurrence from a ZERO state (a chain of M - 1 FMAs), publishes the value that
    // leaves it, and after the exchange row j is corrected by z^(j+1) times the state that really entered the segment.
    // (A variant that applied the same split once more to sub-segments of 4 rows -- shorter chains, 6 constants per
    // stage -- was slower: 1.32 vs 1.15 ms at order 11; the kernel is bound by instruction issue, not by latency.)
#pragma unroll
    for (int k = 0; k < 2 * H; ++k) {
        const bool causal = k < H;           // causal stages y[i] = x[i] + z y[i-1], then anticausal y[i] = x[i] + z y[i+1]
        const int kk = causal ? k : k - H;
#define SLB_IX(t) (causal ? (t) : (M - 1 - (t)))  // the anticausal stages are the causal ones on the reversed segment
        const double z = fa.tab.z[kk];
#pragma unroll
        for (int j = 1; j < M; ++j) v[SLB_IX(j)] = fma(z, v[SLB_IX(j - 1)], v[SLB_IX(j)]);
        double* Lb = Lbuf + (k & 1) * S * 32;
        Lb[w * 32 + lane] = v[SLB_IX(M - 1)];
        __syncthreads();
        if (k == 0) {
            // stencil weights: the S warps share the Horner evaluations (the shift became visible with this barrier)
            const double tt = tts[lane];
            for (int j = w; j < P1; j += S) wsm[j * 32 + lane] = bspseg_weight(ct, fa.nc, j, tt) * fa.tab.invC;
        }
        // state entering this segment: geometric sum over the preceding (causal) / following (anticausal) segments
        double c = 0.0;
#pragma unroll
        for (int m = 0; m < S; ++m) {
            int ws = causal ? w - 1 - m : w + 1 + m;
            ws = ws < 0 ? ws + S : (ws >= S ? ws - S : ws);
            c = fma(czs[kk * S + m], Lb[ws * 32 + lane], c);
        }
        const double2* zp2 = reinterpret_cast<const double2*>(zpl + kk * M);
        const int cut = slb_seg_cut(H, kk, M);   // folds per unrolled stage: rows beyond it get corrections below 1e-19
#pragma unroll
        for (int j = 0; j < M; j += 2) {
            if (j < cut) {
                const double2 pz = zp2[j >> 1];
                v[SLB_IX(j)] = fma(c, pz.x, v[SLB_IX(j)]);
                v[SLB_IX(j + 1)] = fma(c, pz.y, v[SLB_IX(j + 1)]);
            }
        }
#undef SLB_IX
    }
    // ---- stencil on the segment: d[i'] = sum_q w_q c[i' + q]; the overlap comes from the following segment(s) ----
#pragma unroll
    for (int j = 0; j < HM; ++j) hal[(w * HM + j) * 32 + lane] = v[j];
    __syncthreads();
    double x[M + HALO];
#pragma unroll
    for (int j = 0; j < M; ++j) x[j] = v[j];
#pragma unroll
    for (int j = 0; j < HALO; ++j) {
        int wn = w + 1 + j / M;
        wn = wn >= S ? wn % S : wn;
        x[M + j] = hal[(wn * HM + (j % M)) * 32 + lane];
    }
    double wt[P1];
#pragma unroll
    for (int q = 0; q < P1; ++q) wt[q] = wsm[q * 32 + lane];
#pragma unroll
    for (int j = 0; j < M; ++j) {
        double acc = x[j] * wt[0];
#pragma unroll
        for (int q = 1; q < P1; ++q) acc = fma(x[j + q], wt[q], acc);
        v[j] = acc;
    }
    // output j of this segment is row (w M + j - s0) mod n:  out[i] = sum_q c[(i + s0 + q) mod n] w_q
    int i0 = w * M - s0s[lane];
    i0 = i0 < 0 ? i0 + n : i0;
    if (CONTIG) {
        // outputs go back into the tile at their shifted rows (once everybody is done with the exchange buffers that
        // share its memory), then out coalesced along the lines
        __syncthreads();
        int ii = i0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            tile[ii * TP + lane] = v[j];
            ii = ii + 1 == n ? 0 : ii + 1;
        }
        __syncthreads();
        for (int j = w; j < 32; j += S) {
            const long long lj = line0 + j;
            if (lj < fa.nlines) {
                double* dst = fa.out + lj * n;
#pragma unroll
                for (int k = lane; k < n; k += 32) __stcs(dst + k, tile[k * TP + j]);
            }
        }
    } else {
        const long long obase = b * fa.om.bstride + a;
        if (fa.om.kc >= n) {  // plain layout: a running pointer with the periodic wrap
            double* const row0 = fa.out + obase;
            double* po = row0 + (long long)i0 * fa.inner;
            int left = n - i0;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                if (active) __stcs(po, v[j]);
                po += fa.inner;
                if (--left == 0) po = row0;
            }
        } else {  // block-major / peer blocks (multi-GPU re-shard fused into the stores)
            int q = i0 / fa.om.kc, r = i0 - q * fa.om.kc;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                double* blk = fa.om.npeer > 0 ? fa.om.blk[q] : fa.out + (long long)q * fa.om.kblk;
                if (active) __stcs(blk + obase + (long long)r * fa.inner, v[j]);
                if (++r == fa.om.kc) {
                    r = 0;
                    q = (q + 1) * fa.om.kc >= n ? 0 : q + 1;
                }
            }
        }
        if (fa.linesum) {
            double lsum = 0.0;
#pragma unroll
            for (int j = 0; j < M; ++j) lsum += v[j];
            lsb[w * 32 + lane] = lsum;
            __syncthreads();
            if (w == 0 && active) {
                double sacc = lsb[lane];
#pragma unroll
                for (int q = 1; q < S; ++q) sacc += lsb[q * 32 + lane];
                fa.linesum[line] = sacc;
            }
        }
    }
}

// ---- long lines: one warp per line, lane = segment of M rows ------------------------------------------------------
__device__ __forceinline__ double bspseg_shfl(double x, int src)
{
    return __shfl_sync(0xffffffffu, x, src);
}

template <int H, int M>
__global__ void __launch_bounds__(128) k_bspline_wline(const __grid_constant__ BspSegArgs fa, const __grid_constant__ CoefTab ct)
{
    constexpr int P1 = 2 * H + 2, HALO = P1 - 1;
    static_assert(HALO <= M, "the stencil overlap must come from the next segment only");
    const int lane = threadIdx.x & 31;
    const int n = fa.n;
    const long long line = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (line >= fa.nlines) return;  // whole warps leave together
    const long long b = line / fa.inner, a = line - b * fa.inner;
    const long long st = fa.inner;
    double v[M];
    {
        const double* src = fa.in + (b * n) * st + a + (long long)(lane * M) * st;
        if (st == 1) {
#pragma unroll
            for (int j = 0; j < M; j += 2) {
                const double2 t2 = __ldg(reinterpret_cast<const double2*>(src + j));
                v[j] = t2.x;
                v[j + 1] = t2.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < M; ++j) v[j] = __ldg(src + (long long)j * st);
        }
    }
    int s0;
    double wt[P1];
    {
        const double alpha = fa.am.scale * __ldg(fa.am.tab + slb_alpha_off(fa.am, (unsigned)a, (unsigned)b));
        double tt;
        slb_split(alpha, n, (P1 - 1) / 2, tt, s0);
        double mine = 0.0;  // lane q < P1 evaluates weight q, then everybody collects them
        if (lane < P1) mine = bspseg_weight(ct, fa.nc, lane, tt) * fa.tab.invC;
#pragma unroll
        for (int q = 0; q < P1; ++q) wt[q] = bspseg_shfl(mine, q);
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const double z = fa.tab.z[k];
#pragma unroll
        for (int j = 1; j < M; ++j) v[j] = fma(z, v[j - 1], v[j]);
        const double L = v[M - 1];
        double c = 0.0;
        const int nm = fa.tab.nm[k];
        for (int m = 0; m < nm; ++m) c = fma(fa.tab.cz[k][m], bspseg_shfl(L, (lane - 1 - m) & 31), c);
#pragma unroll
        for (int j = 0; j < M; ++j) v[j] = fma(c, fa.tab.zp[k][j], v[j]);
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const double z = fa.tab.z[k];
#pragma unroll
        for (int j = M - 2; j >= 0; --j) v[j] = fma(z, v[j + 1], v[j]);
        const double L = v[0];
        double c = 0.0;
        const int nm = fa.tab.nm[k];
        for (int m = 0; m < nm; ++m) c = fma(fa.tab.cz[k][m], bspseg_shfl(L, (lane + 1 + m) & 31), c);
#pragma unroll
        for (int j = 0; j < M; ++j) v[j] = fma(c, fa.tab.zp[k][M - 1 - j], v[j]);
    }
    double x[M + HALO];
#pragma unroll
    for (int j = 0; j < M; ++j) x[j] = v[j];
#pragma unroll
    for (int j = 0; j < HALO; ++j) x[M + j] = bspseg_shfl(v[j], (lane + 1) & 31);
    int i = lane * M - s0;
    i = i < 0 ? i + n : i;
    double* po = fa.out + (b * n) * st + a;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        double acc = x[j] * wt[0];
#pragma unroll
        for (int q = 1; q < P1; ++q) acc = fma(x[j + q], wt[q], acc);
        po[(long long)i * st] = acc;
        i = i + 1 == n ? 0 : i + 1;
    }
}
#endif  // SLB_BSPSEG_IMPL
