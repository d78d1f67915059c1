// slb_bspfused.cuh -- K2+K1 fused: one periodic B-spline sweep (pre-solve c = A^{-1} u AND the
// order+1 point stencil) in ONE pass over HBM: 16 B of traffic per cell-update, as SURVEY.md 8(d)
// assumes for every interpolation kind.
//
// Reference: advection! for a B-spline state = sol(interp, line) (src/bsplinelu.jl:179-220,
// :275-284 | src/bsplinefft.jl:49-58) followed by the periodic stencil
// (src/interpolation.jl:175-193), per line, after a permutedims! (src/advection.jl:372-386).
//
// One WARP owns a tile of 32 lines and keeps it in shared memory for the whole sweep:
//   1. load   : cp.async, 8 bytes per lane and row, every row of the tile in flight at once
//               (strided dims: lanes = 32 neighbouring lines, each row one coalesced 256 B segment,
//                tile pitch 32; dim 0: lines are contiguous, lanes run along the line and the tile
//                is stored transposed with pitch 33 -- conflict-free both ways);
//   2. solve  : thread-per-line bordered banded LU substitution, in place in shared memory
//               (the formulation of slb_bspline.cuh; 4h+1 FMAs per cell).  The factor tables sit in
//               shared memory too, one copy per (persistent) block; each row's record is fetched
//               with 16-byte broadcast loads ONE ROW AHEAD of its use, so the recurrence never
//               waits for a table read (the first version read them from global memory through L1
//               and ran at L1 latency; a second one kept them in the kernel parameters, where the
//               63 uniform registers cannot hold a row ahead);
//   3. stencil: thread-per-line march with a rotating register window fed from shared memory;
//               strided dims store each output row straight to HBM (coalesced); dim 0 writes the
//               outputs back into the tile in place (output i overwrites the oldest window entry,
//               the first order values are kept in registers for the periodic wrap) and
//   4. (dim 0) the warp copies the tile out transposed, coalesced along the lines.
// FP64 work: (4h + 1) + (2h + 2) FMAs per cell -- 33 for order 11, i.e. ~0.5 ms of pure DFMA issue
// for a 128^4 sweep, next to 0.66 ms of HBM time: order 11 is balanced between the FP64 pipe and
// HBM, lower orders are HBM-bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slb_sweep.cuh"
#include "slb_bsprf.cuh"

// Factor tables (device memory, copied to shared memory by every block):
//   forward  record i < N : { L[i][0..h), Ri[i][0..h) }                        2h doubles
//   backward record i < N : { 1/d_i, U[i][0..h)/d_i, G[i][0..h), pad }         2h+2 doubles
//   Sinv[h][h]
// (U is pre-scaled by 1/diag so that the backward recurrence is one FMA deep per row)
struct BspFusedTab {
    int h, n, N;
    int o_bwd, o_S, ndoubles;   // offsets (doubles) of the backward records and of Sinv; total size
};

struct BspFusedArgs {
    const double* in;
    double* out;
    long long inner;   // element stride of the swept index (1: contiguous variant)
    long long nlines;
    int n;
    int nc;            // polynomial coefficients per stencil weight
    int warps;         // warps (tiles in flight) per block
    AlphaMap am;
    OutMap om;         // strided variant: re-shard fused into the stores (plain: kc >= n)
    InMap im;          // contiguous variant: block-major input lines (plain: c == 0)
    double* linesum;   // strided variant, optional: per-line sums of the outputs
    const double* tab_dev;
    BspFusedTab tab;
    int stagger;       // cycles by which consecutive warps of a block delay their first tile (0: none): de-phases the
                       // load / compute cycles of the warps, which otherwise stay in lockstep
    int use_rf;        // 1: the solve is the constant-coefficient recursive-filter cascade (slb_bsprf.cuh);
    BspRfTab rf;       //    tab_dev then holds its table (poles, gain, start-up responses)
};

int slb_bspfused_launch(const BspFusedArgs& a, const CoefTab& ct, int sm_count, cudaStream_t stream);
// warps per block that fit the shared memory next to the tables; 0: unsupported
int slb_bspfused_warps(int h, int n, bool contig);
int slb_bspfused_warps_rf(int ndoubles, int n, bool contig);
bool slb_bspfused_supported(int h, int n);
// host table in the kernel's layout
void slb_bspfused_fill(BspFusedTab* tab, double* v, int h, int n, int N, const double* L, const double* U, const double* invd,
                       const double* Ri, const double* G, const double* Sinv);
int slb_bspfused_tab_doubles(int h, int n);

#ifdef SLB_BSPF_IMPL
__device__ __forceinline__ void bspf_cp_async8(unsigned smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

template <int NV>
__device__ __forceinline__ void bspf_ldrec(double (&dst)[NV], const double* src)
{
    static_assert(NV % 2 == 0, "records are whole 16-byte words");
#pragma unroll
    for (int q = 0; q < NV; q += 2) {
        const double2 v = *reinterpret_cast<const double2*>(src + q);
        dst[q] = v.x;
        dst[q + 1] = v.y;
    }
}

template <int H, bool CONTIG, bool RF>
__global__ void __launch_bounds__(256, 1)
k_bspline_fused(const __grid_constant__ BspFusedArgs fa, const __grid_constant__ CoefTab ct)
{
    constexpr int P1 = 2 * H + 2;           // order + 1 stencil points, order = 2h + 1
    constexpr int PITCH = CONTIG ? 33 : 32;
    constexpr int FR = 2 * H, BR = 2 * H + 2;  // doubles per forward / backward record
    extern __shared__ __align__(16) double bsm[];
    __shared__ int s0s_all[8][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = fa.n, N = fa.tab.N;
    // ---- factor tables: one copy per block --------------------------------------------------------
    const int ntab = RF ? fa.rf.ndoubles : fa.tab.ndoubles;
    double* tabs = bsm;
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) tabs[i] = __ldg(fa.tab_dev + i);
    __syncthreads();
    const double* tF = tabs;
    const double* tB = tabs + fa.tab.o_bwd;
    const double* tS = tabs + fa.tab.o_S;
    double* tile = bsm + ((ntab + 1) & ~1) + (size_t)wid * n * PITCH;
    int* s0s = s0s_all[wid];
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(tile);
    double* col = tile + lane;
    const long long ntiles = (fa.nlines + 31) / 32;
    if (fa.stagger > 0 && wid > 0) {
        const long long t0 = clock64(), wait = (long long)fa.stagger * wid;
        while (clock64() - t0 < wait) __nanosleep(200);
    }

    for (long long t = (long long)blockIdx.x * fa.warps + wid; t < ntiles; t += (long long)gridDim.x * fa.warps) {
    const long long line0 = t * 32;
    const long long line = line0 + lane;
    const bool active = line < fa.nlines;
    const long long lc = active ? line : fa.nlines - 1;

    // ---- 1. load the tile --------------------------------------------------------------------
    long long a = 0, b = 0;
    if (CONTIG) {
        for (int j = 0; j < 32; ++j) {
            const long long lj = line0 + j;
            if (lj < fa.nlines) {
                const double* src = fa.in + slb_in_line(fa.im, lj) * n;
                for (int k = lane; k < n; k += 32) bspf_cp_async8(sbase + 8u * (unsigned)(k * PITCH + j), src + k);
            }
        }
    } else {
        b = lc / fa.inner;
        a = lc - b * fa.inner;
        const double* src = fa.in + (b * n) * fa.inner + a;
        if (active) {
#pragma unroll 8
            for (int k = 0; k < n; ++k) bspf_cp_async8(sbase + 8u * (unsigned)(k * PITCH + lane), src + (long long)k * fa.inner);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // stencil shift and weights of this lane's line (overlaps the loads)
    double w[P1];
    int s0;
    {
        const double alpha = fa.am.scale * __ldg(fa.am.tab + (CONTIG ? slb_alpha_off(fa.am, 0u, (unsigned)lc)
                                                                      : slb_alpha_off(fa.am, (unsigned)a, (unsigned)b)));
        double tt;
        slb_split(alpha, n, (P1 - 1) / 2, tt, s0);
        const int nc = fa.nc;
#pragma unroll
        for (int j = 0; j < P1; ++j) w[j] = ct.c[j * SLB_NCMAX + nc - 1];
        for (int k = nc - 2; k >= 0; --k) {
#pragma unroll
            for (int j = 0; j < P1; ++j) w[j] = fma(tt, w[j], ct.c[j * SLB_NCMAX + k]);
        }
        if (RF) {  // the cascade returns C A^{-1} u: the gain goes into the stencil weights
            const double invC = tabs[H];
#pragma unroll
            for (int j = 0; j < P1; ++j) w[j] *= invC;
        }
    }
    if (CONTIG) s0s[lane] = s0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    if (RF) {
        // ---- 2'. recursive-filter cascade, in place, thread per line (slb_bsprf.cuh): 2h FMAs per cell,
        // poles in registers, no per-row tables
        bsprf_solve_line<H>(fa.rf, tabs, col, PITCH);
    } else {
    // ---- 2. bordered banded LU solve, in place, thread per line ---------------------------------
    // Both recurrences are arranged so that the newest dependency enters LAST: the terms that use
    // older results are summed first, the right-hand side is added, and only one FMA per row waits
    // for the previous row.  Right-hand sides are fetched a group of h rows ahead, table records one
    // row ahead.
    double x2[H];
    {
        double yw[H], acc[H], un[H], Tn[FR];
#pragma unroll
        for (int s = 0; s < H; ++s) {
            yw[s] = 0.0;
            acc[s] = 0.0;
            un[s] = s < N ? col[s * PITCH] : 0.0;
        }
        bspf_ldrec<FR>(Tn, tF);
#define BSPF_FWD_ROW(r, GUARD)                                                                        \
    {                                                                                                 \
        const int i = i0 + (r);                                                                       \
        if (!(GUARD) || i < N) {                                                                      \
            double T[FR];                                                                             \
            _Pragma("unroll") for (int q = 0; q < FR; ++q) T[q] = Tn[q];                              \
            if (!(GUARD) || i + 1 < N) bspf_ldrec<FR>(Tn, tF + (i + 1) * FR);                         \
            double y;                                                                                 \
            if (H >= 2) {                                                                             \
                double sacc = -T[H - 1] * yw[((r) - H + 2 * H) % H];                                  \
                _Pragma("unroll") for (int j = H - 1; j >= 2; --j)                                    \
                    sacc = fma(-T[j - 1], yw[((r) - j + 2 * H) % H], sacc);                           \
                y = fma(-T[0], yw[((r) - 1 + 2 * H) % H], u[(r)] + sacc);                             \
            } else {                                                                                  \
                y = fma(-T[0], yw[0], u[(r)]);                                                        \
            }                                                                                         \
            yw[(r)] = y;                                                                              \
            _Pragma("unroll") for (int q = 0; q < H; ++q) acc[q] = fma(T[H + q], y, acc[q]);          \
            col[i * PITCH] = y;                                                                       \
        }                                                                                             \
    }
        int i0 = 0;
        for (; i0 + 2 * H <= N; i0 += H) {  // whole groups, the next group's right-hand sides exist: no guards
            double u[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                u[r] = un[r];
                un[r] = col[(i0 + H + r) * PITCH];
            }
#pragma unroll
            for (int r = 0; r < H; ++r) BSPF_FWD_ROW(r, false)
        }
        for (; i0 < N; i0 += H) {
            double u[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                u[r] = un[r];
                const int inext = i0 + H + r;
                un[r] = inext < N ? col[inext * PITCH] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < H; ++r) BSPF_FWD_ROW(r, true)
        }
#undef BSPF_FWD_ROW
        double rhs[H];
#pragma unroll
        for (int r = 0; r < H; ++r) rhs[r] = col[(N + r) * PITCH] - acc[r];
#pragma unroll
        for (int q = 0; q < H; ++q) {
            double sq = 0.0;
#pragma unroll
            for (int r = 0; r < H; ++r) sq = fma(tS[q * H + r], rhs[r], sq);
            x2[q] = sq;
        }
#pragma unroll
        for (int q = 0; q < H; ++q) col[(N + q) * PITCH] = x2[q];
    }
    {
        double ww[H], un[H], Tn[BR];
        const int ilast = ((N - 1) / H) * H;
#pragma unroll
        for (int s = 0; s < H; ++s) {
            ww[s] = 0.0;
            un[s] = ilast + s < N ? col[(ilast + s) * PITCH] : 0.0;
        }
        bspf_ldrec<BR>(Tn, tB + (N - 1) * BR);
        // record layout: T[0] = 1/d, T[1..h] = U/d, T[1+h..2h] = G
#define BSPF_BWD_ROW(r, GUARD)                                                                        \
    {                                                                                                 \
        const int i = i0 + (r);                                                                       \
        if (!(GUARD) || i < N) {                                                                      \
            double T[BR];                                                                             \
            _Pragma("unroll") for (int q = 0; q < BR; ++q) T[q] = Tn[q];                              \
            bspf_ldrec<BR>(Tn, tB + (i > 0 ? i - 1 : 0) * BR); /* clamped, not predicated: no MOVs */ \
            double v; /* w_i = (y_i - sum_j U[i][j-1] w_{i+j}) / d_i */                               \
            if (H >= 2) {                                                                             \
                double sacc = -T[1 + H - 1] * ww[((r) + H) % H];                                      \
                _Pragma("unroll") for (int j = H - 1; j >= 2; --j)                                    \
                    sacc = fma(-T[1 + j - 1], ww[((r) + j) % H], sacc);                               \
                v = fma(-T[1], ww[((r) + 1) % H], fma(T[0], u[(r)], sacc));                           \
            } else {                                                                                  \
                v = fma(-T[1], ww[0], T[0] * u[(r)]);                                                 \
            }                                                                                         \
            ww[(r)] = v;                                                                              \
            double x = v;                                                                             \
            _Pragma("unroll") for (int q = 0; q < H; ++q) x = fma(-T[1 + H + q], x2[q], x);           \
            col[i * PITCH] = x;                                                                       \
        }                                                                                             \
    }
        const int ngroups = ilast / H + 1;
        {   // the last (possibly ragged) group
            const int i0 = ilast;
            double u[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                u[r] = un[r];
                const int inext = i0 - H + r;
                un[r] = inext >= 0 ? col[inext * PITCH] : 0.0;
            }
#pragma unroll
            for (int r = H - 1; r >= 0; --r) BSPF_BWD_ROW(r, true)
        }
        for (int gq = 1; gq < ngroups; ++gq) {  // whole groups
            const int i0 = (ngroups - 1 - gq) * H;
            double u[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                u[r] = un[r];
                un[r] = col[((i0 >= H ? i0 - H : 0) + r) * PITCH];  // clamped: the last group's prefetch is unused
            }
#pragma unroll
            for (int r = H - 1; r >= 0; --r) BSPF_BWD_ROW(r, false)
        }
#undef BSPF_BWD_ROW
    }
    }  // !RF

    // ---- 3. stencil, thread per line ---------------------------------------------------------------
    // window slot convention of slb_dot: logical element j of output i lives in win[(i + j) % P1]
    double win[P1];
    int kk = s0;
#define BSPF_NEXT(dst)                   \
    {                                    \
        dst = col[kk * PITCH];           \
        kk = kk + 1 == n ? 0 : kk + 1;   \
    }
#pragma unroll
    for (int j = 0; j < P1 - 1; ++j) BSPF_NEXT(win[j]);
    if (!CONTIG) {
        const long long ooff = b * fa.om.bstride + a;
        double* po = (fa.om.npeer > 0 ? fa.om.blk[0] : fa.out) + ooff;
        double lsum = 0.0;
        if (fa.om.kc >= n) {  // plain layout
            int i0 = 0;
            for (; i0 + P1 <= n; i0 += P1) {  // whole groups: no guards, the order+1 dot products overlap
                double nx[P1];
#pragma unroll
                for (int r = 0; r < P1; ++r) BSPF_NEXT(nx[r]);
#pragma unroll
                for (int r = 0; r < P1; ++r) {
                    win[(r + P1 - 1) % P1] = nx[r];
                    const double acc = slb_dot<P1, false>(win, w, r);
                    lsum += acc;
                    if (active) po[(long long)r * fa.inner] = acc;
                }
                po += (long long)P1 * fa.inner;
            }
            for (; i0 < n; i0 += P1) {
#pragma unroll
                for (int r = 0; r < P1; ++r) {
                    if (i0 + r < n) {
                        BSPF_NEXT(win[(r + P1 - 1) % P1]);
                        const double acc = slb_dot<P1, false>(win, w, r);
                        lsum += acc;
                        if (active) *po = acc;
                        po += fa.inner;
                    }
                }
            }
        } else {              // output blocked along the swept index (multi-GPU re-shard)
            int ko = 0, kb = 0;
            for (int i0 = 0; i0 < n; i0 += P1) {
#pragma unroll
                for (int r = 0; r < P1; ++r) {
                    if (i0 + r < n) {
                        BSPF_NEXT(win[(r + P1 - 1) % P1]);
                        const double acc = slb_dot<P1, false>(win, w, r);
                        lsum += acc;
                        if (active) *po = acc;
                        po += fa.inner;
                        if (++ko == fa.om.kc) {
                            ko = 0;
                            ++kb;
                            if (fa.om.npeer > 0)
                                po = fa.om.blk[kb < fa.om.npeer ? kb : 0] + ooff;
                            else
                                po += fa.om.kblk - (long long)fa.om.kc * fa.inner;
                        }
                    }
                }
            }
        }
        if (fa.linesum && active) fa.linesum[line] = lsum;
    } else {
        // in place: output i replaces the oldest window entry c[(i + s0) mod n]; the first order
        // entries are needed again by the last outputs (periodic wrap) and stay in registers
        double head[P1 - 1];
#pragma unroll
        for (int j = 0; j < P1 - 1; ++j) head[j] = win[j];
        int kw = s0;  // where output i goes
        const int nmain = n - (P1 - 1);
        int i0 = 0;
        for (; i0 + P1 <= nmain; i0 += P1) {  // whole groups: no guards, the dot products overlap
            double nx[P1];
#pragma unroll
            for (int r = 0; r < P1; ++r) BSPF_NEXT(nx[r]);
#pragma unroll
            for (int r = 0; r < P1; ++r) {
                win[(r + P1 - 1) % P1] = nx[r];
                col[kw * PITCH] = slb_dot<P1, false>(win, w, r);
                kw = kw + 1 == n ? 0 : kw + 1;
            }
        }
        for (; i0 < nmain; i0 += P1) {
#pragma unroll
            for (int r = 0; r < P1; ++r) {
                if (i0 + r < nmain) {
                    BSPF_NEXT(win[(r + P1 - 1) % P1]);
                    col[kw * PITCH] = slb_dot<P1, false>(win, w, r);
                    kw = kw + 1 == n ? 0 : kw + 1;
                }
            }
        }
        // the last order outputs read their inputs afresh: entries not yet overwritten come from the
        // tile, the wrapped ones (the first order entries of the line's window) from `head`
        {
            int kr = kw;  // position of c[(s0 + nmain) mod n], the oldest input of output nmain
#pragma unroll
            for (int q = 0; q < P1 - 1; ++q) {
                double xs[P1];
                int kq = kr;
#pragma unroll
                for (int j = 0; j < P1; ++j) {
                    if (q + j < P1 - 1) {
                        xs[j] = col[kq * PITCH];
                        kq = kq + 1 == n ? 0 : kq + 1;
                    } else {
                        xs[j] = head[q + j - (P1 - 1)];
                    }
                }
                col[kr * PITCH] = slb_dot<P1, false>(xs, w, 0);
                kr = kr + 1 == n ? 0 : kr + 1;
            }
        }
        __syncwarp();
        // ---- 4. transposed, coalesced write-out ------------------------------------------------------
        for (int j = 0; j < 32; ++j) {
            const long long lj = line0 + j;
            if (lj < fa.nlines) {
                double* dst = fa.out + lj * n;
                int pos = s0s[j] + lane;
                while (pos >= n) pos -= n;
                const int step = 32 % n;  // == 32 for every n > 32: no division in the loop
                for (int k = lane; k < n; k += 32) {
                    __stcs(dst + k, tile[pos * PITCH + j]);
                    pos += step;
                    pos -= pos >= n ? n : 0;
                }
            }
        }
    }
#undef BSPF_NEXT
    __syncwarp();  // the tile is reused by this warp's next iteration
    }  // persistent loop over tiles
}
#endif  // SLB_BSPF_IMPL
