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

// alpha = scale * tab[ sum_d idx_d * stride_d ] over the dims below (lo) / above (hi) the swept one.
// The host compresses the per-dim (extent, stride) list into terms idx = (x / div) % ext (zero-stride
// dims dropped, layout-compatible neighbours merged, ext == 0: no modulo needed), so the usual
// Vlasov cases cost at most one 32-bit division per line.
struct AlphaTerm {
    unsigned div, ext;
    long long stride;
};
struct AlphaMap {
    const double* tab;
    double scale;
    int nlo, nhi;
    AlphaTerm lo[SLB_MAXD], hi[SLB_MAXD];
};

// Generalised addressing of the swept index k on the OUTPUT side (multi-GPU re-shard fused
// into the sweep epilogue): k -> (k % kc) * inner + (k / kc) * kblk.  kc == n: plain layout.
#define SLB_MAX_PEERS 16
struct OutMap {
    int kc;
    long long kblk;
    long long bstride;  // offset between consecutive outer indices b (plain: n * inner)
    // Fused sweep + all-to-all: when npeer > 0, k-block q of the output is not written to
    // out + q*kblk but to blk[q], a pointer into the destination rank's buffer (its own HBM or a
    // peer's, mapped through CUDA IPC): the sweep's coalesced row stores travel over NVLink
    // directly and no separate collective or pack pass exists.
    int npeer;
    double* blk[SLB_MAX_PEERS];
};

// Generalised addressing of the INPUT lines of a contiguous-dim sweep (multi-GPU re-shard fused
// into the sweep prologue): the input is stored block-major along a non-swept dim `bdim` (as an
// all-to-all leaves it): line (lo, q, hi) with q the index along bdim lives at line index
//   (q / c) * blk_lines + lo + L * ((q % c) + c * hi).          c == 0: plain layout.
struct InMap {
    unsigned L, q_ext, c;
    long long blk_lines;
};

__device__ __forceinline__ long long slb_in_line(const InMap& im, long long line)
{
    if (im.c == 0) return line;
    unsigned l = (unsigned)line;
    unsigned lo = im.L > 1 ? l % im.L : 0u;
    unsigned t = im.L > 1 ? l / im.L : l;
    unsigned q = t % im.q_ext;
    unsigned hi = t / im.q_ext;
    unsigned blk = q / im.c;
    unsigned qi = q - blk * im.c;
    return (long long)blk * im.blk_lines + lo + (long long)im.L * (qi + (long long)im.c * hi);
}

__device__ __forceinline__ long long slb_alpha_off(const AlphaMap& m, unsigned a, unsigned b)
{
    long long off = 0;
#pragma unroll 1
    for (int d = 0; d < m.nlo; ++d) {
        unsigned q = m.lo[d].div > 1 ? a / m.lo[d].div : a;
        if (m.lo[d].ext) q %= m.lo[d].ext;
        off += (long long)q * m.lo[d].stride;
    }
#pragma unroll 1
    for (int d = 0; d < m.nhi; ++d) {
        unsigned q = m.hi[d].div > 1 ? b / m.hi[d].div : b;
        if (m.hi[d].ext) q %= m.hi[d].ext;
        off += (long long)q * m.hi[d].stride;
    }
    return off;
}

// floor/frac split of src/interpolation.jl:381-389 and the start index of the periodic window
__device__ __forceinline__ void slb_split(double alpha, int n, int half, double& t, int& s0)
{
    double fl = floor(alpha);
    t = alpha - fl;
    if (fabs(fl) < 1.0e9) {  // the usual case: 32-bit arithmetic
        int r = ((int)fl - half) % n;
        s0 = r < 0 ? r + n : r;
    } else {
        long long d = (long long)fl - half;
        long long r = d % n;
        s0 = (int)(r < 0 ? r + n : r);
    }
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
                AlphaMap am, CoefTab ct, int nc, OutMap om, double* __restrict__ linesum)
{
    long long gid = (long long)blockIdx.x * 128 + threadIdx.x;
    if (gid >= nlines) return;
    long long b = gid / inner;
    long long a = gid - b * inner;

    double alpha = am.scale * __ldg(am.tab + slb_alpha_off(am, (unsigned)a, (unsigned)b));
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
    const long long ooff = b * om.bstride + a;
    double* pout = (om.npeer > 0 ? om.blk[0] : out) + ooff;
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
    double lsum = 0.0;  // sum of this line's outputs: feeds the charge density without another pass over f
    double* po = pout;
    int ko = 0;  // position inside the current output k-block
    int kb = 0;  // current output k-block

#define SLB_STORE(val)                                  \
    {                                                   \
        *po = (val);                                    \
        po += inner;                                    \
        if (!plain) {                                   \
            if (++ko == om.kc) {                        \
                ko = 0;                                 \
                ++kb;                                   \
                if (om.npeer > 0)                       \
                    po = om.blk[kb < om.npeer ? kb : 0] + ooff; \
                else                                    \
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
            lsum += acc;                                    \
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
    if (linesum) linesum[gid] = lsum;
#undef SLB_COMPUTE_GROUP
#undef SLB_STORE
#undef SLB_LOAD_NEXT
}

// ------------------------------------------------------------------------------------------
// K1 strided, FEW lines (small 2-D grids: the 1D1V example has 128 lines of 256 points): one thread per line leaves
// a B200 with a single block of 128 threads marching 256 rows one after the other (37 us per sweep, 70 % of a
// 128 x 256 Vlasov-Poisson step).  A Lagrange / Hermite stencil has no recurrence, so the line is cut into chunks of
// CH outputs: block = 32 neighbouring lines x NCH chunks, every thread loads its CH + order inputs at once and
// evaluates CH outputs from registers.  Same operation order per output as k_sweep_strided (bit-identical results);
// the line sums are reduced over the chunks through shared memory in a fixed order.
// ------------------------------------------------------------------------------------------
template <int P1, bool EXACT, int CH>
__global__ void __launch_bounds__(512)
k_sweep_strided_chunk(const double* __restrict__ in, double* __restrict__ out, long long inner, int n, long long nlines,
                      AlphaMap am, CoefTab ct, int nc, double* __restrict__ linesum)
{
    __shared__ double lsb[16][33];
    const int lane = threadIdx.x, ch = threadIdx.y, nch = blockDim.y;  // nch <= 16
    const long long gid = (long long)blockIdx.x * 32 + lane;
    const bool active = gid < nlines;
    const long long gc = active ? gid : nlines - 1;
    const long long b = gc / inner, a = gc - b * inner;
    const double alpha = am.scale * __ldg(am.tab + slb_alpha_off(am, (unsigned)a, (unsigned)b));
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
    const int i0 = ch * CH;                 // first output of this chunk
    const int cnt = n - i0 < CH ? n - i0 : CH;
    const double* pin = in + (b * n) * inner + a;
    double x[CH + P1 - 1];
    int kk = (s0 + i0) % n;
#pragma unroll
    for (int j = 0; j < CH + P1 - 1; ++j) {
        x[j] = (j < cnt + P1 - 1) ? __ldg(pin + (long long)kk * inner) : 0.0;
        kk = kk + 1 == n ? 0 : kk + 1;
    }
    double* po = out + (b * n) * inner + a + (long long)i0 * inner;
    double lsum = 0.0;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
        double acc;
        if (EXACT) {
            acc = __dmul_rn(x[j], w[0]);
#pragma unroll
            for (int q = 1; q < P1; ++q) acc = __dadd_rn(acc, __dmul_rn(x[j + q], w[q]));
        } else {
            acc = x[j] * w[0];
#pragma unroll
            for (int q = 1; q < P1; ++q) acc = fma(x[j + q], w[q], acc);
        }
        if (j < cnt) {
            lsum += acc;
            if (active) po[(long long)j * inner] = acc;
        }
    }
    if (linesum) {
        lsb[ch][lane] = lsum;
        __syncthreads();
        if (ch == 0 && active) {
            double sacc = lsb[0][lane];
            for (int q = 1; q < nch; ++q) sacc += lsb[q][lane];
            linesum[gid] = sacc;
        }
    }
}

// ------------------------------------------------------------------------------------------
// K1 contiguous: warp per line segment
// ------------------------------------------------------------------------------------------
template <int R>
struct ContigCfg {
    // row pitch (doubles) of the [e mod R][e div R] staging tile.  64-bit shared accesses are
    // served per half-warp: the 16 lanes e = 16h..16h+15 of a staging store hit rows e % R and
    // columns e / R, so PITCH = 16/R * (odd) (mod 16) spreads them over all 16 bank pairs
    // (ncu on the first version, PITCH 40 for R = 4: 3.7 wavefronts per store instead of 2).
    // Window loads read consecutive columns of one row and are conflict-free for any pitch.
    static constexpr int PITCH = (R == 4) ? 36 : (R == 2 ? 40 : 32 + SLB_P1MAX);
};

template <int P1, int R, bool EXACT>
__global__ void __launch_bounds__(256)
k_sweep_contig(const double* __restrict__ in, double* __restrict__ out, int n, long long nlines, AlphaMap am,
               const double* __restrict__ coef, int nc, InMap im, int lpw)
{
    constexpr int WARPS = 8;
    constexpr int SEG = 32 * R;           // outputs per work item
    constexpr int W = SEG + P1 - 1;       // inputs staged per work item
    constexpr int NLD = (W + 31) / 32;
    constexpr int LASTW = W - 32 * (NLD - 1);  // active lanes of the last staging load
    constexpr int PITCH = ContigCfg<R>::PITCH;
    constexpr int NX = R + P1 - 1;
    __shared__ double sm[WARPS][R * PITCH];
    __shared__ double scoef[SLB_NCMAX * P1];  // weight polynomials, [k][j]: lane j reads conflict-free

    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    double* tile = sm[wid];
    for (int i = threadIdx.x; i < nc * P1; i += blockDim.x) scoef[i] = __ldg(coef + (i % P1) * nc + i / P1);
    __syncthreads();

    // each warp owns `lpw` consecutive lines; a work item is one segment of SEG outputs of one
    // line, and the inputs of the next item are prefetched into registers while the current
    // one is computed.  Consecutive lines often share alpha (Vlasov space sweeps: alpha depends
    // on one velocity index only), so weights are recomputed only when alpha changes -- the
    // device counterpart of the reference's CachePrecal memo (src/interpolation.jl:381-389).
    const long long l0 = ((long long)blockIdx.x * WARPS + wid) * lpw;
    if (l0 >= nlines) return;
    const long long l1 = l0 + lpw < nlines ? l0 + lpw : nlines;
    const bool wrap2 = (W <= 2 * n);  // two conditional subtractions bring any staged index into [0, n)

    auto getalpha = [&](long long ln) { return am.scale * __ldg(am.tab + slb_alpha_off(am, 0u, (unsigned)ln)); };
    auto load = [&](long long ln, int sg0, int s0_, double (&v)[NLD]) {
        const double* lin = in + slb_in_line(im, ln) * n;
        int g0 = s0_ + sg0;
        g0 -= (g0 >= n) ? n : 0;
        g0 += lane;
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
            int g = g0 + 32 * q;
            if (wrap2) {
                g -= (g >= n) ? n : 0;
                g -= (g >= n) ? n : 0;
            } else {
                g %= n;
            }
            if (q < NLD - 1 || lane < LASTW)
                v[q] = __ldg(lin + g);
            else
                v[q] = 0.0;
        }
    };

    long long line = l0;
    int seg0 = 0, s0;
    double alpha = getalpha(line), t;
    slb_split(alpha, n, (P1 - 1) / 2, t, s0);
    double v[NLD];
    load(line, seg0, s0, v);
    double w[P1];
    double memo_alpha = 0.0;
    bool have_w = false;

    while (true) {
        long long nline = line;
        int nseg0 = seg0 + SEG, ns0 = s0;
        double nalpha = alpha, nt = t;
        if (nseg0 >= n) {
            nseg0 = 0;
            nline = line + 1;
        }
        const bool has_next = nline < l1;
        double vn[NLD];
        if (has_next) {
            if (nline != line) {
                nalpha = getalpha(nline);
                if (nalpha != alpha) slb_split(nalpha, n, (P1 - 1) / 2, nt, ns0);
            }
            load(nline, nseg0, ns0, vn);
        }
        if (!have_w || alpha != memo_alpha) {
            // lane j evaluates weight polynomial j (Horner, FMA), then the warp broadcasts
            const int jl = lane < P1 ? lane : 0;
            double wl = scoef[(nc - 1) * P1 + jl];
            for (int k = nc - 2; k >= 0; --k) wl = fma(t, wl, scoef[k * P1 + jl]);
#pragma unroll
            for (int j = 0; j < P1; ++j) w[j] = __shfl_sync(0xffffffffu, wl, j);
            memo_alpha = alpha;
            have_w = true;
        }
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
            const int e = lane + 32 * q;
            if (q < NLD - 1 || lane < LASTW) tile[(e % R) * PITCH + e / R] = v[q];
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
        const int i0 = seg0 + R * lane;
        double* dst = out + line * n + i0;
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
        if (!has_next) break;
        line = nline;
        seg0 = nseg0;
        s0 = ns0;
        alpha = nalpha;
        t = nt;
#pragma unroll
        for (int q = 0; q < NLD; ++q) v[q] = vn[q];
    }
}

// ------------------------------------------------------------------------------------------
// K1c tile (round 2): persistent blocks; a block stages LT whole neighbouring lines -- ONE contiguous run of LT * n
// doubles along dim 0 -- in shared memory with 16-byte cp.async (no per-element index arithmetic, loads in flight cost
// no registers; double-buffered: tile t + 1 is in flight while tile t is swept; every row is followed by a copy of its
// first order + 1 points, so no stencil wraps).  A lane computes PAIRS of neighbouring outputs (2 l, 2 l + 1), (2 l + 64,
// 2 l + 65), ...: the order + 2 inputs of a pair are fetched as (order + 3) / 2 aligned 16-byte shared-memory loads
// (the parity of the window start is uniform per line: two unrolled variants) and the pair leaves as one 16-byte store.
// ncu, 128^4 order 7: k_sweep_contig executed 300 warp instructions per 128-point line (60 % issue-bound at 73 % of the
// HBM peak), a first tile version with one 8-byte load per stencil input 330 (the shared-memory pipe at 69 %).
// Same operation order per output: bit-identical results.  n even, order odd, plain input layout, 16-byte aligned
// buffers; anything else keeps k_sweep_contig.
// ------------------------------------------------------------------------------------------
template <int P1, bool EXACT, int OFF>
__device__ __forceinline__ void contig_tile_pair(const double* __restrict__ row, int ga, const double (&w)[P1], double& o0, double& o1)
{
    constexpr int NV = P1 + 2;
    double xv[NV];
    const double2* p2 = reinterpret_cast<const double2*>(row + ga);
#pragma unroll
    for (int q = 0; q < NV / 2; ++q) {
        const double2 t = p2[q];
        xv[2 * q] = t.x;
        xv[2 * q + 1] = t.y;
    }
    if (EXACT) {
        o0 = __dmul_rn(xv[OFF], w[0]);
        o1 = __dmul_rn(xv[OFF + 1], w[0]);
#pragma unroll
        for (int j = 1; j < P1; ++j) {
            o0 = __dadd_rn(o0, __dmul_rn(xv[OFF + j], w[j]));
            o1 = __dadd_rn(o1, __dmul_rn(xv[OFF + 1 + j], w[j]));
        }
    } else {
        o0 = xv[OFF] * w[0];
        o1 = xv[OFF + 1] * w[0];
#pragma unroll
        for (int j = 1; j < P1; ++j) {
            o0 = fma(xv[OFF + j], w[j], o0);
            o1 = fma(xv[OFF + 1 + j], w[j], o1);
        }
    }
}

// stage the rows [l0, l0 + nl) of the grid (n doubles each, contiguous) at smem byte address dst0, row pitch `pitch`
// doubles, each followed by a copy of its first P1 points: warp per row, 16 bytes per lane and copy
template <int P1>
__device__ __forceinline__ void contig_tile_fetch(const double* __restrict__ in, long long l0, int nl, int n, int pitch, unsigned dst0,
                                                  int wid, int lane)
{
    const char* src = reinterpret_cast<const char*>(in + (l0 + wid) * n) + 16 * lane;
    unsigned dst = dst0 + 8u * (unsigned)(wid * pitch) + 16u * (unsigned)lane;
    const long long sstep = 64ll * n;            // 8 rows further
    const unsigned dstep = 64u * (unsigned)pitch;
    const int nb = 8 * n;                        // bytes per row
    for (int l = wid; l < nl; l += 8) {
        for (int kb = 16 * lane; kb < nb; kb += 512)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)(kb - 16 * lane)), "l"(src + (kb - 16 * lane)) : "memory");
        if (lane < P1 / 2)   // the periodic padding: the first order + 1 points again
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)nb), "l"(src) : "memory");
        src += sstep;
        dst += dstep;
    }
}

#define SLB_TILE_NSLOT 5   // 16-byte copies per thread and FULL tile (the host keeps LT * (n / 2 + P1 / 2) <= 5 * 256)
template <int P1, bool EXACT>
__global__ void __launch_bounds__(256)   // 4 blocks per SM (capped at 48 registers for 5 blocks: 0.74 -> 0.83 ms)
k_sweep_contig_tile(const double* __restrict__ in, double* __restrict__ out, int n, long long nlines, AlphaMap am,
                    const double* __restrict__ coef, int nc, int LT)
{
    static_assert(P1 % 2 == 0, "pairs of outputs share order + 2 inputs fetched as 16-byte words");
    extern __shared__ __align__(16) double tsm[];  // [2][LT][n + P1]: the next tile is in flight while this one is swept; [2][LT] shifts
    __shared__ double scoef[SLB_NCMAX * P1];       // weight polynomials, [k][j]: lane j reads conflict-free
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pitch = n + P1;                      // even: every row starts on a 16-byte boundary
    const long long ntiles = (nlines + LT - 1) / LT;
    // every block sweeps a CONTIGUOUS range of tiles: neighbouring lines usually share their shift, so the weights are
    // re-evaluated only where it changes
    const long long per = (ntiles + gridDim.x - 1) / gridDim.x;
    const long long tbeg = (long long)blockIdx.x * per;
    const long long tend = tbeg + per < ntiles ? tbeg + per : ntiles;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(tsm);
    const unsigned tile_b = (unsigned)LT * (unsigned)pitch * 8u;
    double* const alsm = tsm + (size_t)2 * LT * pitch;   // [2][LT] shifts of the lines of the two tiles
    // copy slots of a FULL tile, computed once: copy c = tid + 256 i moves 16 bytes of row c / cpr (the row's n / 2 words,
    // then its first P1 / 2 words again as the periodic padding) -- per tile only the tile's base address is added
    // (the per-row loops this replaces were 77 of the 280 warp instructions per line)
    const int nh = n >> 1, cpr = nh + P1 / 2;
    unsigned soff[SLB_TILE_NSLOT], doff[SLB_TILE_NSLOT];
    unsigned smask = 0;
#pragma unroll
    for (int i = 0; i < SLB_TILE_NSLOT; ++i) {
        const int c = threadIdx.x + 256 * i;
        const int r = c / cpr, k = c - r * cpr;
        soff[i] = (unsigned)(r * n) * 8u + 16u * (unsigned)(k < nh ? k : k - nh);
        doff[i] = (unsigned)(r * pitch) * 8u + 16u * (unsigned)k;
        if (c < LT * cpr) smask |= 1u << i;
    }
    auto fetch = [&](long long t, unsigned b) {   // tile t (< tend) into buffer b
        const long long l0 = t * LT;
        const int nl = (int)(nlines - l0 < LT ? nlines - l0 : LT);
        if (nl == LT) {
            const char* src = reinterpret_cast<const char*>(in + l0 * n);
            const unsigned dst = sbase + b * tile_b;
#pragma unroll
            for (int i = 0; i < SLB_TILE_NSLOT; ++i)
                if (smask & (1u << i))
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + doff[i]), "l"(src + soff[i]) : "memory");
        } else {
            contig_tile_fetch<P1>(in, l0, nl, n, pitch, sbase + b * tile_b, wid, lane);   // the grid's last, partial tile
        }
    };
    // shifts of a tile's lines: one thread per line, ONE TILE AHEAD like the tile itself (a dependent L2 round trip per
    // tile would otherwise sit in front of every sweep phase; evaluated by every warp for its own lines they were 16 more
    // instructions per line)
    auto tile_alpha = [&](long long t, unsigned b) {
        const long long ln = t * LT + threadIdx.x;
        if ((int)threadIdx.x < LT && ln < nlines) alsm[b * LT + threadIdx.x] = am.scale * __ldg(am.tab + slb_alpha_off(am, 0u, (unsigned)ln));
    };
    if (tbeg < tend) {
        fetch(tbeg, 0u);
        tile_alpha(tbeg, 0u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = threadIdx.x; i < nc * P1; i += 256) scoef[i] = __ldg(coef + (i % P1) * nc + i / P1);
    double w[P1];
    double memo_alpha = 0.0;
    bool have_w = false;
    int s0 = 0;
    unsigned buf = 0;
    for (long long t = tbeg; t < tend; ++t, buf ^= 1u) {
        const long long line0 = t * LT;
        const int nl = (int)(nlines - line0 < LT ? nlines - line0 : LT);
        if (t + 1 < tend) {
            fetch(t + 1, buf ^ 1u);
            tile_alpha(t + 1, buf ^ 1u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const double* tile = tsm + (size_t)buf * LT * pitch;
        for (int l = wid; l < nl; l += 8) {
            const double alpha = alsm[buf * LT + l];
            if (!have_w || alpha != memo_alpha) {
                // lane j evaluates weight polynomial j (Horner, FMA), then the warp broadcasts; consecutive lines often
                // share alpha (the device counterpart of the reference's CachePrecal memo, src/interpolation.jl:381-389)
                double tt;
                slb_split(alpha, n, (P1 - 1) / 2, tt, s0);
                const int jl = lane < P1 ? lane : 0;
                double wl = scoef[(nc - 1) * P1 + jl];
                for (int q = nc - 2; q >= 0; --q) wl = fma(tt, wl, scoef[q * P1 + jl]);
#pragma unroll
                for (int j = 0; j < P1; ++j) w[j] = __shfl_sync(0xffffffffu, wl, j);
                memo_alpha = alpha;
                have_w = true;
            }
            const double* row = tile + (size_t)l * pitch;
            double* dst = out + (line0 + l) * n;
            const int off = s0 & 1;   // parity of every window start of this line
            int g = s0 + 2 * lane;    // window start of the first pair
            g -= g >= n ? n : 0;
            g -= g >= n ? n : 0;      // 2 * lane may exceed n on short lines
#pragma unroll 2
            for (int i = 2 * lane; i < n; i += 64) {
                double o0, o1;
                if (off)
                    contig_tile_pair<P1, EXACT, 1>(row, g - 1, w, o0, o1);
                else
                    contig_tile_pair<P1, EXACT, 0>(row, g, w, o0, o1);
                asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(dst + i), "d"(o0), "d"(o1) : "memory");
                g += 64;
                while (g >= n) g -= n;
            }
        }
        __syncthreads();  // everybody is done with this buffer: the fetch of the next iteration may overwrite it
    }
}

// ------------------------------------------------------------------------------------------
// generic fallback: thread per line, any order <= 63
// ------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(128)
k_sweep_generic(const double* __restrict__ in, double* __restrict__ out, long long inner, int n, long long nlines,
                AlphaMap am, const double* __restrict__ coef, int np, int nc, int exact)
{
    long long gid = (long long)blockIdx.x * 128 + threadIdx.x;
    if (gid >= nlines) return;
    long long b = gid / inner;
    long long a = gid - b * inner;
    double alpha = am.scale * __ldg(am.tab + slb_alpha_off(am, (unsigned)a, (unsigned)b));
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

// ------------------------------------------------------------------------------------------
// InsideEdge (non-periodic) sweep -- "a marginal case" in the reference's words
// (src/interpolation.jl:123-132 get_allprecal, :250-286 interpolate!(..., interp::InsideEdge)): the
// stencil window never leaves the line.  With indbeg = decint - order/2 (must satisfy
// -order <= indbeg <= 0), 1-based output i uses
//     i <= -indbeg              : res[1 : order+1]            weights tabfct(t + indbeg + i - 1)
//     -indbeg < i <= n-decint-order/2-1 : res[i+indbeg : i+indbeg+order]   weights tabfct(t)  (the usual ones)
//     else                      : res[n-order : n]            weights tabfct(t + indbeg + order - (n - i))
// i.e. near the ends the same polynomial is evaluated outside [0, 1] (one-sided stencils).  Thread per
// line, any dim, run-time order; lines whose shift violates the bound are filled with NaN (the reference
// indexes out of bounds there).  Not a hot path: kernel-seam calls only.
// ------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(128)
k_sweep_inside(const double* __restrict__ in, double* __restrict__ out, long long inner, int n, long long nlines,
               AlphaMap am, const double* __restrict__ coef, int np, int nc, int exact)
{
    long long gid = (long long)blockIdx.x * 128 + threadIdx.x;
    if (gid >= nlines) return;
    long long b = gid / inner;
    long long a = gid - b * inner;
    const double alpha = am.scale * __ldg(am.tab + slb_alpha_off(am, (unsigned)a, (unsigned)b));
    const double fl = floor(alpha);
    const double t = alpha - fl;
    const int order = np - 1;
    const double* pin = in + (b * n) * inner + a;
    double* pout = out + (b * n) * inner + a;
    const double dib = fl - (double)(order / 2);  // indbeg
    if (!(dib <= 0.0 && dib >= -(double)order)) {
        for (int i = 0; i < n; ++i) pout[(long long)i * inner] = nan("");
        return;
    }
    const int indbeg = (int)dib;
    // 1-based bounds of the three regimes: borne1 = -decint - origin, borne2 = lg - decint + origin - 1 (:264-265)
    const int borne1 = -indbeg, borne2 = n - (int)fl - order / 2 - 1;
    double wmid[64], we[64];
    for (int j = 0; j < np; ++j) {
        const double* c = coef + j * nc;
        double ex = __ldg(c + nc - 1);
        for (int k = nc - 2; k >= 0; --k) ex = fma(t, ex, __ldg(c + k));
        wmid[j] = ex;
    }
    for (int i = 1; i <= n; ++i) {
        int first;
        const double* w = wmid;
        if (i <= borne1 || i > borne2) {
            int ind;
            if (i <= borne1) {
                first = 1;
                ind = i;
            } else {
                first = n - order;
                ind = np - (n - i);
            }
            const double targ = t + (double)(indbeg + ind - 1);
            for (int j = 0; j < np; ++j) {
                const double* c = coef + j * nc;
                double ex = __ldg(c + nc - 1);
                for (int k = nc - 2; k >= 0; --k) ex = fma(targ, ex, __ldg(c + k));
                we[j] = ex;
            }
            w = we;
        } else {
            first = i - borne1;
        }
        double acc = 0.0;
        for (int j = 0; j < np; ++j) {
            const double v = __ldg(pin + (long long)(first - 1 + j) * inner);
            if (exact) {
                const double pr = __dmul_rn(v, w[j]);
                acc = (j == 0) ? pr : __dadd_rn(acc, pr);
            } else {
                acc = (j == 0) ? v * w[j] : fma(v, w[j], acc);
            }
        }
        pout[(long long)(i - 1) * inner] = acc;
    }
}
