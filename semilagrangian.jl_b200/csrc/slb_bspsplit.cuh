// slb_bspsplit.cuh -- K2+K1 fused B-spline sweep along a strided dim with TWO warps per tile of 32
// lines: the periodic line is cut into two halves by two separators of h entries each, so that the
// banded substitutions of the halves are independent recurrences.
//
// Why: the one-warp-per-tile kernel (slb_bspfused.cuh) is bound by dependent-instruction latency.  A
// line must stay in shared memory between the forward and the backward substitution, which limits an
// SM to 6 tiles = 6 warps of thread-per-line chains (1.5 warps per scheduler; ncu: issue slots 45 %
// busy, FP64 pipe 30 %, DRAM 28 %).  Shared memory cannot hold more lines -- but a line can be worked
// on by more than one thread.
//
// Formulation.  n even, h = (order-1)/2, Na = n/2 - h.  Order the unknowns as
//     half a = [0, Na) | sep1 = [Na, n/2) | half b = [n/2, n/2 + Na) | sep2 = [n - h, n)
// and move both separators to the border of the bordered system of slb_bspline.cuh:
//     [ Ta  0  Ca ] [xa]   [ua]        Ta = Tb: Na x Na banded Toeplitz, T = L U (no pivoting, SPD),
//     [ 0   Tb Cb ] [xb] = [ub]        C: couplings to the 2h separator unknowns, R = C^T,
//     [ Ra  Rb D  ] [x2]   [u2]        S = D - Ra Ta^-1 Ca - Rb Tb^-1 Cb   (2h x 2h).
// The halves never touch each other (the separators cut the band) and, the matrix being circulant, the
// cyclic shift by n/2 maps half a to half b, sep1 to sep2 and sep2 to sep1: in its LOCAL frame --
// rows 0..Na-1, border slots 0..h-1 = the separator that FOLLOWS the half, slots h..2h-1 = the one
// that PRECEDES it -- each half has the SAME tables.  Per half (one warp, thread per line):
//     forward   y_i = u_i - sum_j L[i][j] y_{i-j} ;  acc_q += Ri[i][q] y_i          (h + 2h FMAs)
//     border    rhs_q = u2_q - acc_q(own) - acc_{sigma(q)}(other),  sigma(q) = (q + h) mod 2h
//               x2 = Sinv rhs                                                     (4h^2 FMAs per line)
//     backward  w_i = (y_i - sum_j U[i][j] w_{i+j}) / d_i ;  x_i = w_i - sum_q G[i][q] x2_q
// i.e. 6h + 1 FMAs per cell instead of 4h + 1, traded for twice as many independent recurrences per
// line, half as long.  The two warps of a tile meet at named barriers: to combine the border sums
// (through the separator rows of the tile itself), before the stencil (each warp then produces half of
// the outputs from the whole solved line) and before the tile is reloaded.
// Same solution as slb_bspline.cuh's solver (and the reference's cyclic LU, src/bsplinelu.jl:179-220)
// up to rounding: tests/test_host_logic.py checks the host build of the same routine.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "../../include/slb200.h"
#include "slb_sweep.cuh"
#include "slb_bsprf.cuh"

struct BspSplitHost {
    int h, n, Na;
    std::vector<double> L, U, invd;  // Na x h, Na x h, Na
    std::vector<double> Ri, G;       // Na x 2h (local border slots)
    std::vector<double> Sinv;        // 2h x 2h (local slot order of half a == of half b)
};

// global row of local border slot q of half s (s = 0: a, 1: b)
__host__ __device__ __forceinline__ int bsps_border_row(int s, int q, int h, int n)
{
    const int base = s * (n / 2);
    if (q < h) return base + (n / 2 - h) + q;  // the separator that follows the half
    int r = base - h + (q - h);                // the separator that precedes it
    return r < 0 ? r + n : r;
}

static int bspsplit_factor(int order, int64_t n64, const double* node_vals, BspSplitHost* out, std::string& msg)
{
    typedef long double ld;
    const int h = (order - 1) / 2;
    const int n = (int)n64;
    if (h < 1 || h > 6) { msg = "split B-spline solver: order must be odd and in [3,13]"; return SLB_E_UNSUPPORTED; }
    if (n % 2 != 0) { msg = "split B-spline solver: n must be even"; return SLB_E_UNSUPPORTED; }
    const int Na = n / 2 - h;
    if (Na < 2 * h + 1) { msg = "split B-spline solver: line too short"; return SLB_E_UNSUPPORTED; }
    const int B = 2 * h;
    std::vector<ld> a(h + 1);
    for (int m = 0; m <= h; ++m) a[m] = (ld)node_vals[h + m];
    auto Aent = [&](int i, int j) -> ld {  // cyclic entry A[i][j] = a_{|i-j| cyclic}
        int m = ((j - i) % n + n) % n;
        if (n - m < m) m = n - m;
        return m <= h ? a[m] : 0.0L;
    };
    // banded LU of Ta (rows 0..Na-1): Lm[i][j-1] multiplies row i-j, Um[i][j] = U(i, i+j)
    std::vector<ld> Lm((size_t)Na * h, 0.0L), Um((size_t)Na * (h + 1), 0.0L);
    for (int i = 0; i < Na; ++i) {
        std::vector<ld> row(2 * h + 1, 0.0L);
        for (int m = -h; m <= h; ++m) {
            int j = i + m;
            if (j >= 0 && j < Na) row[m + h] = a[m < 0 ? -m : m];
        }
        for (int j = h; j >= 1; --j) {
            int pr = i - j;
            if (pr < 0) continue;
            ld mult = row[h - j] / Um[(size_t)pr * (h + 1)];
            Lm[(size_t)i * h + (j - 1)] = mult;
            for (int q = 0; q <= h; ++q) {
                int off = pr + q - i + h;
                if (off >= 0 && off <= 2 * h) row[off] -= mult * Um[(size_t)pr * (h + 1) + q];
            }
        }
        for (int q = 0; q <= h; ++q) Um[(size_t)i * (h + 1) + q] = (i + q < Na) ? row[h + q] : 0.0L;
        if (Um[(size_t)i * (h + 1)] == 0.0L) { msg = "B-spline: singular collocation matrix"; return SLB_E_ARG; }
    }
    // local frame of half a: row i <-> global i, border slot q <-> global bsps_border_row(0, q)
    std::vector<int> bg(B);
    for (int q = 0; q < B; ++q) bg[q] = bsps_border_row(0, q, h, n);
    // Ri = R U^-1 (Na x B), Lc = L^-1 C, G = U^-1 Lc (Na x B)
    std::vector<ld> Ri((size_t)Na * B, 0.0L), Lc((size_t)Na * B, 0.0L), G((size_t)Na * B, 0.0L);
    for (int r = 0; r < B; ++r)
        for (int c = 0; c < Na; ++c) {
            ld s = Aent(bg[r], c);
            for (int j = 1; j <= h; ++j)
                if (c - j >= 0) s -= Ri[(size_t)(c - j) * B + r] * Um[(size_t)(c - j) * (h + 1) + j];
            Ri[(size_t)c * B + r] = s / Um[(size_t)c * (h + 1)];
        }
    for (int q = 0; q < B; ++q) {
        for (int i = 0; i < Na; ++i) {
            ld s = Aent(i, bg[q]);
            for (int j = 1; j <= h; ++j)
                if (i - j >= 0) s -= Lm[(size_t)i * h + (j - 1)] * Lc[(size_t)(i - j) * B + q];
            Lc[(size_t)i * B + q] = s;
        }
        for (int i = Na - 1; i >= 0; --i) {
            ld s = Lc[(size_t)i * B + q];
            for (int j = 1; j <= h; ++j)
                if (i + j < Na) s -= Um[(size_t)i * (h + 1) + j] * G[(size_t)(i + j) * B + q];
            G[(size_t)i * B + q] = s / Um[(size_t)i * (h + 1)];
        }
    }
    // M = Ra Ta^-1 Ca in a's local order; half b contributes M[sigma(r)][sigma(q)]
    std::vector<ld> M((size_t)B * B), S((size_t)B * B), Si((size_t)B * B, 0.0L);
    for (int r = 0; r < B; ++r)
        for (int q = 0; q < B; ++q) {
            ld s = 0.0L;
            for (int i = 0; i < Na; ++i) s += Ri[(size_t)i * B + r] * Lc[(size_t)i * B + q];
            M[(size_t)r * B + q] = s;
        }
    auto sig = [&](int q) { return (q + h) % B; };
    for (int r = 0; r < B; ++r)
        for (int q = 0; q < B; ++q) S[(size_t)r * B + q] = Aent(bg[r], bg[q]) - M[(size_t)r * B + q] - M[(size_t)sig(r) * B + sig(q)];
    for (int r = 0; r < B; ++r) Si[(size_t)r * B + r] = 1.0L;
    for (int c = 0; c < B; ++c) {  // Gauss-Jordan with partial pivoting
        int piv = c;
        for (int r = c + 1; r < B; ++r)
            if (fabsl(S[(size_t)r * B + c]) > fabsl(S[(size_t)piv * B + c])) piv = r;
        if (S[(size_t)piv * B + c] == 0.0L) { msg = "B-spline: singular Schur complement"; return SLB_E_ARG; }
        if (piv != c)
            for (int q = 0; q < B; ++q) {
                std::swap(S[(size_t)c * B + q], S[(size_t)piv * B + q]);
                std::swap(Si[(size_t)c * B + q], Si[(size_t)piv * B + q]);
            }
        ld inv = 1.0L / S[(size_t)c * B + c];
        for (int q = 0; q < B; ++q) { S[(size_t)c * B + q] *= inv; Si[(size_t)c * B + q] *= inv; }
        for (int r = 0; r < B; ++r)
            if (r != c) {
                ld f = S[(size_t)r * B + c];
                if (f != 0.0L)
                    for (int q = 0; q < B; ++q) { S[(size_t)r * B + q] -= f * S[(size_t)c * B + q]; Si[(size_t)r * B + q] -= f * Si[(size_t)c * B + q]; }
            }
    }
    out->h = h; out->n = n; out->Na = Na;
    out->L.resize((size_t)Na * h); out->U.resize((size_t)Na * h); out->invd.resize(Na);
    out->Ri.resize((size_t)Na * B); out->G.resize((size_t)Na * B); out->Sinv.resize((size_t)B * B);
    for (int i = 0; i < Na; ++i) {
        out->invd[i] = (double)(1.0L / Um[(size_t)i * (h + 1)]);
        for (int j = 0; j < h; ++j) {
            out->L[(size_t)i * h + j] = (double)Lm[(size_t)i * h + j];
            // pre-scaled by 1/diag: the backward recurrence is then one FMA deep per row
            out->U[(size_t)i * h + j] = (double)(Um[(size_t)i * (h + 1) + j + 1] / Um[(size_t)i * (h + 1)]);
        }
        for (int q = 0; q < B; ++q) {
            out->Ri[(size_t)i * B + q] = (double)Ri[(size_t)i * B + q];
            out->G[(size_t)i * B + q] = (double)G[(size_t)i * B + q];
        }
    }
    for (int i = 0; i < B * B; ++i) out->Sinv[i] = (double)Si[i];
    return SLB_OK;
}

// Kernel table layout (doubles): forward record i: { L[h], Ri[2h], pad } (FR), backward record i:
// { 1/d, U/d [h], G[2h], pad } (BR), then Sinv[2h][2h]; records are whole 16-byte words.
struct BspSplitTab {
    int h, n, Na;
    int FR, BR;
    int o_bwd, o_S, ndoubles;
};
__host__ __device__ constexpr int bsps_FR(int h) { return (3 * h + 1) / 2 * 2; }
__host__ __device__ constexpr int bsps_BR(int h) { return (3 * h + 2) / 2 * 2; }

static void bspsplit_fill(BspSplitTab* tab, std::vector<double>& v, const BspSplitHost& hb)
{
    const int h = hb.h, Na = hb.Na, B = 2 * h;
    tab->h = h; tab->n = hb.n; tab->Na = Na;
    tab->FR = bsps_FR(h); tab->BR = bsps_BR(h);
    tab->o_bwd = Na * tab->FR;
    tab->o_S = tab->o_bwd + Na * tab->BR;
    tab->ndoubles = tab->o_S + B * B;
    v.assign((size_t)tab->ndoubles, 0.0);
    for (int i = 0; i < Na; ++i) {
        double* F = v.data() + (size_t)i * tab->FR;
        double* Bk = v.data() + tab->o_bwd + (size_t)i * tab->BR;
        for (int j = 0; j < h; ++j) F[j] = hb.L[(size_t)i * h + j];
        for (int q = 0; q < B; ++q) F[h + q] = hb.Ri[(size_t)i * B + q];
        Bk[0] = hb.invd[i];
        for (int j = 0; j < h; ++j) Bk[1 + j] = hb.U[(size_t)i * h + j];
        for (int q = 0; q < B; ++q) Bk[1 + h + q] = hb.G[(size_t)i * B + q];
    }
    memcpy(v.data() + tab->o_S, hb.Sinv.data(), (size_t)B * B * sizeof(double));
}

// Reference implementation of the split solve on one line (host and device): the arithmetic the kernel
// performs, without its staging.  x and u may alias.
__host__ __device__ inline void bspsplit_solve_line(const BspSplitTab& t, const double* tab, const double* u, double* x)
{
    const int h = t.h, n = t.n, Na = t.Na, B = 2 * h;
    const double* tF = tab;
    const double* tB = tab + t.o_bwd;
    const double* tS = tab + t.o_S;
    double acc[2][12], x2[2][12];
    for (int s = 0; s < 2; ++s) {
        const int base = s * (n / 2);
        for (int q = 0; q < B; ++q) acc[s][q] = 0.0;
        for (int i = 0; i < Na; ++i) {
            const double* T = tF + (size_t)i * t.FR;
            double y = u[base + i];
            for (int j = 1; j <= h; ++j)
                if (i - j >= 0) y = fma(-T[j - 1], x[base + i - j], y);
            x[base + i] = y;
            for (int q = 0; q < B; ++q) acc[s][q] = fma(T[h + q], y, acc[s][q]);
        }
    }
    // the separator rows collect u2 - acc: first the half that precedes each separator, then the other
    double rhs[2][12];
    for (int s = 0; s < 2; ++s)
        for (int q = 0; q < h; ++q) x[bsps_border_row(s, q, h, n)] = u[bsps_border_row(s, q, h, n)] - acc[s][q];
    for (int s = 0; s < 2; ++s)
        for (int q = h; q < B; ++q) x[bsps_border_row(s, q, h, n)] -= acc[s][q];
    for (int s = 0; s < 2; ++s)
        for (int q = 0; q < B; ++q) rhs[s][q] = x[bsps_border_row(s, q, h, n)];
    for (int s = 0; s < 2; ++s)
        for (int q = 0; q < B; ++q) {
            double sq = 0.0;
            for (int r = 0; r < B; ++r) sq = fma(tS[q * B + r], rhs[s][r], sq);
            x2[s][q] = sq;
        }
    for (int s = 0; s < 2; ++s) {
        const int base = s * (n / 2);
        for (int q = 0; q < h; ++q) x[bsps_border_row(s, q, h, n)] = x2[s][q];
        for (int i = Na - 1; i >= 0; --i) {
            const double* T = tB + (size_t)i * t.BR;
            double w = T[0] * x[base + i];
            for (int j = 1; j <= h; ++j)
                if (i + j < Na) w = fma(-T[j], x[base + i + j] /* holds w_{i+j}, see below */, w);
            x[base + i] = w;
        }
        // x_i = w_i - G_i x2 in a second pass (the kernel fuses it: it keeps the w window in registers)
        for (int i = 0; i < Na; ++i) {
            const double* T = tB + (size_t)i * t.BR;
            double v = x[base + i];
            for (int q = 0; q < B; ++q) v = fma(-T[1 + h + q], x2[s][q], v);
            x[base + i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
struct BspSplitArgs {
    const double* in;
    double* out;
    long long inner;    // element stride of the swept index (> 1: strided dims only)
    long long nlines;
    long long bstride;  // offset between consecutive outer indices of the output (n * inner)
    int n;
    int nc;             // polynomial coefficients per stencil weight
    int tiles;          // tiles in flight per block (two warps each)
    AlphaMap am;
    double* linesum;    // optional: per-line sums of the outputs
    const double* tab_dev;
    BspSplitTab tab;
    int use_rf;         // 1: each warp runs the recursive-filter cascade (slb_bsprf.cuh) on its half of the rows,
    BspRfTab rf;        //    started from the periodic look-back sums; tab_dev then holds the RF table
};

int slb_bspsplit_tiles_rf(int ndoubles, int n);
int slb_bspsplit_launch(const BspSplitArgs& a, const CoefTab& ct, int sm_count, cudaStream_t stream);
int slb_bspsplit_tiles(int h, int n);   // tiles per block that fit the shared memory next to the tables; 0: unsupported

#ifdef SLB_BSPS_IMPL
#define SLB_BSPS_MAXTILES 6      // LU form: 163 registers x 384 threads
#define SLB_BSPS_MAXTILES_RF 7   // recursive-filter form: lighter threads, smaller tables

__device__ __forceinline__ void bsps_cp_async8(unsigned smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

template <int NV>
__device__ __forceinline__ void bsps_ldrec(double (&dst)[NV], const double* src)
{
    static_assert(NV % 2 == 0, "records are whole 16-byte words");
#pragma unroll
    for (int q = 0; q < NV; q += 2) {
        const double2 v = *reinterpret_cast<const double2*>(src + q);
        dst[q] = v.x;
        dst[q + 1] = v.y;
    }
}

template <int H, bool RF>
__global__ void __launch_bounds__(RF ? 64 * SLB_BSPS_MAXTILES_RF : 64 * SLB_BSPS_MAXTILES, 1)
k_bspline_split(const __grid_constant__ BspSplitArgs fa, const __grid_constant__ CoefTab ct)
{
    constexpr int P1 = 2 * H + 2;  // order + 1 stencil points, order = 2h + 1
    constexpr int B = 2 * H;       // border slots per half
    constexpr int PITCH = 32;
    constexpr int FR = bsps_FR(H), BR = bsps_BR(H);
    extern __shared__ __align__(16) double bsm[];
    __shared__ double lsx[SLB_BSPS_MAXTILES_RF][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int ts = wid >> 1;  // tile slot of this warp pair
    const int s = wid & 1;    // which half of the line this warp owns
    const int n = fa.n, Na = fa.tab.Na, nh = n >> 1;
    // ---- factor tables: one copy per block ------------------------------------------------------
    const int ntab = RF ? fa.rf.ndoubles : fa.tab.ndoubles;
    double* tabs = bsm;
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) tabs[i] = __ldg(fa.tab_dev + i);
    __syncthreads();
    const double* tF = tabs;
    const double* tB = tabs + fa.tab.o_bwd;
    const double* tS = tabs + fa.tab.o_S;
    double* tile = bsm + ((ntab + 1) & ~1) + (size_t)ts * n * PITCH;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(tile);
    double* col = tile + lane;                   // this lane's line: element k at col[k * PITCH]
    double* colh = col + (size_t)s * nh * PITCH;  // its own half: local row i at colh[i * PITCH]
    const int row_next = s * nh + Na;            // first row of the separator that follows the half
    const int row_prev = (s ? nh : n) - H;       // first row of the separator that precedes it
    const int bar_id = 1 + ts;
#define BSPS_PAIR_SYNC() asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory")
    const long long ntiles = (fa.nlines + 31) / 32;

    for (long long t = (long long)blockIdx.x * fa.tiles + ts; t < ntiles; t += (long long)gridDim.x * fa.tiles) {
    const long long line = t * 32 + lane;
    const bool active = line < fa.nlines;
    const long long lc = active ? line : fa.nlines - 1;
    const long long b = lc / fa.inner, a = lc - b * fa.inner;

    // ---- 1. load this warp's rows [s n/2, (s+1) n/2): its half and the separator after it ----------
    {
        const double* src = fa.in + (b * n + (long long)s * nh) * fa.inner + a;
        const unsigned dst = sbase + 8u * (unsigned)(s * nh * PITCH + lane);
        if (active) {
#pragma unroll 8
            for (int k = 0; k < nh; ++k) bsps_cp_async8(dst + 8u * (unsigned)(k * PITCH), src + (long long)k * fa.inner);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // stencil shift and weights of this lane's line (overlaps the loads)
    double w[P1];
    int s0;
    {
        const double alpha = fa.am.scale * __ldg(fa.am.tab + slb_alpha_off(fa.am, (unsigned)a, (unsigned)b));
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
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();

    if (RF) {
        // ---- 2'. recursive-filter cascade on this warp's rows [s n/2, (s+1) n/2) (slb_bsprf.cuh).  The state
        // of the cascade before any row m is a short periodic look-back sum, so the two halves of a line
        // are independent; the barriers only order the in-place updates against the look-back reads.
        double z[H], st[H];
#pragma unroll
        for (int k = 0; k < H; ++k) z[k] = tabs[k];
        const int m0 = s * nh;
        BSPS_PAIR_SYNC();  // every row of the line has arrived
        bsprf_init_wrap<H>(fa.rf, tabs, col, PITCH, n, m0 - 1, -1, st);
        BSPS_PAIR_SYNC();  // all causal start-up sums are done: the inputs may be overwritten
        bsprf_pass<H>(z, st, col, PITCH, nh, m0, +1);
        BSPS_PAIR_SYNC();  // causal output complete
        bsprf_init_wrap<H>(fa.rf, tabs, col, PITCH, n, m0 + nh, +1, st);
        BSPS_PAIR_SYNC();
        bsprf_pass<H>(z, st, col, PITCH, nh, m0 + nh - 1, -1);
    } else {
    // ---- 2. forward substitution on the own half; border sums --------------------------------------
    double x2[B];
    {
        double yw[H], acc[B], un[H], Tn[FR];
#pragma unroll
        for (int q = 0; q < B; ++q) acc[q] = 0.0;
#pragma unroll
        for (int q = 0; q < H; ++q) {
            yw[q] = 0.0;
            un[q] = colh[q * PITCH];   // Na >= 2h + 1
        }
        bsps_ldrec<FR>(Tn, tF);
        // record: T[0..h) = L, T[h..3h) = Ri.  The terms that use older results are summed first; only
        // one FMA per row waits for the previous row.
#define BSPS_FWD_ROW(r, GUARD)                                                                        \
    {                                                                                                 \
        const int i = i0 + (r);                                                                       \
        if (!(GUARD) || i < Na) {                                                                     \
            double T[FR];                                                                             \
            _Pragma("unroll") for (int q = 0; q < FR; ++q) T[q] = Tn[q];                              \
            bsps_ldrec<FR>(Tn, tF + ((GUARD) ? min(i + 1, Na - 1) : i + 1) * FR);                     \
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
            _Pragma("unroll") for (int q = 0; q < B; ++q) acc[q] = fma(T[H + q], y, acc[q]);          \
            colh[i * PITCH] = y;                                                                      \
        }                                                                                             \
    }
        int i0 = 0;
        for (; i0 + 2 * H <= Na; i0 += H) {  // whole groups whose successor's right-hand sides and records exist
            double u[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                u[r] = un[r];
                un[r] = colh[(i0 + H + r) * PITCH];
            }
#pragma unroll
            for (int r = 0; r < H; ++r) BSPS_FWD_ROW(r, false)
        }
        for (; i0 < Na; i0 += H) {
            double u[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                u[r] = un[r];
                un[r] = colh[min(i0 + H + r, Na - 1) * PITCH];
            }
#pragma unroll
            for (int r = 0; r < H; ++r) BSPS_FWD_ROW(r, true)
        }
#undef BSPS_FWD_ROW
        // the separator rows of the tile collect u2 - acc: each warp first reduces the separator that
        // follows its half (rows it loaded itself), then the one that precedes it
#pragma unroll
        for (int q = 0; q < H; ++q) col[(row_next + q) * PITCH] -= acc[q];
        BSPS_PAIR_SYNC();
#pragma unroll
        for (int q = 0; q < H; ++q) col[(row_prev + q) * PITCH] -= acc[H + q];
        BSPS_PAIR_SYNC();
        double rhs[B];
#pragma unroll
        for (int q = 0; q < H; ++q) {
            rhs[q] = col[(row_next + q) * PITCH];
            rhs[H + q] = col[(row_prev + q) * PITCH];
        }
#pragma unroll
        for (int q = 0; q < B; ++q) {
            double sq = 0.0;
#pragma unroll
            for (int r = 0; r < B; ++r) sq = fma(tS[q * B + r], rhs[r], sq);
            x2[q] = sq;
        }
        BSPS_PAIR_SYNC();  // both warps hold the right-hand sides: the separator rows may now take x2
#pragma unroll
        for (int q = 0; q < H; ++q) col[(row_next + q) * PITCH] = x2[q];
    }
    // ---- 3. backward substitution on the own half ------------------------------------------------
    {
        double ww[H], un[H], Tn[BR];
        const int ilast = ((Na - 1) / H) * H;
#pragma unroll
        for (int q = 0; q < H; ++q) {
            ww[q] = 0.0;
            un[q] = colh[min(ilast + q, Na - 1) * PITCH];
        }
        bsps_ldrec<BR>(Tn, tB + (Na - 1) * BR);
        // record: T[0] = 1/d, T[1..h] = U/d, T[1+h..1+3h) = G
#define BSPS_BWD_ROW(r, GUARD)                                                                        \
    {                                                                                                 \
        const int i = i0 + (r);                                                                       \
        if (!(GUARD) || i < Na) {                                                                     \
            double T[BR];                                                                             \
            _Pragma("unroll") for (int q = 0; q < BR; ++q) T[q] = Tn[q];                              \
            bsps_ldrec<BR>(Tn, tB + (i > 0 ? i - 1 : 0) * BR);                                        \
            double v; /* w_i = y_i / d_i - sum_j (U[i][j-1] / d_i) w_{i+j} */                         \
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
            _Pragma("unroll") for (int q = 0; q < B; ++q) x = fma(-T[1 + H + q], x2[q], x);           \
            colh[i * PITCH] = x;                                                                      \
        }                                                                                             \
    }
        const int ngroups = ilast / H + 1;
        {   // the last (possibly ragged) group
            const int i0 = ilast;
            double u[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                u[r] = un[r];
                un[r] = colh[max(i0 - H + r, 0) * PITCH];
            }
#pragma unroll
            for (int r = H - 1; r >= 0; --r) BSPS_BWD_ROW(r, true)
        }
        for (int gq = 1; gq < ngroups; ++gq) {  // whole groups
            const int i0 = (ngroups - 1 - gq) * H;
            double u[H];
#pragma unroll
            for (int r = 0; r < H; ++r) {
                u[r] = un[r];
                un[r] = colh[((i0 >= H ? i0 - H : 0) + r) * PITCH];
            }
#pragma unroll
            for (int r = H - 1; r >= 0; --r) BSPS_BWD_ROW(r, false)
        }
#undef BSPS_BWD_ROW
    }
    }  // !RF
    BSPS_PAIR_SYNC();  // the whole line is solved

    // ---- 4. stencil: this warp produces outputs [s n/2, (s+1) n/2) from the whole line --------------
    // window slot convention of slb_dot: logical element j of output i lives in win[(i + j) % P1]
    {
        double win[P1];
        int kk = s0 + s * nh;
        kk -= kk >= n ? n : 0;
#define BSPS_NEXT(dst)                   \
    {                                    \
        dst = col[kk * PITCH];           \
        kk = kk + 1 == n ? 0 : kk + 1;   \
    }
#pragma unroll
        for (int j = 0; j < P1 - 1; ++j) BSPS_NEXT(win[j]);
        double* po = fa.out + b * fa.bstride + a + (long long)s * nh * fa.inner;
        double lsum = 0.0;
        int i0 = 0;
        for (; i0 + P1 <= nh; i0 += P1) {  // whole groups: no guards, the order+1 dot products overlap
            double nx[P1];
#pragma unroll
            for (int r = 0; r < P1; ++r) BSPS_NEXT(nx[r]);
#pragma unroll
            for (int r = 0; r < P1; ++r) {
                win[(r + P1 - 1) % P1] = nx[r];
                const double acc = slb_dot<P1, false>(win, w, r);
                lsum += acc;
                if (active) *po = acc;
                po += fa.inner;
            }
        }
        for (; i0 < nh; i0 += P1) {
#pragma unroll
            for (int r = 0; r < P1; ++r) {
                if (i0 + r < nh) {
                    BSPS_NEXT(win[(r + P1 - 1) % P1]);
                    const double acc = slb_dot<P1, false>(win, w, r);
                    lsum += acc;
                    if (active) *po = acc;
                    po += fa.inner;
                }
            }
        }
#undef BSPS_NEXT
        if (fa.linesum) {  // uniform over the block
            if (s == 1) lsx[ts][lane] = lsum;
            BSPS_PAIR_SYNC();
            if (s == 0 && active) fa.linesum[line] = lsum + lsx[ts][lane];
        }
    }
    BSPS_PAIR_SYNC();  // the tile is reloaded by the next iteration: the other warp must be done reading
    }  // persistent loop over tiles
#undef BSPS_PAIR_SYNC
}
#endif  // SLB_BSPS_IMPL
