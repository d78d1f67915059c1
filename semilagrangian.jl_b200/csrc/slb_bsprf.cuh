// slb_bsprf.cuh -- the periodic B-spline pre-solve  c = A^{-1} u  as a cascade of first-order
// recursive filters with CONSTANT coefficients (host factorisation + per-line arithmetic; the fused
// sweep kernel of slb_bspfused.cuh runs it when its RF template flag is set).
//
// Why: the banded-LU substitution of slb_bspline.cuh needs a record of per-row factors (L, U, R U^-1,
// T^-1 C: 4h + 2 doubles) for every row of every line, and those records -- broadcast shared-memory
// loads -- are what bounds the fused B-spline sweep (ncu: shared-memory pipe 78 % busy, 32 of 45
// wavefronts per warp-row are table records; profiles/r1_ncu_full_bspline11_split.txt).  The
// collocation matrix is circulant and symmetric, so its symbol factors over its h poles inside the
// unit circle (all real, negative and simple for cardinal B-splines):
//     a(z) = a_0 + sum_m a_m (z^m + z^-m) = C prod_k (1 - z_k z^-1)(1 - z_k z),   C = a_h prod_k (-1/z_k)
// and  c = (1/C) [prod_k 1/(1 - z_k z)] [prod_k 1/(1 - z_k z^-1)] u :
//     causal      stage k:  y[i] = x[i] + z_k y[i-1]      i = 0 .. n-1
//     anticausal  stage k:  y[i] = x[i] + z_k y[i+1]      i = n-1 .. 0
// 2h FMAs per cell (the LU form: 4h + 1), h + 1 constants in registers, no per-row tables at all.
// Periodicity enters only through the initial states: the state of stage k before row 0 is
//     s_k = sum_{r >= 0} G_k[r] x[n-1-r],   G_k = impulse response of stages 1..k, aliased modulo n,
// truncated where it has decayed below 1e-19 of its maximum (the stages are ordered by increasing |z_k|,
// so only the last one needs ~ log(1e-19)/log|z_h| terms: order 11, n = 128: 6 + 11 + 19 + 34 + 108
// FMAs per line and direction).  The same table serves the anticausal cascade (read forwards).
// The gain 1/C is folded into the stencil weights.
// This is the classical B-spline prefilter (Unser et al.) with exact periodic start-up; it solves the
// same system as the reference's cyclic LU (src/bsplinelu.jl:179-220) and FFT (src/bsplinefft.jl:49-58)
// forms: tests/test_host_logic.py compares the host build of this routine with the oracle's LU solve
// (<= 1e-13 relative), the GPU parity tests run through it.
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/slb200.h"

#define SLB_BSPRF_HMAX 6

struct BspRfHost {
    int h, n;
    double z[SLB_BSPRF_HMAX];  // poles inside the unit circle, increasing magnitude
    double invC;
    int K[SLB_BSPRF_HMAX];     // truncation length of G_k (<= n)
    std::vector<double> G;     // concatenated G_0[0..K_0), G_1[0..K_1), ...
    int oG[SLB_BSPRF_HMAX];    // offsets into G
};

static int bsprf_factor(int order, int64_t n64, const double* node_vals, BspRfHost* out, std::string& msg)
{
    typedef long double ld;
    const int h = (order - 1) / 2;
    const int n = (int)n64;
    if (order % 2 == 0 || h < 1 || h > SLB_BSPRF_HMAX) { msg = "recursive-filter B-spline solver: order must be odd and in [3,13]"; return SLB_E_UNSUPPORTED; }
    if (n < 2 * h + 2) { msg = "B-spline: line too short for the stencil (n >= order + 1 required)"; return SLB_E_ARG; }
    ld a[SLB_BSPRF_HMAX + 1];
    for (int m = 0; m <= h; ++m) a[m] = (ld)node_vals[h + m];
    // q(w) = a_0 + sum_m a_m p_m(w), w = z + 1/z, p_0 = 2, p_1 = w, p_{m+1} = w p_m - p_{m-1}
    std::vector<std::vector<ld>> P(h + 1);
    P[0] = {2.0L};
    P[1] = {0.0L, 1.0L};
    for (int m = 1; m < h; ++m) {
        std::vector<ld> nx(m + 2, 0.0L);
        for (size_t i = 0; i < P[m].size(); ++i) nx[i + 1] += P[m][i];
        for (size_t i = 0; i < P[m - 1].size(); ++i) nx[i] -= P[m - 1][i];
        P[m + 1] = nx;
    }
    std::vector<ld> q(h + 1, 0.0L);
    q[0] += a[0];
    for (int m = 1; m <= h; ++m)
        for (size_t i = 0; i < P[m].size(); ++i) q[i] += a[m] * P[m][i];
    auto eval = [](const std::vector<ld>& c, ld w, ld& d) {
        ld r = 0.0L;
        d = 0.0L;
        for (int i = (int)c.size() - 1; i >= 0; --i) {
            d = d * w + r;
            r = r * w + c[i];
        }
        return r;
    };
    // all roots of q are real and < -2 (the symbol is positive on the unit circle): Newton from w = -2
    // converges monotonically to the largest one; deflate, repeat, polish on q itself
    std::vector<ld> zs;
    std::vector<ld> c = q;
    for (int k = 0; k < h; ++k) {
        ld w = -2.0L, d;
        bool ok = false;
        for (int it = 0; it < 500; ++it) {
            ld r = eval(c, w, d);
            if (d == 0.0L) break;
            ld dw = r / d;
            w -= dw;
            if (fabsl(dw) <= fabsl(w) * 1e-16L) { ok = true; break; }  // the polish below finishes the job
        }
        for (int it = 0; it < 4; ++it) {
            ld r = eval(q, w, d);
            if (d != 0.0L) w -= r / d;
        }
        if (!ok || !(w < -2.0L)) { msg = "recursive-filter B-spline solver: pole search failed (symbol not positive?)"; return SLB_E_ARG; }
        zs.push_back(2.0L / (w - sqrtl(w * w - 4.0L)));  // the root of z + 1/z = w inside the unit circle
        std::vector<ld> nc(c.size() - 1, 0.0L);
        ld rem = 0.0L;
        for (int i = (int)c.size() - 1; i >= 1; --i) {
            nc[i - 1] = c[i] + rem;
            rem = nc[i - 1] * w;
        }
        c = nc;
    }
    std::sort(zs.begin(), zs.end(), [](ld x, ld y) { return fabsl(x) < fabsl(y); });
    ld C = a[h];
    for (int k = 0; k < h; ++k) C *= (-1.0L / zs[k]);
    // aliased impulse responses of the cascades 1..k
    const ld zmax = fabsl(zs[h - 1]);
    int J = (int)ceill(-70.0L / log10l(zmax)) + 8;  // the slowest stage has decayed below 1e-70 / (poly factor)
    J = ((J + n - 1) / n + 1) * n;
    std::vector<ld> cur(J, 0.0L), nxt(J);
    cur[0] = 1.0L;
    out->h = h;
    out->n = n;
    out->G.clear();
    for (int k = 0; k < h; ++k) {
        ld s = 0.0L;
        for (int j = 0; j < J; ++j) {
            s = cur[j] + zs[k] * s;
            nxt[j] = s;
        }
        cur = nxt;
        std::vector<ld> Gk(n, 0.0L);
        for (int j = 0; j < J; ++j) Gk[j % n] += cur[j];
        ld gmax = 0.0L;
        for (int r = 0; r < n; ++r) gmax = std::max(gmax, fabsl(Gk[r]));
        int K = 1;
        for (int r = 0; r < n; ++r)
            if (fabsl(Gk[r]) > 1e-19L * gmax) K = r + 1;
        out->K[k] = K;
        out->oG[k] = (int)out->G.size();
        for (int r = 0; r < K; ++r) out->G.push_back((double)Gk[r]);
        out->z[k] = (double)zs[k];
    }
    out->invC = (double)(1.0L / C);
    return SLB_OK;
}

// Kernel table (doubles): [ z[0..h) | invC | pad to even | G_0 | G_1 | ... ]
struct BspRfTab {
    int h, n;
    int K[SLB_BSPRF_HMAX], oG[SLB_BSPRF_HMAX];  // oG: offsets from the start of the table
    int ndoubles;
};

static void bsprf_fill(BspRfTab* tab, std::vector<double>& v, const BspRfHost& hb)
{
    const int h = hb.h;
    const int head = (h + 1 + 1) / 2 * 2;
    tab->h = h;
    tab->n = hb.n;
    v.assign((size_t)head + hb.G.size(), 0.0);
    for (int k = 0; k < h; ++k) {
        v[k] = hb.z[k];
        tab->K[k] = hb.K[k];
        tab->oG[k] = head + hb.oG[k];
    }
    for (int k = h; k < SLB_BSPRF_HMAX; ++k) tab->K[k] = tab->oG[k] = 0;
    v[h] = hb.invC;
    std::copy(hb.G.begin(), hb.G.end(), v.begin() + head);
    tab->ndoubles = (int)v.size();
}

// Start-up states of the cascade: s_k = sum_{r < K_k} G_k[r] x[r0 + dir * r] (two partial sums).
template <int H>
__host__ __device__ __forceinline__ void bsprf_init(const BspRfTab& t, const double* tab, const double* x, int pitch, int r0,
                                                    int dir, double (&s)[H])
{
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const double* G = tab + t.oG[k];
        const int K = t.K[k];
        const double* p = x + (long long)r0 * pitch;
        const int st = dir * pitch;
        double a0 = 0.0, a1 = 0.0;
        int r = 0;
        for (; r + 1 < K; r += 2) {
            a0 = fma(G[r], p[0], a0);
            a1 = fma(G[r + 1], p[st], a1);
            p += 2 * st;
        }
        if (r < K) a0 = fma(G[r], p[0], a0);
        s[k] = a0 + a1;
    }
}

// Same with periodic indexing: x[(r0 + dir * r) mod n] -- start-up of a cascade that begins in the middle
// of the line (the state before ANY row m is sum_r G_k[r] x[(m-1-r) mod n], which is what lets several
// threads work on segments of one line independently).  Four partial sums: the loop is a latency chain.
template <int H>
__host__ __device__ __forceinline__ void bsprf_init_wrap(const BspRfTab& t, const double* tab, const double* x, int pitch, int n,
                                                         int r0, int dir, double (&s)[H])
{
    r0 = r0 < 0 ? r0 + n : (r0 >= n ? r0 - n : r0);
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const double* G = tab + t.oG[k];
        const int K = t.K[k];
        double a[4] = {0.0, 0.0, 0.0, 0.0};
        int idx = r0;
        int r = 0;
        for (; r + 3 < K; r += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                a[q] = fma(G[r + q], x[(long long)idx * pitch], a[q]);
                idx += dir;
                idx = idx < 0 ? idx + n : (idx >= n ? idx - n : idx);
            }
        }
        for (; r < K; ++r) {
            a[0] = fma(G[r], x[(long long)idx * pitch], a[0]);
            idx += dir;
            idx = idx < 0 ? idx + n : (idx >= n ? idx - n : idx);
        }
        s[k] = (a[0] + a[1]) + (a[2] + a[3]);
    }
}

// One pass of the cascade over the line, rows i0, i0 + dir, ... (n rows), in place.  Rows are handled
// in groups of 8 whose inputs are fetched one group ahead; within a group the h stages of consecutive
// rows overlap (row r+1 stage k only waits for row r stage k and row r+1 stage k-1), so the chain is one
// FMA per row deep.
template <int H>
__host__ __device__ __forceinline__ void bsprf_pass(const double (&z)[H], double (&s)[H], double* x, int pitch, int n, int i0,
                                                    int dir)
{
    constexpr int R = 8;
    const int st = dir * pitch;
    double* p = x + (long long)i0 * pitch;
    double nx[R];
#pragma unroll
    for (int r = 0; r < R; ++r) nx[r] = p[(r < n ? r : n - 1) * st];
    int i = 0;
    for (; i + R <= n; i += R) {
        double v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            v[r] = nx[r];
            const int inext = i + R + r;
            nx[r] = p[(inext < n ? inext : n - 1) * st];  // clamped, not predicated
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
#pragma unroll
            for (int k = 0; k < H; ++k) {
                v[r] = fma(z[k], s[k], v[r]);
                s[k] = v[r];
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) p[(i + r) * st] = v[r];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {  // remainder (static indices: nx stays in registers)
        if (i + r < n) {
            double v = nx[r];
#pragma unroll
            for (int k = 0; k < H; ++k) {
                v = fma(z[k], s[k], v);
                s[k] = v;
            }
            p[(i + r) * st] = v;
        }
    }
}

// The per-line arithmetic (host and device): in place on x[0..n) with element stride `pitch`; the
// result is C * A^{-1} u -- the caller folds 1/C (tab[h]) into whatever consumes it.
template <int H>
__host__ __device__ __forceinline__ void bsprf_solve_line(const BspRfTab& t, const double* tab, double* x, int pitch)
{
    const int n = t.n;
    double z[H], s[H];
#pragma unroll
    for (int k = 0; k < H; ++k) z[k] = tab[k];
    bsprf_init<H>(t, tab, x, pitch, n - 1, -1, s);  // causal start-up: s_k = sum_r G_k[r] x[n-1-r]
    bsprf_pass<H>(z, s, x, pitch, n, 0, +1);
    bsprf_init<H>(t, tab, x, pitch, 0, +1, s);      // anticausal start-up on the causal output
    bsprf_pass<H>(z, s, x, pitch, n, n - 1, -1);
}
