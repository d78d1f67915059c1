// Host-only check of the bordered-LU B-spline solver (factorisation + the per-line solve
// routine the kernels run), callable from tests/ through ctypes without a GPU.
// Not part of libslb200.so's ABI: built as a separate test helper (lib/libslb200_hosttest.so).
#include "slb_bspline.cuh"
#include "slb_points.cuh"
#include "slb_bspsplit.cuh"
#include "slb_bsprf.cuh"

extern "C" int slbt_bspline_solve_host(int order, long long n, const double* node_vals, const double* b, double* x)
{
    BsplineHost hb;
    std::string msg;
    int rc = bspline_factor(order, n, node_vals, &hb, msg);
    if (rc) return rc;
    BsplineDev f;
    f.h = hb.h; f.n = hb.n; f.N = hb.N;
    f.L = hb.L.data(); f.U = hb.U.data(); f.invd = hb.invd.data(); f.Ri = hb.Ri.data(); f.G = hb.G.data(); f.Sinv = hb.Sinv.data();
    for (long long i = 0; i < n; ++i) x[i] = b[i];
    auto LD = [&](int k) { return x[k]; };
    auto ST = [&](int k, double v) { x[k] = v; };
    switch (hb.h) {
#define X(H) case H: bspline_solve_line<H>(f, LD, ST); break;
        X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13)
#undef X
        default: return SLB_E_UNSUPPORTED;
    }
    return 0;
}

// Host run of the per-point 2-D interpolation body (slb_point_eval, the code k_interp2d_points runs
// per thread) over a whole [n1, n2, ncomp] field.  coefA / coefB: (order+1) x nc rows, row-major.
// templated != 0 exercises the unrolled instantiation (needs equal orders, order + 1 <= 14).
extern "C" int slbt_points_host(const double* res, const double* dec, double* out, int n1, int n2, int ncomp, int pA, int ncA,
                                const double* coefA, int pB, int ncB, const double* coefB, int exact, int templated)
{
    PointsArgs pa;
    pa.res = res; pa.dec = dec; pa.out = out;
    pa.n1 = n1; pa.n2 = n2; pa.ncomp = ncomp;
    pa.pA = pA; pa.pB = pB; pa.ncA = ncA; pa.ncB = ncB;
    if (pA > SLB_POINTS_MAXP1 || pB > SLB_POINTS_MAXP1) return SLB_E_ARG;
    for (int j = 0; j < n2; ++j)
        for (int i = 0; i < n1; ++i) {
            if (templated) {
                if (pA != pB) return SLB_E_ARG;
                switch (pA) {
#define X(P) case P: if (exact) slb_point_eval<P, true>(pa, coefA, ncA, coefB, ncB, i, j); else slb_point_eval<P, false>(pa, coefA, ncA, coefB, ncB, i, j); break;
                    X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14)
#undef X
                    default: return SLB_E_UNSUPPORTED;
                }
            } else if (exact)
                slb_point_eval<0, true>(pa, coefA, ncA, coefB, ncB, i, j);
            else
                slb_point_eval<0, false>(pa, coefA, ncA, coefB, ncB, i, j);
        }
    return 0;
}

// Host run of the split-line (two halves + two separators) B-spline solver: factorisation, the
// kernel's table layout and the reference form of its per-line arithmetic (slb_bspsplit.cuh).
extern "C" int slbt_bspsplit_solve_host(int order, long long n, const double* node_vals, const double* b, double* x)
{
    BspSplitHost hb;
    std::string msg;
    int rc = bspsplit_factor(order, n, node_vals, &hb, msg);
    if (rc) return rc;
    BspSplitTab tab;
    std::vector<double> v;
    bspsplit_fill(&tab, v, hb);
    for (long long i = 0; i < n; ++i) x[i] = b[i];
    bspsplit_solve_line(tab, v.data(), x, x);
    return 0;
}

// Host run of the recursive-filter B-spline solver (slb_bsprf.cuh): pole search, start-up tables, the
// kernel's table layout and its per-line arithmetic; x = A^{-1} b.
extern "C" int slbt_bsprf_solve_host(int order, long long n, const double* node_vals, const double* b, double* x, int* K_out)
{
    BspRfHost hb;
    std::string msg;
    int rc = bsprf_factor(order, n, node_vals, &hb, msg);
    if (rc) return rc;
    BspRfTab tab;
    std::vector<double> v;
    bsprf_fill(&tab, v, hb);
    for (long long i = 0; i < n; ++i) x[i] = b[i];
    switch (hb.h) {
#define X(H) case H: bsprf_solve_line<H>(tab, v.data(), x, 1); break;
        X(1) X(2) X(3) X(4) X(5) X(6)
#undef X
        default: return SLB_E_UNSUPPORTED;
    }
    for (long long i = 0; i < n; ++i) x[i] *= v[hb.h];
    if (K_out) for (int k = 0; k < hb.h; ++k) K_out[k] = hb.K[k];
    return 0;
}

// Barrier placement of a step program (slb_program_host.h): op k reads nr[k] and writes nw[k] byte ranges, given one
// after the other in (addr, len, b0) with the reads of an op before its writes; flags_out[k] = 1: a grid barrier
// precedes op k.
#include "slb_program_host.h"
extern "C" int slbt_program_barriers(int nops, const int* nr, const int* nw, const long long* addr, const long long* len, const int* b0,
                                     int* flags_out)
{
    std::vector<ProgAccess> ops((size_t)nops);
    size_t q = 0;
    for (int k = 0; k < nops; ++k) {
        for (int i = 0; i < nr[k]; ++i, ++q) ops[k].reads.push_back(prog_range((const void*)addr[q], (size_t)len[q], b0[q] != 0));
        for (int i = 0; i < nw[k]; ++i, ++q) ops[k].writes.push_back(prog_range((const void*)addr[q], (size_t)len[q], b0[q] != 0));
    }
    const std::vector<int> f = prog_place_barriers(ops);
    for (int k = 0; k < nops; ++k) flags_out[k] = f[k];
    return 0;
}

// the compile-time table of correction counts of the segmented B-spline kernel (slb_bspseg.cuh: slb_seg_cut)
#include "slb_bspseg.cuh"
extern "C" int slbt_seg_cut(int H, int k, int M) { return slb_seg_cut(H, k, M); }
