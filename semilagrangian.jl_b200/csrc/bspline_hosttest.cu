// Host-only check of the bordered-LU B-spline solver (factorisation + the per-line solve
// routine the kernels run), callable from tests/ through ctypes without a GPU.
// Not part of libslb200.so's ABI: built as a separate test helper (lib/libslb200_hosttest.so).
#include "slb_bspline.cuh"

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
