// slb_bspfused.cu -- instantiations and host launcher of the fused B-spline sweep (slb_bspfused.cuh).
#define SLB_BSPF_IMPL
#include "slb_bspfused.cuh"

#include <string.h>

#define SLB_BSPF_FOR_H(X) X(1) X(2) X(3) X(4) X(5) X(6)
#define SLB_BSPF_SMEM_MAX (226 * 1024)   // dynamic shared memory available to one block on sm_100

int slb_bspfused_tab_doubles(int h, int n)
{
    const int N = n - h;
    return N * (2 * h) + N * (2 * h + 2) + h * h;
}

int slb_bspfused_warps(int h, int n, bool contig)
{
    if (h < 1 || h > 6 || n < 2 * h + 2) return 0;   // orders 3 .. 13
    const size_t tab = ((size_t)slb_bspfused_tab_doubles(h, n) + 1) / 2 * 2 * sizeof(double);
    const size_t tile = (size_t)n * (contig ? 33 : 32) * sizeof(double);
    if (tab + tile > SLB_BSPF_SMEM_MAX) return 0;
    size_t w = (SLB_BSPF_SMEM_MAX - tab) / tile;
    return (int)(w > 8 ? 8 : w);
}

int slb_bspfused_warps_rf(int ndoubles, int n, bool contig)
{
    const size_t tab = ((size_t)ndoubles + 1) / 2 * 2 * sizeof(double);
    const size_t tile = (size_t)n * (contig ? 33 : 32) * sizeof(double);
    if (tab + tile > SLB_BSPF_SMEM_MAX) return 0;
    size_t w = (SLB_BSPF_SMEM_MAX - tab) / tile;
    return (int)(w > 8 ? 8 : w);
}

bool slb_bspfused_supported(int h, int n) { return slb_bspfused_warps(h, n, true) > 0; }

void slb_bspfused_fill(BspFusedTab* tab, double* v, int h, int n, int N, const double* L, const double* U, const double* invd,
                       const double* Ri, const double* G, const double* Sinv)
{
    const int FR = 2 * h, BR = 2 * h + 2;
    tab->h = h;
    tab->n = n;
    tab->N = N;
    tab->o_bwd = N * FR;
    tab->o_S = tab->o_bwd + N * BR;
    tab->ndoubles = tab->o_S + h * h;
    for (int i = 0; i < N; ++i) {
        double* F = v + (size_t)i * FR;
        double* B = v + tab->o_bwd + (size_t)i * BR;
        B[0] = invd[i];
        for (int j = 0; j < h; ++j) {
            F[j] = L[(size_t)i * h + j];
            F[h + j] = Ri[(size_t)i * h + j];
            B[1 + j] = (double)((long double)U[(size_t)i * h + j] * (long double)invd[i]);
            B[1 + h + j] = G[(size_t)i * h + j];
        }
        B[2 * h + 1] = 0.0;
    }
    memcpy(v + tab->o_S, Sinv, (size_t)h * h * sizeof(double));
}

template <int H, bool CONTIG, bool RF>
static int launch1(const BspFusedArgs& a, const CoefTab& ct, int sm_count, cudaStream_t stream)
{
    auto kern = k_bspline_fused<H, CONTIG, RF>;
    const size_t tab = ((size_t)(RF ? a.rf.ndoubles : a.tab.ndoubles) + 1) / 2 * 2 * sizeof(double);
    const size_t smem = tab + (size_t)a.warps * a.n * (CONTIG ? 33 : 32) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long tiles = (a.nlines + 31) / 32;
    long long blocks = (tiles + a.warps - 1) / a.warps;
    if (blocks > sm_count) blocks = sm_count;   // persistent: one block per SM, each warp loops over tiles
    kern<<<(unsigned)blocks, 32 * a.warps, smem, stream>>>(a, ct);
    return (int)cudaGetLastError();
}

int slb_bspfused_launch(const BspFusedArgs& a, const CoefTab& ct, int sm_count, cudaStream_t stream)
{
    const bool contig = (a.inner == 1);
    if (a.warps < 1 || a.warps > 8) return -1;
    if (a.use_rf) {
        switch (a.rf.h) {
#define X(H) \
    case H:  \
        return contig ? launch1<H, true, true>(a, ct, sm_count, stream) : launch1<H, false, true>(a, ct, sm_count, stream);
            SLB_BSPF_FOR_H(X)
#undef X
        }
        return -1;
    }
    switch (a.tab.h) {
#define X(H) \
    case H:  \
        return contig ? launch1<H, true, false>(a, ct, sm_count, stream) : launch1<H, false, false>(a, ct, sm_count, stream);
        SLB_BSPF_FOR_H(X)
#undef X
    }
    return -1;
}
