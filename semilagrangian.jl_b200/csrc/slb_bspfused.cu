// slb_bspfused.cu -- instantiations and host launcher of the fused B-spline sweep (slb_bspfused.cuh).
#define SLB_BSPF_IMPL
#include "slb_bspfused.cuh"

#include <string.h>

#define SLB_BSPF_FOR_H(X) X(1) X(2) X(3) X(4) X(5) X(6)

bool slb_bspfused_supported(int h, int n)
{
    if (h < 1 || h > 6) return false;                      // orders 3 .. 13
    if (n < 2 * h + 2) return false;
    if ((size_t)n * 33 * sizeof(double) > 200 * 1024) return false;   // one tile of 32 lines per warp in shared memory
    const int N = n - h;
    return (long long)N * (1 + 4 * h) + h * h <= SLB_BSPF_TAB;
}

bool slb_bspfused_fill(BspParamTab* tab, int h, int n, int N, const double* L, const double* U, const double* invd,
                       const double* Ri, const double* G, const double* Sinv)
{
    if (!slb_bspfused_supported(h, n)) return false;
    memset(tab, 0, sizeof(*tab));
    tab->h = h;
    tab->n = n;
    tab->N = N;
    const int TS = 4 * h + 1;
    for (int i = 0; i < N; ++i) {
        double* T = tab->v + (size_t)i * TS;
        T[0] = invd[i];
        for (int j = 0; j < h; ++j) {
            T[1 + j] = L[(size_t)i * h + j];
            T[1 + h + j] = (double)((long double)U[(size_t)i * h + j] * (long double)invd[i]);
            T[1 + 2 * h + j] = Ri[(size_t)i * h + j];
            T[1 + 3 * h + j] = G[(size_t)i * h + j];
        }
    }
    tab->o_S = N * TS;
    memcpy(tab->v + tab->o_S, Sinv, (size_t)h * h * sizeof(double));
    return true;
}

template <int H, bool CONTIG>
static int launch1(const BspFusedArgs& a, const BspParamTab& tab, const CoefTab& ct, cudaStream_t stream)
{
    auto kern = k_bspline_fused<H, CONTIG>;
    const size_t smem = (size_t)a.n * (CONTIG ? 33 : 32) * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const long long tiles = (a.nlines + 31) / 32;
    if (tiles >= 0x7fffffffLL) return -1;
    kern<<<(unsigned)tiles, 32, smem, stream>>>(a, tab, ct);
    return (int)cudaGetLastError();
}

int slb_bspfused_launch(const BspFusedArgs& a, const BspParamTab& tab, const CoefTab& ct, int sm_count, cudaStream_t stream)
{
    (void)sm_count;
    const bool contig = (a.inner == 1);
    switch (tab.h) {
#define X(H) \
    case H:  \
        return contig ? launch1<H, true>(a, tab, ct, stream) : launch1<H, false>(a, tab, ct, stream);
        SLB_BSPF_FOR_H(X)
#undef X
    }
    return -1;
}
