// slb_bspsplit.cu -- instantiations and host launcher of the split-line fused B-spline sweep
// (slb_bspsplit.cuh): two warps per tile of 32 lines, strided dims.
#define SLB_BSPS_IMPL
#include "slb_bspsplit.cuh"

#define SLB_BSPS_FOR_H(X) X(1) X(2) X(3) X(4) X(5) X(6)
#define SLB_BSPS_SMEM_MAX (224 * 1024)   // dynamic shared memory budget of one block (static: 1.8 KB)

static int bsps_tab_doubles(int h, int n)
{
    const int Na = n / 2 - h;
    return Na * (bsps_FR(h) + bsps_BR(h)) + 4 * h * h;
}

int slb_bspsplit_tiles(int h, int n)
{
    if (h < 1 || h > 6 || n % 2 != 0 || n / 2 - h < 2 * h + 1) return 0;
    const size_t tab = ((size_t)bsps_tab_doubles(h, n) + 1) / 2 * 2 * sizeof(double);
    const size_t tile = (size_t)n * 32 * sizeof(double);
    if (tab + tile > SLB_BSPS_SMEM_MAX) return 0;
    size_t w = (SLB_BSPS_SMEM_MAX - tab) / tile;
    return (int)(w > SLB_BSPS_MAXTILES ? SLB_BSPS_MAXTILES : w);
}

int slb_bspsplit_tiles_rf(int ndoubles, int n)
{
    if (n % 2 != 0 || n < 4) return 0;
    const size_t tab = ((size_t)ndoubles + 1) / 2 * 2 * sizeof(double);
    const size_t tile = (size_t)n * 32 * sizeof(double);
    if (tab + tile > SLB_BSPS_SMEM_MAX) return 0;
    size_t w = (SLB_BSPS_SMEM_MAX - tab) / tile;
    return (int)(w > SLB_BSPS_MAXTILES_RF ? SLB_BSPS_MAXTILES_RF : w);
}

template <int H, bool RF>
static int launch1(const BspSplitArgs& a, const CoefTab& ct, int sm_count, cudaStream_t stream)
{
    auto kern = k_bspline_split<H, RF>;
    const size_t tab = ((size_t)(RF ? a.rf.ndoubles : a.tab.ndoubles) + 1) / 2 * 2 * sizeof(double);
    const size_t smem = tab + (size_t)a.tiles * a.n * 32 * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long tiles = (a.nlines + 31) / 32;
    long long blocks = (tiles + a.tiles - 1) / a.tiles;
    if (blocks > sm_count) blocks = sm_count;   // persistent: one block per SM, each warp pair loops over tiles
    kern<<<(unsigned)blocks, 64 * a.tiles, smem, stream>>>(a, ct);
    return (int)cudaGetLastError();
}

int slb_bspsplit_launch(const BspSplitArgs& a, const CoefTab& ct, int sm_count, cudaStream_t stream)
{
    if (a.use_rf) {
        if (a.tiles < 1 || a.tiles > SLB_BSPS_MAXTILES_RF) return -1;
        switch (a.rf.h) {
#define X(H) \
    case H:  \
        return launch1<H, true>(a, ct, sm_count, stream);
            SLB_BSPS_FOR_H(X)
#undef X
        }
        return -1;
    }
    if (a.tiles < 1 || a.tiles > SLB_BSPS_MAXTILES) return -1;
    switch (a.tab.h) {
#define X(H) \
    case H:  \
        return launch1<H, false>(a, ct, sm_count, stream);
        SLB_BSPS_FOR_H(X)
#undef X
    }
    return -1;
}
