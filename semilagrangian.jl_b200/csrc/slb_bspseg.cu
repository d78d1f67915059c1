// slb_bspseg.cu -- instantiations, host tables and launchers of the segmented B-spline sweep (slb_bspseg.cuh).
#define SLB_BSPSEG_IMPL
#include "slb_bspseg.cuh"

#include <math.h>
#include <stdlib.h>

bool slb_bspseg_plan(const BspRfHost& hr, bool wline, bool contig, BspSegTab* tab)
{
    typedef long double ld;
    const int h = hr.h, n = hr.n;
    if (h < 1 || h > SLB_SEG_HMAX) return false;
    int M = 0, S = 0;
    if (wline) {
        if (n % 32 != 0) return false;
        M = n / 32;
        S = 32;
        if (!(M == 16 || M == 32 || M == 64) || 2 * h + 1 > M) return false;
    } else {
        switch (n) {  // the instantiated (M, S) pairs of slb_bspseg_launch
        case 16: M = 8; break;
        case 32: M = 16; break;
        case 64: M = 16; break;
        case 128:
            // measured at 128^4 (ms per sweep, M = 16 x 8 warps | M = 32 x 4 warps): strided order 11 1.42 | 1.15, order 5
            // 0.80 | 0.70, order 3 0.70 | 0.66; dim 0, with the tile sharing the exchange buffers' memory and 5 blocks per SM:
            // order 9 1.40 | 1.11, order 7 1.18 | 1.00, order 5 1.03 | 0.97, order 3 0.80 | 0.89
            M = (!contig || h >= 2) ? 32 : 16;
            if (getenv("SLB_SEG_M128")) M = atoi(getenv("SLB_SEG_M128"));
            break;
        case 256: M = 32; break;
        default: return false;
        }
        S = n / M;
    }
    memset(tab, 0, sizeof(*tab));
    tab->h = h;
    tab->n = n;
    tab->M = M;
    tab->S = S;
    tab->invC = hr.invC;
    for (int k = 0; k < h; ++k) {
        const ld z = (ld)hr.z[k];
        tab->z[k] = hr.z[k];
        ld p = z;
        for (int j = 0; j < M; ++j) {
            tab->zp[k][j] = (double)p;  // z^(j+1)
            // the block kernel drops the corrections of rows j >= slb_seg_cut(h, k, M): they must be below 1e-19
            if (!wline && j >= slb_seg_cut(h, k, M) && fabsl(p) >= 1e-19L) return false;
            p *= z;
        }
        const ld zM = powl(z, (ld)M), zn = powl(z, (ld)n);
        // terms of the carry sum: all S segments of the ring (exact periodic closure), or fewer once z^(M m) has
        // decayed below 1e-19 (then 1 / (1 - z^n) differs from 1 by less than that, too)
        int nm = S;
        ld q = 1.0L;
        for (int m = 0; m < S; ++m) {
            if (fabsl(q) < 1e-19L) {
                nm = m;
                break;
            }
            q *= zM;
        }
        if (nm < 1) nm = 1;
        tab->nm[k] = nm;
        q = 1.0L / (1.0L - zn);
        for (int m = 0; m < nm; ++m) {
            tab->cz[k][m] = (double)q;
            q *= zM;
        }
    }
    return true;
}

static size_t seg_smem(int h, int M, int S, bool contig)
{
    const int P1 = 2 * h + 2, HALO = P1 - 1, HM = HALO < M ? HALO : M;
    const size_t exch = (size_t)2 * S * 32 + (size_t)P1 * 32 + (size_t)S * HM * 32 + (size_t)S * 32;
    const size_t tile = contig ? (size_t)M * S * 33 : 0;   // shares its memory with the exchange buffers
    const size_t d = (size_t)h * S + (size_t)h * M + 64 + (tile > exch ? tile : exch);
    return d * sizeof(double);
}

template <int H, int M, int S, bool CONTIG, int MINB = 0>
static int seg_launch1(const BspSegArgs& a, const CoefTab& ct, cudaStream_t stream)
{
    auto kern = k_bspline_seg<H, M, S, CONTIG, MINB>;
    const size_t smem = seg_smem(H, M, S, CONTIG);
    if (smem > 200 * 1024) return -1;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const long long nb = (a.nlines + 31) / 32;
    if (nb > 0x7fffffffLL) return -1;
    kern<<<(unsigned)nb, 32 * S, smem, stream>>>(a, ct);
    return (int)cudaGetLastError();
}

// line lengths on this path: n = M S with (M, S) one of the pairs below (powers of two from 16 to 256)
template <int H>
static int seg_launch_h(const BspSegArgs& a, const CoefTab& ct, bool contig, cudaStream_t stream)
{
#define SLB_SEG_CASE(MM, SS)                                                                                    \
    if (a.tab.M == MM && a.tab.S == SS)                                                                         \
        return contig ? seg_launch1<H, MM, SS, true>(a, ct, stream) : seg_launch1<H, MM, SS, false>(a, ct, stream);
    SLB_SEG_CASE(8, 2)
    SLB_SEG_CASE(16, 2)
    SLB_SEG_CASE(16, 4)
    SLB_SEG_CASE(16, 8)
    if (a.tab.M == 32 && a.tab.S == 4) {
        // 128-point lines: capping the registers for 5 resident blocks per SM (96 registers, 20 warps instead of 16)
        // pays at the high orders (strided, ms per 128^4 sweep: order 11 1.05 -> 0.96, order 9 0.90 -> 0.82, order 7 0.78 -> 0.70,
        // order 5 0.66 -> 0.64; 6 blocks = 80 registers spill: 1.33); SLB_SEG_MINB (strided) / SLB_SEG_MINB_C (dim 0) = 0 | 5 | 6
        static const int mbs = getenv("SLB_SEG_MINB") ? atoi(getenv("SLB_SEG_MINB")) : 5;
        static const int mbc = getenv("SLB_SEG_MINB_C") ? atoi(getenv("SLB_SEG_MINB_C")) : 5;
        const int mb = contig ? mbc : mbs;
        if (mb == 5) return contig ? seg_launch1<H, 32, 4, true, 5>(a, ct, stream) : seg_launch1<H, 32, 4, false, 5>(a, ct, stream);
        if (mb == 6) return contig ? seg_launch1<H, 32, 4, true, 6>(a, ct, stream) : seg_launch1<H, 32, 4, false, 6>(a, ct, stream);
    }
    SLB_SEG_CASE(32, 4)
    SLB_SEG_CASE(32, 8)
#undef SLB_SEG_CASE
    return -1;
}

int slb_bspseg_launch(const BspSegArgs& a, const CoefTab& ct, bool contig, cudaStream_t stream)
{
    if (a.tab.S * a.tab.M != a.n) return -1;
    switch (a.tab.h) {
    case 1: return seg_launch_h<1>(a, ct, contig, stream);
    case 2: return seg_launch_h<2>(a, ct, contig, stream);
    case 3: return seg_launch_h<3>(a, ct, contig, stream);
    case 4: return seg_launch_h<4>(a, ct, contig, stream);
    case 5: return seg_launch_h<5>(a, ct, contig, stream);
    }
    return -1;
}

template <int H, int M>
static int wline_launch1(const BspSegArgs& a, const CoefTab& ct, cudaStream_t stream)
{
    if constexpr (2 * H + 1 > M) return -1;
    else {
        const long long nb = (a.nlines + 3) / 4;
        if (nb > 0x7fffffffLL) return -1;
        k_bspline_wline<H, M><<<(unsigned)nb, 128, 0, stream>>>(a, ct);
        return (int)cudaGetLastError();
    }
}

template <int H>
static int wline_launch_h(const BspSegArgs& a, const CoefTab& ct, cudaStream_t stream)
{
    switch (a.tab.M) {
    case 16: return wline_launch1<H, 16>(a, ct, stream);
    case 32: return wline_launch1<H, 32>(a, ct, stream);
    case 64: return wline_launch1<H, 64>(a, ct, stream);
    }
    return -1;
}

int slb_bspwline_launch(const BspSegArgs& a, const CoefTab& ct, cudaStream_t stream)
{
    if (a.tab.S != 32 || 32 * a.tab.M != a.n) return -1;
    switch (a.tab.h) {
    case 1: return wline_launch_h<1>(a, ct, stream);
    case 2: return wline_launch_h<2>(a, ct, stream);
    case 3: return wline_launch_h<3>(a, ct, stream);
    case 4: return wline_launch_h<4>(a, ct, stream);
    case 5: return wline_launch_h<5>(a, ct, stream);
    }
    return -1;
}
