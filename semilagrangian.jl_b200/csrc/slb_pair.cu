// slb_pair.cu -- instantiations and host launcher of the pair-fused sweep kernel (slb_pair.cuh).
// A separate translation unit so that the library builds in parallel.
#define SLB_PAIR_IMPL
#include "slb_pair.cuh"

// orders on the fused path: odd Lagrange orders 3..11 and Hermite 5, 9 (order + 1 even)
#define SLB_FUSED_FOR_P1(X) X(4) X(6) X(8) X(10) X(12)

bool slb_fused_supported(int P1, bool cc, int g)
{
    bool okp = false;
#define X(P) okp = okp || (P1 == P);
    SLB_FUSED_FOR_P1(X)
#undef X
    if (!okp) return false;
    if (cc) return g >= 1;
    return g == 32 || g == 16 || g == 4 || g == 1;
}

size_t slb_fused_smem_bytes(int nrows_max, int g)
{
    return ((size_t)SLB_FUSED_STAGES * SLB_FUSED_ROWS * nrows_max * g + g) * sizeof(double);
}

template <int P1, bool EXACT, bool CC, int G, bool W16>
static int launch1(const FusedArgs& fa, const CoefTab& ctA, const CoefTab& ctB, unsigned nblocks, unsigned nthreads,
                   size_t smem, cudaStream_t stream)
{
    auto kern = k_sweep_fused<P1, EXACT, CC, G, W16>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<nblocks, nthreads, smem, stream>>>(fa, ctA, ctB);
    return (int)cudaGetLastError();
}

template <int P1, bool EXACT>
static int launch2(const FusedArgs& fa, const CoefTab& ctA, const CoefTab& ctB, bool cc, unsigned nblocks,
                   unsigned nthreads, size_t smem, cudaStream_t stream)
{
    if (cc)
        return fa.w16 ? launch1<P1, EXACT, true, 0, true>(fa, ctA, ctB, nblocks, nthreads, smem, stream)
                      : launch1<P1, EXACT, true, 0, false>(fa, ctA, ctB, nblocks, nthreads, smem, stream);
    switch (fa.g) {  // g = 16 and 4 imply even strides: always 16-byte fetches (the host checks the base pointer)
    case 32: return fa.w16 ? launch1<P1, EXACT, false, 32, true>(fa, ctA, ctB, nblocks, nthreads, smem, stream) : -1;
    case 16: return fa.w16 ? launch1<P1, EXACT, false, 16, true>(fa, ctA, ctB, nblocks, nthreads, smem, stream) : -1;
    case 4: return fa.w16 ? launch1<P1, EXACT, false, 4, true>(fa, ctA, ctB, nblocks, nthreads, smem, stream) : -1;
    case 1: return fa.w16 ? -1 : launch1<P1, EXACT, false, 1, false>(fa, ctA, ctB, nblocks, nthreads, smem, stream);
    }
    return -1;
}

int slb_fused_launch(const FusedArgs& fa, const CoefTab& ctA, const CoefTab& ctB, int P1, bool exact, bool cc,
                     unsigned nblocks, unsigned nthreads, size_t smem, cudaStream_t stream)
{
    switch (P1) {
#define X(P)                                                                                      \
    case P:                                                                                       \
        return exact ? launch2<P, true>(fa, ctA, ctB, cc, nblocks, nthreads, smem, stream)        \
                     : launch2<P, false>(fa, ctA, ctB, cc, nblocks, nthreads, smem, stream);
        SLB_FUSED_FOR_P1(X)
#undef X
    }
    return -1;
}
