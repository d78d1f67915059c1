// slb_pair.cu -- instantiations and host launcher of the pair-fused sweep kernel (slb_pair.cuh).
// Compiled once per stencil width (-DSLB_PAIR_P1=4|6|8|10|12: the kernels of that order) and once without
// the macro (the dispatcher), so that the library builds in parallel.
#define SLB_PAIR_IMPL
#include "slb_pair.cuh"

// orders on the fused path: odd Lagrange orders 3..11 and Hermite 5, 9 (order + 1 even)
#define SLB_FUSED_FOR_P1(X) X(4) X(6) X(8) X(10) X(12)

#define SLB_CAT2(a, b) a##b
#define SLB_CAT(a, b) SLB_CAT2(a, b)

#ifdef SLB_PAIR_P1
template <int P1, bool EXACT, bool CC, int G, bool W16, int MODE>
static int launch1(const FusedArgs& fa, const CoefTab& ctA, const CoefTab& ctB, unsigned nblocks, unsigned nthreads,
                   size_t smem, cudaStream_t stream)
{
    auto kern = k_sweep_fused<P1, EXACT, CC, G, W16, MODE>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    kern<<<nblocks, nthreads, smem, stream>>>(fa, ctA, ctB);
    return (int)cudaGetLastError();
}

template <int P1, bool EXACT, int MODE>
static int launch2(const FusedArgs& fa, const CoefTab& ctA, const CoefTab& ctB, bool cc, unsigned nblocks,
                   unsigned nthreads, size_t smem, cudaStream_t stream)
{
    if (!cc && MODE == SLB_FUSED_RHO) return -1;
    if (cc) {
        if constexpr (MODE == SLB_FUSED_WIN) return -1;  // the march dim is never dim 0's neighbour pass here
        else
            return fa.w16 ? launch1<P1, EXACT, true, 0, true, MODE>(fa, ctA, ctB, nblocks, nthreads, smem, stream)
                          : launch1<P1, EXACT, true, 0, false, MODE>(fa, ctA, ctB, nblocks, nthreads, smem, stream);
    }
    if constexpr (MODE == SLB_FUSED_PSH || MODE == SLB_FUSED_RHO) return -1;  // built for the CC (x1 x2) pass only
    else {
        switch (fa.g) {  // g = 16 and 4 imply even strides: always 16-byte fetches (the host checks the base pointer)
        case 32: return fa.w16 ? launch1<P1, EXACT, false, 32, true, MODE>(fa, ctA, ctB, nblocks, nthreads, smem, stream) : -1;
        case 16: return fa.w16 ? launch1<P1, EXACT, false, 16, true, MODE>(fa, ctA, ctB, nblocks, nthreads, smem, stream) : -1;
        case 4: return fa.w16 ? launch1<P1, EXACT, false, 4, true, MODE>(fa, ctA, ctB, nblocks, nthreads, smem, stream) : -1;
        case 1: return fa.w16 ? -1 : launch1<P1, EXACT, false, 1, false, MODE>(fa, ctA, ctB, nblocks, nthreads, smem, stream);
        }
        return -1;
    }
}

int SLB_CAT(slb_fused_launch_p, SLB_PAIR_P1)(const FusedArgs& fa, const CoefTab& ctA, const CoefTab& ctB, bool exact, bool cc, int mode,
                                              unsigned nblocks, unsigned nthreads, size_t smem, cudaStream_t stream)
{
    constexpr int P = SLB_PAIR_P1;
    if (mode == SLB_FUSED_PLAIN)
        return exact ? launch2<P, true, SLB_FUSED_PLAIN>(fa, ctA, ctB, cc, nblocks, nthreads, smem, stream)
                     : launch2<P, false, SLB_FUSED_PLAIN>(fa, ctA, ctB, cc, nblocks, nthreads, smem, stream);
    if (exact) return -1;  // the halo-sharded modes are instantiated for the production (FMA) arithmetic
    if (mode == SLB_FUSED_WIN) return launch2<P, false, SLB_FUSED_WIN>(fa, ctA, ctB, cc, nblocks, nthreads, smem, stream);
    if (mode == SLB_FUSED_PSH) return launch2<P, false, SLB_FUSED_PSH>(fa, ctA, ctB, cc, nblocks, nthreads, smem, stream);
    if (mode == SLB_FUSED_RHO) return launch2<P, false, SLB_FUSED_RHO>(fa, ctA, ctB, cc, nblocks, nthreads, smem, stream);
    return -1;
}

#else  // dispatcher

#define X(P)                                                                                                             \
    int slb_fused_launch_p##P(const FusedArgs& fa, const CoefTab& ctA, const CoefTab& ctB, bool exact, bool cc, int mode, \
                              unsigned nblocks, unsigned nthreads, size_t smem, cudaStream_t stream);
SLB_FUSED_FOR_P1(X)
#undef X

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

int slb_fused_launch(const FusedArgs& fa, const CoefTab& ctA, const CoefTab& ctB, int P1, bool exact, bool cc, int mode,
                     unsigned nblocks, unsigned nthreads, size_t smem, cudaStream_t stream)
{
    switch (P1) {
#define X(P) \
    case P: return slb_fused_launch_p##P(fa, ctA, ctB, exact, cc, mode, nblocks, nthreads, smem, stream);
        SLB_FUSED_FOR_P1(X)
#undef X
    }
    return -1;
}
#endif
