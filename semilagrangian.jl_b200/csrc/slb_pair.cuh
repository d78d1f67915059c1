// slb_pair.cuh -- K7: two consecutive 1-D sweeps fused into ONE pass over HBM.
//
// A Strang step of a 2D2V grid runs its sweeps in pairs (v1 v2 | x1 x2 | v1 v2,
// examples/vlasov-poisson-2d2v.jl:126-131); the reference executes each as a separate
// advection! call (src/advection.jl:594-657), i.e. one full read + write of f per sweep.
// Here a pair (A along the "cross" dim, then B along the "march" dim) is one kernel that reads f
// once and writes it once -- 16 B of HBM traffic per cell for TWO cell-updates:
//
//   * a CTA owns a tile of `ta` consecutive cross indices x `g` points of the remaining
//     ("passive") dims and marches over the whole march dim, two input rows per block barrier;
//     every thread produces TWO neighbouring cross outputs (their stencils share order of the
//     order+2 staged inputs, and all per-row bookkeeping is paid once for both);
//   * sweep A is a cross-THREAD stencil: the step's input rows (tile + periodic halo) are staged
//     in shared memory by cp.async (a ring of D stages, filled D-1 steps ahead, so the loads in
//     flight cost no registers), and thread (p, a) forms T = sum_q w1[q] * raw[a + s0A(p) + q];
//   * sweep B runs along the march direction in REGISTERS: every thread keeps a rotating window
//     of its last order+1 values of T and emits one output per step,
//         out[i] = sum_q w2[q] * T[(i + s0B + q) mod n],   i = (step - order - s0B) mod n.
//
// The intermediate T is rounded to Float64 exactly as when a separate sweep stores it, and both
// stencils use the same operation order as k_sweep_strided / k_sweep_contig, so the result is
// BIT-IDENTICAL to two slb_sweep calls.  Lane order follows the memory-contiguous index:
// CC = true  (cross dim is dim 0, e.g. x1 x2): lanes run along the cross index;
// CC = false (dim 0 is passive, e.g. v1 v2)  : lanes run along the passive index (g = 32: 256 B rows).
// Shifts may differ between the passive points of a tile: the staged row range is the union over
// the tile (up to SLB_FUSED_SPREAD_MAX extra rows); beyond that the CTA reads its stencil inputs
// straight from global memory (correct, slower).  alpha_A must not depend on the march index
// (it would change w1 every step); alpha_B may depend on everything but the march index.
//
// (A first version staged T of whole chunks through a ring buffer kept resident in the L2:
// ncu showed the ring does stay in the L2 up to ~24 MB -- DRAM traffic halved -- but one thread
// per line keeps >= 50 MB of intermediate in flight at full occupancy, so it never beat two
// separate sweeps; see profiles/r1_pair_l2_experiment.txt.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slb_sweep.cuh"

#define SLB_FUSED_SPREAD_MAX 16
#define SLB_FUSED_STAGES 5       // shared-memory ring: D stages ...
#define SLB_FUSED_ROWS 2         // ... of R march rows each; one block barrier per stage
#define SLB_FUSED_MAXTHREADS 256

struct FusedArgs {
    const double* in;
    double* out;
    int ncross, nmarch;          // extents of the cross (sweep A) and march (sweep B) dims
    unsigned elo, ehi;           // extents of the (up to two) passive index groups
    // element strides of (cross, march, passive lo, passive hi) on the input and on the output side.
    // Either side may be BLOCK-MAJOR along the march dim (multi-GPU re-shard fused into the pass,
    // SURVEY.md 8e): march index b lives in block b / kc at position b % kc; input blocks are
    // iblk elements apart, output block q starts at oblk[q] (this GPU's HBM or a peer's, mapped
    // through CUDA IPC).  kc == nmarch: plain layout.
    long long isc, ism, islo, ishi;
    long long osc, osm, oslo, oshi;
    int ikc, okc;
    long long iblk;
    int march0;                  // march index of the first input row (a multiple of okc): the line is
                                 // periodic, so the march may start anywhere -- ranks start at different
                                 // blocks so that their peer stores hit different destinations at any time
    double* oblk[SLB_MAX_PEERS];
    int g, ta;                   // passive points / cross outputs (even) per tile; blockDim = g * ta / 2
    int ntile_c;                 // tiles along the cross dim
    int full;                    // 1: a tile is the whole periodic cross line
    int nrows_max;               // staged rows per march row (shared-memory pitch)
    int spread_max;              // largest shift spread inside a tile that is still staged
    int w16;                     // rows fetched with 16-byte cp.async (alignment verified by the host)
    const double* tabA;          // alpha_A = scaleA * tabA[plo*aAlo + phi*aAhi]
    double scaleA;
    long long aAlo, aAhi;
    const double* tabB;          // alpha_B = scaleB * tabB[plo*aBlo + phi*aBhi]
    double scaleB;
    long long aBlo, aBhi;
    int ncA, ncB;                // polynomial coefficients per weight
    double* linesum;             // optional: per (passive, cross) sum over the march dim of the outputs
    long long lslo, lshi, lsc;
    // ---- halo-sharded grids (MODE != 0): one dim is split over the ranks in slabs of win_c
    // points and every rank keeps win_h halo points of its two neighbours on either side of its slab.
    // MODE 1 (SLB_FUSED_WIN): the MARCH dim is the sharded one.  The arrays hold nmarch = win_c + 2 win_h rows
    //   [low halo | slab | high halo]; the march is not periodic: it runs once over those rows and emits the
    //   win_c outputs of the slab (rows win_h .. win_h + win_c).  A shift whose stencil leaves the halo sets *err.
    // MODE 2 (SLB_FUSED_PSH): a PASSIVE dim is the sharded one (the arrays are the slab, no halo rows involved
    //   in the pass itself).
    // In both modes the outputs that lie within win_h of a slab boundary are ALSO stored into the neighbour's
    // halo (pushL / pushR != NULL): pushL is the address in the lower neighbour's output array that
    // corresponds to this rank's output element 0 shifted by +win_c along the sharded dim, pushR the upper
    // neighbour's shifted by -win_c, so that `own address - oblk[0] + push?` is the halo element.
    int win_h, win_c;
    int push_on_lo;              // MODE 2: the sharded passive index is plo (else phi)
    double* pushL;
    double* pushR;
    int* err;
    // MODE 3 (SLB_FUSED_RHO, CC = true, a tile is the whole cross line): the g passive points of a block share the
    // shift of sweep B, so they emit the same march rows at the same time; every two rows the block adds its g
    // outputs per cross index through shared memory (fixed order) and stores the sums as partial plane `block group`
    // of rhopart[group][march][cross] -- the charge density after a space pass without another pass over f:
    // 1/g of the bytes are re-read instead of all of them.
    double* rhopart;
};
#define SLB_FUSED_PLAIN 0
#define SLB_FUSED_WIN 1
#define SLB_FUSED_PSH 2
#define SLB_FUSED_RHO 3

// host launcher (slb_pair.cu); returns cudaGetLastError() of the launch, or -1 when (P1, G) is not instantiated
int slb_fused_launch(const FusedArgs& fa, const CoefTab& ctA, const CoefTab& ctB, int P1, bool exact, bool cc, int mode,
                     unsigned nblocks, unsigned nthreads, size_t smem_bytes, cudaStream_t stream);
bool slb_fused_supported(int P1, bool cc, int g);
size_t slb_fused_smem_bytes(int nrows_max, int g);

#ifdef SLB_PAIR_IMPL
__device__ __forceinline__ void fused_cp_async8(unsigned smem_dst, const void* gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void fused_cp_async16(unsigned smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void fused_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fused_cp_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// predicated streaming stores (no branch around them: the flags are per-thread constants)
__device__ __forceinline__ void fused_st_v2(bool on, double* p, double a, double b)
{
    // no "memory" clobber: nothing in the kernel reads the output array, and a clobber would pin
    // the shared-memory loads of the next rows behind every store
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q st.global.cs.v2.f64 [%0], {%1, %2}; }" ::"l"(p), "d"(a), "d"(b),
                 "r"((unsigned)on));
}
__device__ __forceinline__ void fused_st(bool on, double* p, double a)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q st.global.cs.f64 [%0], %1; }" ::"l"(p), "d"(a), "r"((unsigned)on));
}

// one term of slb_dot: same rounding as its EXACT / FMA branches
template <bool EXACT>
__device__ __forceinline__ double slb_mul(double x, double w)
{
    return EXACT ? __dmul_rn(x, w) : x * w;
}
template <bool EXACT>
__device__ __forceinline__ double slb_acc(double acc, double x, double w)
{
    return EXACT ? __dadd_rn(acc, __dmul_rn(x, w)) : fma(x, w, acc);
}
// sum_{j < P1} x[OFF + j] * w[j], left to right (== slb_dot on the shifted window)
template <int P1, bool EXACT, int OFF, int NX>
__device__ __forceinline__ double fused_dot(const double (&x)[NX], const double (&w)[P1])
{
    double acc = slb_mul<EXACT>(x[OFF], w[0]);
#pragma unroll
    for (int j = 1; j < P1; ++j) acc = slb_acc<EXACT>(acc, x[OFF + j], w[j]);
    return acc;
}
// the first NOLD terms of the march-direction dot product whose window starts at slot `rot`
template <int P1, bool EXACT, int NOLD>
__device__ __forceinline__ double fused_partial(const double (&win)[P1], const double (&w)[P1], int rot)
{
    double acc = slb_mul<EXACT>(win[rot % P1], w[0]);
#pragma unroll
    for (int j = 1; j < NOLD; ++j) acc = slb_acc<EXACT>(acc, win[(rot + j) % P1], w[j]);
    return acc;
}

template <int P1>
__device__ __forceinline__ void fused_weights(const CoefTab& ct, int nc, double t, double (&w)[P1])
{
#pragma unroll
    for (int j = 0; j < P1; ++j) w[j] = ct.c[j * SLB_NCMAX + nc - 1];
    for (int k = nc - 2; k >= 0; --k) {
#pragma unroll
        for (int j = 0; j < P1; ++j) w[j] = fma(t, w[j], ct.c[j * SLB_NCMAX + k]);
    }
}

// G > 0: passive points per tile fixed at compile time (CC = false); G == 0: run-time fa.g (CC = true)
// W16: rows are fetched with 16-byte cp.async (pairs of doubles; the host checks alignment)
// MODE: SLB_FUSED_PLAIN | SLB_FUSED_WIN (CC = false only) | SLB_FUSED_PSH | SLB_FUSED_RHO (CC = true only) (see FusedArgs)
template <int P1, bool EXACT, bool CC, int G, bool W16, int MODE>
__global__ void __launch_bounds__(SLB_FUSED_MAXTHREADS, (P1 <= 10 ? 2 : 1))  // orders <= 9: 128 registers, two blocks per SM
k_sweep_fused(const __grid_constant__ FusedArgs fa, const __grid_constant__ CoefTab ctA, const __grid_constant__ CoefTab ctB)
{
    constexpr int D = SLB_FUSED_STAGES;
    constexpr int R = SLB_FUSED_ROWS;
    constexpr int HALF = (P1 - 1) / 2;
    static_assert(P1 % 2 == 0 && R == 2, "two march rows per barrier; order + 1 must be even");
    extern __shared__ __align__(16) double fsm[];
    const int NT = blockDim.x, tid = threadIdx.x;
    const int g = G > 0 ? G : fa.g;
    const int ta = fa.ta, ta2 = ta >> 1, nc_ = fa.ncross, nm = fa.nmarch;
    int p, a;  // passive point and FIRST cross output (local index, even) of this thread
    if (CC) {
        a = 2 * (tid % ta2);
        p = tid / ta2;
    } else {
        p = tid % g;
        a = 2 * (tid / g);
    }
    const int tc = (int)(blockIdx.x % (unsigned)fa.ntile_c);
    const long long tp = blockIdx.x / (unsigned)fa.ntile_c;
    const int a0 = tc * ta;
    const long long np = (long long)fa.elo * fa.ehi;
    const long long P = tp * g + p;
    const bool act0 = (P < np) && (a0 + a < nc_);
    const bool act1 = (P < np) && (a0 + a + 1 < nc_);
    const long long Pc = P < np ? P : np - 1;  // idle threads mimic a valid point (they never store)
    const int ac = (a0 + a < nc_) ? a0 + a : nc_ - 1;
    const unsigned plo = (unsigned)(Pc % fa.elo), phi = (unsigned)(Pc / fa.elo);
    const long long obase = (long long)plo * fa.oslo + (long long)phi * fa.oshi;

    // ---- shifts and weights of this thread's passive point: sweep A (cross), sweep B (march) ----
    double w1[P1], w2[P1];
    int s0A, s0B;
    long long dA;
    {
        const double alpha = fa.scaleA * __ldg(fa.tabA + (long long)plo * fa.aAlo + (long long)phi * fa.aAhi);
        double t;
        slb_split(alpha, nc_, HALF, t, s0A);
        const double fl = floor(alpha);
        dA = fabs(fl) < 4.0e18 ? (long long)fl : (fl < 0 ? -4000000000000000000LL : 4000000000000000000LL);
        fused_weights<P1>(ctA, fa.ncA, t, w1);
    }
    int lo = 0;  // MODE WIN: row (in the haloed array) of the next output this thread emits
    {
        const double alpha = fa.scaleB * __ldg(fa.tabB + (long long)plo * fa.aBlo + (long long)phi * fa.aBhi);
        double t;
        slb_split(alpha, nm, HALF, t, s0B);
        fused_weights<P1>(ctB, fa.ncB, t, w2);
        if (MODE == SLB_FUSED_WIN) {
            const double fl = floor(alpha);
            const int di = fabs(fl) < 1.0e9 ? (int)fl : (fl < 0 ? -1000000000 : 1000000000);
            lo = HALF - di;  // output row emitted at step P1 - 1: its stencil starts at input row 0
            // the slab's outputs need input rows win_h + di - HALF .. win_h + win_c + di + HALF of [0, nmarch)
            if (fa.win_h + di - HALF < 0 || di + HALF + 1 > fa.win_h) atomicOr(fa.err, 1);
        }
    }

    // ---- staged row range: union over the tile's passive points --------------------------------
    const int row_elems = fa.nrows_max * g;       // one march row of the tile in shared memory
    long long* dsh = reinterpret_cast<long long*>(fsm + (size_t)D * R * row_elems);
    double* const redbuf =   // MODE RHO: [2 parities][2 rows][g][ta], 16-byte aligned
        reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(dsh + g) + 15) & ~static_cast<uintptr_t>(15));
    if (a == 0) dsh[p] = dA;
    __syncthreads();
    long long dmin = dsh[0], dmax = dsh[0];
    for (int q = 1; q < g; ++q) {
        long long v = dsh[q];
        dmin = v < dmin ? v : dmin;
        dmax = v > dmax ? v : dmax;
    }
    const bool full = fa.full != 0;
    const bool direct = !full && (dmax - dmin > fa.spread_max);  // uniform over the CTA
    const int spread = full || direct ? 0 : (int)(dmax - dmin);
    int nrows = full ? nc_ + P1 : ta + P1 - 1 + spread;
    if (W16 && CC) nrows = (nrows + 2) & ~1;  // even, with room for the alignment element
    int rowbase = 0, rowpad = 0;
    if (!full) {
        long long rb = ((long long)a0 + dmin - HALF) % nc_;
        rowbase = (int)(rb < 0 ? rb + nc_ : rb);
        if (W16 && CC) {  // 16-byte fetches along the cross dim start at an even index
            rowpad = rowbase & 1;
            rowbase -= rowpad;
        }
    }
    int roff;  // first staged row of this thread's stencils (inputs roff .. roff + P1)
    if (full) {
        roff = ac + s0A;
        roff -= roff >= nc_ ? nc_ : 0;
    } else {
        roff = a + rowpad + (direct ? 0 : (int)(dA - dmin));
    }
    constexpr int QS = CC ? 1 : (G > 0 ? G : 1);  // shared-memory stride between consecutive staged rows
    const int sread = CC ? p * fa.nrows_max + roff : roff * g + p;

    const int nsteps = MODE == SLB_FUSED_WIN ? nm : nm + P1 - 1;
    const long long smel = fa.ism, scel = fa.isc, osmel = fa.osm, oscel = fa.osc;
    // march index of the outputs emitted at step P1-1, as (block, position in block)
    const int okc = fa.okc;
    int oq, ol;
    {
        int iout = fa.march0 + nm - s0B;
        iout -= iout >= nm ? nm : 0;
        oq = iout / okc;
        ol = iout - oq * okc;
    }
    const long long othr = obase + (long long)ac * oscel;
    // running output pointer: advances by one march row per emitted output; `oleft` outputs remain
    // in the current output block (plain layout: until the periodic wrap)
    double* po = fa.oblk[oq] + othr + (long long)ol * osmel;
    int oleft = okc - ol;
    if (MODE == SLB_FUSED_WIN) {  // no blocks, no wrap: row `lo` of the haloed output array (stores are predicated on the window)
        oq = 0;
        po = fa.oblk[0] + othr + (long long)lo * osmel;
        oleft = 0x7fffffff;
    }
    const bool st_pair = CC && W16 && act1;          // 16-byte store of both outputs
    const bool st_a = act0 && !st_pair, st_b = act1 && !st_pair;
    // halo pushes: element offsets from this rank's output array to the neighbours' (see FusedArgs)
    const long long pdL = (MODE != SLB_FUSED_PLAIN && fa.pushL) ? fa.pushL - fa.oblk[0] : 0;
    const long long pdR = (MODE != SLB_FUSED_PLAIN && fa.pushR) ? fa.pushR - fa.oblk[0] : 0;
    // MODE PSH: per-thread constants -- this thread's passive point lies in the lower or the upper boundary layer
    // of the slab (the host guarantees win_c >= 2 win_h: never both), pd is the offset to that neighbour's halo
    bool psh = false;
    long long pd = 0;
    if (MODE == SLB_FUSED_PSH) {
        const int ps = (int)(fa.push_on_lo ? plo : phi);
        const bool toL = fa.pushL != nullptr && ps < fa.win_h, toR = fa.pushR != nullptr && ps >= fa.win_c - fa.win_h;
        psh = toL || toR;
        pd = toL ? pdL : pdR;
    }
    double winA[P1], winB[P1];           // last order+1 values of T for the two cross outputs
#pragma unroll
    for (int j = 0; j < P1; ++j) winA[j] = winB[j] = 0.0;
    double lsumA = 0.0, lsumB = 0.0;

    int rho_par = 0, rho_k = 0, rho_nprev = 0, rho_rowprev = 0, rho_row = 0;  // MODE RHO bookkeeping (block-uniform)
    if (MODE == SLB_FUSED_RHO) {
        rho_row = fa.march0 + nm - s0B;   // march index of the first emitted output (same for the whole block)
        rho_row -= rho_row >= nm ? nm : 0;
    }
    auto emit = [&](double accA, double accB) {
        if (MODE == SLB_FUSED_RHO) {
            double2 v2;
            v2.x = act0 ? accA : 0.0;
            v2.y = act1 ? accB : 0.0;
            *reinterpret_cast<double2*>(redbuf + ((size_t)((rho_par * 2 + rho_k) * g + p) * ta + a)) = v2;
            ++rho_k;
        }
        if (MODE == SLB_FUSED_WIN) {
            // only the slab's rows are this rank's outputs; those within win_h of a slab boundary also go to the
            // neighbour's halo rows (NVLink peer stores riding inside the pass)
            const bool inw = (unsigned)(lo - fa.win_h) < (unsigned)fa.win_c;
            lsumA += inw ? accA : 0.0;
            lsumB += inw ? accB : 0.0;
            fused_st(st_a && inw, po, accA);
            fused_st(st_b && inw, po + oscel, accB);
            if (fa.pushL) {
                const bool pl = inw && lo < 2 * fa.win_h;
                fused_st(st_a && pl, po + pdL, accA);
                fused_st(st_b && pl, po + pdL + oscel, accB);
            }
            if (fa.pushR) {
                const bool pr = inw && lo >= fa.win_c;
                fused_st(st_a && pr, po + pdR, accA);
                fused_st(st_b && pr, po + pdR + oscel, accB);
            }
            ++lo;
            po += osmel;
            return;
        }
        lsumA += accA;
        lsumB += accB;
        if (CC && W16) {  // neighbours in memory, 16-byte aligned: one store; the odd last column is rare
            fused_st_v2(st_pair, po, accA, accB);
            if (st_a) __stcs(po, accA);
            if (MODE == SLB_FUSED_PSH) {
                fused_st_v2(st_pair && psh, po + pd, accA, accB);
                if (st_a && psh) __stcs(po + pd, accA);
            }
        } else {
            fused_st(st_a, po, accA);
            fused_st(st_b, po + oscel, accB);
            if (MODE == SLB_FUSED_PSH) {
                fused_st(st_a && psh, po + pd, accA);
                fused_st(st_b && psh, po + pd + oscel, accB);
            }
        }
        po += osmel;
        if (--oleft == 0) {  // next output block (plain layout: wrap around the periodic line) -- rare
            oq = (oq + 1) * okc >= nm ? 0 : oq + 1;
            po = fa.oblk[oq] + othr;
            oleft = okc;
        }
    };

    if (!direct) {
        // ---- load slots: what this thread fetches for every march row ---------------------------
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(fsm);
        constexpr int NSLOT = W16 ? 2 : 4;
        const char* gsrc[NSLOT];
        unsigned sdst[NSLOT];
        bool lval[NSLOT];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            const int e = tid + s * NT;
            int pe, j;  // passive point / staged row of the (first) element
            if constexpr (W16) {
                if (CC) {
                    const int h = nrows >> 1;
                    j = 2 * (e % h);
                    pe = e / h;
                    lval[s] = pe < g;
                } else {
                    const int h = g >> 1;
                    pe = 2 * (e % h);
                    j = e / h;
                    lval[s] = j < nrows;
                }
            } else {
                lval[s] = e < nrows * g;
                if (CC) {
                    j = e % nrows;
                    pe = e / nrows;
                } else {
                    pe = e % g;
                    j = e / g;
                }
            }
            const long long Pe = tp * g + pe;
            if (Pe >= np) lval[s] = false;
            const long long Pq = Pe < np ? Pe : np - 1;
            const int rc = (rowbase + j) % nc_;
            gsrc[s] = reinterpret_cast<const char*>(fa.in) +
                      8 * ((long long)(Pq % fa.elo) * fa.islo + (long long)(Pq / fa.elo) * fa.ishi + (long long)rc * scel);
            sdst[s] = sbase + 8u * (unsigned)(CC ? pe * fa.nrows_max + j : j * g + pe);
        }
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) asm volatile("" : "+l"(gsrc[s]));  // keep base + offset folded into one pointer
        const unsigned row_b = 8u * (unsigned)row_elems, ring_b = row_b * R * D;
        // block-uniform bookkeeping of the fetch pipeline (kept in the uniform datapath)
        const int ikc = fa.ikc;
        const long long smb = 8 * smel, blkjump = 8 * (fa.iblk - (long long)ikc * smel);
        int b_iss = fa.march0, k_iss = 0;   // march index / step number of the next row to fetch
        int ileft = ikc - b_iss % ikc;      // rows left in its input block
        long long boff = 8 * ((long long)(b_iss / ikc) * fa.iblk + (long long)(b_iss % ikc) * smel);  // its byte offset
        unsigned off_iss = 0;               // byte offset of that row's slot in the ring
        auto issue_stage = [&]() {
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                if (k_iss < nsteps) {
#pragma unroll
                    for (int s = 0; s < NSLOT; ++s) {
                        if (lval[s]) {
                            if constexpr (W16)
                                fused_cp_async16(sdst[s] + off_iss, gsrc[s] + boff);
                            else
                                fused_cp_async8(sdst[s] + off_iss, gsrc[s] + boff);
                        }
                    }
                    boff += smb;
                    ++b_iss;
                    if (--ileft == 0) {  // input block boundary or periodic wrap -- rare
                        if (b_iss == nm) {
                            b_iss = 0;
                            boff = 0;
                        } else {
                            boff += blkjump;
                        }
                        ileft = ikc;
                    }
                }
                ++k_iss;
                off_iss += row_b;
                off_iss = off_iss == ring_b ? 0u : off_iss;
            }
            fused_cp_commit();
        };
#pragma unroll 1
        for (int st = 0; st < D - 1; ++st) issue_stage();

        const double* sp = fsm + sread;
        const double* const sp_end = sp + (size_t)R * D * row_elems;

        // One block barrier per two march rows (window slots r, r+1).  The terms of the four
        // march-direction dot products that only involve OLD window entries are summed before the
        // barrier (same left-to-right order as slb_dot, hence bit-identical), so that after it only
        // the cross stencils and two or three dependent operations per output remain.
        // MODE RHO: the rows the block emitted in the previous interval are complete in redbuf (the barrier just
        // passed): add the g passive points per cross index in a fixed order and store the partial-plane rows
        auto rho_reduce = [&]() {
            if (MODE == SLB_FUSED_RHO) {
                const int parp = rho_par ^ 1;
                double* const plane = fa.rhopart + (long long)tp * ((long long)nc_ * nm);
                for (int q = tid; q < rho_nprev * ta; q += NT) {
                    const int k = q / ta, x = q - k * ta;
                    if (a0 + x < nc_) {
                        const double* src = redbuf + (size_t)((parp * 2 + k) * g) * ta + x;
                        double sacc = src[0];
                        for (int pp = 1; pp < g; ++pp) sacc += src[(size_t)pp * ta];
                        int row = rho_rowprev + k;
                        row -= row >= nm ? nm : 0;
                        plane[(long long)row * nc_ + a0 + x] = sacc;
                    }
                }
            }
        };
        auto rho_advance = [&]() {
            if (MODE == SLB_FUSED_RHO) {
                rho_nprev = rho_k;
                rho_rowprev = rho_row;
                rho_row += rho_k;
                rho_row -= rho_row >= nm ? nm : 0;
                rho_k = 0;
                rho_par ^= 1;
            }
        };
#define SLB_FUSED_INTERVAL(r, EMIT0, EMIT1)                                                             \
    {                                                                                                   \
        const double pA0 = fused_partial<P1, EXACT, P1 - 1>(winA, w2, (r) + 1);                         \
        const double pB0 = fused_partial<P1, EXACT, P1 - 1>(winB, w2, (r) + 1);                         \
        const double pA1 = fused_partial<P1, EXACT, P1 - 2>(winA, w2, (r) + 2);                         \
        const double pB1 = fused_partial<P1, EXACT, P1 - 2>(winB, w2, (r) + 2);                         \
        fused_cp_wait<D - 2>();                                                                         \
        __syncthreads();                                                                                \
        issue_stage();                                                                                  \
        rho_reduce();                                                                                   \
        double xa[R][P1 + 1];                                                                           \
        _Pragma("unroll") for (int rr = 0; rr < R; ++rr)                                                \
            _Pragma("unroll") for (int q = 0; q <= P1; ++q) xa[rr][q] = sp[rr * row_elems + q * QS];    \
        sp += R * row_elems;                                                                            \
        sp = sp == sp_end ? sp - (size_t)R * D * row_elems : sp;                                        \
        const double TA0 = fused_dot<P1, EXACT, 0>(xa[0], w1);                                          \
        const double TB0 = fused_dot<P1, EXACT, 1>(xa[0], w1);                                          \
        const double TA1 = fused_dot<P1, EXACT, 0>(xa[1], w1);                                          \
        const double TB1 = fused_dot<P1, EXACT, 1>(xa[1], w1);                                          \
        winA[(r)] = TA0;                                                                                \
        winB[(r)] = TB0;                                                                                \
        winA[(r) + 1] = TA1;                                                                            \
        winB[(r) + 1] = TB1;                                                                            \
        if (EMIT0) emit(slb_acc<EXACT>(pA0, TA0, w2[P1 - 1]), slb_acc<EXACT>(pB0, TB0, w2[P1 - 1]));    \
        if (EMIT1)                                                                                      \
            emit(slb_acc<EXACT>(slb_acc<EXACT>(pA1, TA0, w2[P1 - 2]), TA1, w2[P1 - 1]),                 \
                 slb_acc<EXACT>(slb_acc<EXACT>(pB1, TB0, w2[P1 - 2]), TB1, w2[P1 - 1]));                \
        rho_advance();                                                                                  \
    }
        // first block: the windows fill up (nsteps >= P1, so no bounds checks)
#pragma unroll
        for (int r = 0; r < P1; r += R) SLB_FUSED_INTERVAL(r, (r >= P1 - 1), (r + 1 >= P1 - 1))
        int k0 = P1;
        for (; k0 + P1 <= nsteps; k0 += P1) {
#pragma unroll
            for (int r = 0; r < P1; r += R) SLB_FUSED_INTERVAL(r, true, true)
        }
        // tail block
#pragma unroll
        for (int r = 0; r < P1; r += R) {
            if (k0 + r < nsteps) SLB_FUSED_INTERVAL(r, true, (k0 + r + 1 < nsteps))
        }
#undef SLB_FUSED_INTERVAL
        if (MODE == SLB_FUSED_RHO) {  // the last interval's rows
            __syncthreads();
            rho_reduce();
        }
    } else {
        // shifts inside the tile are too far apart to stage a common row range: every thread reads
        // its own stencil inputs from global memory (rare; correct, slower)
        const double* pdir = fa.in + (long long)plo * fa.islo + (long long)phi * fa.ishi;
        int b = fa.march0;
        for (int k0 = 0; k0 < nsteps; k0 += P1) {
#pragma unroll
            for (int r = 0; r < P1; ++r) {
                const int k = k0 + r;
                if (k < nsteps) {
                    const double* src = pdir + (long long)(b / fa.ikc) * fa.iblk + (long long)(b % fa.ikc) * smel;
                    b = (MODE != SLB_FUSED_WIN && b + 1 == nm) ? 0 : b + 1;
                    int rc = ac + s0A;
                    rc -= rc >= nc_ ? nc_ : 0;
                    double xa[P1 + 1];
#pragma unroll
                    for (int q = 0; q <= P1; ++q) {
                        xa[q] = __ldg(src + (long long)rc * scel);
                        rc = rc + 1 == nc_ ? 0 : rc + 1;
                    }
                    winA[r] = fused_dot<P1, EXACT, 0>(xa, w1);
                    winB[r] = fused_dot<P1, EXACT, 1>(xa, w1);
                    if (k >= P1 - 1) emit(slb_dot<P1, EXACT>(winA, w2, (r + 1) % P1), slb_dot<P1, EXACT>(winB, w2, (r + 1) % P1));
                }
            }
        }
    }
    if (fa.linesum) {
        double* ls = fa.linesum + (long long)plo * fa.lslo + (long long)phi * fa.lshi + (long long)ac * fa.lsc;
        if (act0) ls[0] = lsumA;
        if (act1) ls[fa.lsc] = lsumB;
    }
}
#endif  // SLB_PAIR_IMPL
