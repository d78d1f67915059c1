// slb_program.cuh -- step programs: whole time steps of TINY grids as ONE persistent kernel.
//
// The reference's 1D1V example (128 x 256, examples/vlasov-poisson-1d1v.jl:60-64) moves 256 KB per split stage; on a
// B200 a step is pure launch latency: 7 kernels, 66 us stepwise, 35 us as a CUDA graph (slb_capture_*).  A step
// program records the same library calls (slb_sweep, slb_vp_field_solve, slb_reduce_sumsq_async) as a list of ops
// instead of launching them, and one cooperative kernel interprets the list -- `nrep` times over -- with grid
// barriers only where an op reads what an earlier one wrote (hazard analysis on the host).  The grid lives in the L2
// for the whole run.
//
// Every op keeps the ARITHMETIC of the kernel it replaces, operation by operation:
//   OP_SWEEP  dim 0  : k_sweep_contig   (weights by Horner per line, dot product left to right)
//   OP_SWEEP  dim > 0: k_sweep_strided_chunk (same)
//   OP_CHARGE        : k_charge_partial (four accumulators per thread, warps combined in a fixed order)
//   OP_FIELD1D       : k_field_fft's one-space-dim branch (field_fft_1d_warp, shared code)
//   OP_SUMSQ         : k_reduce_partial<1> + k_reduce_final (the same shuffle trees)
// so a program's results are BIT-IDENTICAL to the stepwise calls (tests/test_gpu_driver.py::test_step_program_*).
// Data written inside the kernel is read with plain (coherent) loads, never through the non-coherent path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slb_devfn.cuh"
#include "slb_sweep.cuh"

#define SLB_PROG_THREADS 256
#define SLB_PROG_CH 16         // outputs per line-sum chunk of a strided sweep (== k_sweep_strided_chunk's: the same chunk sums);
                               // a thread computes 8 of them (two threads per chunk), 16 on lines longer than 2048 points
#define SLB_PROG_MAXSUMSQ_BLOCKS 8

enum { SLB_OP_SWEEP = 1, SLB_OP_CHARGE = 2, SLB_OP_FIELD1D = 3, SLB_OP_SUMSQ = 4 };

struct ProgOp {
    int kind;
    int barrier_before;      // a grid barrier separates this op from the ones before it
    // ---- OP_SWEEP: out = sweep(in) along a dim viewed as [inner, n, outer]
    const double* in;
    double* out;
    long long inner, outer;
    int n, P1, nc;
    const double* coef;      // (order + 1) x nc weight polynomials, [j * nc + k]
    double* linesum;         // strided sweeps: optional per-line sums of the outputs (slb_grid_set_linesum)
    AlphaMap am;
    int tab_local;           // the shift table is the E field of the last OP_FIELD1D: read this block's shared-memory copy
    // ---- OP_CHARGE: partial[c][a] = sum over chunk c of f[a + ns * b]   (k_charge_partial's decomposition)
    const double* f;
    long long ns, nv, chunk;
    int nchunk;
    double* partial;
    // ---- OP_FIELD1D
    FieldFftArgs ffa;
    // ---- OP_SUMSQ: out[slot + rep * stride] = scale * sum(x .^ 2)
    const double* x;
    int x_local;             // x is the E field of the last OP_FIELD1D: reduce this block's shared-memory copy
    long long nx;
    double scale;
    double* outp;
};

int slb_program_run(const ProgOp* ops_dev, int nops, int nrep, long long out_stride, int nmax_field, unsigned* bar_ctr, unsigned bar_base,
                    int use_cg, unsigned long long* prof, int nblocks, size_t smem_bytes, cudaStream_t stream);
size_t slb_program_smem_bytes(int nops, int nmax_field);
bool slb_program_supports_p1(int P1);

#ifdef SLB_PROGRAM_IMPL
#include <cooperative_groups.h>

// Loads of data that an earlier op of the same launch may have written (f, E, partial sums) are ordinary coherent
// loads -- never __ldg / const __restrict__ (the non-coherent path) -- and are ordered by the grid barrier.
__device__ __forceinline__ double prog_ld(const double* p) { return *p; }

template <int P1, int CH>
__device__ __forceinline__ void prog_sweep(const ProgOp& op, const double* tab, const double* scoef, double* lsb)
{
    const int n = op.n, nc = op.nc;
    const int gthreads = gridDim.x * SLB_PROG_THREADS;
    const int gtid = blockIdx.x * SLB_PROG_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    if (op.inner == 1) {
        // ---- contiguous lines: one warp per line, a lane computes R = 4 consecutive outputs of a 128-output segment from
        // R + order inputs that are all loaded before the first store (k_sweep_contig's register blocking); the weights
        // are evaluated by P1 lanes and broadcast
        constexpr int R = 4, NX = R + P1 - 1;
        const long long nlines = op.outer;
        const double* const gin = op.in;
        double* const gout = op.out;
        const int gw = gtid >> 5, nw = gthreads >> 5;
        for (long long ln = gw; ln < nlines; ln += nw) {
            const double alpha = op.am.scale * prog_ld(tab + slb_alpha_off(op.am, 0u, (unsigned)ln));
            double t;
            int s0;
            slb_split(alpha, n, (P1 - 1) / 2, t, s0);
            const int jl = lane < P1 ? lane : 0;
            double wl = scoef[(nc - 1) * P1 + jl];
            for (int k = nc - 2; k >= 0; --k) wl = fma(t, wl, scoef[k * P1 + jl]);
            double w[P1];
#pragma unroll
            for (int j = 0; j < P1; ++j) w[j] = __shfl_sync(0xffffffffu, wl, j);
            const double* lin = gin + ln * n;
            double* lout = gout + ln * n;
            for (int seg = 0; seg < n; seg += 32 * R) {
                const int i0 = seg + R * lane;
                if (i0 < n) {
                    int g = s0 + i0;
                    g -= g >= n ? n : 0;
                    double x[NX];
#pragma unroll
                    for (int j = 0; j < NX; ++j) {
                        x[j] = prog_ld(lin + g);
                        g = g + 1 == n ? 0 : g + 1;
                    }
                    double o[R];
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        double acc = x[m] * w[0];
#pragma unroll
                        for (int j = 1; j < P1; ++j) acc = fma(x[m + j], w[j], acc);
                        o[m] = acc;
                    }
#pragma unroll
                    for (int m = 0; m < R; ++m)
                        if (i0 + m < n) lout[i0 + m] = o[m];
                }
            }
        }
    } else {
        // ---- strided lines: virtual block = LB neighbouring lines x nch chunks of CH outputs (k_sweep_strided_chunk's
        // decomposition, CH = 16), thread = (line, chunk), lanes along the contiguous inner index.  The weights of a line
        // are evaluated ONCE (chunk thread j evaluates weight j, same Horner sequence) and shared through shared memory;
        // the line sums are combined over the chunks in the same fixed order as the stand-alone kernel.
        const long long nlines = op.inner * op.outer;
        const int nch = (n + CH - 1) / CH;                 // <= 256 (host check)
        int LB = 1;
        while (2 * LB * nch <= SLB_PROG_THREADS && LB < 32) LB *= 2;
        while (LB > 4 && (nlines + LB - 1) / LB < (long long)gridDim.x) LB >>= 1;   // few lines: spread them over the blocks
        const int li = threadIdx.x % LB, ch = threadIdx.x / LB;
        const long long nvb = (nlines + LB - 1) / LB;
        double* wsm = lsb + SLB_PROG_THREADS;              // [P1][LB]
        for (long long vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
            const long long gid = vb * LB + li;
            const bool active = ch < nch && gid < nlines;
            double t = 0.0;
            double o[CH];
            int s0 = 0, ocnt = 0;
#pragma unroll
            for (int j = 0; j < CH; ++j) o[j] = 0.0;
            long long a = 0, b = 0;
            if (active) {
                b = gid / op.inner;
                a = gid - b * op.inner;
                const double alpha = op.am.scale * prog_ld(tab + slb_alpha_off(op.am, (unsigned)a, (unsigned)b));
                slb_split(alpha, n, (P1 - 1) / 2, t, s0);
                for (int j = ch; j < P1; j += nch) {
                    double wl = scoef[(nc - 1) * P1 + j];
                    for (int k = nc - 2; k >= 0; --k) wl = fma(t, wl, scoef[k * P1 + j]);
                    wsm[j * LB + li] = wl;
                }
            }
            __syncthreads();
            if (active) {
                double w[P1];
#pragma unroll
                for (int j = 0; j < P1; ++j) w[j] = wsm[j * LB + li];
                const int i0 = ch * CH;
                const int cnt = n - i0 < CH ? n - i0 : CH;
                const double* pin = op.in + (b * n) * op.inner + a;
                double x[CH + P1 - 1];
                int kk = (s0 + i0) % n;
#pragma unroll
                for (int j = 0; j < CH + P1 - 1; ++j) {
                    x[j] = (j < cnt + P1 - 1) ? prog_ld(pin + (long long)kk * op.inner) : 0.0;
                    kk = kk + 1 == n ? 0 : kk + 1;
                }
                double* po = op.out + (b * n) * op.inner + a + (long long)i0 * op.inner;
                ocnt = cnt;
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    double acc = x[j] * w[0];
#pragma unroll
                    for (int q = 1; q < P1; ++q) acc = fma(x[j + q], w[q], acc);
                    o[j] = acc;
                    if (j < cnt) po[(long long)j * op.inner] = acc;
                }
            }
            if (op.linesum) {  // block-uniform
                // the stand-alone kernel sums the outputs of a chunk of 16 one after the other, then the chunks of a line one
                // after the other; with SUB threads per chunk the first adds its outputs starting from 0 and hands the sum on
                constexpr int SUB = SLB_PROG_CH / CH;
                const bool first = (ch % SUB) == 0;
                double part = 0.0;
                if (first) {
#pragma unroll
                    for (int j = 0; j < CH; ++j)
                        if (j < ocnt) part += o[j];
                }
                if (SUB > 1) {
                    lsb[threadIdx.x] = part;   // [ch][li]
                    __syncthreads();           // (everybody has its weights in registers: wsm is free from here on)
                    if (!first && ch < nch) {
                        part = lsb[(ch - 1) * LB + li];
#pragma unroll
                        for (int j = 0; j < CH; ++j)
                            if (j < ocnt) part += o[j];
                    }
                }
                const bool last = (SUB == 1) || !first || ch + 1 >= nch;   // this thread holds the complete sum of its chunk
                double* csum = SUB > 1 ? wsm : lsb;                        // [chunk][li]
                if (last && ch < nch) csum[(ch / SUB) * LB + li] = part;
                __syncthreads();
                if (ch == 0 && gid < nlines) {
                    const int nch16 = (nch + SUB - 1) / SUB;
                    double sacc = csum[li];
                    for (int q = 1; q < nch16; ++q) sacc += csum[q * LB + li];
                    op.linesum[gid] = sacc;
                }
            }
            __syncthreads();   // wsm / lsb are free again
        }
    }
}

// Grid barrier on a monotonically increasing counter: thread 0 of every block arrives with a release (the block's
// earlier writes, ordered before it by the block barrier, become visible at GPU scope) and polls with acquire loads
// until all blocks of this barrier generation have arrived (`target`, compared modulo 2^32).  The blocks are
// co-resident (cooperative launch), so the spin terminates.  Measured equal to cooperative_groups' grid.sync() here
// (C1 step 24.8 vs 25.1 us): the cost of a barrier is the release / acquire round trip either way.
__device__ __forceinline__ void prog_barrier(unsigned* ctr, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while ((int)(v - target) < 0);
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned long long prog_timer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// dynamic shared memory: [ops: nops ProgOp][tws: nf double2][xa: nf][xb: nf][E copy: nf doubles][multiplier: nf doubles]
__global__ void __launch_bounds__(SLB_PROG_THREADS) k_program(const ProgOp* __restrict__ ops, int nops, int nrep, long long out_stride,
                                                              int nf, unsigned* bar_ctr, unsigned bar_base, int use_cg,
                                                              unsigned long long* prof)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(16) unsigned char psm_raw[];
    __shared__ double scoef[SLB_NCMAX * SLB_P1MAX];
    __shared__ double csm[8][33];
    __shared__ double redsm[32];
    __shared__ double lsb[SLB_PROG_THREADS + SLB_P1MAX * 32];   // line-sum exchange, then the shared weights [P1][LB]
    __shared__ double partsm[SLB_PROG_MAXSUMSQ_BLOCKS];
    ProgOp* sops = reinterpret_cast<ProgOp*>(psm_raw);
    double2* tws = reinterpret_cast<double2*>(psm_raw + (((size_t)nops * sizeof(ProgOp) + 15) & ~(size_t)15));
    double2* xa = tws + nf;
    double2* xb = xa + nf;
    double* esm = reinterpret_cast<double*>(xb + nf);   // this block's copy of the field of the last OP_FIELD1D
    double* msm = esm + nf;                             // the field solve's multiplier in bit-reversed order
    const int tid = threadIdx.x;
    {   // all descriptors once: no global reads of them inside the loop
        const int* src = reinterpret_cast<const int*>(ops);
        int* dst = reinterpret_cast<int*>(sops);
        for (int i = tid; i < nops * (int)(sizeof(ProgOp) / sizeof(int)); i += SLB_PROG_THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const double* coef_loaded = nullptr;   // block-uniform caches
    const double2* tw_loaded = nullptr;
    const double* mult_loaded = nullptr;
    unsigned bar_target = bar_base;
    for (int rep = 0; rep < nrep; ++rep) {
        for (int k = 0; k < nops; ++k) {
            const ProgOp& op = sops[k];
            // optional time stamps of block 0 (slb_program_profile): before the barrier, after it, after the op
            const bool stamp = prof != nullptr && rep == nrep - 1 && blockIdx.x == 0 && tid == 0;
            if (stamp) prof[3 * k] = prog_timer();
            if (op.barrier_before) {
                if (use_cg) {
                    cg::this_grid().sync();
                } else {
                    bar_target += gridDim.x;
                    prog_barrier(bar_ctr, bar_target);
                }
            }
            if (stamp) prof[3 * k + 1] = prog_timer();
            switch (op.kind) {
            case SLB_OP_SWEEP: {
                if (op.coef != coef_loaded) {
                    __syncthreads();  // nobody still evaluates weights from the old table
                    for (int i = tid; i < op.nc * op.P1; i += SLB_PROG_THREADS) scoef[i] = __ldg(op.coef + (i % op.P1) * op.nc + i / op.P1);
                    __syncthreads();
                    coef_loaded = op.coef;
                }
                const double* tab = op.tab_local ? esm : op.am.tab;
#define SLB_PROG_SWEEP(P)                                                         \
    case P:                                                                       \
        if (op.n > 2048)   /* chunks of 8 need n / 8 <= 256 threads */             \
            prog_sweep<P, SLB_PROG_CH>(op, tab, scoef, lsb);                      \
        else                                                                      \
            prog_sweep<P, SLB_PROG_CH / 2>(op, tab, scoef, lsb);                  \
        break;
                switch (op.P1) {
                    SLB_PROG_SWEEP(4)
                    SLB_PROG_SWEEP(6)
                    SLB_PROG_SWEEP(8)
                    SLB_PROG_SWEEP(10)
                    SLB_PROG_SWEEP(12)
                }
#undef SLB_PROG_SWEEP
                break;
            }
            case SLB_OP_CHARGE: {
                // virtual blocks (x tile of 32 space points, velocity chunk) of k_charge_partial, 32 x 8 threads each
                const int tx = tid & 31, ty = tid >> 5;
                const long long xt = (op.ns + 31) / 32;
                const long long nvb = xt * op.nchunk;
                for (long long vb = blockIdx.x; vb < nvb; vb += gridDim.x) {
                    const long long bx = vb % xt, by = vb / xt;
                    const long long a = bx * 32 + tx;
                    const long long b0 = by * op.chunk;
                    const long long b1 = b0 + op.chunk < op.nv ? b0 + op.chunk : op.nv;
                    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                    if (a < op.ns) {
                        const double* p = op.f + a;
                        long long b = b0 + ty;
                        for (; b + 24 < b1; b += 32) {
                            acc0 += prog_ld(p + op.ns * b);
                            acc1 += prog_ld(p + op.ns * (b + 8));
                            acc2 += prog_ld(p + op.ns * (b + 16));
                            acc3 += prog_ld(p + op.ns * (b + 24));
                        }
                        for (; b < b1; b += 8) acc0 += prog_ld(p + op.ns * b);
                    }
                    __syncthreads();  // the previous virtual block's sums have been read
                    csm[ty][tx] = (acc0 + acc1) + (acc2 + acc3);
                    __syncthreads();
                    if (ty == 0 && a < op.ns) {
                        double s = csm[0][tx];
#pragma unroll
                        for (int i = 1; i < 8; ++i) s += csm[i][tx];
                        op.partial[by * op.ns + a] = s;
                    }
                }
                break;
            }
            case SLB_OP_FIELD1D: {
                // EVERY block solves the (tiny) field problem itself, into its own shared-memory copy of E: the sweeps
                // that read E as their shift table (tab_local) then need no grid barrier after the solve.  Block 0 also
                // stores rho and E to global memory (compute_ee, the caller's E_dev).
                const int n1 = op.ffa.n1;
                __syncthreads();  // earlier sweeps of this block are done with the old E copy
                if (tid < 32) {
                    if (op.ffa.tw1 != tw_loaded) field_fill_tws(tws, op.ffa.tw1, n1, tid, 32);
                    if (op.ffa.mult[0] != mult_loaded)
                        for (int p = tid; p < n1; p += 32) msm[p] = __ldg(op.ffa.mult[0] + field_brev(p, op.ffa.l1));
                    __syncwarp();
                    const FieldFftArgs fa = op.ffa;   // registers: stores to global memory cannot invalidate it
                    field_fft_1d_warp(fa, tws, xa, xb, tid, esm, blockIdx.x == 0, msm);
                }
                tw_loaded = op.ffa.tw1;
                mult_loaded = op.ffa.mult[0];
                __syncthreads();
                break;
            }
            case SLB_OP_SUMSQ: {
                // x_local: x is the field of the last OP_FIELD1D, of which every block holds a copy -- the LAST block reduces
                // its copy (it is the one most likely to be idle in the sweeps around); else block 0 reads global memory
                const double* xs = op.x_local ? esm : op.x;
                if (blockIdx.x == (op.x_local ? gridDim.x - 1 : 0u)) {
                    // k_reduce_partial<1> with nb blocks of 256 threads, then k_reduce_final
                    long long nb = (op.nx + 256 * 8 - 1) / (256 * 8);
                    nb = nb < 1 ? 1 : nb;
                    __syncthreads();
                    for (int vb = 0; vb < (int)nb; ++vb) {
                        double acc = 0.0;
                        for (long long i = (long long)vb * 256 + tid; i < op.nx; i += nb * 256) {
                            const double v = prog_ld(xs + i);
                            acc += v * v;
                        }
                        const double r = slb_block_reduce(acc, redsm);
                        if (tid == 0) partsm[vb] = r;
                        __syncthreads();
                    }
                    double acc = 0.0;
                    for (int i = tid; i < (int)nb; i += 256) acc += partsm[i];
                    const double r = slb_block_reduce(acc, redsm);
                    if (tid == 0) op.outp[(long long)rep * out_stride] = op.scale * r;
                }
                break;
            }
            }
            if (stamp) prof[3 * k + 2] = prog_timer();
        }
    }
}

bool slb_program_supports_p1(int P1) { return P1 == 4 || P1 == 6 || P1 == 8 || P1 == 10 || P1 == 12; }

size_t slb_program_smem_bytes(int nops, int nmax_field)
{
    const size_t nf = (size_t)(nmax_field > 0 ? nmax_field : 1);
    return (((size_t)nops * sizeof(ProgOp) + 15) & ~(size_t)15) + 3 * nf * sizeof(double2) + 2 * nf * sizeof(double);
}

int slb_program_run(const ProgOp* ops_dev, int nops, int nrep, long long out_stride, int nmax_field, unsigned* bar_ctr, unsigned bar_base,
                    int use_cg, unsigned long long* prof, int nblocks, size_t smem_bytes, cudaStream_t stream)
{
    int nf = nmax_field > 0 ? nmax_field : 1;
    if (smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_program, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return (int)e;
    }
    void* args[] = {(void*)&ops_dev, (void*)&nops, (void*)&nrep, (void*)&out_stride, (void*)&nf, (void*)&bar_ctr, (void*)&bar_base, (void*)&use_cg, (void*)&prof};
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_program, dim3((unsigned)nblocks), dim3(SLB_PROG_THREADS), args, smem_bytes, stream);
    return (int)e;
}
#endif  // SLB_PROGRAM_IMPL
