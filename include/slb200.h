/*
 * slb200.h -- C ABI of libslb200.so, the B200 (sm_100a) implementation of the
 * SemiLagrangian.jl hot path: the 1-D interpolation sweep inside the split advection!
 * driver, the charge-density reduction and the Fourier Poisson solve that feed it.
 *
 * The reference (JuliaVlasov/SemiLagrangian.jl v0.1.2) is pure Julia and has no FFI; these
 * are the entry points its Julia host layer binds with `ccall` (see INTEGRATION.md and
 * semilagrangian.jl_b200/julia/SemiLagrangianB200.jl).  Each entry point cites the
 * reference interface (file:line under the reference's root) it replaces.
 *
 * Conventions
 *   - every function returns int: 0 = ok, negative = error (SLB_E_*); the message is
 *     available from slb_last_error().  Nothing throws across the ABI.
 *   - arrays are COLUMN-MAJOR, dim 0 fastest (Julia's layout), extents are int64_t.
 *   - opaque handles own their device memory; host pointers are borrowed for the call.
 *   - one host thread per context; calls enqueue on the context's stream and return;
 *     slb_sync() waits.  Functions that return host scalars synchronise themselves.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     SLB_E_CUDA.
 */
#ifndef SLB200_H
#define SLB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLB_OK 0
#define SLB_E_ARG (-1)     /* invalid argument (ArgumentError / DomainError in the reference) */
#define SLB_E_CUDA (-2)    /* CUDA runtime error, or no device */
#define SLB_E_ALLOC (-3)   /* out of memory */
#define SLB_E_UNSUPPORTED (-4)

/* interpolation kinds: AbstractInterpolation subtypes, src/interpolation.jl:18 */
#define SLB_LAGRANGE 0     /* src/lagrange.jl:58-72    */
#define SLB_BSPLINE_LU 1   /* src/bsplinelu.jl:253-270 */
#define SLB_BSPLINE_FFT 2  /* src/bsplinefft.jl:25-45  */
#define SLB_HERMITE 3      /* src/hermite.jl:99-132    */

/* slb_sweep flags */
#define SLB_SWEEP_EXACT 1  /* stencil as rounded products summed left to right (bitwise the
                              reference's `sum(res[...] .* precal)`, src/interpolation.jl:190)
                              instead of an FMA chain */

#define SLB_SWEEP_INSIDE_EDGE 2  /* non-periodic InsideEdge interpolation (src/interpolation.jl:123-132, :250-286):
                                    one-sided stencils near the ends of the line instead of the periodic wrap;
                                    Lagrange / Hermite kinds, -order <= floor(alpha) - order/2 <= 0 required
                                    (SLB_E_ARG for host tables, NaN-filled lines for device tables) */

#define SLB_MAX_DIMS 6
#define SLB_MAX_ORDER 63

typedef struct slb_ctx slb_ctx;
typedef struct slb_grid slb_grid;
typedef struct slb_interp slb_interp;
typedef struct slb_poisson slb_poisson;

/* ---- context ------------------------------------------------------------------------ */
/* One context per process and GPU.  `stream` is a cudaStream_t to adopt (e.g. torch's
 * current stream) or NULL to create a private non-blocking stream.
 * Replaces: the `timeopt` back-end selection of Advection (src/advection.jl:2,95). */
int slb_ctx_create(int device_id, void* stream, slb_ctx** out);
void slb_ctx_destroy(slb_ctx* ctx);
const char* slb_last_error(void);
int slb_sync(slb_ctx* ctx);
int slb_device_count(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t slb_launch_count(const slb_ctx* ctx);
/* CUDA-event timer on the context's stream */
int slb_timer_start(slb_ctx* ctx);
int slb_timer_stop(slb_ctx* ctx, float* elapsed_ms); /* records, synchronises, returns ms */

/* ---- whole time steps as one launch (CUDA graph) ----------------------------------------------------------
 * Small grids (the 1D1V example, 128 x 256: 256 KB, examples/vlasov-poisson-1d1v.jl:60-64) are bound by launch
 * latency, not by HBM.  Between slb_capture_begin and slb_capture_end the calls of this library on the context's
 * stream are RECORDED instead of executed; slb_graph_launch replays them as one launch.  The recorded sequence must
 * consist of kernel launches on device-resident tables (no host alpha tables, no allocation, no call that returns a
 * host scalar), and it must leave every grid's front/back roles as it found them (an even number of sweeps per grid),
 * so that a replay continues where the last one ended.  slb_launch_count keeps counting the replayed kernels. */
typedef struct slb_graph slb_graph;
int slb_capture_begin(slb_ctx* ctx);
int slb_capture_end(slb_ctx* ctx, slb_graph** out);
int slb_graph_launch(slb_graph* g);
void slb_graph_destroy(slb_graph* g);

/* ---- whole time steps as ONE persistent kernel (step programs) ------------------------------------------------
 * A CUDA graph still pays one kernel launch per split stage (the 1D1V example: 7 nodes, 35 us per step for 256 KB of
 * data).  Between slb_program_begin and slb_program_end the calls slb_sweep (plain Lagrange / Hermite kinds of odd
 * order 3..11, device-resident shift tables, no flags), slb_vp_field_solve (one power-of-two space dim) and
 * slb_reduce_sumsq_async are recorded as a list of ops; NOTHING executes and any other compute call fails with
 * SLB_E_UNSUPPORTED (the recording is then unusable: slb_program_end reports it and the caller falls back to
 * slb_capture_* or to stepwise calls).  slb_program_launch runs the list `nrep` times inside one cooperative kernel
 * with grid barriers only between dependent ops; repetition r stores the result of every recorded
 * slb_reduce_sumsq_async at out_dev + r * out_stride (doubles).  Results are bit-identical to the stepwise calls.
 * Like a captured graph, the recorded sequence must leave every grid's front/back roles as it found them.
 * A program keeps the device pointers it recorded (grid buffers, shift tables, interpolation objects, E / rho, the
 * energy slots): destroy it before any of those objects.  One launch at a time per program (launches on the context's
 * stream are ordered anyway).
 * Replaces: the time loop of examples/vlasov-poisson-1d1v.jl:60-64 around advection! (src/advection.jl:594-704). */
typedef struct slb_program slb_program;
int slb_program_begin(slb_ctx* ctx);
int slb_program_end(slb_ctx* ctx, slb_program** out);
int slb_program_launch(slb_program* p, int nrep, int64_t out_stride);
int slb_program_info(const slb_program* p, int* nops, int* nbarriers, int* nblocks);
/* like slb_program_launch (the data advances), with time stamps taken by block 0 in the last repetition: per recorded op
 * its kind (1 sweep, 2 charge partial sums, 3 field solve, 4 sum of squares), the ns block 0 waited at the barrier
 * before it and the ns the op took in block 0; cap >= number of ops.  Synchronises. */
int slb_program_profile(slb_program* p, int nrep, int64_t out_stride, int cap, int* kind_out, double* wait_ns_out, double* run_ns_out);
void slb_program_destroy(slb_program* p);

/* extra CUDA events on the context's stream (per-kernel timing inside bench.py) */
int slb_event_create(slb_ctx* ctx, void** ev_out);
int slb_event_record(slb_ctx* ctx, void* ev);
int slb_event_elapsed_ms(void* ev_start, void* ev_stop, float* elapsed_ms); /* waits for ev_stop */
/* everything enqueued on ctx's stream after this call waits for `ev` (recorded on ANOTHER context's stream of the
 * same device): lets a host layer run uploads and read-backs on their own streams next to the sweeps */
int slb_stream_wait_event(slb_ctx* ctx, void* ev);
int slb_event_destroy(void* ev);

/* raw device / pinned-host buffers (E fields, rho, shift tables) */
int slb_malloc(slb_ctx* ctx, int64_t bytes, void** dev_out);
int slb_free(slb_ctx* ctx, void* dev);
int slb_host_alloc(int64_t bytes, void** host_out); /* pinned */
int slb_host_free(void* host);
int slb_memcpy_h2d(slb_ctx* ctx, void* dev, const void* host, int64_t bytes); /* async on the stream */
int slb_memcpy_d2h(slb_ctx* ctx, void* host, const void* dev, int64_t bytes); /* async on the stream */

/* ---- grid: the device-resident distribution function -------------------------------- */
/* Replaces AdvectionData.data + bufdata (src/advection.jl:229-267): `f` and an equally
 * sized scratch array.  Sweeps are out-of-place front -> back, then the roles swap, so f
 * never leaves HBM between split stages and no permutedims! is ever needed
 * (src/advection.jl:372-386). */
int slb_grid_create(slb_ctx* ctx, int ndims, const int64_t* extents, slb_grid** out);
/* same, over caller-owned device buffers (e.g. torch tensors); not freed on destroy */
int slb_grid_create_external(slb_ctx* ctx, int ndims, const int64_t* extents, double* dev_front,
                             double* dev_back, slb_grid** out);
void slb_grid_destroy(slb_grid* g);
int slb_grid_upload(slb_grid* g, const double* host);         /* AdvectionData ctor copy, src/advection.jl:264-267 */
int slb_grid_download(const slb_grid* g, double* host);       /* getdata, src/advection.jl:317 */
double* slb_grid_front(const slb_grid* g);                     /* current f (device pointer) */
double* slb_grid_back(const slb_grid* g);                      /* scratch (device pointer)   */
int slb_grid_swap(slb_grid* g);

/* ---- interpolation object ------------------------------------------------------------ */
/* coef: (order+1) rows x ncoef ascending Float64 coefficients, row-major: row j is
 * `tabfct[j+1]` (src/interpolation.jl:18, :96-98) already rounded to Float64 by the host
 * (which keeps the reference's exact-rational constructors).  node_vals: B(1..order) for
 * the B-spline kinds (src/bsplinelu.jl:264, src/bsplinefft.jl:35), else NULL.  n: line
 * length the object is bound to (B-splines; ignored otherwise).
 * Errors mirror the reference: even order for SLB_BSPLINE_LU (src/bsplinelu.jl:257-261),
 * n not a power of two for SLB_BSPLINE_FFT (src/fftbig.jl:57). */
int slb_interp_create(slb_ctx* ctx, int kind, int order, int64_t n, const double* coef, int ncoef,
                      const double* node_vals, slb_interp** out);
void slb_interp_destroy(slb_interp* it);

/* ---- the sweep: one advection! call of a const-shift 1-D state ------------------------ */
/* Replaces advection! (src/advection.jl:594-657) = getformdata + per-line
 * getprecal/interpolate! (src/interpolation.jl:175-193, :381-396) + copydata!.
 * For every line along `dim` with other-dim indices idx[]:
 *     alpha = alpha_scale * alpha_tab[ sum_d idx[d] * alpha_strides[d] ]      (grid units)
 *     decint = floor(alpha); w_j = tabfct[j](alpha - decint)
 *     c = sol(interp, line)   (identity | cyclic banded LU | FFT-diagonal)
 *     out[i] = sum_j c[(i + decint - order/2 + j) mod n] * w_j
 * alpha_strides[dim] is ignored; 0 = broadcast.  alpha_tab is a device pointer when
 * alpha_on_device != 0 (E fields, mesh points kept on the GPU), else a host array of
 * alpha_len doubles that is copied first.  The (alpha_scale, table) split is exactly how
 * the plugins build bufcur: (dt/step) * E  (src/poisson.jl:178-189),
 * (-dt/step) * v  (:191-203), (sign*dt/step) * points (src/rotation.jl:21-31). */
int slb_sweep(slb_grid* g, int dim, const slb_interp* it, const double* alpha_tab, int64_t alpha_len,
              const int64_t* alpha_strides, double alpha_scale, int alpha_on_device, int flags);

/* ---- two consecutive sweeps in one pass over HBM ---------------------------------------------- */
/* Equivalent to slb_sweep(dimA, ...) followed by slb_sweep(dimB, ...) -- two advection! calls of a
 * split step (src/advection.jl:594-657; e.g. the v1 v2 and x1 x2 pairs of
 * examples/vlasov-poisson-2d2v.jl:126-131) -- with BIT-IDENTICAL results, but f is read from and
 * written to HBM once: sweep A is a cross-thread stencil on rows staged in shared memory, sweep B
 * runs along the march direction in registers (csrc/slb_pair.cuh).  Both alpha tables follow
 * slb_sweep's convention and are evaluated before either sweep, like bufcur in
 * src/poisson.jl:178-203.  A line-sum buffer set with slb_grid_set_linesum receives the sums of
 * sweep B's outputs.
 * Returns SLB_E_UNSUPPORTED for combinations that are not pair-fused (dimB == 0, alpha_A depending
 * on dimB or alpha_B on dimA, B-spline pre-solves, different or even orders, more than 4 dims):
 * callers then issue the two sweeps separately. */
int slb_sweep_pair(slb_grid* g, int dimA, const slb_interp* itA, const double* alphaA_tab, int64_t alphaA_len,
                   const int64_t* alphaA_strides, double alphaA_scale, int dimB, const slb_interp* itB,
                   const double* alphaB_tab, int64_t alphaB_len, const int64_t* alphaB_strides, double alphaB_scale,
                   int alpha_on_device, int flags);

/* slb_sweep_pair with the multi-GPU re-shard fused in (replaces mpibroadcast,
 * src/mpiinterface.jl:17-38).  Both sides may be BLOCK-MAJOR along dimB: nblocks consecutive
 * sub-arrays, each with extent[dimB] / nblocks along dimB.
 *   in_nblocks  > 1: the front buffer holds the blocks an all-to-all delivered (block r from rank r);
 *   out_nblocks > 1: block q of the result is stored at out_block_bases[q] -- a pointer into the
 *                    buffer of the rank that owns block q after the exchange (this GPU's HBM or a
 *                    peer's, mapped with slb_ipc_open_handle): the stores travel over NVLink inside
 *                    the sweep and no separate collective or pack pass exists.  With
 *                    out_block_bases == NULL the blocks go to this grid's back buffer (for an NCCL
 *                    all-to-all) and front/back swap; otherwise the grid's roles do not change and
 *                    callers synchronise the ranks before reading the result.
 *   first_block    : the march along dimB is periodic and may start at any output block; rank r of P
 *                    passes (r + 1) % P so that at any time the P ranks store into P different
 *                    destinations (no incast on one GPU's NVLink ingress). */
int slb_sweep_pair_ex(slb_grid* g, int dimA, const slb_interp* itA, const double* alphaA_tab, int64_t alphaA_len,
                      const int64_t* alphaA_strides, double alphaA_scale, int dimB, const slb_interp* itB,
                      const double* alphaB_tab, int64_t alphaB_len, const int64_t* alphaB_strides, double alphaB_scale,
                      int alpha_on_device, int flags, int in_nblocks, int out_nblocks, double* const* out_block_bases,
                      int first_block);

/* ---- pair passes on a HALO-SHARDED grid (replaces the MPI mode's replicate + broadcast of
 * every sweep, src/mpiinterface.jl:17-38, src/advection.jl:116-123) -----------------------------------------
 * A 2D2V grid f[x1,x2,v1,v2] is split over P ranks in slabs of c = n4 / P points along ONE dim (v2) for the whole
 * step; every rank stores [low halo | slab | high halo] = c + 2 halo planes along that dim.  No transposes:
 *   - a pass that does not sweep the sharded dim (x1 x2) runs on the slab view (SLB_HALO_PASSIVE);
 *   - a pass whose SECOND sweep runs along the sharded dim (v1 v2) marches once over the c + 2 halo rows and
 *     emits the slab's c rows (SLB_HALO_MARCH): shifts need floor(alpha) + order/2 + 1 <= halo and
 *     order/2 - floor(alpha) <= halo, else bit 0 of the error word is set (slb_halo_error) -- for
 *     Vlasov-Poisson velocity sweeps |alpha| = dt/dv |E| is about one cell.
 * Either pass can also store the outputs that lie within `halo` of a slab boundary into the neighbours' halo
 * planes (push_lo / push_hi: the address, in the rank below / above, of the array that corresponds to this
 * grid's back buffer; NULL = that side is not pushed by this pass): the halo exchange rides inside the pass as NVLink
 * peer stores, there is no separate collective.  Callers order the ranks with slb_comm_* before the halos are read.
 * (A pass whose NVLink time exceeds its HBM time can push one side itself and leave the other to a copy on a second
 * stream that overlaps the field solve: slb_memcpy_d2d + slb_comm_signal / slb_comm_wait.)
 * Results are bit-identical to slb_sweep_pair on the unsharded grid.  Lagrange / Hermite kinds (B-spline
 * pre-solves couple the whole line: they use the transposing driver, slb_sweep_pair_ex / slb_sweep_peer). */
#define SLB_HALO_MARCH 1
#define SLB_HALO_PASSIVE 2
typedef struct slb_halo {
    int mode;        /* SLB_HALO_MARCH: the grid's extent along dimB is c + 2 halo | SLB_HALO_PASSIVE: the grid is the slab */
    int halo;        /* halo planes on either side */
    int shard_dim;   /* SLB_HALO_PASSIVE: the sharded grid dim (not dimA, not dimB) */
    double* push_lo; /* neighbour arrays (device pointers, possibly peer memory), or NULL */
    double* push_hi;
    int* err_flag;   /* device word for the shift-range check, or NULL: the context's (slb_halo_error) */
} slb_halo;
int slb_sweep_pair_halo(slb_grid* g, int dimA, const slb_interp* itA, const double* alphaA_tab, int64_t alphaA_len,
                        const int64_t* alphaA_strides, double alphaA_scale, int dimB, const slb_interp* itB,
                        const double* alphaB_tab, int64_t alphaB_len, const int64_t* alphaB_strides, double alphaB_scale,
                        int alpha_on_device, int flags, const slb_halo* halo);
/* reads and clears the context's error word (synchronises): bit 0 = a shift exceeded the halo */
int slb_halo_error(slb_ctx* ctx, int* flags_out);

/* ---- rank-to-rank plumbing without MPI / NCCL in the data path ------------------------------
 * One process per GPU.  Every rank owns a small device "mailbox" (flags + nslot_doubles * nranks doubles) that
 * its peers map (CUDA IPC across processes, plain peer access inside one process).  The host language only
 * moves the opaque handles once at start-up (MPI.Allgather in Julia, any all-gather elsewhere); afterwards
 * every exchange is a kernel on the context's stream: peer stores + a flag, no host synchronisation.
 * Replaces MPI.Bcast / mpibroadcast (src/mpiinterface.jl:17-38) and the MPIOpt fields of Advection
 * (src/advection.jl:116-123). */
#define SLB_COMM_HANDLE_BYTES 128
typedef struct slb_comm slb_comm;
int slb_comm_create(slb_ctx* ctx, int rank, int nranks, int64_t nslot_doubles, slb_comm** out);
void slb_comm_destroy(slb_comm* cm);
int slb_comm_export(slb_comm* cm, void* handle128);                 /* this rank's mailbox handle */
int slb_comm_connect(slb_comm* cm, const void* all_handles);        /* nranks * 128 bytes, rank order */
/* any slb_malloc'ed buffer: export / map / unmap (same process: the pointer itself, peer access enabled) */
int slb_comm_export_buffer(slb_comm* cm, void* dev, void* handle128);
int slb_comm_open_buffer(slb_comm* cm, const void* handle128, void** dev_out);
int slb_comm_close_buffer(slb_comm* cm, void* dev);
/* stream-ordered barrier over all ranks (one small kernel: flag stores to every peer, spin on the own flags) */
int slb_comm_barrier(slb_comm* cm);
/* all-gather of n <= nslot_doubles doubles per rank: afterwards *slots_out (device pointer into this rank's
 * mailbox) holds [nranks][n] contiguous, rank order, identical on every rank.  Doubles as a barrier: it returns
 * (in stream order) only when every rank's contribution has arrived, i.e. every rank has reached this call.
 * Charge-density slabs (transposing driver: the gathered array IS rho) and per-rank partial charge densities
 * (halo driver: slb_poisson_solve_partial sums them in rank order) travel this way. */
int slb_comm_allgather(slb_comm* cm, const double* local_dev, int64_t n, const double** slots_out);

/* point-to-point ordering between two ranks (4 independent slots per pair): slb_comm_signal raises a flag in `peer`'s
 * mailbox once everything enqueued before it on the stream of `on_ctx` (NULL: the comm's own context; e.g. a second
 * context used as a copy stream for halo planes) is complete; slb_comm_wait holds the stream until the matching
 * signal of `peer` has arrived.  Signals and waits of a (pair, slot) must match one to one. */
int slb_comm_signal(slb_comm* cm, slb_ctx* on_ctx, int peer, int slot);
int slb_comm_wait(slb_comm* cm, slb_ctx* on_ctx, int peer, int slot);

/* ---- sweeps fused with the multi-GPU re-shard -------------------------------- */
/* A 2D2V grid sharded over P ranks alternates between two slab layouts; the all-to-all between
 * them (replacing mpibroadcast, src/mpiinterface.jl:17-38) moves contiguous blocks only when the
 * sweep before it writes, or the sweep after it reads, a BLOCK-MAJOR array: nblocks consecutive
 * sub-arrays, each with extent[bdim] / nblocks along bdim.
 *   SLB_RESHARD_OUT_BLOCKED: the output is written block-major along the swept dim (bdim == dim > 0);
 *                            after the exchange that dim is the sharded one.
 *   SLB_RESHARD_IN_BLOCKED : the input (front buffer) is block-major along bdim != dim; dim must be 0. */
#define SLB_RESHARD_NONE 0
#define SLB_RESHARD_OUT_BLOCKED 1
#define SLB_RESHARD_IN_BLOCKED 2
int slb_sweep_ex(slb_grid* g, int dim, const slb_interp* it, const double* alpha_tab, int64_t alpha_len,
                 const int64_t* alpha_strides, double alpha_scale, int alpha_on_device, int flags,
                 int reshard_mode, int bdim, int nblocks);

/* Fused sweep + all-to-all over NVLink peer memory: like SLB_RESHARD_OUT_BLOCKED along the swept
 * dim, but k-block q of the output is stored to block_bases[q] -- a pointer into the buffer of the
 * rank that owns block q after the exchange (this rank's own HBM, or a peer's mapped with
 * slb_ipc_open_handle) -- instead of this grid's back buffer.  The grid's front/back roles do not
 * change.  Callers synchronise the ranks (any stream-ordered barrier) before reading the result. */
int slb_sweep_peer(slb_grid* g, int dim, const slb_interp* it, const double* alpha_tab, int64_t alpha_len,
                   const int64_t* alpha_strides, double alpha_scale, int alpha_on_device, int flags,
                   int nblocks, double* const* block_bases);
/* CUDA IPC plumbing for the above (handles are 64 bytes; memory must come from slb_malloc) */
int slb_ipc_get_handle(slb_ctx* ctx, void* dev, void* handle64);
int slb_ipc_open_handle(slb_ctx* ctx, const void* handle64, void** dev_out);
int slb_ipc_close_handle(slb_ctx* ctx, void* dev);

/* sol(interp, b) applied to every line along dim (src/interpolation.jl:40,
 * src/bsplinelu.jl:275-284, src/bsplinefft.jl:49-58); in place on the grid (front buffer
 * after the call).  Exposed for tests. */
int slb_presolve(slb_grid* g, int dim, const slb_interp* it);

/* ---- Vlasov-Poisson field solve ------------------------------------------------------ */
/* compute_charge! (src/util_poisson.jl:68-79): rho = dv * sum over the trailing
 * (ndims - nsp) dims of f, minus its mean.  rho_dev holds prod(extents[0..nsp)) doubles. */
int slb_charge_density(slb_grid* g, int nsp, double dv, double* rho_dev);
/* same without the mean subtraction and over a sub-range; used by the sharded driver,
 * which all-gathers slabs of rho before removing the mean */
int slb_charge_density_raw(slb_grid* g, int nsp, double dv, double* rho_dev);
int slb_subtract_mean(slb_ctx* ctx, double* dev, int64_t n);
/* Charge density without a dedicated pass over f: when `linesum_dev` is set (nlines doubles), every
 * fast-path sweep along a dim > 0 also stores, per line, the sum of the line's outputs.  After a
 * sweep along a VELOCITY dim those sums are f reduced over that dim, so
 * slb_charge_density_from(linesums viewed as [nsp_total, nv_total / n_dim]) equals compute_charge!
 * of the swept array (src/util_poisson.jl:68-79) at 1/n_dim of the traffic.  NULL disables. */
int slb_grid_set_linesum(slb_grid* g, double* linesum_dev);
int slb_charge_density_from(slb_ctx* ctx, const double* f_dev, int64_t nsp_total, int64_t nv_total, double dv,
                            double* rho_dev, int subtract_mean);

/* Charge density after a SPACE pass without another pass over f: when `rhopart_dev` is set, the next pair-fused
 * pass whose first sweep runs along dim 0 (x1 x2 of a 2D2V grid) processes several passive (velocity) points per
 * thread block and also stores their sums per space point: slb_grid_rhopart_planes() partial planes
 * [plane][prod(space extents)], 1/4 of the bytes of f.  slb_vp_field_solve(plan, rhopart_dev, planes, dv, ...) then
 * reduces those instead of f (fixed summation order: reproducible).  The pass falls back to the plain kernel (0
 * planes) when the combination is not supported or the buffer is too small (capacity: numel / 2 doubles suffices). */
int slb_grid_set_rhopart(slb_grid* g, double* rhopart_dev, int64_t capacity_doubles);
int64_t slb_grid_rhopart_planes(const slb_grid* g);

/* PoissonConst (src/poisson.jl:35-59): fctv_imag[x] = imag part of fctv_k[x]
 * (src/poisson.jl:7-15), each prod(extents) doubles, column-major, host pointers. */
int slb_poisson_create(slb_ctx* ctx, int nsp, const int64_t* extents, const double* const* fctv_imag,
                       slb_poisson** out);
void slb_poisson_destroy(slb_poisson* p);
/* compute_elfield! (src/poisson.jl:139-144): E_x = real(ifft(fctv_k[x] .* fft(rho))).
 * E_dev: nsp device pointers, each prod(extents) doubles. */
int slb_poisson_solve(slb_poisson* p, const double* rho_dev, double* const* E_dev);

/* The field solve of one initcoef! (src/poisson.jl:171-176) with two launches instead of thirteen:
 * the charge-density pass over f_dev ([prod(extents), nv_total] doubles: the grid's front buffer, or
 * the line sums of the last velocity sweep) and ONE cooperative kernel that finishes rho (mean
 * removed, src/util_poisson.jl:77) and runs every DFT pass of compute_elfield!.  rho_dev and E_dev
 * as above.  Falls back to the separate kernels where cooperative launch is unavailable. */
int slb_vp_field_solve(slb_poisson* p, const double* f_dev, int64_t nv_total, double dv, double* rho_dev,
                       double* const* E_dev);
/* same, starting from a charge density that is already reduced (e.g. all-gathered slabs of the
 * sharded driver); subtract_mean != 0 removes its mean first.  rho_dev is updated in place. */
int slb_poisson_solve_raw(slb_poisson* p, double* rho_dev, int subtract_mean, double* const* E_dev);

/* same as slb_poisson_solve_raw, from `nparts` partial charge densities stored [nparts][prod(extents)]
 * (e.g. the slots of slb_comm_allgather): rho = scale * (part_0 + part_1 + ...) summed in that order, mean
 * removed when subtract_mean != 0, then compute_elfield! -- all in the one cooperative kernel. */
int slb_poisson_solve_partial(slb_poisson* p, const double* partial_dev, int nparts, double scale, int subtract_mean,
                              double* rho_dev, double* const* E_dev);

/* sum(x .^ 2) for compute_ee (src/util_poisson.jl:156-162); deterministic; synchronises */
int slb_reduce_sumsq(slb_ctx* ctx, const double* dev, int64_t n, double* host_out);
/* same without synchronising: out_dev[0] = scale * sum(x .^ 2), a device location (usable inside a capture) */
int slb_reduce_sumsq_async(slb_ctx* ctx, const double* dev, int64_t n, double scale, double* out_dev);
/* sum(x) (deterministic; synchronises) */
int slb_reduce_sum(slb_ctx* ctx, const double* dev, int64_t n, double* host_out);
/* compute_ke (src/util_poisson.jl:41-53): (dsp*dv) * sum(v_square .* sum_sp f); v_square_dev
 * holds prod(extents[nsp..)) doubles on the device; scale = dsp*dv.  Synchronises. */
int slb_kinetic_energy(slb_grid* g, int nsp, const double* v_square_dev, double scale, double* host_out);

/* ---- N-D per-point interpolation: the unsplit 2-D solvers ------------------- */
/* interpolate!(fp, fi, bufdec::Array{OpTuple{2}}, interp_t) (src/interpolation.jl:561-621) and its
 * closure form interpolate!(fp, fi, dec::Function, interp_t) (:401-429) for N = 2: every point
 * (i, j) of an [n1, n2] slice has its own displacement (dec[i,j,0], dec[i,j,1]) in grid units,
 *     d_x = floor(a_x), w^x = tabfct_x(a_x - d_x),
 *     out[i,j,c] = sum_{a,b} res[(i + d_1 - p_1/2 + a) mod n1, (j + d_2 - p_2/2 + b) mod n2, c] * w^1_a * w^2_b,
 * res = sol(interp_t, in) (the 1-D B-spline solves along each dim, :48-94).  `ncomp` component
 * planes share the displacements: 1 for a scalar field, 2 for the OpTuple{2} displacement fields the
 * Adams-Bashforth time algorithms interpolate (autointerp!/interpbufc!, :626-682).  All arrays are
 * device pointers, column-major [n1, n2, ncomp] / [n1, n2, 2].  out must not alias in.  work_dev
 * ([n1, n2, ncomp]) is required when an interpolation is a B-spline and in_dev is then overwritten
 * by intermediate solves.  SLB_SWEEP_EXACT in flags: rounded products summed in column-major order
 * (the reference's sum(res[...] .* tab)) instead of FMA chains.
 * Replaces the single-state branch of advection! (src/advection.jl:607-619). */
int slb_interp2d_points(slb_ctx* ctx, const slb_interp* it1, const slb_interp* it2, int64_t n1, int64_t n2, int ncomp,
                        double* in_dev, const double* dec_dev, double* out_dev, double* work_dev, int flags);
/* dec[i,j,0] = scale_j * tab_j[j], dec[i,j,1] = scale_i * tab_i[i]: the bufcur fill loops of
 * initcoef! for StdPoisson2d (src/poisson.jl:229-247: tab_j = v nodes, tab_i = E) and the unsplit
 * rotation (src/rotation.jl:36-54).  Device pointers. */
int slb_fill_dec2d(slb_ctx* ctx, double* dec_dev, int64_t n1, int64_t n2, const double* tab_j_dev, double scale_j,
                   const double* tab_i_dev, double scale_i);
/* out = sum_k coefs[k] * x[k] on device arrays of n doubles, products rounded and summed left to
 * right: sum(map(k -> c(abcoef, k, ord) * t_bufc[k], 1:ord)) (src/advection.jl:431, :479, :506,
 * :540).  coefs and the pointer list are host arrays; nterms <= 8; out may be one of the x[k]. */
int slb_lincomb(slb_ctx* ctx, double* out_dev, int nterms, const double* coefs, const double* const* x_dev, int64_t n);
/* device-to-device copy on the context's stream (copy(buf), fmr .= to in src/interpolation.jl:636-655) */
int slb_memcpy_d2d(slb_ctx* ctx, void* dst_dev, const void* src_dev, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* SLB200_H */
