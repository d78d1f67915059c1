// slb_api.cu -- C ABI of libslb200.so (see include/slb200.h).  Host-side object management
// and kernel dispatch; all arithmetic lives in the kernels of slb_sweep.cuh, slb_bspline.cuh
// and slb_field.cuh.  No CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/slb200.h"
#include "slb_internal.h"
#include "slb_sweep.cuh"
#include "slb_pair.cuh"
#include "slb_bspline.cuh"
#include "slb_bspfused.cuh"
#include "slb_bspsplit.cuh"
#include "slb_bspseg.cuh"
#include "slb_field.cuh"
#include "slb_points.cuh"
#include "slb_program.cuh"
#include "slb_program_host.h"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

int slb_fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define fail slb_fail

// ------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------
struct slb_grid {
    slb_ctx* ctx;
    int nd;
    int64_t ext[SLB_MAX_DIMS];
    int64_t numel;
    double *front, *back;
    bool owned;
    double* linesum;  // optional: per-line sums of the next strided sweeps' outputs
    double* rhopart;  // optional: partial charge planes written by the next fused space pass (slb_grid_set_rhopart)
    int64_t rhopart_cap, rhopart_planes;
};

struct slb_interp {
    slb_ctx* ctx;
    int kind, order, nc;
    int64_t n;
    std::vector<double> coef;  // (order+1) x nc
    bool fast;                 // fits the templated kernels
    CoefTab tab;
    double* coef_dev;
    BsplineDev bsp;            // LU factors / circulant symbol on the device (B-spline kinds)
    double* bsptab_dev;        // the same factors in the fused sweep's record layout (device), or NULL
    BspFusedTab bsptab;
    double* bsprf_dev;         // table of the recursive-filter form of the pre-solve (slb_bsprf.cuh), or NULL
    BspRfTab bsprf;
    double* bspstab_dev;       // tables of the split-line fused sweep (two warps per tile, slb_bspsplit.cuh), or NULL
    BspSplitTab bspstab;
    int seg_ok, wl_ok;         // segmented sweeps (slb_bspseg.cuh): block-per-tile / warp-per-line plans exist
    BspSegTab segtab, segtab_c, wltab;  // strided dims, dim 0, long lines
};

struct slb_poisson {
    slb_ctx* ctx;
    int nsp;
    int64_t ext[SLB_MAX_DIMS];
    int64_t ntot;
    double2* tw[SLB_MAX_DIMS];
    double* mult[SLB_MAX_DIMS];
    double2 *wa, *wb, *wc;
    double2* wx[2 * SLB_FIELD_MAXDIM];  // per-component work buffers of the one-kernel field solve
    double* red;                        // its block sums
    int coop_blocks;                    // cooperative grid size (0: cooperative launch unavailable)
    int fft_ok;                         // power-of-two extents within the cluster FFT kernel's limits (k_field_fft)
    int fft_log[2];
    double* fft_mean;
};

static long long env_ll(const char* name, long long dflt)
{
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return atoll(v);
}

static int ensure_scratch(slb_ctx* c, size_t bytes)
{
    if (bytes <= c->scratch_bytes) return SLB_OK;
    if (c->scratch) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaFree(c->scratch));
        c->scratch = nullptr;
        c->scratch_bytes = 0;
    }
    size_t want = bytes + bytes / 4 + 4096;
    CUDA_TRY(cudaMalloc(&c->scratch, want));
    c->scratch_bytes = want;
    return SLB_OK;
}

// ------------------------------------------------------------------------------------------
// step programs (slb_program.cuh): recording state
// ------------------------------------------------------------------------------------------
struct ProgOpHost : ProgAccess {   // reads / writes: the byte ranges the op touches (slb_program_host.h)
    ProgOp op;
};
struct slb_prog_rec {
    std::vector<ProgOpHost> ops;
    std::vector<std::pair<slb_grid*, double*>> grids;  // grids touched and their front buffer when first seen
    std::vector<void*> owned;                          // device buffers the program owns (partial sums)
    int nmax_field;
    const double* last_E;   // E output of the last recorded field solve: sweeps that use it as their shift table read the block-local copy
    bool failed;
};
// compute entry points that cannot be part of a step program refuse to run while one is being recorded (nothing
// executes during a recording: the caller falls back to stepwise calls or to a CUDA graph)
#define NOT_RECORDABLE(ctx, name)                                                                                   \
    do {                                                                                                            \
        if ((ctx) && (ctx)->prog_rec) {                                                                             \
            (ctx)->prog_rec->failed = true;                                                                         \
            return fail(SLB_E_UNSUPPORTED, name ": this call cannot be recorded in a step program");               \
        }                                                                                                           \
    } while (0)
static void prog_note_grid(slb_prog_rec* r, slb_grid* g)
{
    for (auto& e : r->grids)
        if (e.first == g) return;
    r->grids.push_back({g, g->front});
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
extern "C" const char* slb_last_error(void) { return g_err.c_str(); }

extern "C" int slb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int slb_ctx_create(int device_id, void* stream, slb_ctx** out)
{
    if (!out) return fail(SLB_E_ARG, "slb_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SLB_E_CUDA, "slb_ctx_create: no CUDA device (%s); libslb200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    }
    if (device_id < 0 || device_id >= n) return fail(SLB_E_ARG, "slb_ctx_create: device %d out of range [0,%d)", device_id, n);
    CUDA_TRY(cudaSetDevice(device_id));
    slb_ctx* c = new slb_ctx();
    memset(c, 0, sizeof(*c));
    c->device = device_id;
    if (stream) {
        c->stream = (cudaStream_t)stream;
        c->own_stream = false;
    } else {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    CUDA_TRY(cudaMalloc(&c->red_partial, 1024 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c->red_out, 8 * sizeof(double)));
    CUDA_TRY(cudaMallocHost(&c->host_out, 8 * sizeof(double)));
    CUDA_TRY(cudaMalloc(&c->err_word, sizeof(int)));
    CUDA_TRY(cudaMemset(c->err_word, 0, sizeof(int)));
    CUDA_TRY(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device_id));
    *out = c;
    return SLB_OK;
}

extern "C" void slb_ctx_destroy(slb_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaFree(c->red_partial);
    cudaFree(c->red_out);
    cudaFreeHost(c->host_out);
    cudaFree(c->err_word);
    if (c->scratch) cudaFree(c->scratch);
    if (c->prog_rec) {  // a recording that was never closed
        for (void* p : c->prog_rec->owned) cudaFree(p);
        delete c->prog_rec;
    }
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int slb_sync(slb_ctx* c)
{
    if (!c) return fail(SLB_E_ARG, "slb_sync: ctx is NULL");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return SLB_OK;
}

extern "C" int64_t slb_launch_count(const slb_ctx* c) { return c ? c->launches : 0; }

extern "C" int slb_timer_start(slb_ctx* c)
{
    if (!c) return fail(SLB_E_ARG, "ctx is NULL");
    CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
    return SLB_OK;
}

extern "C" int slb_timer_stop(slb_ctx* c, float* ms)
{
    if (!c || !ms) return fail(SLB_E_ARG, "ctx/ms is NULL");
    CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
    CUDA_TRY(cudaEventSynchronize(c->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return SLB_OK;
}

extern "C" int slb_event_create(slb_ctx* c, void** ev)
{
    if (!c || !ev) return fail(SLB_E_ARG, "slb_event_create: NULL argument");
    cudaEvent_t e;
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventCreate(&e));
    *ev = (void*)e;
    return SLB_OK;
}

extern "C" int slb_event_record(slb_ctx* c, void* ev)
{
    if (!c || !ev) return fail(SLB_E_ARG, "slb_event_record: NULL argument");
    CUDA_TRY(cudaEventRecord((cudaEvent_t)ev, c->stream));
    return SLB_OK;
}

extern "C" int slb_stream_wait_event(slb_ctx* c, void* ev)
{
    if (!c || !ev) return fail(SLB_E_ARG, "slb_stream_wait_event: NULL argument");
    CUDA_TRY(cudaStreamWaitEvent(c->stream, (cudaEvent_t)ev, 0));
    return SLB_OK;
}

extern "C" int slb_event_elapsed_ms(void* e0, void* e1, float* ms)
{
    if (!e0 || !e1 || !ms) return fail(SLB_E_ARG, "slb_event_elapsed_ms: NULL argument");
    CUDA_TRY(cudaEventSynchronize((cudaEvent_t)e1));
    CUDA_TRY(cudaEventElapsedTime(ms, (cudaEvent_t)e0, (cudaEvent_t)e1));
    return SLB_OK;
}

extern "C" int slb_event_destroy(void* ev)
{
    if (ev) CUDA_TRY(cudaEventDestroy((cudaEvent_t)ev));
    return SLB_OK;
}

extern "C" int slb_malloc(slb_ctx* c, int64_t bytes, void** out)
{
    if (!c || !out || bytes < 0) return fail(SLB_E_ARG, "slb_malloc: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMalloc(out, (size_t)(bytes > 0 ? bytes : 8)));
    return SLB_OK;
}

extern "C" int slb_free(slb_ctx* c, void* dev)
{
    if (!c) return fail(SLB_E_ARG, "ctx is NULL");
    if (dev) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        CUDA_TRY(cudaFree(dev));
    }
    return SLB_OK;
}

extern "C" int slb_host_alloc(int64_t bytes, void** out)
{
    if (!out || bytes < 0) return fail(SLB_E_ARG, "slb_host_alloc: bad argument");
    CUDA_TRY(cudaMallocHost(out, (size_t)(bytes > 0 ? bytes : 8)));
    return SLB_OK;
}

extern "C" int slb_host_free(void* host)
{
    if (host) CUDA_TRY(cudaFreeHost(host));
    return SLB_OK;
}

extern "C" int slb_memcpy_h2d(slb_ctx* c, void* dev, const void* host, int64_t bytes)
{
    if (!c || !dev || !host || bytes < 0) return fail(SLB_E_ARG, "slb_memcpy_h2d: bad argument");
    CUDA_TRY(cudaMemcpyAsync(dev, host, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    return SLB_OK;
}

extern "C" int slb_memcpy_d2h(slb_ctx* c, void* host, const void* dev, int64_t bytes)
{
    if (!c || !dev || !host || bytes < 0) return fail(SLB_E_ARG, "slb_memcpy_d2h: bad argument");
    CUDA_TRY(cudaMemcpyAsync(host, dev, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    return SLB_OK;
}

// ------------------------------------------------------------------------------------------
// grid
// ------------------------------------------------------------------------------------------
static int grid_init(slb_ctx* c, int nd, const int64_t* ext, slb_grid** out)
{
    if (!c || !ext || !out) return fail(SLB_E_ARG, "slb_grid_create: NULL argument");
    if (nd < 1 || nd > SLB_MAX_DIMS) return fail(SLB_E_ARG, "slb_grid_create: ndims=%d not in [1,%d]", nd, SLB_MAX_DIMS);
    int64_t numel = 1;
    for (int d = 0; d < nd; ++d) {
        if (ext[d] < 1 || ext[d] > (int64_t)1 << 30) return fail(SLB_E_ARG, "slb_grid_create: extent[%d]=%lld invalid", d, (long long)ext[d]);
        numel *= ext[d];
        if (numel > ((int64_t)1 << 40)) return fail(SLB_E_ARG, "slb_grid_create: grid too large");
    }
    slb_grid* g = new slb_grid();
    g->ctx = c;
    g->nd = nd;
    for (int d = 0; d < nd; ++d) g->ext[d] = ext[d];
    g->numel = numel;
    g->front = g->back = nullptr;
    g->owned = false;
    g->linesum = nullptr;
    g->rhopart = nullptr;
    g->rhopart_cap = g->rhopart_planes = 0;
    *out = g;
    return SLB_OK;
}

extern "C" int slb_grid_create(slb_ctx* c, int nd, const int64_t* ext, slb_grid** out)
{
    int rc = grid_init(c, nd, ext, out);
    if (rc) return rc;
    slb_grid* g = *out;
    cudaSetDevice(c->device);
    cudaError_t e = cudaMalloc(&g->front, g->numel * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&g->back, g->numel * sizeof(double));
    if (e != cudaSuccess) {
        const long long nbytes = (long long)(g->numel * 8);
        if (g->front) cudaFree(g->front);
        delete g;
        *out = nullptr;
        cudaGetLastError();
        return fail(SLB_E_ALLOC, "slb_grid_create: cudaMalloc of 2 x %lld bytes failed: %s", nbytes, cudaGetErrorString(e));
    }
    g->owned = true;
    return SLB_OK;
}

extern "C" int slb_grid_create_external(slb_ctx* c, int nd, const int64_t* ext, double* front, double* back, slb_grid** out)
{
    if (!front || !back || front == back) return fail(SLB_E_ARG, "slb_grid_create_external: need two distinct device buffers");
    int rc = grid_init(c, nd, ext, out);
    if (rc) return rc;
    (*out)->front = front;
    (*out)->back = back;
    return SLB_OK;
}

extern "C" void slb_grid_destroy(slb_grid* g)
{
    if (!g) return;
    if (g->owned) {
        cudaStreamSynchronize(g->ctx->stream);
        cudaFree(g->front);
        cudaFree(g->back);
    }
    delete g;
}

extern "C" int slb_grid_upload(slb_grid* g, const double* host)
{
    if (!g || !host) return fail(SLB_E_ARG, "slb_grid_upload: NULL argument");
    NOT_RECORDABLE(g->ctx, "slb_grid_upload");
    CUDA_TRY(cudaMemcpyAsync(g->front, host, g->numel * sizeof(double), cudaMemcpyHostToDevice, g->ctx->stream));
    return SLB_OK;
}

extern "C" int slb_grid_download(const slb_grid* g, double* host)
{
    if (!g || !host) return fail(SLB_E_ARG, "slb_grid_download: NULL argument");
    NOT_RECORDABLE(g->ctx, "slb_grid_download");
    CUDA_TRY(cudaMemcpyAsync(host, g->front, g->numel * sizeof(double), cudaMemcpyDeviceToHost, g->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(g->ctx->stream));
    return SLB_OK;
}

extern "C" int slb_grid_set_linesum(slb_grid* g, double* dev)
{
    if (!g) return fail(SLB_E_ARG, "grid is NULL");
    g->linesum = dev;
    return SLB_OK;
}

extern "C" int slb_grid_set_rhopart(slb_grid* g, double* dev, int64_t capacity_doubles)
{
    if (!g) return fail(SLB_E_ARG, "grid is NULL");
    g->rhopart = dev;
    g->rhopart_cap = dev ? capacity_doubles : 0;
    g->rhopart_planes = 0;
    return SLB_OK;
}

extern "C" int64_t slb_grid_rhopart_planes(const slb_grid* g) { return g ? g->rhopart_planes : 0; }

extern "C" double* slb_grid_front(const slb_grid* g) { return g ? g->front : nullptr; }
extern "C" double* slb_grid_back(const slb_grid* g) { return g ? g->back : nullptr; }
extern "C" int slb_grid_swap(slb_grid* g)
{
    if (!g) return fail(SLB_E_ARG, "grid is NULL");
    double* t = g->front;
    g->front = g->back;
    g->back = t;
    return SLB_OK;
}

// ------------------------------------------------------------------------------------------
// interpolation object
// ------------------------------------------------------------------------------------------
extern "C" int slb_interp_create(slb_ctx* c, int kind, int order, int64_t n, const double* coef, int nc,
                                 const double* node_vals, slb_interp** out)
{
    if (!c || !coef || !out) return fail(SLB_E_ARG, "slb_interp_create: NULL argument");
    *out = nullptr;
    if (kind < SLB_LAGRANGE || kind > SLB_HERMITE) return fail(SLB_E_ARG, "slb_interp_create: unknown kind %d", kind);
    if (order < 1 || order > SLB_MAX_ORDER) return fail(SLB_E_ARG, "slb_interp_create: order=%d not in [1,%d]", order, SLB_MAX_ORDER);
    if (nc < 1 || nc > 2 * SLB_MAX_ORDER + 4) return fail(SLB_E_ARG, "slb_interp_create: ncoef=%d invalid", nc);
    bool bs = (kind == SLB_BSPLINE_LU || kind == SLB_BSPLINE_FFT);
    if (bs) {
        if (!node_vals) return fail(SLB_E_ARG, "slb_interp_create: B-spline kinds need node_vals");
        // src/bsplinelu.jl:257-261; even orders are singular for BSplineFFT as well
        if (order % 2 == 0) return fail(SLB_E_ARG, "order=%d BSpline for even order is not implemented n=%lld", order, (long long)n);
        if (kind == SLB_BSPLINE_FFT && (n < 1 || (n & (n - 1)) != 0))
            return fail(SLB_E_ARG, "BSplineFFT: n=%lld must be a power of two (src/fftbig.jl:57)", (long long)n);
        if (n < order + 1) return fail(SLB_E_ARG, "B-spline: n=%lld must exceed order=%d", (long long)n, order);
    }
    slb_interp* it = new slb_interp();
    it->ctx = c;
    it->kind = kind;
    it->order = order;
    it->nc = nc;
    it->n = n;
    it->coef.assign(coef, coef + (size_t)(order + 1) * nc);
    it->fast = (order + 1 <= SLB_P1MAX && order + 1 >= 2 && nc <= SLB_NCMAX);
    memset(&it->tab, 0, sizeof(it->tab));
    if (it->fast)
        for (int j = 0; j <= order; ++j)
            for (int k = 0; k < nc; ++k) it->tab.c[j * SLB_NCMAX + k] = coef[(size_t)j * nc + k];
    it->coef_dev = nullptr;
    memset(&it->bsp, 0, sizeof(it->bsp));
    cudaSetDevice(c->device);
    cudaError_t e = cudaMalloc(&it->coef_dev, it->coef.size() * sizeof(double));
    if (e == cudaSuccess)
        e = cudaMemcpy(it->coef_dev, it->coef.data(), it->coef.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        delete it;
        cudaGetLastError();
        return fail(SLB_E_CUDA, "slb_interp_create: %s", cudaGetErrorString(e));
    }
    it->bsptab_dev = nullptr;
    memset(&it->bsptab, 0, sizeof(it->bsptab));
    it->bspstab_dev = nullptr;
    memset(&it->bspstab, 0, sizeof(it->bspstab));
    it->bsprf_dev = nullptr;
    memset(&it->bsprf, 0, sizeof(it->bsprf));
    it->seg_ok = it->wl_ok = 0;
    if (bs) {
        std::string msg;
        BsplineHost hb;
        int rc = bspline_build(kind, order, n, node_vals, &it->bsp, msg, &hb);
        if (rc) {
            cudaFree(it->coef_dev);
            delete it;
            return fail(rc, "slb_interp_create: %s", msg.c_str());
        }
        if (it->fast && slb_bspfused_supported(hb.h, hb.n)) {
            std::vector<double> v((size_t)slb_bspfused_tab_doubles(hb.h, hb.n));
            slb_bspfused_fill(&it->bsptab, v.data(), hb.h, hb.n, hb.N, hb.L.data(), hb.U.data(), hb.invd.data(), hb.Ri.data(),
                              hb.G.data(), hb.Sinv.data());
            cudaError_t e2 = cudaMalloc(&it->bsptab_dev, v.size() * sizeof(double));
            if (e2 == cudaSuccess) e2 = cudaMemcpy(it->bsptab_dev, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
            if (e2 != cudaSuccess) {
                if (it->bsptab_dev) cudaFree(it->bsptab_dev);
                it->bsptab_dev = nullptr;
                cudaGetLastError();
            }
        }
        if (it->fast && hb.h <= SLB_BSPRF_HMAX) {
            BspRfHost hr;
            std::string msg2;
            if (bsprf_factor(order, n, node_vals, &hr, msg2) == SLB_OK) {
                it->seg_ok = (slb_bspseg_plan(hr, false, false, &it->segtab) && slb_bspseg_plan(hr, false, true, &it->segtab_c)) ? 1 : 0;
                it->wl_ok = slb_bspseg_plan(hr, true, false, &it->wltab) ? 1 : 0;
                std::vector<double> v;
                bsprf_fill(&it->bsprf, v, hr);
                if (slb_bspfused_warps_rf(it->bsprf.ndoubles, hb.n, true) > 0) {
                    cudaError_t e2 = cudaMalloc(&it->bsprf_dev, v.size() * sizeof(double));
                    if (e2 == cudaSuccess) e2 = cudaMemcpy(it->bsprf_dev, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
                    if (e2 != cudaSuccess) {
                        if (it->bsprf_dev) cudaFree(it->bsprf_dev);
                        it->bsprf_dev = nullptr;
                        cudaGetLastError();
                    }
                }
            }
        }
        if (it->fast && slb_bspsplit_tiles(hb.h, hb.n) > 0) {
            BspSplitHost hs;
            std::string msg2;
            if (bspsplit_factor(order, n, node_vals, &hs, msg2) == SLB_OK) {
                std::vector<double> v;
                bspsplit_fill(&it->bspstab, v, hs);
                cudaError_t e2 = cudaMalloc(&it->bspstab_dev, v.size() * sizeof(double));
                if (e2 == cudaSuccess) e2 = cudaMemcpy(it->bspstab_dev, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
                if (e2 != cudaSuccess) {
                    if (it->bspstab_dev) cudaFree(it->bspstab_dev);
                    it->bspstab_dev = nullptr;
                    cudaGetLastError();
                }
            }
        }
    }
    *out = it;
    return SLB_OK;
}

extern "C" void slb_interp_destroy(slb_interp* it)
{
    if (!it) return;
    cudaStreamSynchronize(it->ctx->stream);
    cudaFree(it->coef_dev);
    bspline_free(&it->bsp);
    if (it->bsptab_dev) cudaFree(it->bsptab_dev);
    if (it->bspstab_dev) cudaFree(it->bspstab_dev);
    if (it->bsprf_dev) cudaFree(it->bsprf_dev);
    delete it;
}

// ------------------------------------------------------------------------------------------
// sweep dispatch
// ------------------------------------------------------------------------------------------
struct View {
    long long inner, outer;
    int n;
};

static View make_view(const slb_grid* g, int dim)
{
    View v;
    v.inner = 1;
    v.outer = 1;
    for (int d = 0; d < dim; ++d) v.inner *= g->ext[d];
    for (int d = dim + 1; d < g->nd; ++d) v.outer *= g->ext[d];
    v.n = (int)g->ext[dim];
    return v;
}

template <int P1>
static void launch_strided(slb_ctx* c, const double* in, double* out, const View& v, const AlphaMap& am,
                           const slb_interp* it, const OutMap& om, bool exact, double* linesum)
{
    long long nlines = v.inner * v.outer;
    if (nlines < 32768 && v.n <= 256 && v.n >= 32 && om.kc >= v.n && om.npeer == 0 && env_ll("SLB_SWEEP_CHUNK", 1) != 0) {
        // few lines: cut them into chunks of 16 outputs so that the sweep fills more than a handful of warps
        constexpr int CH = 16;
        dim3 blk(32, (unsigned)((v.n + CH - 1) / CH));
        unsigned nb = (unsigned)((nlines + 31) / 32);
        if (exact)
            k_sweep_strided_chunk<P1, true, CH><<<nb, blk, 0, c->stream>>>(in, out, v.inner, v.n, nlines, am, it->tab, it->nc, linesum);
        else
            k_sweep_strided_chunk<P1, false, CH><<<nb, blk, 0, c->stream>>>(in, out, v.inner, v.n, nlines, am, it->tab, it->nc, linesum);
        return;
    }
    unsigned blocks = (unsigned)((nlines + 127) / 128);
    if (exact)
        k_sweep_strided<P1, true><<<blocks, 128, 0, c->stream>>>(in, out, v.inner, v.n, nlines, am, it->tab, it->nc, om, linesum);
    else
        k_sweep_strided<P1, false><<<blocks, 128, 0, c->stream>>>(in, out, v.inner, v.n, nlines, am, it->tab, it->nc, om, linesum);
}

template <int P1, int R>
static void launch_contig_r(slb_ctx* c, const double* in, double* out, const View& v, const AlphaMap& am,
                            const slb_interp* it, bool exact, const InMap& im)
{
    long long nlines = v.outer;
    // lines per warp: up to 16 consecutive lines, but keep >= ~8 resident-warp waves of work
    long long lpw = nlines / ((long long)c->sm_count * 32 * 8);
    if (lpw > 16) lpw = 16;
    if (lpw < 1) lpw = 1;
    long long nwarps = (nlines + lpw - 1) / lpw;
    unsigned blocks = (unsigned)((nwarps + 7) / 8);
    if (exact)
        k_sweep_contig<P1, R, true><<<blocks, 256, 0, c->stream>>>(in, out, v.n, nlines, am, it->coef_dev, it->nc, im, (int)lpw);
    else
        k_sweep_contig<P1, R, false><<<blocks, 256, 0, c->stream>>>(in, out, v.n, nlines, am, it->coef_dev, it->nc, im, (int)lpw);
}

template <int P1>
static void launch_contig(slb_ctx* c, const double* in, double* out, const View& v, const AlphaMap& am,
                          const slb_interp* it, bool exact, const InMap& im)
{
    if constexpr (P1 % 2 == 0) {
    if (im.c == 0 && v.n % 2 == 0 && v.n >= P1 && 8 * (v.n / 2 + P1 / 2) <= SLB_TILE_NSLOT * 256 && ((uintptr_t)in % 16 == 0) && ((uintptr_t)out % 16 == 0) &&
        env_ll("SLB_CONTIG_TILE", 1) != 0) {
        // whole lines staged by cp.async (k_sweep_contig_tile): persistent blocks, two tiles of LT lines (a multiple of the
        // 8 warps, ~17 KB) each; small grids get smaller tiles so that every SM has work
        const long long nlines = v.outer;
        long long LT = env_ll("SLB_CONTIG_TILE_BYTES", 18000) / (8 * (v.n + P1));   // 128-point lines: 16 per tile (0.74 ms at 128^4; 24: 0.80)
        const long long spread = nlines / (2 * (long long)c->sm_count);
        if (LT > spread) LT = spread;
        LT = LT < 8 ? 8 : (LT / 8) * 8;
        if (LT > 256) LT = 256;
        while (LT > 8 && LT * (v.n / 2 + P1 / 2) > SLB_TILE_NSLOT * 256) LT -= 8;   // the kernel's per-thread copy slots
        const size_t smem = ((size_t)2 * LT * (size_t)(v.n + P1) + (size_t)2 * LT) * sizeof(double);
        const long long ntiles = (nlines + LT - 1) / LT;
        long long per_sm = (long long)(200 * 1024) / (long long)(smem + 2048);
        const long long ctas = env_ll("SLB_CONTIG_TILE_CTAS", 4);
        if (per_sm > ctas) per_sm = ctas;
        if (per_sm < 1) per_sm = 1;
        long long nb = per_sm * c->sm_count;
        if (nb > ntiles) nb = ntiles;
        const unsigned blocks = (unsigned)nb;
        if (exact) {
            if (smem > 48 * 1024) cudaFuncSetAttribute(k_sweep_contig_tile<P1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_sweep_contig_tile<P1, true><<<blocks, 256, smem, c->stream>>>(in, out, v.n, nlines, am, it->coef_dev, it->nc, (int)LT);
        } else {
            if (smem > 48 * 1024) cudaFuncSetAttribute(k_sweep_contig_tile<P1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k_sweep_contig_tile<P1, false><<<blocks, 256, smem, c->stream>>>(in, out, v.n, nlines, am, it->coef_dev, it->nc, (int)LT);
        }
        return;
    }
    }
    if (v.n >= 96)
        launch_contig_r<P1, 4>(c, in, out, v, am, it, exact, im);
    else if (v.n >= 48)
        launch_contig_r<P1, 2>(c, in, out, v, am, it, exact, im);
    else
        launch_contig_r<P1, 1>(c, in, out, v, am, it, exact, im);
}

#define SLB_FOR_P1(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14)

// stencil pass in -> out along `dim`; `in` already holds sol(interp, f)
static int launch_stencil(slb_grid* g, int dim, const slb_interp* it, const double* in, double* out,
                          const AlphaMap& am, int flags, const OutMap* omp, const InMap* imp = nullptr)
{
    slb_ctx* c = g->ctx;
    View v = make_view(g, dim);
    bool exact = (flags & SLB_SWEEP_EXACT) != 0;
    OutMap om;
    if (omp)
        om = *omp;
    else {
        memset(&om, 0, sizeof(om));
        om.kc = v.n;
        om.kblk = 0;
        om.bstride = (long long)v.n * v.inner;
    }
    int P1 = it->order + 1;
    if (g->linesum && (dim == 0 || !it->fast))
        return fail(SLB_E_UNSUPPORTED, "line sums are produced by fast-path sweeps along dim > 0 only");
    InMap im;
    memset(&im, 0, sizeof(im));
    if (imp) {
        if (dim != 0 || !it->fast || omp) return fail(SLB_E_UNSUPPORTED, "blocked input is implemented for fast-path sweeps along dim 0 only");
        im = *imp;
    }
    if (it->fast) {
        if (dim == 0 && !omp) {
            switch (P1) {
#define X(P) case P: launch_contig<P>(c, in, out, v, am, it, exact, im); break;
                SLB_FOR_P1(X)
#undef X
            }
        } else {
            switch (P1) {
#define X(P) case P: launch_strided<P>(c, in, out, v, am, it, om, exact, g->linesum); break;
                SLB_FOR_P1(X)
#undef X
            }
        }
    } else {
        if (omp) return fail(SLB_E_UNSUPPORTED, "re-shard output mapping needs order+1 <= %d", SLB_P1MAX);
        long long nlines = v.inner * v.outer;
        unsigned blocks = (unsigned)((nlines + 127) / 128);
        k_sweep_generic<<<blocks, 128, 0, c->stream>>>(in, out, v.inner, v.n, nlines, am, it->coef_dev, P1, it->nc, exact ? 1 : 0);
    }
    LAUNCH_CHECK(c);
    return SLB_OK;
}

// Compress dims [d0, d1) into terms idx = (x / div) % ext: zero-stride dims are dropped (their
// extent only advances div), neighbours whose strides are layout-compatible are merged, and
// a term that reaches the top of the group needs no modulo (ext = 0).
static int compress_terms(const slb_grid* g, const int64_t* astr, int d0, int d1, AlphaTerm* out)
{
    int nt = 0;
    unsigned long long div = 1;
    int d = d0;
    while (d < d1) {
        if (astr[d] == 0) {
            div *= (unsigned long long)g->ext[d];
            ++d;
            continue;
        }
        unsigned long long ext = (unsigned long long)g->ext[d];
        long long stride = astr[d];
        int e = d + 1;
        while (e < d1 && astr[e] == stride * (long long)ext) {
            ext *= (unsigned long long)g->ext[e];
            ++e;
        }
        out[nt].div = (unsigned)div;
        out[nt].stride = stride;
        out[nt].ext = (e >= d1) ? 0u : (unsigned)ext;  // top of the group: (x / div) is already < ext
        nt++;
        div *= ext;
        d = e;
    }
    return nt;
}

static int build_alpha_map(slb_grid* g, int dim, const double* alpha_tab, int64_t alpha_len,
                           const int64_t* astr, double scale, int on_device, AlphaMap* am)
{
    slb_ctx* c = g->ctx;
    if (!alpha_tab || alpha_len < 1 || !astr) return fail(SLB_E_ARG, "slb_sweep: alpha table missing");
    int64_t maxoff = 0;
    for (int d = 0; d < g->nd; ++d) {
        if (d == dim) continue;
        if (astr[d] < 0) return fail(SLB_E_ARG, "slb_sweep: negative alpha stride");
        maxoff += astr[d] * (g->ext[d] - 1);
    }
    if (maxoff >= alpha_len) return fail(SLB_E_ARG, "slb_sweep: alpha table too short (%lld needed, %lld given)", (long long)maxoff + 1, (long long)alpha_len);
    const double* tab = alpha_tab;
    if (!on_device) {
        int rc = ensure_scratch(c, (size_t)alpha_len * sizeof(double));
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(c->scratch, alpha_tab, (size_t)alpha_len * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        tab = (const double*)c->scratch;
    }
    memset(am, 0, sizeof(*am));
    am->tab = tab;
    am->scale = scale;
    am->nlo = compress_terms(g, astr, 0, dim, am->lo);
    am->nhi = compress_terms(g, astr, dim + 1, g->nd, am->hi);
    return SLB_OK;
}

static int sweep_impl(slb_grid* g, int dim, const slb_interp* it, const double* alpha_tab, int64_t alpha_len,
                      const int64_t* astr, double scale, int on_device, int flags, const OutMap* omp,
                      const InMap* imp = nullptr)
{
    if (!g || !it) return fail(SLB_E_ARG, "slb_sweep: NULL argument");
    if (dim < 0 || dim >= g->nd) return fail(SLB_E_ARG, "slb_sweep: dim=%d out of range", dim);
    slb_ctx* c = g->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    View v = make_view(g, dim);
    bool bs = (it->kind == SLB_BSPLINE_LU || it->kind == SLB_BSPLINE_FFT);
    if (bs && it->n != v.n) return fail(SLB_E_ARG, "slb_sweep: B-spline object built for n=%lld, line length is %d", (long long)it->n, v.n);
    if (v.inner >= ((long long)1 << 31) || v.outer >= ((long long)1 << 31)) return fail(SLB_E_UNSUPPORTED, "slb_sweep: view too large");
    if (c->prog_rec) {
        // step program: record the sweep (front -> back, roles swap) instead of launching it
        slb_prog_rec* r = c->prog_rec;
        if (bs || !it->fast || !slb_program_supports_p1(it->order + 1) || flags != 0 || omp || imp || (g->linesum && dim == 0) || !on_device ||
            v.n > 4096 || g->numel > ((int64_t)1 << 24)) {
            r->failed = true;
            return fail(SLB_E_UNSUPPORTED, "slb_sweep: step programs record plain Lagrange / Hermite sweeps (odd orders 3..11, device-resident "
                                           "shift tables, no flags) on small grids only");
        }
        ProgOpHost h;
        memset(&h.op, 0, sizeof(h.op));
        int rc = build_alpha_map(g, dim, alpha_tab, alpha_len, astr, scale, 1, &h.op.am);
        if (rc) {
            r->failed = true;
            return rc;
        }
        h.op.kind = SLB_OP_SWEEP;
        h.op.in = g->front;
        h.op.out = g->back;
        h.op.inner = v.inner;
        h.op.outer = v.outer;
        h.op.n = v.n;
        h.op.P1 = it->order + 1;
        h.op.nc = it->nc;
        h.op.coef = it->coef_dev;
        h.reads.push_back(prog_range(g->front, (size_t)g->numel * sizeof(double)));
        if (alpha_tab == r->last_E)
            h.op.tab_local = 1;  // every block holds the field of the last solve in shared memory: no grid barrier after the solve
        else
            h.reads.push_back(prog_range(alpha_tab, (size_t)alpha_len * sizeof(double)));
        h.writes.push_back(prog_range(g->back, (size_t)g->numel * sizeof(double)));
        if (g->linesum) {
            h.op.linesum = g->linesum;
            h.writes.push_back(prog_range(g->linesum, (size_t)(v.inner * v.outer) * sizeof(double)));
        }
        prog_note_grid(r, g);
        r->ops.push_back(h);
        return slb_grid_swap(g);
    }
    AlphaMap am;
    int rc = build_alpha_map(g, dim, alpha_tab, alpha_len, astr, scale, on_device, &am);
    if (rc) return rc;
    const double* src = g->front;
    if (flags & SLB_SWEEP_INSIDE_EDGE) {
        // InsideEdge: one-sided stencils near the line ends (kernel-seam feature, not a hot path)
        if (bs) return fail(SLB_E_UNSUPPORTED, "slb_sweep: InsideEdge is implemented for Lagrange / Hermite interpolations");
        if (omp || imp || g->linesum) return fail(SLB_E_UNSUPPORTED, "slb_sweep: InsideEdge does not combine with re-shard maps or line sums");
        if (v.n < it->order + 1) return fail(SLB_E_ARG, "slb_sweep: InsideEdge needs n >= order + 1");
        if (!on_device) {  // the reference's loops index out of bounds for such shifts: refuse them
            for (int64_t k = 0; k < alpha_len; ++k) {
                double ib = floor(scale * alpha_tab[k]) - (double)(it->order / 2);
                if (!(ib <= 0.0 && ib >= -(double)it->order))
                    return fail(SLB_E_ARG, "slb_sweep: InsideEdge shift %g moves the stencil window outside the line (order %d)",
                                scale * alpha_tab[k], it->order);
            }
        }
        long long nlines = v.inner * v.outer;
        unsigned blocks = (unsigned)((nlines + 127) / 128);
        k_sweep_inside<<<blocks, 128, 0, c->stream>>>(g->front, g->back, v.inner, v.n, nlines, am, it->coef_dev, it->order + 1, it->nc,
                                                       (flags & SLB_SWEEP_EXACT) ? 1 : 0);
        LAUNCH_CHECK(c);
        return slb_grid_swap(g);
    }
    if (bs && !(flags & SLB_SWEEP_EXACT) && (it->seg_ok || it->wl_ok) && env_ll("SLB_BSPLINE_SEG", 1) != 0) {
        // segmented sweeps (slb_bspseg.cuh): S threads per line, line data in registers
        const bool contig = v.inner == 1;
        const bool wl = it->wl_ok && !omp && !imp && !g->linesum;
        const bool sg = it->seg_ok && !(contig && (g->linesum || omp)) && !(imp && !contig);
        if (wl || sg) {
            BspSegArgs a;
            memset(&a, 0, sizeof(a));
            a.in = g->front;
            a.out = g->back;
            a.inner = v.inner;
            a.nlines = v.inner * v.outer;
            a.n = v.n;
            a.nc = it->nc;
            a.am = am;
            if (omp)
                a.om = *omp;
            else {
                a.om.kc = v.n;
                a.om.bstride = (long long)v.n * v.inner;
            }
            if (imp) a.im = *imp;
            a.linesum = g->linesum;
            a.tab = sg ? (contig ? it->segtab_c : it->segtab) : it->wltab;
            const int lrc = sg ? slb_bspseg_launch(a, it->tab, contig, c->stream) : slb_bspwline_launch(a, it->tab, c->stream);
            if (lrc > 0) return fail(SLB_E_CUDA, "slb_sweep: segmented B-spline launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
            if (lrc == 0) {
                c->launches++;
                return slb_grid_swap(g);
            }
            // lrc < 0: not instantiated for this shape: fall through to the thread-per-line kernels
        }
    }
    const bool use_rf = bs && it->bsprf_dev && env_ll("SLB_BSPLINE_RF", 1) != 0;
    // two warps per line pay off while the look-back start-up of a half line is short (measured at 128^4: order 3
    // 0.94 -> 0.80 ms, order 5 1.05 -> 0.98 ms, order 11 1.55 -> 1.92 ms): orders 3 and 5 only
    const bool rf_split = use_rf && dim > 0 && !omp && !imp && it->bsprf.h <= (int)env_ll("SLB_BSPLINE_RFSPLIT_HMAX", 2) &&
                          slb_bspsplit_tiles_rf(it->bsprf.ndoubles, v.n) > 0;
    if (bs && (rf_split || (!use_rf && it->bspstab_dev)) && dim > 0 && !omp && !imp && !(flags & SLB_SWEEP_EXACT) &&
        env_ll("SLB_BSPLINE_FUSED", 1) != 0 && env_ll("SLB_BSPLINE_SPLIT", 1) != 0) {
        // strided dims: pre-solve + stencil in one pass with two warps per tile of lines (slb_bspsplit.cuh)
        BspSplitArgs a;
        memset(&a, 0, sizeof(a));
        a.in = g->front;
        a.out = g->back;
        a.inner = v.inner;
        a.nlines = v.inner * v.outer;
        a.bstride = (long long)v.n * v.inner;
        a.n = v.n;
        a.nc = it->nc;
        a.am = am;
        a.linesum = g->linesum;
        if (rf_split) {
            a.use_rf = 1;
            a.rf = it->bsprf;
            a.tab_dev = it->bsprf_dev;
            a.tiles = slb_bspsplit_tiles_rf(it->bsprf.ndoubles, v.n);
        } else {
            a.tiles = slb_bspsplit_tiles(it->bspstab.h, v.n);
            a.tab_dev = it->bspstab_dev;
            a.tab = it->bspstab;
        }
        int lrc = slb_bspsplit_launch(a, it->tab, c->sm_count, c->stream);
        if (lrc != 0) return fail(lrc < 0 ? SLB_E_UNSUPPORTED : SLB_E_CUDA, "slb_sweep: split B-spline launch failed (%d)", lrc);
        c->launches++;
        return slb_grid_swap(g);
    }
    if (bs && (use_rf || it->bsptab_dev) && !(flags & SLB_SWEEP_EXACT) && !(g->linesum && dim == 0) && !(omp && dim == 0) && !(imp && dim != 0) &&
        env_ll("SLB_BSPLINE_FUSED", 1) != 0) {
        // pre-solve + stencil in one pass over HBM (slb_bspfused.cuh): front -> back, then swap
        BspFusedArgs a;
        memset(&a, 0, sizeof(a));
        a.in = g->front;
        a.out = g->back;
        a.inner = v.inner;
        a.nlines = v.inner * v.outer;
        a.n = v.n;
        a.nc = it->nc;
        a.am = am;
        if (omp)
            a.om = *omp;
        else {
            a.om.kc = v.n;
            a.om.bstride = (long long)v.n * v.inner;
        }
        if (imp) a.im = *imp;
        a.linesum = g->linesum;
        a.stagger = (int)env_ll("SLB_BSPLINE_STAGGER", 0);
        if (use_rf) {
            a.use_rf = 1;
            a.rf = it->bsprf;
            a.tab_dev = it->bsprf_dev;
            a.warps = slb_bspfused_warps_rf(it->bsprf.ndoubles, v.n, v.inner == 1);
        } else {
            a.tab_dev = it->bsptab_dev;
            a.tab = it->bsptab;
            a.warps = slb_bspfused_warps(it->bsptab.h, v.n, v.inner == 1);
        }
        int lrc = slb_bspfused_launch(a, it->tab, c->sm_count, c->stream);
        if (lrc != 0) return fail(lrc < 0 ? SLB_E_UNSUPPORTED : SLB_E_CUDA, "slb_sweep: fused B-spline launch failed (%d)", lrc);
        c->launches++;
        return slb_grid_swap(g);
    }
    if (bs) {
        // c = sol(interp, line) for every line: front -> back -> (stencil) -> front
        rc = bspline_presolve(c->stream, &it->bsp, g->front, g->back, v.inner, v.n, v.outer, &c->launches);
        if (rc) return fail(rc, "slb_sweep: B-spline pre-solve launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        if (omp || imp) return fail(SLB_E_UNSUPPORTED, "re-shard mappings with B-splines are not implemented");
        rc = launch_stencil(g, dim, it, g->back, g->front, am, flags, nullptr);
        return rc;  // result is in front: no swap
    }
    rc = launch_stencil(g, dim, it, src, g->back, am, flags, omp, imp);
    if (rc) return rc;
    return slb_grid_swap(g);
}

extern "C" int slb_sweep(slb_grid* g, int dim, const slb_interp* it, const double* alpha_tab, int64_t alpha_len,
                         const int64_t* astr, double scale, int on_device, int flags)
{
    return sweep_impl(g, dim, it, alpha_tab, alpha_len, astr, scale, on_device, flags, nullptr);
}

extern "C" int slb_sweep_ex(slb_grid* g, int dim, const slb_interp* it, const double* alpha_tab, int64_t alpha_len,
                            const int64_t* astr, double scale, int on_device, int flags, int reshard_mode, int bdim,
                            int nblocks)
{
    if (reshard_mode == SLB_RESHARD_NONE || nblocks <= 1)
        return sweep_impl(g, dim, it, alpha_tab, alpha_len, astr, scale, on_device, flags, nullptr);
    if (!g) return fail(SLB_E_ARG, "slb_sweep_ex: NULL grid");
    if (dim < 0 || dim >= g->nd || bdim < 0 || bdim >= g->nd) return fail(SLB_E_ARG, "slb_sweep_ex: dim out of range");
    if (g->ext[bdim] % nblocks != 0) return fail(SLB_E_ARG, "slb_sweep_ex: extent %lld of dim %d is not divisible by %d blocks", (long long)g->ext[bdim], bdim, nblocks);
    View v = make_view(g, dim);
    if (reshard_mode == SLB_RESHARD_OUT_BLOCKED) {
        if (bdim != dim || dim == 0) return fail(SLB_E_UNSUPPORTED, "SLB_RESHARD_OUT_BLOCKED: the blocked dim must be the swept dim, and not dim 0");
        OutMap om;
        memset(&om, 0, sizeof(om));
        om.kc = (int)(g->ext[dim] / nblocks);
        om.kblk = g->numel / nblocks;
        om.bstride = (long long)om.kc * v.inner;
        return sweep_impl(g, dim, it, alpha_tab, alpha_len, astr, scale, on_device, flags, &om);
    }
    if (reshard_mode == SLB_RESHARD_IN_BLOCKED) {
        if (dim != 0 || bdim == 0) return fail(SLB_E_UNSUPPORTED, "SLB_RESHARD_IN_BLOCKED: implemented for sweeps along dim 0 with another dim blocked");
        InMap im;
        unsigned long long L = 1;
        for (int d = 1; d < bdim; ++d) L *= (unsigned long long)g->ext[d];
        im.L = (unsigned)L;
        im.q_ext = (unsigned)g->ext[bdim];
        im.c = (unsigned)(g->ext[bdim] / nblocks);
        im.blk_lines = (g->numel / g->ext[0]) / nblocks;
        return sweep_impl(g, dim, it, alpha_tab, alpha_len, astr, scale, on_device, flags, nullptr, &im);
    }
    return fail(SLB_E_ARG, "slb_sweep_ex: unknown reshard mode %d", reshard_mode);
}

extern "C" int slb_sweep_peer(slb_grid* g, int dim, const slb_interp* it, const double* alpha_tab, int64_t alpha_len,
                              const int64_t* astr, double scale, int on_device, int flags, int nblocks,
                              double* const* block_bases)
{
    if (!g || !block_bases) return fail(SLB_E_ARG, "slb_sweep_peer: NULL argument");
    if (dim < 1 || dim >= g->nd) return fail(SLB_E_ARG, "slb_sweep_peer: the swept (blocked) dim must be > 0");
    if (nblocks < 1 || nblocks > SLB_MAX_PEERS) return fail(SLB_E_ARG, "slb_sweep_peer: nblocks=%d not in [1,%d]", nblocks, SLB_MAX_PEERS);
    if (g->ext[dim] % nblocks != 0) return fail(SLB_E_ARG, "slb_sweep_peer: extent %lld not divisible by %d", (long long)g->ext[dim], nblocks);
    View v = make_view(g, dim);
    OutMap om;
    memset(&om, 0, sizeof(om));
    om.kc = (int)(g->ext[dim] / nblocks);
    om.kblk = g->numel / nblocks;
    om.bstride = (long long)om.kc * v.inner;
    om.npeer = nblocks;
    for (int q = 0; q < nblocks; ++q) {
        if (!block_bases[q]) return fail(SLB_E_ARG, "slb_sweep_peer: block_bases[%d] is NULL", q);
        om.blk[q] = block_bases[q];
    }
    // the output does not land in this grid's back buffer: do not swap (sweep_impl swaps; undo)
    int rc = sweep_impl(g, dim, it, alpha_tab, alpha_len, astr, scale, on_device, flags, &om);
    if (rc) return rc;
    return slb_grid_swap(g);
}

// ---- CUDA IPC: map another process's device buffer (one process per GPU) --------
extern "C" int slb_ipc_get_handle(slb_ctx* c, void* dev, void* handle64)
{
    if (!c || !dev || !handle64) return fail(SLB_E_ARG, "slb_ipc_get_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, dev));
    memcpy(handle64, &h, 64);
    return SLB_OK;
}

extern "C" int slb_ipc_open_handle(slb_ctx* c, const void* handle64, void** dev_out)
{
    if (!c || !handle64 || !dev_out) return fail(SLB_E_ARG, "slb_ipc_open_handle: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(dev_out, h, cudaIpcMemLazyEnablePeerAccess));
    return SLB_OK;
}

extern "C" int slb_ipc_close_handle(slb_ctx* c, void* dev)
{
    if (!c) return fail(SLB_E_ARG, "ctx is NULL");
    if (dev) {
        CUDA_TRY(cudaSetDevice(c->device));
        CUDA_TRY(cudaIpcCloseMemHandle(dev));
    }
    return SLB_OK;
}


// ------------------------------------------------------------------------------------------
// pair-fused sweeps (slb_pair.cuh): two advection! stages in one pass over HBM
// ------------------------------------------------------------------------------------------
static int check_alpha_table(const slb_grid* g, int dim, const double* tab, int64_t len, const int64_t* astr)
{
    if (!tab || len < 1 || !astr) return fail(SLB_E_ARG, "slb_sweep_pair: alpha table missing");
    int64_t maxoff = 0;
    for (int d = 0; d < g->nd; ++d) {
        if (d == dim) continue;
        if (astr[d] < 0) return fail(SLB_E_ARG, "slb_sweep_pair: negative alpha stride");
        maxoff += astr[d] * (g->ext[d] - 1);
    }
    if (maxoff >= len) return fail(SLB_E_ARG, "slb_sweep_pair: alpha table too short (%lld needed, %lld given)", (long long)maxoff + 1, (long long)len);
    return SLB_OK;
}


static int sweep_pair_impl(slb_grid* g, int dimA, const slb_interp* itA, const double* alphaA, int64_t alenA,
                           const int64_t* astrA, double scaleA, int dimB, const slb_interp* itB, const double* alphaB,
                           int64_t alenB, const int64_t* astrB, double scaleB, int on_device, int flags, int in_nblocks,
                           int out_nblocks, double* const* out_bases, int first_block, const slb_halo* halo = nullptr)
{
    if (!g || !itA || !itB) return fail(SLB_E_ARG, "slb_sweep_pair: NULL argument");
    slb_ctx* c = g->ctx;
    NOT_RECORDABLE(c, "slb_sweep_pair");
    const int nd = g->nd;
    if (dimA < 0 || dimA >= nd || dimB < 0 || dimB >= nd || dimA == dimB) return fail(SLB_E_ARG, "slb_sweep_pair: need two distinct dims in range");
    int rc = check_alpha_table(g, dimA, alphaA, alenA, astrA);
    if (rc) return rc;
    rc = check_alpha_table(g, dimB, alphaB, alenB, astrB);
    if (rc) return rc;
    if (flags & SLB_SWEEP_INSIDE_EDGE) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: InsideEdge sweeps are not pair-fused");
    if (nd > 4) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: grids with more than 4 dims are not pair-fused");
    if (dimB == 0) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: the second sweep must not run along dim 0");
    int mode = SLB_FUSED_PLAIN;
    if (halo) {
        if (in_nblocks > 1 || out_nblocks > 1 || out_bases || first_block) return fail(SLB_E_ARG, "slb_sweep_pair_halo: does not combine with block-major re-shards");
        if (flags & SLB_SWEEP_EXACT) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair_halo: SLB_SWEEP_EXACT is not built for the halo-sharded passes");
        if (halo->halo < 1) return fail(SLB_E_ARG, "slb_sweep_pair_halo: halo=%d must be positive", halo->halo);
        if (halo->mode == SLB_HALO_MARCH) {
            mode = SLB_FUSED_WIN;
            if (dimA == 0) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair_halo: SLB_HALO_MARCH needs a first sweep along a dim > 0");
            if (g->ext[dimB] - 2 * (int64_t)halo->halo < halo->halo)
                return fail(SLB_E_ARG, "slb_sweep_pair_halo: a slab of %lld rows is shorter than the halo %d (halos reach the next neighbour only)",
                            (long long)(g->ext[dimB] - 2 * (int64_t)halo->halo), halo->halo);
            if (halo->halo < (itB->order + 1) / 2) return fail(SLB_E_ARG, "slb_sweep_pair_halo: halo=%d is narrower than half the stencil (order %d)", halo->halo, itB->order);
        } else if (halo->mode == SLB_HALO_PASSIVE) {
            mode = SLB_FUSED_PSH;
            if (dimA != 0) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair_halo: SLB_HALO_PASSIVE is built for passes whose first sweep runs along dim 0");
            if (halo->shard_dim < 0 || halo->shard_dim >= nd || halo->shard_dim == dimA || halo->shard_dim == dimB)
                return fail(SLB_E_ARG, "slb_sweep_pair_halo: shard_dim must be a dim other than the two swept ones");
            if (g->ext[halo->shard_dim] < 2 * (int64_t)halo->halo)
                return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair_halo: a slab of %lld planes is shorter than two halos (%d each)",
                            (long long)g->ext[halo->shard_dim], halo->halo);
        } else
            return fail(SLB_E_ARG, "slb_sweep_pair_halo: unknown mode %d", halo->mode);
    }
    auto plain = [](const slb_interp* it) { return it->fast && it->kind != SLB_BSPLINE_LU && it->kind != SLB_BSPLINE_FFT; };
    if (!plain(itA) || !plain(itB) || itA->order != itB->order)
        return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: both stages need the same order (<= %d) and an identity pre-solve", SLB_P1MAX - 1);
    if (astrA[dimB] != 0) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: the first sweep's shift must not depend on the second sweep's dim");
    const int P1 = itA->order + 1;
    const int64_t ncross = g->ext[dimA], nmarch = g->ext[dimB];
    if (ncross < P1 || ncross >= ((int64_t)1 << 30) || nmarch >= ((int64_t)1 << 30))
        return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: line length out of range for the fused kernel");
    CUDA_TRY(cudaSetDevice(c->device));
    if (in_nblocks < 1) in_nblocks = 1;
    if (out_nblocks < 1) out_nblocks = 1;
    if (in_nblocks > SLB_MAX_PEERS || out_nblocks > SLB_MAX_PEERS || nmarch % in_nblocks != 0 || nmarch % out_nblocks != 0)
        return fail(SLB_E_ARG, "slb_sweep_pair: the block counts (%d in, %d out) must divide extent %lld and not exceed %d", in_nblocks,
                    out_nblocks, (long long)nmarch, SLB_MAX_PEERS);
    // strides of the plain layout (gs), of an input block (is) and of an output block (os): a block is
    // the sub-array with extent[dimB] / nblocks along dimB, stored contiguously
    int64_t gs[SLB_MAX_DIMS], is[SLB_MAX_DIMS], os[SLB_MAX_DIMS], lsstr[SLB_MAX_DIMS];
    int64_t run = 1, irun = 1, orun = 1, lsrun = 1;
    for (int q = 0; q < nd; ++q) {
        gs[q] = run;
        is[q] = irun;
        os[q] = orun;
        run *= g->ext[q];
        irun *= (q == dimB ? g->ext[q] / in_nblocks : g->ext[q]);
        orun *= (q == dimB ? g->ext[q] / out_nblocks : g->ext[q]);
        lsstr[q] = lsrun;  // index of the plain kernel's line id over the dims other than dimB
        if (q != dimB) lsrun *= g->ext[q];
    }
    FusedArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.in = g->front;
    fa.out = g->back;
    fa.ncross = (int)ncross;
    fa.nmarch = (int)nmarch;
    fa.isc = is[dimA];
    fa.ism = is[dimB];
    fa.osc = os[dimA];
    fa.osm = os[dimB];
    fa.ikc = (int)(nmarch / in_nblocks);
    fa.okc = (int)(nmarch / out_nblocks);
    fa.iblk = g->numel / in_nblocks;
    if (first_block < 0 || first_block >= out_nblocks) return fail(SLB_E_ARG, "slb_sweep_pair: first_block out of range");
    fa.march0 = first_block * fa.okc;
    for (int q = 0; q < out_nblocks; ++q) {
        fa.oblk[q] = out_bases ? out_bases[q] : g->back + (int64_t)q * (g->numel / out_nblocks);
        if (!fa.oblk[q]) return fail(SLB_E_ARG, "slb_sweep_pair: out_block_bases[%d] is NULL", q);
    }
    fa.elo = fa.ehi = 1;
    int npass = 0, pd[2] = {-1, -1};
    for (int q = 0; q < nd; ++q) {
        if (q == dimA || q == dimB) continue;
        pd[npass++] = q;
    }
    // halo pushes along a passive dim: make it the FAST passive index, so that consecutive thread blocks belong to
    // different slab planes and the boundary layers' NVLink stores are spread over the whole pass
    if (mode == SLB_FUSED_PSH && npass == 2 && pd[1] == halo->shard_dim) {
        pd[1] = pd[0];
        pd[0] = halo->shard_dim;
    }
    for (int x = 0; x < npass; ++x) {
        const int q = pd[x];
        if (x == 0) {
            fa.elo = (unsigned)g->ext[q];
            fa.islo = is[q];
            fa.oslo = os[q];
            fa.aAlo = astrA[q];
            fa.aBlo = astrB[q];
            fa.lslo = lsstr[q];
        } else {
            fa.ehi = (unsigned)g->ext[q];
            fa.ishi = is[q];
            fa.oshi = os[q];
            fa.aAhi = astrA[q];
            fa.aBhi = astrB[q];
            fa.lshi = lsstr[q];
        }
    }
    if (astrB[dimA] != 0) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: the second sweep's shift must not depend on the first sweep's dim");
    fa.lsc = lsstr[dimA];
    const int64_t np = (int64_t)fa.elo * fa.ehi;
    if (np >= ((int64_t)1 << 31)) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: grid too large");
    // tiles: `ta` (even) cross outputs x `gg` passive points, two cross outputs per thread
    const bool cc = (dimA == 0);
    int gg, ta, full;
    const int64_t nc_even = (ncross + 1) & ~(int64_t)1;
    g->rhopart_planes = 0;
    if (cc) {
        if (nc_even / 2 <= SLB_FUSED_MAXTHREADS) {
            full = 1;
            ta = (int)nc_even;
            int64_t want = 2 * env_ll("SLB_FUSED_CC_THREADS", 64) / ta;  // lines per block
            gg = (int)(want < 1 ? 1 : want);
            if (gg > np) gg = (int)np;
            while ((int64_t)gg * ta / 2 > SLB_FUSED_MAXTHREADS) --gg;
            // partial charge planes (slb_grid_set_rhopart): blocks of several passive points that share sweep B's shift
            if (g->rhopart && mode == SLB_FUSED_PLAIN && !(flags & SLB_SWEEP_EXACT) && in_nblocks == 1 && out_nblocks == 1 && !out_bases &&
                fa.aBlo == 0 && env_ll("SLB_FUSED_RHO", 1) != 0) {
                int gr = (int)env_ll("SLB_FUSED_RHO_G", 4);
                while (gr > 1 && ((int64_t)gr * ta / 2 > SLB_FUSED_MAXTHREADS || fa.elo % (unsigned)gr != 0)) gr >>= 1;
                const int64_t planes = np / (gr > 0 ? gr : 1);
                if (gr >= 2 && planes * ncross * nmarch <= g->rhopart_cap) {
                    mode = SLB_FUSED_RHO;
                    gg = gr;
                }
            }
        } else {
            full = 0;
            ta = 512;
            gg = 1;
        }
    } else {
        gg = (int)env_ll("SLB_FUSED_G", P1 >= 12 ? 16 : 32);  // 256 B rows (128^4 L7: 0.786 ms vs 0.805 with 128 B); smaller when dim 0 is not a multiple
        while (gg > 1 && (fa.elo % gg != 0 || !slb_fused_supported(P1, false, gg))) gg >>= 1;
        int64_t want = 2 * env_ll("SLB_FUSED_THREADS", P1 >= 12 ? 128 : 256) / gg;  // cross outputs per tile (order 11: 188 registers)
        if (want > 2 * SLB_FUSED_MAXTHREADS / gg) want = 2 * SLB_FUSED_MAXTHREADS / gg;
        if (want < 16) want = 16;
        want &= ~(int64_t)1;
        if (want >= nc_even) {
            full = 1;
            ta = (int)nc_even;
        } else {
            full = 0;
            ta = (int)want;
        }
        if ((int64_t)gg * ta / 2 > SLB_FUSED_MAXTHREADS) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: tile does not fit a thread block");
    }
    if (!slb_fused_supported(P1, cc, gg)) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: order %d is not on the fused path", P1 - 1);
    fa.g = gg;
    fa.ta = ta;
    fa.full = full;
    fa.ntile_c = (int)((ncross + ta - 1) / ta);
    // 16-byte fetches (pairs of doubles along the contiguous index) when everything is 16-byte aligned
    bool aligned = ((uintptr_t)g->front % 16 == 0);
    for (int q = 0; q < out_nblocks; ++q) aligned = aligned && ((uintptr_t)fa.oblk[q] % 16 == 0);
    auto even = [](long long v) { return (v & 1) == 0; };
    const bool even_strides = even(fa.ism) && even(fa.osm) && even(fa.iblk) && (fa.ehi == 1 || (even(fa.ishi) && even(fa.oshi)));
    if (cc) {
        fa.w16 = aligned && even(ncross) && even_strides && even(fa.islo) && even(fa.oslo) && fa.isc == 1 && fa.osc == 1;
    } else {
        fa.w16 = gg > 1;
        if (fa.w16 && !(aligned && fa.islo == 1 && even(fa.elo) && even(fa.isc) && even_strides))
            return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: buffers are not 16-byte aligned");
    }
    // every thread fetches at most two pairs (four doubles) per march row: staged rows <= 2 * ta
    for (;;) {
        const int extra = (cc && fa.w16) ? 2 : 0;  // even pitch + room for the alignment element
        fa.spread_max = full ? 0 : 2 * ta - (ta + P1 - 1) - extra;
        if (fa.spread_max > SLB_FUSED_SPREAD_MAX) fa.spread_max = SLB_FUSED_SPREAD_MAX;
        fa.nrows_max = (full ? (int)ncross + P1 : ta + P1 - 1 + fa.spread_max) + extra;
        if (extra) fa.nrows_max &= ~1;
        if (fa.spread_max >= 0 && fa.nrows_max <= 2 * ta) break;
        if (cc && fa.w16) {
            fa.w16 = 0;  // short lines: 8-byte fetches need no alignment slack
            continue;
        }
        return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: lines too short for the fused kernel");
    }
    const double *tabA = alphaA, *tabB = alphaB;
    if (!on_device) {
        rc = ensure_scratch(c, (size_t)(alenA + alenB) * sizeof(double));
        if (rc) return rc;
        double* s = (double*)c->scratch;
        CUDA_TRY(cudaMemcpyAsync(s, alphaA, (size_t)alenA * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(s + alenA, alphaB, (size_t)alenB * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        tabA = s;
        tabB = s + alenA;
    }
    fa.tabA = tabA;
    fa.scaleA = scaleA;
    fa.tabB = tabB;
    fa.scaleB = scaleB;
    fa.ncA = itA->nc;
    fa.ncB = itB->nc;
    fa.linesum = g->linesum;
    if (halo) {
        fa.win_h = halo->halo;
        fa.err = halo->err_flag ? halo->err_flag : c->err_word;
        if (mode == SLB_FUSED_WIN) {
            fa.win_c = (int)nmarch - 2 * halo->halo;
            if (halo->push_lo) fa.pushL = halo->push_lo + (int64_t)fa.win_c * os[dimB];
            if (halo->push_hi) fa.pushR = halo->push_hi - (int64_t)fa.win_c * os[dimB];
        } else {
            const int sd = halo->shard_dim;
            fa.win_c = (int)g->ext[sd];
            fa.push_on_lo = (pd[0] == sd);
            if (halo->push_lo) fa.pushL = halo->push_lo + (int64_t)fa.win_c * os[sd];
            if (halo->push_hi) fa.pushR = halo->push_hi - (int64_t)fa.win_c * os[sd];
        }
    }
    const int64_t nblk = (int64_t)fa.ntile_c * ((np + gg - 1) / gg);
    if (nblk >= 0x7fffffffLL) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: too many blocks");
    size_t smem = slb_fused_smem_bytes(fa.nrows_max, gg);
    if (mode == SLB_FUSED_RHO) {
        fa.rhopart = g->rhopart;
        smem += ((size_t)2 * 2 * gg * ta + 2) * sizeof(double);
    }
    if (smem > 200 * 1024) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: tile does not fit shared memory");
    const bool exact = (flags & SLB_SWEEP_EXACT) != 0;
    if ((mode == SLB_FUSED_WIN && cc) || (mode == SLB_FUSED_PSH && !cc))
        return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair_halo: this dim combination has no halo-sharded kernel");
    int lrc = slb_fused_launch(fa, itA->tab, itB->tab, P1, exact, cc, mode, (unsigned)nblk, (unsigned)(gg * ta / 2), smem, c->stream);
    if (lrc < 0) return fail(SLB_E_UNSUPPORTED, "slb_sweep_pair: no fused kernel for order %d, tile width %d, mode %d", P1 - 1, gg, mode);
    if (lrc != 0) return fail(SLB_E_CUDA, "slb_sweep_pair: launch failed: %s", cudaGetErrorString((cudaError_t)lrc));
    c->launches++;
    if (mode == SLB_FUSED_RHO) g->rhopart_planes = np / gg;
    if (out_bases) return SLB_OK;  // the result left this grid (slb_sweep_peer's convention): roles unchanged
    return slb_grid_swap(g);
}

extern "C" int slb_sweep_pair(slb_grid* g, int dimA, const slb_interp* itA, const double* alphaA, int64_t alenA,
                              const int64_t* astrA, double scaleA, int dimB, const slb_interp* itB, const double* alphaB,
                              int64_t alenB, const int64_t* astrB, double scaleB, int on_device, int flags)
{
    return sweep_pair_impl(g, dimA, itA, alphaA, alenA, astrA, scaleA, dimB, itB, alphaB, alenB, astrB, scaleB, on_device, flags, 1, 1,
                           nullptr, 0);
}

extern "C" int slb_sweep_pair_ex(slb_grid* g, int dimA, const slb_interp* itA, const double* alphaA, int64_t alenA,
                                 const int64_t* astrA, double scaleA, int dimB, const slb_interp* itB, const double* alphaB,
                                 int64_t alenB, const int64_t* astrB, double scaleB, int on_device, int flags, int in_nblocks,
                                 int out_nblocks, double* const* out_block_bases, int first_block)
{
    return sweep_pair_impl(g, dimA, itA, alphaA, alenA, astrA, scaleA, dimB, itB, alphaB, alenB, astrB, scaleB, on_device, flags,
                           in_nblocks, out_nblocks, out_block_bases, first_block);
}

extern "C" int slb_sweep_pair_halo(slb_grid* g, int dimA, const slb_interp* itA, const double* alphaA, int64_t alenA,
                                   const int64_t* astrA, double scaleA, int dimB, const slb_interp* itB, const double* alphaB,
                                   int64_t alenB, const int64_t* astrB, double scaleB, int alpha_on_device, int flags,
                                   const slb_halo* halo)
{
    if (!halo) return fail(SLB_E_ARG, "slb_sweep_pair_halo: halo is NULL");
    return sweep_pair_impl(g, dimA, itA, alphaA, alenA, astrA, scaleA, dimB, itB, alphaB, alenB, astrB, scaleB, alpha_on_device, flags,
                           1, 1, nullptr, 0, halo);
}

extern "C" int slb_halo_error(slb_ctx* c, int* flags_out)
{
    if (!c || !flags_out) return fail(SLB_E_ARG, "slb_halo_error: NULL argument");
    CUDA_TRY(cudaSetDevice(c->device));
    int* h = reinterpret_cast<int*>(c->host_out + 4);
    CUDA_TRY(cudaMemcpyAsync(h, c->err_word, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemsetAsync(c->err_word, 0, sizeof(int), c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *flags_out = *h;
    return SLB_OK;
}

extern "C" int slb_presolve(slb_grid* g, int dim, const slb_interp* it)
{
    if (!g || !it) return fail(SLB_E_ARG, "slb_presolve: NULL argument");
    NOT_RECORDABLE(g->ctx, "slb_presolve");
    if (dim < 0 || dim >= g->nd) return fail(SLB_E_ARG, "slb_presolve: dim out of range");
    bool bs = (it->kind == SLB_BSPLINE_LU || it->kind == SLB_BSPLINE_FFT);
    if (!bs) return SLB_OK;  // sol(interp, b) = b, src/interpolation.jl:40
    View v = make_view(g, dim);
    if (it->n != v.n) return fail(SLB_E_ARG, "slb_presolve: B-spline object built for n=%lld, line length is %d", (long long)it->n, v.n);
    slb_ctx* c = g->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = bspline_presolve(c->stream, &it->bsp, g->front, g->back, v.inner, v.n, v.outer, &c->launches);
    if (rc) return fail(rc, "slb_presolve: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return slb_grid_swap(g);
}

// ------------------------------------------------------------------------------------------
// charge density, reductions
// ------------------------------------------------------------------------------------------
static int reduce_to_dev(slb_ctx* c, const double* dev, int64_t n, int mode, double scale, double* out_dev)
{
    long long nb = (n + 256 * 8 - 1) / (256 * 8);
    if (nb > 1024) nb = 1024;
    if (nb < 1) nb = 1;
    if (mode == 1)
        k_reduce_partial<1><<<(unsigned)nb, 256, 0, c->stream>>>(dev, n, c->red_partial);
    else
        k_reduce_partial<0><<<(unsigned)nb, 256, 0, c->stream>>>(dev, n, c->red_partial);
    LAUNCH_CHECK(c);
    k_reduce_final<<<1, 256, 0, c->stream>>>(c->red_partial, (int)nb, scale, out_dev);
    LAUNCH_CHECK(c);
    return SLB_OK;
}

static int reduce_to_host(slb_ctx* c, const double* dev, int64_t n, int mode, double* host_out)
{
    if (!c || !dev || !host_out || n < 1) return fail(SLB_E_ARG, "slb_reduce: bad argument");
    NOT_RECORDABLE(c, "slb_reduce");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = reduce_to_dev(c, dev, n, mode, 1.0, c->red_out);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->host_out, c->red_out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *host_out = c->host_out[0];
    return SLB_OK;
}

extern "C" int slb_reduce_sumsq(slb_ctx* c, const double* dev, int64_t n, double* host_out) { return reduce_to_host(c, dev, n, 1, host_out); }

extern "C" int slb_reduce_sumsq_async(slb_ctx* c, const double* dev, int64_t n, double scale, double* out_dev)
{
    if (!c || !dev || !out_dev || n < 1) return fail(SLB_E_ARG, "slb_reduce_sumsq_async: bad argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->prog_rec) {
        slb_prog_rec* r = c->prog_rec;
        if (n > (int64_t)SLB_PROG_MAXSUMSQ_BLOCKS * 2048) {
            r->failed = true;
            return fail(SLB_E_UNSUPPORTED, "slb_reduce_sumsq_async: step programs reduce at most %d values", SLB_PROG_MAXSUMSQ_BLOCKS * 2048);
        }
        ProgOpHost h;
        memset(&h.op, 0, sizeof(h.op));
        h.op.kind = SLB_OP_SUMSQ;
        h.op.x = dev;
        h.op.nx = n;
        h.op.scale = scale;
        h.op.outp = out_dev;
        if (dev == r->last_E) {
            h.op.x_local = 1;   // the last block reduces its own copy of the field: no global read, no grid barrier
        } else {
            h.reads.push_back(prog_range(dev, (size_t)n * sizeof(double), true));   // block 0 reduces
        }
        h.writes.push_back(prog_range(out_dev, sizeof(double), true));
        r->ops.push_back(h);
        return SLB_OK;
    }
    return reduce_to_dev(c, dev, n, 1, scale, out_dev);
}

// ------------------------------------------------------------------------------------------
// CUDA graphs: whole time steps of small grids as one launch
// ------------------------------------------------------------------------------------------
struct slb_graph {
    slb_ctx* ctx;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    int64_t launches;  // kernels recorded in the graph (added to the context's count at every replay)
};

extern "C" int slb_capture_begin(slb_ctx* c)
{
    if (!c) return fail(SLB_E_ARG, "slb_capture_begin: ctx is NULL");
    NOT_RECORDABLE(c, "slb_capture_begin");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->capture_launches0 = c->launches;
    CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    return SLB_OK;
}

extern "C" int slb_capture_end(slb_ctx* c, slb_graph** out)
{
    if (!c || !out) return fail(SLB_E_ARG, "slb_capture_end: NULL argument");
    *out = nullptr;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (e != cudaSuccess || !g) {
        cudaGetLastError();
        return fail(SLB_E_CUDA, "slb_capture_end: the captured sequence is not replayable (%s): it must consist of kernel launches on "
                                "device-resident tables only -- no allocation, host table or synchronising call", cudaGetErrorString(e));
    }
    cudaGraphExec_t x = nullptr;
    e = cudaGraphInstantiate(&x, g, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(g);
        cudaGetLastError();
        return fail(SLB_E_CUDA, "slb_capture_end: cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    slb_graph* gr = new slb_graph();
    gr->ctx = c;
    gr->graph = g;
    gr->exec = x;
    gr->launches = c->launches - c->capture_launches0;
    c->launches = c->capture_launches0;  // nothing ran yet
    *out = gr;
    return SLB_OK;
}

extern "C" int slb_graph_launch(slb_graph* g)
{
    if (!g) return fail(SLB_E_ARG, "slb_graph_launch: graph is NULL");
    CUDA_TRY(cudaSetDevice(g->ctx->device));
    CUDA_TRY(cudaGraphLaunch(g->exec, g->ctx->stream));
    g->ctx->launches += g->launches;
    return SLB_OK;
}

extern "C" void slb_graph_destroy(slb_graph* g)
{
    if (!g) return;
    cudaStreamSynchronize(g->ctx->stream);
    cudaGraphExecDestroy(g->exec);
    cudaGraphDestroy(g->graph);
    delete g;
}
// ------------------------------------------------------------------------------------------
// step programs: whole time steps of tiny grids as ONE persistent kernel (slb_program.cuh)
// ------------------------------------------------------------------------------------------
struct slb_program {
    slb_ctx* ctx;
    ProgOp* ops_dev;
    int nops, nbarriers, nsumsq;
    int nblocks, nmax_field, use_cg;
    size_t smem;
    unsigned* bar_ctr;    // grid-barrier counter (device), never reset: generation targets are tracked in bar_done
    unsigned bar_done;    // arrivals of all launches so far (modulo 2^32)
    std::vector<void*> owned;
};

extern "C" int slb_program_begin(slb_ctx* c)
{
    if (!c) return fail(SLB_E_ARG, "slb_program_begin: ctx is NULL");
    if (c->prog_rec) return fail(SLB_E_ARG, "slb_program_begin: a recording is already open on this context");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    slb_prog_rec* r = new slb_prog_rec();
    r->nmax_field = 0;
    r->last_E = nullptr;
    r->failed = false;
    c->prog_rec = r;
    return SLB_OK;
}

extern "C" int slb_program_end(slb_ctx* c, slb_program** out)
{
    if (!c || !out) return fail(SLB_E_ARG, "slb_program_end: NULL argument");
    *out = nullptr;
    slb_prog_rec* r = c->prog_rec;
    if (!r) return fail(SLB_E_ARG, "slb_program_end: no recording is open on this context");
    c->prog_rec = nullptr;
    // nothing ran: every grid gets its roles back; a program must leave them as it found them
    bool odd = false;
    for (auto& e : r->grids) {
        if (e.first->front != e.second) {
            odd = true;
            slb_grid_swap(e.first);
        }
    }
    auto drop = [&]() {
        for (void* p : r->owned) cudaFree(p);
        delete r;
    };
    if (r->failed || r->ops.empty()) {
        const bool f = r->failed;
        drop();
        return fail(SLB_E_UNSUPPORTED, f ? "slb_program_end: the recording contains a call that step programs do not support"
                                         : "slb_program_end: nothing was recorded");
    }
    if (odd) {
        drop();
        return fail(SLB_E_ARG, "slb_program_end: the recorded steps swap a grid's buffers an odd number of times: record an even number of steps");
    }
    // barriers only where an op reads what an earlier one wrote (or overwrites what it read): slb_program_host.h
    const int n = (int)r->ops.size();
    {
        std::vector<ProgAccess> acc(r->ops.begin(), r->ops.end());
        const std::vector<int> flags = prog_place_barriers(acc);
        for (int k = 0; k < n; ++k) r->ops[k].op.barrier_before = flags[k];
    }
    slb_program* pr = new slb_program();
    pr->ctx = c;
    pr->nops = n;
    pr->nbarriers = pr->nsumsq = 0;
    pr->bar_ctr = nullptr;
    pr->ops_dev = nullptr;
    std::vector<ProgOp> flat((size_t)n);
    for (int k = 0; k < n; ++k) {
        flat[k] = r->ops[k].op;
        pr->nbarriers += flat[k].barrier_before;
        pr->nsumsq += flat[k].kind == SLB_OP_SUMSQ;
    }
    pr->owned = r->owned;
    r->owned.clear();
    delete r;
    pr->ops_dev = nullptr;
    cudaError_t e = cudaMalloc(&pr->ops_dev, (size_t)n * sizeof(ProgOp));
    if (e == cudaSuccess) e = cudaMemcpy(pr->ops_dev, flat.data(), (size_t)n * sizeof(ProgOp), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaGetLastError();
        slb_program_destroy(pr);
        return fail(SLB_E_CUDA, "slb_program_end: %s", cudaGetErrorString(e));
    }
    int nmaxf = 0;
    for (int k = 0; k < n; ++k)
        if (flat[k].kind == SLB_OP_FIELD1D && flat[k].ffa.n1 > nmaxf) nmaxf = flat[k].ffa.n1;
    pr->nmax_field = nmaxf;
    pr->smem = slb_program_smem_bytes(n, nmaxf);
    pr->use_cg = env_ll("SLB_PROGRAM_CGSYNC", 0) != 0;
    pr->bar_done = 0;
    pr->bar_ctr = nullptr;
    e = cudaMalloc(&pr->bar_ctr, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(pr->bar_ctr, 0, sizeof(unsigned));
    if (e != cudaSuccess || pr->smem > 160 * 1024) {
        cudaGetLastError();
        slb_program_destroy(pr);
        return fail(e != cudaSuccess ? SLB_E_CUDA : SLB_E_UNSUPPORTED, "slb_program_end: %s", e != cudaSuccess ? cudaGetErrorString(e) : "too many recorded ops for one program");
    }
    long long nb = env_ll("SLB_PROGRAM_BLOCKS", 40);   // a few more than the 32 virtual blocks of a 128-line sweep: the rest reduce / idle
    if (nb > c->sm_count) nb = c->sm_count;  // cooperative launch: every block resident (one per SM is always possible)
    if (nb < 1) nb = 1;
    pr->nblocks = (int)nb;
    *out = pr;
    return SLB_OK;
}

extern "C" int slb_program_launch(slb_program* p, int nrep, int64_t out_stride)
{
    if (!p || nrep < 1 || out_stride < 0) return fail(SLB_E_ARG, "slb_program_launch: bad argument");
    if ((long long)nrep * p->nbarriers * p->nblocks >= (1LL << 30))   // the barrier counter is compared modulo 2^32
        return fail(SLB_E_ARG, "slb_program_launch: nrep=%d is too large for one launch (%d barriers x %d blocks per repetition)", nrep,
                    p->nbarriers, p->nblocks);
    slb_ctx* c = p->ctx;
    NOT_RECORDABLE(c, "slb_program_launch");
    CUDA_TRY(cudaSetDevice(c->device));
    const int rc = slb_program_run(p->ops_dev, p->nops, nrep, (long long)out_stride, p->nmax_field, p->bar_ctr, p->bar_done, p->use_cg,
                                   nullptr, p->nblocks, p->smem, c->stream);
    if (rc != 0) return fail(SLB_E_CUDA, "slb_program_launch: %s", cudaGetErrorString((cudaError_t)rc));
    if (!p->use_cg) p->bar_done += (unsigned)p->nbarriers * (unsigned)nrep * (unsigned)p->nblocks;
    c->launches++;
    return SLB_OK;
}

// One more launch (nrep repetitions, advancing the data like slb_program_launch) with time stamps taken by block 0 in
// the LAST repetition: per op, the nanoseconds block 0 waited at the barrier before it and the nanoseconds the op took
// in block 0; kind_out[k] = the op's kind (1 sweep, 2 charge partial sums, 3 field solve, 4 sum of squares).
extern "C" int slb_program_profile(slb_program* p, int nrep, int64_t out_stride, int cap, int* kind_out, double* wait_ns_out, double* run_ns_out)
{
    if (!p || nrep < 1 || !kind_out || !wait_ns_out || !run_ns_out || cap < p->nops) return fail(SLB_E_ARG, "slb_program_profile: bad argument");
    slb_ctx* c = p->ctx;
    NOT_RECORDABLE(c, "slb_program_profile");
    CUDA_TRY(cudaSetDevice(c->device));
    unsigned long long* prof = nullptr;
    CUDA_TRY(cudaMalloc(&prof, (size_t)3 * p->nops * sizeof(unsigned long long)));
    const int rc = slb_program_run(p->ops_dev, p->nops, nrep, (long long)out_stride, p->nmax_field, p->bar_ctr, p->bar_done, p->use_cg, prof,
                                   p->nblocks, p->smem, c->stream);
    if (rc != 0) {
        cudaFree(prof);
        return fail(SLB_E_CUDA, "slb_program_profile: %s", cudaGetErrorString((cudaError_t)rc));
    }
    if (!p->use_cg) p->bar_done += (unsigned)p->nbarriers * (unsigned)nrep * (unsigned)p->nblocks;
    c->launches++;
    std::vector<unsigned long long> h((size_t)3 * p->nops);
    std::vector<ProgOp> ops((size_t)p->nops);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(h.data(), prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(ops.data(), p->ops_dev, ops.size() * sizeof(ProgOp), cudaMemcpyDeviceToHost);
    cudaFree(prof);
    if (e != cudaSuccess) return fail(SLB_E_CUDA, "slb_program_profile: %s", cudaGetErrorString(e));
    for (int k = 0; k < p->nops; ++k) {
        kind_out[k] = ops[k].kind;
        wait_ns_out[k] = (double)(h[3 * k + 1] - h[3 * k]);
        run_ns_out[k] = (double)(h[3 * k + 2] - h[3 * k + 1]);
    }
    return SLB_OK;
}

extern "C" int slb_program_info(const slb_program* p, int* nops, int* nbarriers, int* nblocks)
{
    if (!p) return fail(SLB_E_ARG, "slb_program_info: program is NULL");
    if (nops) *nops = p->nops;
    if (nbarriers) *nbarriers = p->nbarriers;
    if (nblocks) *nblocks = p->nblocks;
    return SLB_OK;
}

extern "C" void slb_program_destroy(slb_program* p)
{
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->ops_dev) cudaFree(p->ops_dev);
    if (p->bar_ctr) cudaFree(p->bar_ctr);
    for (void* q : p->owned) cudaFree(q);
    delete p;
}

extern "C" int slb_reduce_sum(slb_ctx* c, const double* dev, int64_t n, double* host_out) { return reduce_to_host(c, dev, n, 0, host_out); }

extern "C" int slb_subtract_mean(slb_ctx* c, double* dev, int64_t n)
{
    if (!c || !dev || n < 1) return fail(SLB_E_ARG, "slb_subtract_mean: bad argument");
    NOT_RECORDABLE(c, "slb_subtract_mean");
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = reduce_to_dev(c, dev, n, 0, 1.0 / (double)n, c->red_out + 1);
    if (rc) return rc;
    k_sub_scalar<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(dev, n, c->red_out + 1);
    LAUNCH_CHECK(c);
    return SLB_OK;
}

static int charge_from(slb_ctx* c, const double* f, long long ns, long long nv, double dv, double* rho_dev);

extern "C" int slb_charge_density_raw(slb_grid* g, int nsp, double dv, double* rho_dev)
{
    if (!g || !rho_dev) return fail(SLB_E_ARG, "slb_charge_density: NULL argument");
    if (nsp < 1 || nsp >= g->nd) return fail(SLB_E_ARG, "slb_charge_density: nsp=%d must be in [1,%d)", nsp, g->nd);
    long long ns = 1, nv = 1;
    for (int d = 0; d < nsp; ++d) ns *= g->ext[d];
    for (int d = nsp; d < g->nd; ++d) nv *= g->ext[d];
    return charge_from(g->ctx, g->front, ns, nv, dv, rho_dev);
}

extern "C" int slb_charge_density_from(slb_ctx* c, const double* f_dev, int64_t nsp_total, int64_t nv_total, double dv,
                                       double* rho_dev, int subtract_mean)
{
    if (!c || !f_dev || !rho_dev || nsp_total < 1 || nv_total < 1) return fail(SLB_E_ARG, "slb_charge_density_from: bad argument");
    int rc = charge_from(c, f_dev, nsp_total, nv_total, dv, rho_dev);
    if (rc || !subtract_mean) return rc;
    return slb_subtract_mean(c, rho_dev, nsp_total);
}

static int charge_from(slb_ctx* c, const double* fsrc, long long ns, long long nv, double dv, double* rho_dev)
{
    NOT_RECORDABLE(c, "slb_charge_density");
    CUDA_TRY(cudaSetDevice(c->device));
    long long xt = (ns + 31) / 32;
    // enough blocks to fill the machine a few times over, at least 64 velocity rows per block
    long long want = (long long)c->sm_count * 16;
    long long nchunk = (want + xt - 1) / xt;
    long long maxchunk = (nv + 63) / 64;
    if (nchunk > maxchunk) nchunk = maxchunk;
    if (nchunk < 1) nchunk = 1;
    if (nchunk > 65535) nchunk = 65535;
    long long chunk = (nv + nchunk - 1) / nchunk;
    nchunk = (nv + chunk - 1) / chunk;
    int rc = ensure_scratch(c, (size_t)(nchunk * ns) * sizeof(double));
    if (rc) return rc;
    double* partial = (double*)c->scratch;
    dim3 grid((unsigned)xt, (unsigned)nchunk), block(32, 8);
    k_charge_partial<<<grid, block, 0, c->stream>>>(fsrc, ns, nv, chunk, partial);
    LAUNCH_CHECK(c);
    k_charge_final<<<(unsigned)((ns + 255) / 256), 256, 0, c->stream>>>(partial, ns, (int)nchunk, dv, rho_dev);
    LAUNCH_CHECK(c);
    return SLB_OK;
}

extern "C" int slb_charge_density(slb_grid* g, int nsp, double dv, double* rho_dev)
{
    int rc = slb_charge_density_raw(g, nsp, dv, rho_dev);
    if (rc) return rc;
    long long ns = 1;
    for (int d = 0; d < nsp; ++d) ns *= g->ext[d];
    return slb_subtract_mean(g->ctx, rho_dev, ns);
}

extern "C" int slb_kinetic_energy(slb_grid* g, int nsp, const double* vsq_dev, double scale, double* host_out)
{
    if (!g || !vsq_dev || !host_out) return fail(SLB_E_ARG, "slb_kinetic_energy: NULL argument");
    if (nsp < 1 || nsp >= g->nd) return fail(SLB_E_ARG, "slb_kinetic_energy: nsp out of range");
    slb_ctx* c = g->ctx;
    NOT_RECORDABLE(c, "slb_kinetic_energy");
    CUDA_TRY(cudaSetDevice(c->device));
    long long ns = 1, nv = 1;
    for (int d = 0; d < nsp; ++d) ns *= g->ext[d];
    for (int d = nsp; d < g->nd; ++d) nv *= g->ext[d];
    if (nv > 0x7fffffffLL) return fail(SLB_E_UNSUPPORTED, "slb_kinetic_energy: velocity grid too large");
    int rc = ensure_scratch(c, (size_t)nv * sizeof(double));
    if (rc) return rc;
    double* partial = (double*)c->scratch;
    k_ke_partial<<<(unsigned)nv, 256, 0, c->stream>>>(g->front, ns, vsq_dev, partial);
    LAUNCH_CHECK(c);
    rc = reduce_to_dev(c, partial, nv, 0, scale, c->red_out);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->host_out, c->red_out, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *host_out = c->host_out[0];
    return SLB_OK;
}

// ------------------------------------------------------------------------------------------
// Poisson
// ------------------------------------------------------------------------------------------
extern "C" void slb_poisson_destroy(slb_poisson* p)
{
    if (!p) return;
    cudaStreamSynchronize(p->ctx->stream);
    for (int d = 0; d < SLB_MAX_DIMS; ++d) {
        if (p->tw[d]) cudaFree(p->tw[d]);
        if (p->mult[d]) cudaFree(p->mult[d]);
    }
    if (p->wa) cudaFree(p->wa);
    if (p->wb) cudaFree(p->wb);
    if (p->wc) cudaFree(p->wc);
    for (int d = 0; d < 2 * SLB_FIELD_MAXDIM; ++d)
        if (p->wx[d]) cudaFree(p->wx[d]);
    if (p->red) cudaFree(p->red);
    if (p->fft_mean) cudaFree(p->fft_mean);
    delete p;
}

extern "C" int slb_poisson_create(slb_ctx* c, int nsp, const int64_t* ext, const double* const* fctv_imag, slb_poisson** out)
{
    if (!c || !ext || !fctv_imag || !out) return fail(SLB_E_ARG, "slb_poisson_create: NULL argument");
    *out = nullptr;
    if (nsp < 1 || nsp > 3) return fail(SLB_E_ARG, "slb_poisson_create: nsp=%d not in [1,3]", nsp);
    CUDA_TRY(cudaSetDevice(c->device));
    slb_poisson* p = new slb_poisson();
    memset(p, 0, sizeof(*p));
    p->ctx = c;
    p->nsp = nsp;
    p->ntot = 1;
    for (int d = 0; d < nsp; ++d) {
        if (ext[d] < 1 || ext[d] > 1024) {  // k_dft_line stages line + twiddles in 48 KB of shared memory
            delete p;
            return fail(SLB_E_ARG, "slb_poisson_create: extent[%d] invalid", d);
        }
        p->ext[d] = ext[d];
        p->ntot *= ext[d];
    }
    cudaError_t e = cudaSuccess;
    const long double PI2 = 6.283185307179586476925286766559005768L;
    for (int d = 0; d < nsp && e == cudaSuccess; ++d) {
        int n = (int)ext[d];
        std::vector<double2> tw(n);
        for (int m = 0; m < n; ++m) {  // tw[m] = exp(-2 pi i m / n)
            long double ang = PI2 * (long double)m / (long double)n;
            tw[m] = make_double2((double)cosl(ang), (double)(-sinl(ang)));
        }
        e = cudaMalloc(&p->tw[d], n * sizeof(double2));
        if (e == cudaSuccess) e = cudaMemcpy(p->tw[d], tw.data(), n * sizeof(double2), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&p->mult[d], p->ntot * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(p->mult[d], fctv_imag[d], p->ntot * sizeof(double), cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMalloc(&p->wa, p->ntot * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&p->wb, p->ntot * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&p->wc, p->ntot * sizeof(double2));
    for (int d = 0; d < 2 * nsp && e == cudaSuccess; ++d) e = cudaMalloc(&p->wx[d], p->ntot * sizeof(double2));
    if (e == cudaSuccess) {
        // one-kernel field solve: a cooperative grid, at most one block per line and per resident slot
        int coop = 0, maxb = 0, nmax = 1;
        for (int d = 0; d < nsp; ++d) nmax = ext[d] > nmax ? (int)ext[d] : nmax;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device);
        size_t smem = 2 * (size_t)nmax * sizeof(double2);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&maxb, k_field_solve, 128, smem) == cudaSuccess && maxb > 0) {
            long long lines = p->ntot / nmax * nsp;
            long long cap = (long long)maxb * c->sm_count;
            p->coop_blocks = (int)(lines < cap ? lines : cap);
            if (p->coop_blocks < 1) p->coop_blocks = 1;
            e = cudaMalloc(&p->red, (size_t)p->coop_blocks * sizeof(double));
        }
        cudaGetLastError();
        // cluster FFT kernel: power-of-two extents, one or two space dims
        auto lg = [](int64_t v) { int l = 0; while (((int64_t)1 << l) < v) ++l; return (((int64_t)1 << l) == v) ? l : -1; };
        const int la = lg(ext[0]), lb = nsp == 2 ? lg(ext[1]) : 0;
        const int64_t lim = nsp == 2 ? SLB_FFT_NMAX : 4096;
        if (nsp <= 2 && la >= 1 && lb >= 0 && ext[0] <= lim && (nsp == 1 || (ext[1] <= lim && lb >= 1 && coop))) {
            const int nth = nsp == 2 ? SLB_FFT_THREADS : 32;
            const int n2 = nsp == 2 ? (int)ext[1] : 1;
            const size_t fsm = ((size_t)ext[0] + n2 + (size_t)2 * (nth / 32) * nmax) * sizeof(double2);
            if (cudaFuncSetAttribute(k_field_fft, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm) == cudaSuccess &&
                cudaMalloc(&p->fft_mean, sizeof(double)) == cudaSuccess) {
                p->fft_ok = 1;
                p->fft_log[0] = la;
                p->fft_log[1] = lb;
            }
            cudaGetLastError();
        }
    }
    if (e != cudaSuccess) {
        slb_poisson_destroy(p);
        cudaGetLastError();
        return fail(SLB_E_CUDA, "slb_poisson_create: %s", cudaGetErrorString(e));
    }
    *out = p;
    return SLB_OK;
}

// one DFT pass over dim d of the space grid
template <bool REAL_IN, bool INVERSE, bool MULT, bool REAL_OUT>
static void dft_pass(slb_poisson* p, int d, const void* in, void* out, const double* mult)
{
    long long inner = 1;
    for (int q = 0; q < d; ++q) inner *= p->ext[q];
    int n = (int)p->ext[d];
    unsigned nlines = (unsigned)(p->ntot / n);
    int threads = n >= 256 ? 256 : (n >= 128 ? 128 : 64);
    size_t smem = 2 * (size_t)n * sizeof(double2);
    k_dft_line<REAL_IN, INVERSE, MULT, REAL_OUT><<<nlines, threads, smem, p->ctx->stream>>>(in, out, inner, n, p->tw[d], mult);
}

extern "C" int slb_poisson_solve(slb_poisson* p, const double* rho_dev, double* const* E_dev)
{
    if (!p || !rho_dev || !E_dev) return fail(SLB_E_ARG, "slb_poisson_solve: NULL argument");
    slb_ctx* c = p->ctx;
    NOT_RECORDABLE(c, "slb_poisson_solve");
    CUDA_TRY(cudaSetDevice(c->device));
    const int nsp = p->nsp;
    for (int x = 0; x < nsp; ++x)
        if (!E_dev[x]) return fail(SLB_E_ARG, "slb_poisson_solve: E_dev[%d] is NULL", x);
    // forward transform over every space dim: rho -> spectrum
    double2* cur = p->wa;
    double2* nxt = p->wb;
    dft_pass<true, false, false, false>(p, 0, rho_dev, cur, nullptr);
    LAUNCH_CHECK(c);
    for (int d = 1; d < nsp; ++d) {
        dft_pass<false, false, false, false>(p, d, cur, nxt, nullptr);
        LAUNCH_CHECK(c);
        double2* t = cur; cur = nxt; nxt = t;
    }
    double2* spec = cur;
    double2* w1 = nxt;
    double2* w2 = p->wc;
    // per component: E_x = real(ifft(i * mult_x .* spectrum)); the multiplier is applied while the
    // first inverse pass loads its input, the last pass stores the real part only
    for (int x = 0; x < nsp; ++x) {
        if (nsp == 1) {
            dft_pass<false, true, true, true>(p, 0, spec, E_dev[x], p->mult[x]);
            LAUNCH_CHECK(c);
            continue;
        }
        dft_pass<false, true, true, false>(p, 0, spec, w1, p->mult[x]);
        LAUNCH_CHECK(c);
        double2 *a = w1, *b = w2;
        for (int d = 1; d < nsp - 1; ++d) {
            dft_pass<false, true, false, false>(p, d, a, b, nullptr);
            LAUNCH_CHECK(c);
            double2* t = a; a = b; b = t;
        }
        dft_pass<false, true, false, true>(p, nsp - 1, a, E_dev[x], nullptr);
        LAUNCH_CHECK(c);
    }
    return SLB_OK;
}

// ------------------------------------------------------------------------------------------
// one-kernel field solve
// ------------------------------------------------------------------------------------------
static int field_from_partial(slb_poisson* p, const double* partial, int nchunk, double scale, int subtract_mean, double* rho_dev,
                              double* const* E_dev)
{
    slb_ctx* c = p->ctx;
    NOT_RECORDABLE(c, "slb_poisson_solve");
    if (p->fft_ok && env_ll("SLB_FIELD_FFT", 1) != 0) {
        FieldFftArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.partial = partial;
        fa.nchunk = nchunk;
        fa.scale = scale;
        fa.subtract_mean = subtract_mean;
        fa.nsp = p->nsp;
        fa.n1 = (int)p->ext[0];
        fa.n2 = p->nsp == 2 ? (int)p->ext[1] : 1;
        fa.l1 = p->fft_log[0];
        fa.l2 = p->fft_log[1];
        fa.tw1 = p->tw[0];
        fa.tw2 = p->nsp == 2 ? p->tw[1] : p->tw[0];
        for (int d = 0; d < p->nsp; ++d) {
            fa.mult[d] = p->mult[d];
            fa.E[d] = E_dev[d];
            fa.wc[d] = p->wx[2 * d];
        }
        fa.rho = rho_dev;
        fa.wa = p->wa;
        fa.mean = p->fft_mean;
        const int nmax = fa.n1 > fa.n2 ? fa.n1 : fa.n2;
        const int nth = p->nsp == 2 ? SLB_FFT_THREADS : 32;
        const size_t fsm = ((size_t)fa.n1 + fa.n2 + (size_t)2 * (nth / 32) * nmax) * sizeof(double2);
        void* kargs[] = {&fa};
        if (p->nsp == 2)
            CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_field_fft, dim3(SLB_FFT_BLOCKS), dim3(nth), kargs, fsm, c->stream));
        else
            k_field_fft<<<1, nth, fsm, c->stream>>>(fa);
        c->launches++;
        return SLB_OK;
    }
    if (!p->coop_blocks) return SLB_E_UNSUPPORTED;
    FieldArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.partial = partial;
    fa.nchunk = nchunk;
    fa.scale = scale;
    fa.subtract_mean = subtract_mean;
    fa.nsp = p->nsp;
    fa.ntot = p->ntot;
    int nmax = 1;
    for (int d = 0; d < p->nsp; ++d) {
        fa.ext[d] = (int)p->ext[d];
        nmax = fa.ext[d] > nmax ? fa.ext[d] : nmax;
        fa.tw[d] = p->tw[d];
        fa.mult[d] = p->mult[d];
        fa.E[d] = E_dev[d];
        fa.wc[d] = p->wx[2 * d];
        fa.wd[d] = p->wx[2 * d + 1];
    }
    fa.rho = rho_dev;
    fa.wa = p->wa;
    fa.wb = p->wb;
    fa.red = p->red;
    void* args[] = {&fa};
    size_t smem = 2 * (size_t)nmax * sizeof(double2);
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_field_solve, dim3((unsigned)p->coop_blocks), dim3(128), args, smem, c->stream));
    c->launches++;
    return SLB_OK;
}

extern "C" int slb_poisson_solve_raw(slb_poisson* p, double* rho_dev, int subtract_mean, double* const* E_dev)
{
    if (!p || !rho_dev || !E_dev) return fail(SLB_E_ARG, "slb_poisson_solve_raw: NULL argument");
    for (int x = 0; x < p->nsp; ++x)
        if (!E_dev[x]) return fail(SLB_E_ARG, "slb_poisson_solve_raw: E_dev[%d] is NULL", x);
    CUDA_TRY(cudaSetDevice(p->ctx->device));
    int rc = field_from_partial(p, rho_dev, 1, 1.0, subtract_mean, rho_dev, E_dev);
    if (rc != SLB_E_UNSUPPORTED) return rc;
    if (subtract_mean) {
        rc = slb_subtract_mean(p->ctx, rho_dev, p->ntot);
        if (rc) return rc;
    }
    return slb_poisson_solve(p, rho_dev, E_dev);
}

extern "C" int slb_poisson_solve_partial(slb_poisson* p, const double* partial_dev, int nparts, double scale, int subtract_mean,
                                         double* rho_dev, double* const* E_dev)
{
    if (!p || !partial_dev || !rho_dev || !E_dev || nparts < 1) return fail(SLB_E_ARG, "slb_poisson_solve_partial: bad argument");
    for (int x = 0; x < p->nsp; ++x)
        if (!E_dev[x]) return fail(SLB_E_ARG, "slb_poisson_solve_partial: E_dev[%d] is NULL", x);
    slb_ctx* c = p->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    int rc = field_from_partial(p, partial_dev, nparts, scale, subtract_mean, rho_dev, E_dev);
    if (rc != SLB_E_UNSUPPORTED) return rc;
    // no cooperative launch: sum the parts with the charge kernel's second stage, then the separate DFT passes
    k_charge_final<<<(unsigned)((p->ntot + 255) / 256), 256, 0, c->stream>>>(partial_dev, p->ntot, nparts, scale, rho_dev);
    LAUNCH_CHECK(c);
    if (subtract_mean) {
        rc = slb_subtract_mean(c, rho_dev, p->ntot);
        if (rc) return rc;
    }
    return slb_poisson_solve(p, rho_dev, E_dev);
}

extern "C" int slb_vp_field_solve(slb_poisson* p, const double* f_dev, int64_t nv_total, double dv, double* rho_dev,
                                  double* const* E_dev)
{
    if (!p || !f_dev || !rho_dev || !E_dev || nv_total < 1) return fail(SLB_E_ARG, "slb_vp_field_solve: bad argument");
    for (int x = 0; x < p->nsp; ++x)
        if (!E_dev[x]) return fail(SLB_E_ARG, "slb_vp_field_solve: E_dev[%d] is NULL", x);
    slb_ctx* c = p->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    const long long ns = p->ntot, nv = nv_total;
    if (c->prog_rec) {
        // step program: the same two stages (K4 partial sums, then the one-warp 1-D field solve) as two ops
        slb_prog_rec* r = c->prog_rec;
        if (p->nsp != 1 || !p->fft_ok || ns > 1024 || env_ll("SLB_FIELD_FFT", 1) == 0) {
            r->failed = true;
            return fail(SLB_E_UNSUPPORTED, "slb_vp_field_solve: step programs record field solves over ONE power-of-two space dim (<= 1024 points)");
        }
        long long xt = (ns + 31) / 32;
        long long want = (long long)c->sm_count * 16;
        long long nchunk = (want + xt - 1) / xt;
        long long maxchunk = (nv + 63) / 64;
        if (nchunk > maxchunk) nchunk = maxchunk;
        if (nchunk < 1) nchunk = 1;
        if (nchunk > 65535) nchunk = 65535;
        long long chunk = (nv + nchunk - 1) / nchunk;
        nchunk = (nv + chunk - 1) / chunk;
        const double* partial = f_dev;   // nv == 1 (line sums of a velocity sweep): K4's partial sums ARE the input (0 + x = x)
        if (nv > 1) {
            double* pbuf = nullptr;
            CUDA_TRY(cudaMalloc(&pbuf, (size_t)(nchunk * ns) * sizeof(double)));
            r->owned.push_back(pbuf);
            partial = pbuf;
            ProgOpHost h;
            memset(&h.op, 0, sizeof(h.op));
            h.op.kind = SLB_OP_CHARGE;
            h.op.f = f_dev;
            h.op.ns = ns;
            h.op.nv = nv;
            h.op.chunk = chunk;
            h.op.nchunk = (int)nchunk;
            h.op.partial = pbuf;
            h.reads.push_back(prog_range(f_dev, (size_t)(ns * nv) * sizeof(double)));
            h.writes.push_back(prog_range(pbuf, (size_t)(nchunk * ns) * sizeof(double)));
            r->ops.push_back(h);
        }
        ProgOpHost q;
        memset(&q.op, 0, sizeof(q.op));
        q.op.kind = SLB_OP_FIELD1D;
        FieldFftArgs& fa = q.op.ffa;
        fa.partial = partial;
        fa.nchunk = (int)nchunk;
        fa.scale = dv;
        fa.subtract_mean = 1;
        fa.nsp = 1;
        fa.n1 = (int)p->ext[0];
        fa.n2 = 1;
        fa.l1 = p->fft_log[0];
        fa.l2 = 0;
        fa.tw1 = p->tw[0];
        fa.tw2 = p->tw[0];
        fa.mult[0] = p->mult[0];
        fa.E[0] = E_dev[0];
        fa.rho = rho_dev;
        q.reads.push_back(prog_range(partial, (size_t)(nchunk * ns) * sizeof(double)));   // by every block (each solves for itself)
        q.writes.push_back(prog_range(rho_dev, (size_t)ns * sizeof(double), true));        // by block 0
        q.writes.push_back(prog_range(E_dev[0], (size_t)ns * sizeof(double), true));
        r->ops.push_back(q);
        r->last_E = E_dev[0];
        if (fa.n1 > r->nmax_field) r->nmax_field = fa.n1;
        return SLB_OK;
    }
    if (!p->coop_blocks) {
        int rc = slb_charge_density_from(c, f_dev, ns, nv, dv, rho_dev, 1);
        if (rc) return rc;
        return slb_poisson_solve(p, rho_dev, E_dev);
    }
    // K4 partial sums (the pass over f or over the line sums), then everything else in one kernel
    long long xt = (ns + 31) / 32;
    long long want = (long long)c->sm_count * 16;
    long long nchunk = (want + xt - 1) / xt;
    long long maxchunk = (nv + 63) / 64;
    if (nchunk > maxchunk) nchunk = maxchunk;
    if (nchunk < 1) nchunk = 1;
    if (nchunk > 65535) nchunk = 65535;
    long long chunk = (nv + nchunk - 1) / nchunk;
    nchunk = (nv + chunk - 1) / chunk;
    int rc = ensure_scratch(c, (size_t)(nchunk * ns) * sizeof(double));
    if (rc) return rc;
    double* partial = (double*)c->scratch;
    dim3 grid((unsigned)xt, (unsigned)nchunk), block(32, 8);
    k_charge_partial<<<grid, block, 0, c->stream>>>(f_dev, ns, nv, chunk, partial);
    LAUNCH_CHECK(c);
    return field_from_partial(p, partial, (int)nchunk, dv, 1, rho_dev, E_dev);
}

// ------------------------------------------------------------------------------------------
// N-D per-point interpolation (unsplit 2-D solvers) and the small array
// operations of the Adams-Bashforth time algorithms
// ------------------------------------------------------------------------------------------
extern "C" int slb_memcpy_d2d(slb_ctx* c, void* dst, const void* src, int64_t bytes)
{
    if (!c || !dst || !src || bytes < 0) return fail(SLB_E_ARG, "slb_memcpy_d2d: bad argument");
    NOT_RECORDABLE(c, "slb_memcpy_d2d");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, c->stream));
    return SLB_OK;
}

template <bool EXACT>
static void launch_points(slb_ctx* c, const PointsArgs& pa, const slb_interp* itA, const slb_interp* itB)
{
    dim3 grid((unsigned)((pa.n1 + 127) / 128), (unsigned)pa.n2);
    if (itA->fast && itB->fast && itA->order == itB->order) {
        switch (pa.pA) {
#define X(P) case P: k_interp2d_points<P, EXACT><<<grid, 128, 0, c->stream>>>(pa, itA->tab, itB->tab); return;
            SLB_FOR_P1(X)
#undef X
        }
    }
    k_interp2d_points_generic<EXACT><<<grid, 128, 0, c->stream>>>(pa, itA->coef_dev, itB->coef_dev);
}

extern "C" int slb_interp2d_points(slb_ctx* c, const slb_interp* it1, const slb_interp* it2, int64_t n1, int64_t n2,
                                   int ncomp, double* in_dev, const double* dec_dev, double* out_dev, double* work_dev,
                                   int flags)
{
    if (!c || !it1 || !it2 || !in_dev || !dec_dev || !out_dev) return fail(SLB_E_ARG, "slb_interp2d_points: NULL argument");
    NOT_RECORDABLE(c, "slb_interp2d_points");
    if (n1 < 1 || n2 < 1 || n1 > 0x7fffffffLL || n2 > 65535 || ncomp < 1 || ncomp > 16)
        return fail(SLB_E_ARG, "slb_interp2d_points: extents (%lld, %lld) x %d components out of range", (long long)n1, (long long)n2, ncomp);
    if (in_dev == out_dev) return fail(SLB_E_ARG, "slb_interp2d_points: fp and fi must not alias");
    if (it1->order + 1 > SLB_POINTS_MAXP1 || it2->order + 1 > SLB_POINTS_MAXP1)
        return fail(SLB_E_ARG, "slb_interp2d_points: order + 1 must not exceed %d", SLB_POINTS_MAXP1);
    const slb_interp* its[2] = {it1, it2};
    const int64_t ns[2] = {n1, n2};
    bool bs[2];
    for (int x = 0; x < 2; ++x) {
        bs[x] = (its[x]->kind == SLB_BSPLINE_LU || its[x]->kind == SLB_BSPLINE_FFT);
        if (bs[x] && its[x]->n != ns[x])
            return fail(SLB_E_ARG, "slb_interp2d_points: B-spline object %d built for n=%lld, extent is %lld", x + 1,
                        (long long)its[x]->n, (long long)ns[x]);
    }
    if ((bs[0] || bs[1]) && (!work_dev || work_dev == in_dev || work_dev == out_dev))
        return fail(SLB_E_ARG, "slb_interp2d_points: B-spline interpolations need a distinct work buffer");
    CUDA_TRY(cudaSetDevice(c->device));
    // res = sol(interp_t, fi): the 1-D solves along each dim, src/interpolation.jl:48-94
    const double* res = in_dev;
    if (bs[0]) {
        int rc = bspline_presolve(c->stream, &it1->bsp, res, work_dev, 1, (int)n1, n2 * ncomp, &c->launches);
        if (rc) return fail(rc, "slb_interp2d_points: pre-solve along dim 1 failed: %s", cudaGetErrorString(cudaGetLastError()));
        res = work_dev;
    }
    if (bs[1]) {
        double* dst = (res == in_dev) ? work_dev : in_dev;
        int rc = bspline_presolve(c->stream, &it2->bsp, res, dst, n1, (int)n2, ncomp, &c->launches);
        if (rc) return fail(rc, "slb_interp2d_points: pre-solve along dim 2 failed: %s", cudaGetErrorString(cudaGetLastError()));
        res = dst;
    }
    PointsArgs pa;
    pa.res = res;
    pa.dec = dec_dev;
    pa.out = out_dev;
    pa.n1 = (int)n1;
    pa.n2 = (int)n2;
    pa.ncomp = ncomp;
    pa.pA = it1->order + 1;
    pa.pB = it2->order + 1;
    pa.ncA = it1->nc;
    pa.ncB = it2->nc;
    if (flags & SLB_SWEEP_EXACT)
        launch_points<true>(c, pa, it1, it2);
    else
        launch_points<false>(c, pa, it1, it2);
    LAUNCH_CHECK(c);
    return SLB_OK;
}

extern "C" int slb_fill_dec2d(slb_ctx* c, double* dec_dev, int64_t n1, int64_t n2, const double* tab_j_dev, double scale_j,
                              const double* tab_i_dev, double scale_i)
{
    if (!c || !dec_dev || !tab_j_dev || !tab_i_dev) return fail(SLB_E_ARG, "slb_fill_dec2d: NULL argument");
    NOT_RECORDABLE(c, "slb_fill_dec2d");
    if (n1 < 1 || n2 < 1 || n1 > 0x7fffffffLL || n2 > 65535) return fail(SLB_E_ARG, "slb_fill_dec2d: extents out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    dim3 grid((unsigned)((n1 + 127) / 128), (unsigned)n2);
    k_fill_dec2d<<<grid, 128, 0, c->stream>>>(dec_dev, (int)n1, (int)n2, tab_j_dev, scale_j, tab_i_dev, scale_i);
    LAUNCH_CHECK(c);
    return SLB_OK;
}

extern "C" int slb_lincomb(slb_ctx* c, double* out_dev, int nterms, const double* coefs, const double* const* x_dev, int64_t n)
{
    if (!c || !out_dev || !coefs || !x_dev || n < 1) return fail(SLB_E_ARG, "slb_lincomb: bad argument");
    NOT_RECORDABLE(c, "slb_lincomb");
    if (nterms < 1 || nterms > SLB_LINCOMB_MAX) return fail(SLB_E_ARG, "slb_lincomb: nterms=%d not in [1,%d]", nterms, SLB_LINCOMB_MAX);
    LincombArgs la;
    memset(&la, 0, sizeof(la));
    la.nterms = nterms;
    for (int k = 0; k < nterms; ++k) {
        if (!x_dev[k]) return fail(SLB_E_ARG, "slb_lincomb: x_dev[%d] is NULL", k);
        la.coef[k] = coefs[k];
        la.x[k] = x_dev[k];
    }
    CUDA_TRY(cudaSetDevice(c->device));
    unsigned blocks = (unsigned)((n + 255) / 256);
    k_lincomb<<<blocks, 256, 0, c->stream>>>(out_dev, n, la);
    LAUNCH_CHECK(c);
    return SLB_OK;
}
