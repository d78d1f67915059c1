// slb_comm.cu -- rank-to-rank plumbing of the sharded drivers (include/slb200.h, slb_comm_*): one process per
// GPU, a device "mailbox" per rank that the peers map (CUDA IPC across processes, plain peer access inside one
// process), and two stream-ordered collectives built from NVLink peer stores + flags: a barrier and an
// all-gather of small arrays (charge densities).  Replaces the reference's MPI layer (src/mpiinterface.jl:1-38:
// MPI.Init / Comm_rank / Bcast) for everything that is not the f array itself -- f moves inside the sweep
// kernels (slb_pair.cuh: re-shard stores, halo pushes).  The host language only carries the opaque handles
// once at start-up.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <vector>

#include "slb_internal.h"

#define SLB_COMM_MAXP 16
#define SLB_COMM_MAGIC 0x534c4243u  // "SLBC"

typedef unsigned long long ull;

struct CommHandle {  // 128 bytes on the wire
    uint32_t magic;
    int32_t pid;
    int32_t device;
    int32_t pad;
    uint64_t ptr;     // the exporting process's device pointer
    uint64_t offset;  // ptr - base of its allocation (cudaMalloc may sub-allocate; IPC handles name allocations)
    cudaIpcMemHandle_t ipc;
    char fill[128 - 4 * 4 - 2 * 8 - 64];
};
static_assert(sizeof(CommHandle) == SLB_COMM_HANDLE_BYTES, "handle size");

#define SLB_COMM_NSIG 4
struct CommDev {  // kernel argument
    int rank, P;
    ull* bar[SLB_COMM_MAXP];     // bar[q]: the barrier flags in rank q's mailbox (bar[rank]: own)
    ull* gat[SLB_COMM_MAXP];     // all-gather flags
    ull* sig[SLB_COMM_MAXP];     // point-to-point signals: [SLB_COMM_NSIG][MAXP]
    double* slots[SLB_COMM_MAXP];
};

struct Opened {
    void* base;    // what cudaIpcOpenMemHandle returned (NULL: same-process pointer, nothing to close)
    void* ptr;     // what the caller got
};

struct slb_comm {
    slb_ctx* ctx;
    int rank, P;
    int64_t nslot;
    char* mbox;
    size_t mbox_bytes;
    CommDev dev;
    bool connected;
    ull epoch_bar, epoch_gat;
    ull sent[SLB_COMM_MAXP][SLB_COMM_NSIG], awaited[SLB_COMM_MAXP][SLB_COMM_NSIG];
    std::vector<Opened> opened;
};

// mailbox layout: [bar flags: MAXP ull][gather flags: MAXP ull][signal flags: NSIG x MAXP ull][slots: 2 x P x nslot doubles]
static const size_t kFlagBytes = (2 + SLB_COMM_NSIG) * SLB_COMM_MAXP * sizeof(ull);

// ------------------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void comm_store_flag(ull* p, ull v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ ull comm_load_flag(const ull* p)
{
    ull v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// thread q: tell rank q that this rank has arrived (everything enqueued before on this stream is complete),
// then wait until rank q has told us the same
__global__ void k_comm_barrier(const __grid_constant__ CommDev cd, ull epoch)
{
    const int q = threadIdx.x;
    if (q < cd.P) {
        __threadfence_system();
        comm_store_flag(cd.bar[q] + cd.rank, epoch);
        while (comm_load_flag(cd.bar[cd.rank] + q) < epoch) {
        }
    }
}

// point-to-point: "what this rank enqueued before on that stream (a halo copy into your memory) is complete"
__global__ void k_comm_signal(ull* flag, ull epoch)
{
    __threadfence_system();
    comm_store_flag(flag, epoch);
}
__global__ void k_comm_wait(const ull* flag, ull epoch)
{
    while (comm_load_flag(flag) < epoch) {
    }
}

// block q: copy this rank's n doubles into slot `rank` of rank q's mailbox, raise its flag there, then wait for
// rank q's contribution to arrive here.  When the kernel ends every slot of this rank's mailbox is filled.
__global__ void __launch_bounds__(1024) k_comm_allgather(const __grid_constant__ CommDev cd, const double* __restrict__ local, long long n,
                                                        long long slot_off, ull epoch)
{
    const int q = blockIdx.x;
    double* dst = cd.slots[q] + slot_off + (long long)cd.rank * n;
    if ((n & 1) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(local) & 15) == 0) {
        const double2* s2 = reinterpret_cast<const double2*>(local);
        double2* d2 = reinterpret_cast<double2*>(dst);
        for (long long i = threadIdx.x; i < n / 2; i += blockDim.x) d2[i] = s2[i];
    } else {
        for (long long i = threadIdx.x; i < n; i += blockDim.x) dst[i] = local[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        comm_store_flag(cd.gat[q] + cd.rank, epoch);
        while (comm_load_flag(cd.gat[cd.rank] + q) < epoch) {
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef int (*cuMemGetAddressRange_t)(unsigned long long* pbase, size_t* psize, unsigned long long dptr);

static int alloc_base(void* dev, uint64_t* offset)
{
    // base of the allocation that contains dev (driver entry point fetched through the runtime: libslb200 does
    // not link libcuda, so that it still loads on hosts without a driver)
    static cuMemGetAddressRange_t fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &qr) != cudaSuccess || !p) {
            cudaGetLastError();
            *offset = 0;
            return 0;
        }
        fn = (cuMemGetAddressRange_t)p;
    }
    unsigned long long base = 0;
    size_t size = 0;
    if (fn(&base, &size, (unsigned long long)(uintptr_t)dev) != 0) {
        *offset = 0;
        return 0;
    }
    *offset = (uint64_t)((uintptr_t)dev - (uintptr_t)base);
    return 0;
}

static int export_ptr(slb_ctx* c, void* dev, void* handle128)
{
    CommHandle h;
    memset(&h, 0, sizeof(h));
    h.magic = SLB_COMM_MAGIC;
    h.pid = (int32_t)getpid();
    h.device = c->device;
    h.ptr = (uint64_t)(uintptr_t)dev;
    alloc_base(dev, &h.offset);
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaIpcGetMemHandle(&h.ipc, (char*)dev - h.offset));
    memcpy(handle128, &h, sizeof(h));
    return SLB_OK;
}

static int open_ptr(slb_comm* cm, const void* handle128, void** out)
{
    CommHandle h;
    memcpy(&h, handle128, sizeof(h));
    if (h.magic != SLB_COMM_MAGIC) return slb_fail(SLB_E_ARG, "slb_comm: not a handle (bad magic)");
    slb_ctx* c = cm->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    Opened o;
    if (h.pid == (int32_t)getpid()) {
        // same process (ranks emulated on one GPU, or one process driving several GPUs): the pointer itself
        if (h.device != c->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return slb_fail(SLB_E_CUDA, "slb_comm: no peer access from device %d to %d: %s", c->device, h.device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        o.base = nullptr;
        o.ptr = (void*)(uintptr_t)h.ptr;
    } else {
        void* base = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&base, h.ipc, cudaIpcMemLazyEnablePeerAccess));
        o.base = base;
        o.ptr = (char*)base + h.offset;
    }
    cm->opened.push_back(o);
    *out = o.ptr;
    return SLB_OK;
}

extern "C" int slb_comm_create(slb_ctx* c, int rank, int nranks, int64_t nslot, slb_comm** out)
{
    if (!c || !out) return slb_fail(SLB_E_ARG, "slb_comm_create: NULL argument");
    *out = nullptr;
    if (nranks < 1 || nranks > SLB_COMM_MAXP || rank < 0 || rank >= nranks)
        return slb_fail(SLB_E_ARG, "slb_comm_create: rank %d of %d out of range (at most %d ranks)", rank, nranks, SLB_COMM_MAXP);
    if (nslot < 1) return slb_fail(SLB_E_ARG, "slb_comm_create: nslot_doubles must be positive");
    CUDA_TRY(cudaSetDevice(c->device));
    slb_comm* cm = new slb_comm();
    cm->ctx = c;
    cm->rank = rank;
    cm->P = nranks;
    cm->nslot = nslot;
    cm->connected = false;
    cm->epoch_bar = cm->epoch_gat = 0;
    memset(cm->sent, 0, sizeof(cm->sent));
    memset(cm->awaited, 0, sizeof(cm->awaited));
    cm->mbox_bytes = kFlagBytes + (size_t)2 * nranks * nslot * sizeof(double);
    if (cm->mbox_bytes < ((size_t)4 << 20)) cm->mbox_bytes = (size_t)4 << 20;  // its own allocation, not a sub-allocated block
    cm->mbox = nullptr;
    cudaError_t e = cudaMalloc(&cm->mbox, cm->mbox_bytes);
    if (e == cudaSuccess) e = cudaMemset(cm->mbox, 0, cm->mbox_bytes);
    if (e != cudaSuccess) {
        if (cm->mbox) cudaFree(cm->mbox);
        delete cm;
        cudaGetLastError();
        return slb_fail(SLB_E_ALLOC, "slb_comm_create: %s", cudaGetErrorString(e));
    }
    memset(&cm->dev, 0, sizeof(cm->dev));
    cm->dev.rank = rank;
    cm->dev.P = nranks;
    *out = cm;
    return SLB_OK;
}

extern "C" int slb_comm_export(slb_comm* cm, void* handle128)
{
    if (!cm || !handle128) return slb_fail(SLB_E_ARG, "slb_comm_export: NULL argument");
    return export_ptr(cm->ctx, cm->mbox, handle128);
}

static void wire(slb_comm* cm, int q, char* mbox)
{
    cm->dev.bar[q] = reinterpret_cast<ull*>(mbox);
    cm->dev.gat[q] = reinterpret_cast<ull*>(mbox) + SLB_COMM_MAXP;
    cm->dev.sig[q] = reinterpret_cast<ull*>(mbox) + 2 * SLB_COMM_MAXP;
    cm->dev.slots[q] = reinterpret_cast<double*>(mbox + kFlagBytes);
}

extern "C" int slb_comm_connect(slb_comm* cm, const void* all_handles)
{
    if (!cm || !all_handles) return slb_fail(SLB_E_ARG, "slb_comm_connect: NULL argument");
    if (cm->connected) return slb_fail(SLB_E_ARG, "slb_comm_connect: already connected");
    for (int q = 0; q < cm->P; ++q) {
        if (q == cm->rank) {
            wire(cm, q, cm->mbox);
            continue;
        }
        void* p = nullptr;
        int rc = open_ptr(cm, (const char*)all_handles + (size_t)q * SLB_COMM_HANDLE_BYTES, &p);
        if (rc) return rc;
        wire(cm, q, (char*)p);
    }
    cm->connected = true;
    return SLB_OK;
}

extern "C" void slb_comm_destroy(slb_comm* cm)
{
    if (!cm) return;
    cudaSetDevice(cm->ctx->device);
    cudaStreamSynchronize(cm->ctx->stream);
    for (auto& o : cm->opened)
        if (o.base) cudaIpcCloseMemHandle(o.base);
    cudaFree(cm->mbox);
    cudaGetLastError();
    delete cm;
}

extern "C" int slb_comm_export_buffer(slb_comm* cm, void* dev, void* handle128)
{
    if (!cm || !dev || !handle128) return slb_fail(SLB_E_ARG, "slb_comm_export_buffer: NULL argument");
    return export_ptr(cm->ctx, dev, handle128);
}

extern "C" int slb_comm_open_buffer(slb_comm* cm, const void* handle128, void** dev_out)
{
    if (!cm || !handle128 || !dev_out) return slb_fail(SLB_E_ARG, "slb_comm_open_buffer: NULL argument");
    return open_ptr(cm, handle128, dev_out);
}

extern "C" int slb_comm_close_buffer(slb_comm* cm, void* dev)
{
    if (!cm) return slb_fail(SLB_E_ARG, "slb_comm_close_buffer: comm is NULL");
    for (size_t i = 0; i < cm->opened.size(); ++i) {
        if (cm->opened[i].ptr == dev) {
            if (cm->opened[i].base) {
                CUDA_TRY(cudaSetDevice(cm->ctx->device));
                CUDA_TRY(cudaStreamSynchronize(cm->ctx->stream));
                CUDA_TRY(cudaIpcCloseMemHandle(cm->opened[i].base));
            }
            cm->opened.erase(cm->opened.begin() + (long)i);
            return SLB_OK;
        }
    }
    return slb_fail(SLB_E_ARG, "slb_comm_close_buffer: pointer was not opened through this comm");
}

extern "C" int slb_comm_barrier(slb_comm* cm)
{
    if (!cm) return slb_fail(SLB_E_ARG, "slb_comm_barrier: comm is NULL");
    if (!cm->connected) return slb_fail(SLB_E_ARG, "slb_comm_barrier: not connected");
    slb_ctx* c = cm->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    if (cm->P == 1) return SLB_OK;
    k_comm_barrier<<<1, 32, 0, c->stream>>>(cm->dev, ++cm->epoch_bar);
    LAUNCH_CHECK(c);
    return SLB_OK;
}

extern "C" int slb_comm_allgather(slb_comm* cm, const double* local_dev, int64_t n, const double** slots_out)
{
    if (!cm || !local_dev || !slots_out) return slb_fail(SLB_E_ARG, "slb_comm_allgather: NULL argument");
    if (!cm->connected) return slb_fail(SLB_E_ARG, "slb_comm_allgather: not connected");
    if (n < 1 || n > cm->nslot) return slb_fail(SLB_E_ARG, "slb_comm_allgather: n=%lld exceeds the mailbox slot (%lld doubles)", (long long)n, (long long)cm->nslot);
    slb_ctx* c = cm->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    const ull epoch = ++cm->epoch_gat;
    // two sets of slots, alternating: a rank that is one collective ahead never overwrites what a slower rank is
    // still reading (it cannot get two ahead: the next all-gather needs the slower rank's contribution)
    const long long slot_off = (long long)(epoch & 1) * cm->P * cm->nslot;
    k_comm_allgather<<<cm->P, 1024, 0, c->stream>>>(cm->dev, local_dev, n, slot_off, epoch);
    LAUNCH_CHECK(c);
    *slots_out = cm->dev.slots[cm->rank] + slot_off;
    return SLB_OK;
}

extern "C" int slb_comm_signal(slb_comm* cm, slb_ctx* on, int peer, int slot)
{
    if (!cm || !cm->connected) return slb_fail(SLB_E_ARG, "slb_comm_signal: comm missing or not connected");
    if (peer < 0 || peer >= cm->P || slot < 0 || slot >= SLB_COMM_NSIG) return slb_fail(SLB_E_ARG, "slb_comm_signal: peer / slot out of range");
    slb_ctx* c = on ? on : cm->ctx;
    if (c->device != cm->ctx->device) return slb_fail(SLB_E_ARG, "slb_comm_signal: the stream's context lives on another device");
    CUDA_TRY(cudaSetDevice(c->device));
    k_comm_signal<<<1, 1, 0, c->stream>>>(cm->dev.sig[peer] + slot * SLB_COMM_MAXP + cm->rank, ++cm->sent[peer][slot]);
    LAUNCH_CHECK(c);
    return SLB_OK;
}

extern "C" int slb_comm_wait(slb_comm* cm, slb_ctx* on, int peer, int slot)
{
    if (!cm || !cm->connected) return slb_fail(SLB_E_ARG, "slb_comm_wait: comm missing or not connected");
    if (peer < 0 || peer >= cm->P || slot < 0 || slot >= SLB_COMM_NSIG) return slb_fail(SLB_E_ARG, "slb_comm_wait: peer / slot out of range");
    slb_ctx* c = on ? on : cm->ctx;
    if (c->device != cm->ctx->device) return slb_fail(SLB_E_ARG, "slb_comm_wait: the stream's context lives on another device");
    CUDA_TRY(cudaSetDevice(c->device));
    k_comm_wait<<<1, 1, 0, c->stream>>>(cm->dev.sig[cm->rank] + slot * SLB_COMM_MAXP + peer, ++cm->awaited[peer][slot]);
    LAUNCH_CHECK(c);
    return SLB_OK;
}
