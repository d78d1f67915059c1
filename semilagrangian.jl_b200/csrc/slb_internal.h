// slb_internal.h -- objects shared between the translation units of libslb200 (not part of the ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/slb200.h"

struct slb_ctx {
    int device;
    cudaStream_t stream;
    bool own_stream;
    int64_t launches;
    cudaEvent_t ev0, ev1;
    double* red_partial;  // 1024 doubles
    double* red_out;      // 8 doubles
    double* host_out;     // pinned, 8 doubles
    void* scratch;        // growable device scratch (alpha tables, partial sums)
    size_t scratch_bytes;
    int sm_count;
    int64_t capture_launches0;  // launch counter when a stream capture began
    int* err_word;        // device word: bit 0 = a shift exceeded the halo of a sharded pass (slb_halo_error)
    struct slb_prog_rec* prog_rec;  // non-NULL while a step program is being recorded (slb_program_begin / _end)
};

int slb_fail(int code, const char* fmt, ...);

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return slb_fail(e_ == cudaErrorMemoryAllocation ? SLB_E_ALLOC : SLB_E_CUDA, "%s: %s",   \
                            #expr, cudaGetErrorString(e_));                                         \
    } while (0)

#define LAUNCH_CHECK(ctx)                                                            \
    do {                                                                             \
        (ctx)->launches++;                                                           \
        cudaError_t e_ = cudaGetLastError();                                         \
        if (e_ != cudaSuccess) return slb_fail(SLB_E_CUDA, "kernel launch: %s", cudaGetErrorString(e_)); \
    } while (0)
