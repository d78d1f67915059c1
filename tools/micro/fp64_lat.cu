// FP64 latency / throughput probe for sm_100a (B200): dependent DFMA chains per warp, ILP 1..8, 1..32 warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP>
void run(int warps, double* out, long long* cyc)
{
    const int iters = 256;
    k<ILP><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    k<ILP><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double n = (double)iters * 16 * ILP;   // DFMAs per thread
    printf("ILP %d warps %2d: %8.2f cycles per DFMA per warp, %6.2f warp-DFMA/clk/SM (%5.1f lanes/clk)\n", ILP, warps, c / n, n * warps / c,
           32.0 * n * warps / c);
}

int main()
{
    double* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&cyc, 1 << 12);
    for (int w : {1, 4, 8, 16, 32}) {
        run<1>(w, out, cyc);
        run<2>(w, out, cyc);
        run<4>(w, out, cyc);
        run<8>(w, out, cyc);
    }
    return 0;
}
