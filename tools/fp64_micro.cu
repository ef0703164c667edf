// fp64_micro.cu -- B200 FP64 pipe characterisation used to model the HeatCool kernel.
// Measures, per SM: DFMA latency (1 warp, dependent chain), and DFMA throughput as a function of warps per SM x independent
// chains per thread (ILP).  build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_micro.cu -o tools/_fp64_micro
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, int iters, double m, double b, long long* cycles) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = 1.0 + threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = __fma_rn(x[i], m, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// mixed chain resembling the division sequence: MUFU.RCP64H + 7 dependent DFMA/DMUL
__global__ void divchain(double* out, int iters, double d0, long long* cycles) {
    double n = 1.0 + threadIdx.x, d = d0 + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { n = n / d; d = d + n; }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = n + d;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void logchain(double* out, int iters, double d0, long long* cycles) {
    double n = 1.0e4 + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) { n = log10(n) + d0; }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = n;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ILP>
void run(int warps, double* out, long long* dcyc) {
    const int iters = 2000;
    chain<ILP><<<148, warps * 32>>>(out, iters, 1.0 + 1e-9, 1e-9, dcyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    chain<ILP><<<148, warps * 32>>>(out, iters, 1.0 + 1e-9, 1e-9, dcyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc; cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
    const double n_inst = (double)iters * 16 * ILP;           // per thread
    const double per_sm_per_clk = n_inst * warps * 32 / (double)cyc;
    printf("warps/SM %2d ILP %d: %.2f cycles per dependent DFMA step, %.1f DFMA lanes/clk/SM (peak 64), %.2f TFLOP/s\n", warps, ILP,
           (double)cyc / (iters * 16.0), per_sm_per_clk, 2.0 * n_inst * warps * 32 * 148 / (ms * 1e-3) * 1e-12);
}

int main() {
    double* out; long long* dcyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&dcyc, 8);
    for (int warps : {1, 4, 8, 12, 16, 24, 32}) { run<1>(warps, out, dcyc); run<2>(warps, out, dcyc); run<4>(warps, out, dcyc); }
    for (int warps : {1, 4, 12, 16, 32}) {
        divchain<<<148, warps * 32>>>(out, 500, 3.0, dcyc); cudaDeviceSynchronize();
        long long cyc; cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        printf("div+add chain warps/SM %2d: %.1f cycles per (div, add) step; %.2f warp-steps per SM-cycle x1000\n", warps, (double)cyc / 2000.0, 1000.0 * warps * 2000.0 / cyc);
        logchain<<<148, warps * 32>>>(out, 500, 3.0, dcyc); cudaDeviceSynchronize();
        cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
        printf("log10+add chain warps/SM %2d: %.1f cycles per step; %.2f warp-steps per SM-cycle x1000\n", warps, (double)cyc / 2000.0, 1000.0 * warps * 2000.0 / cyc);
    }
    return 0;
}
