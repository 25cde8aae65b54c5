// microbench.cu -- FP64 latency/throughput probes that inform the EM kernel design.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS> __global__ void k_dfma(double *out, int iters, double a, double b) {
    double x[CHAINS];
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x + c;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += x[c];
    if (s == 1.2345) out[0] = s;
}
template <int CHAINS> __global__ void k_div(double *out, int iters, double a, double b) {
    double x[CHAINS];
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x + c + 1.5;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = a / x[c] + b;
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += x[c];
    if (s == 1.2345) out[0] = s;
}
template <int CHAINS> __global__ void k_log(double *out, int iters, double a, double b) {
    double x[CHAINS];
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x + c + 1.5;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = log(x[c]) * a + b;
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += x[c];
    if (s == 1.2345) out[0] = s;
}
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
template <int CHAINS> __global__ void k_frcp(double *out, int iters, double a, double b) {
    double x[CHAINS];
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x + c + 1.5;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = a * fast_rcp(x[c]) + b;
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += x[c];
    if (s == 1.2345) out[0] = s;
}

// only the first `active` lanes of each warp do the work: does the FP64 pipe skip idle half-warps?
template <int CHAINS> __global__ void k_dfma_part(double *out, int iters, double a, double b, int active) {
    if ((threadIdx.x & 31) >= active) return;
    double x[CHAINS];
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x + c;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
    double s = 0;
    for (int c = 0; c < CHAINS; c++) s += x[c];
    if (s == 1.2345) out[0] = s;
}

template <class F> float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    double *d;
    cudaMalloc(&d, 64);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk * 1e-6;
    const int iters = 20000;
    printf("clock %.3f GHz (nominal); cycles are at nominal clock\n", ghz);
#define RUN(name, kern, chains, blocks, threads)                                                         \
    {                                                                                                    \
        float ms = timeit([&] { kern<chains><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9); });       \
        printf("%-8s chains=%d grid=%dx%d : %.3f ms  -> %.2f cycles per op per warp-chain-step, %.1f cycles/op/warp\n", \
               name, chains, blocks, threads, ms, ms * 1e-3 * ghz * 1e9 / iters, ms * 1e-3 * ghz * 1e9 / iters / chains); \
    }
    // latency: 1 warp per SM sub-partition (block of 128 = 4 warps, 1 block per SM)
    RUN("dfma", k_dfma, 1, 148, 128)
    RUN("dfma", k_dfma, 2, 148, 128)
    RUN("dfma", k_dfma, 4, 148, 128)
    RUN("dfma", k_dfma, 8, 148, 128)
    RUN("dfma", k_dfma, 1, 148, 32)
    RUN("dfma", k_dfma, 8, 148, 32)
    RUN("dfma", k_dfma, 8, 148, 1024)
    RUN("div", k_div, 1, 148, 128)
    RUN("div", k_div, 4, 148, 128)
    RUN("frcp", k_frcp, 1, 148, 128)
    RUN("frcp", k_frcp, 4, 148, 128)
    RUN("log", k_log, 1, 148, 128)
    RUN("log", k_log, 4, 148, 128)
    // half-filled warps: does a 16-lane warp issue DFMA in half the time?
    {
        float ms = timeit([&] { k_dfma<8><<<148, 128>>>(d, iters, 1.0000001, 1e-9); });
        printf("full warps 8 chains: %.3f ms\n", ms);
    }
    for (int active : {32, 24, 16, 8, 1}) {
        float ms = timeit([&] { k_dfma_part<8><<<148, 128>>>(d, iters, 1.0000001, 1e-9, active); });
        printf("dfma 8 chains, %2d active lanes/warp, 1 warp/SMSP: %.3f ms -> %.2f cycles/op/warp\n", active, ms,
               ms * 1e-3 * ghz * 1e9 / iters / 8);
    }
    for (int active : {32, 16}) {
        float ms = timeit([&] { k_dfma_part<8><<<148, 256>>>(d, iters, 1.0000001, 1e-9, active); });
        printf("dfma 8 chains, %2d active lanes/warp, 2 warps/SMSP: %.3f ms -> %.2f cycles/op/warp\n", active, ms,
               ms * 1e-3 * ghz * 1e9 / iters / 8);
    }
    return 0;
}
