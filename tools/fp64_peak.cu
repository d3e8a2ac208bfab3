// Microbenchmark: sustained DFMA throughput of the B200 as a function of warps/SM and per-thread ILP.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
void run(int threads, int blocks_per_sm, int sms) {
    double* out;
    cudaMalloc(&out, 8);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    dfma_kernel<ILP><<<sms * blocks_per_sm, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    dfma_kernel<ILP><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double fmas = double(sms) * blocks_per_sm * threads * double(iters) * ILP;
    printf("ILP %2d threads/SM %4d : %.2f TDFMA/s (%.2f TFLOP/s), %.2f DFMA/clk/SM @1.965GHz\n", ILP, threads * blocks_per_sm, fmas / ms / 1e9,
           2 * fmas / ms / 1e9, fmas / (ms * 1e-3) / sms / 1.965e9);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    run<1>(256, 2, sms); run<2>(256, 2, sms); run<4>(256, 2, sms); run<8>(256, 2, sms); run<16>(256, 2, sms);
    run<1>(1024, 2, sms); run<4>(1024, 2, sms); run<8>(1024, 1, sms);
    run<8>(128, 1, sms); run<8>(128, 2, sms);
    return 0;
}
