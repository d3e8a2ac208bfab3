// Microbenchmark: do the FP64 tensor-core pipe (mma.sync m8n8k4 f64 = DMMA) and the FP64 CUDA-core pipe (DFMA) of the B200
// run concurrently?  Three launches with the same number of warps: every warp DFMA, every warp DMMA, odd groups DMMA + even
// warps DFMA.  If the mixed launch sustains (close to) the sum of the two rates the pipes are separate -- then half of the
// sweep kernel's gate arithmetic could move to DMMA.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dmma_overlap.bin tools/dmma_overlap.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode 0: every warp DFMA, 1: every warp DMMA, 2: warp groups 0,2,.. DFMA / groups 1,3,.. DMMA, 3: even warps DFMA / odd idle,
// 4: odd warps DMMA / even idle.  The DFMA warps run 8x the iterations so that both halves take about the same time.
__global__ void __launch_bounds__(1024) mix_kernel(double* out, int iters, int mode, double a, double b) {
    const int warp = threadIdx.x >> 5;
    const bool odd = (warp >> 2) & 1;  // groups of four warps alternate: every scheduler (warp % 4) gets both kinds
    const bool use_mma = mode == 1 || ((mode == 2 || mode == 4) && odd);
    const bool idle = (mode == 3 && odd) || (mode == 4 && !odd);
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-9 + i;
    if (idle) return;
    if (use_mma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma(x[i], x[i + 1], a, b);
        }
    } else {
        for (int it = 0; it < 8 * iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}

int main(int argc, char** argv) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, threads = (argc > 1 ? atoi(argv[1]) : 512), blocks = sms * 1, iters = 4000;
    double* out;
    cudaMalloc(&out, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    printf("%s, %d SMs, %d warps/SM\n", p.name, sms, threads / 32);
    for (int mode = 0; mode < 5; ++mode) {
        mix_kernel<<<blocks, threads>>>(out, 100, mode, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        mix_kernel<<<blocks, threads>>>(out, iters, mode, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double warps = double(blocks) * threads / 32;
        // per warp and iteration: DFMA path 8 x 16 x 32 FMAs, DMMA path 8 instructions x 256 FMAs
        const double dfma_warps = mode == 0 ? warps : ((mode == 2 || mode == 3) ? warps / 2 : 0);
        const double dmma_warps = mode == 1 ? warps : ((mode == 2 || mode == 4) ? warps / 2 : 0);
        const double dfma = dfma_warps * iters * 8.0 * 16 * 32, dmma_fma = dmma_warps * iters * 8.0 * 256;
        const char* names[5] = {"all DFMA", "all DMMA", "half DFMA + half DMMA", "half DFMA alone", "half DMMA alone"};
        printf("mode %d (%-22s): %.3f ms  DFMA %.2f TFMA/s  DMMA %.2f TFMA/s  total %.2f TFLOP/s\n", mode, names[mode], ms, dfma / ms / 1e9,
               dmma_fma / ms / 1e9, 2 * (dfma + dmma_fma) / ms / 1e9);
    }
    return 0;
}
