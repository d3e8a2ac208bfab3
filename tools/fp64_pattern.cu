// Microbenchmark: the sweep kernel's in-register 2x2 complex update pattern (16 amplitudes / thread),
// without any memory traffic: what fraction of the DFMA peak can this instruction mix reach?
#include <cstdio>
#include <cuda_runtime.h>

template <int B>
__device__ __forceinline__ void apply(double2 (&a)[16], const double2* m) {
    const double2 m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (j & (1 << B)) continue;
        const double2 x = a[j], y = a[j | (1 << B)];
        double t0 = m01.x * y.x, t1 = m01.x * y.y, t2 = m10.x * x.x, t3 = m10.x * x.y;
        t0 = fma(-m01.y, y.y, t0); t1 = fma(m01.y, y.x, t1); t2 = fma(-m10.y, x.y, t2); t3 = fma(m10.y, x.x, t3);
        t0 = fma(-m00.y, x.y, t0); t1 = fma(m00.y, x.x, t1); t2 = fma(-m11.y, y.y, t2); t3 = fma(m11.y, y.x, t3);
        a[j].x = fma(m00.x, x.x, t0); a[j].y = fma(m00.x, x.y, t1);
        a[j | (1 << B)].x = fma(m11.x, y.x, t2); a[j | (1 << B)].y = fma(m11.x, y.y, t3);
    }
}

__global__ void __launch_bounds__(256, 2) pattern_kernel(double2* out, const double2* mats, int n_ops, int iters, int mode) {
    extern __shared__ double2 sm[];
    for (int i = threadIdx.x; i < n_ops * 4; i += blockDim.x) sm[i] = mats[i];
    __syncthreads();
    double2 a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = make_double2(threadIdx.x * 1e-3 + j, 0.5 * j);
    for (int it = 0; it < iters; ++it) {
        for (int o = 0; o < n_ops; ++o) {
            const double2* m = sm + o * 4;
            if (mode == 0) {
                apply<0>(a, m);
            } else {
                switch ((o + it) & 3) {
                    case 0: apply<0>(a, m); break;
                    case 1: apply<1>(a, m); break;
                    case 2: apply<2>(a, m); break;
                    default: apply<3>(a, m); break;
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j].x + a[j].y;
    if (s == 1.2345) out[0] = make_double2(s, s);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    const int n_ops = 32;
    double2* mats; double2* out;
    cudaMalloc(&mats, n_ops * 4 * sizeof(double2)); cudaMalloc(&out, 16);
    double2 h[n_ops * 4];
    for (int i = 0; i < n_ops * 4; ++i) h[i] = make_double2(0.5 + 1e-3 * i, 0.5 - 1e-3 * i);
    cudaMemcpy(mats, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; ++mode)
        for (int bps = 1; bps <= 2; ++bps) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            const int iters = 200;
            pattern_kernel<<<sms * bps, 256, n_ops * 64>>>(out, mats, n_ops, 4, mode);
            cudaEventRecord(e0);
            pattern_kernel<<<sms * bps, 256, n_ops * 64>>>(out, mats, n_ops, iters, mode);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double instr = double(sms) * bps * 256 * double(iters) * n_ops * 128;
            printf("mode %d (%s) CTAs/SM %d: %.2f T fp64-instr/s  (peak measured ~16.9)\n", mode, mode ? "switch over 4 target bits" : "fixed target bit", bps, instr / ms / 1e9);
        }
    return 0;
}
