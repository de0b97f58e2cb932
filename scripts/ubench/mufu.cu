// micro-benchmark: MUFU.EX2 issue rate per SM sub-partition (cycles per warp instruction)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = ex2(a[i]);                 // MUFU only
            if (MODE == 1) a[i] = ex2(fmaf(a[i], 1.0001f, -0.5f));   // FFMA + MUFU
            if (MODE == 2) a[i] = fmaf(a[i], 1.0001f, -0.5f);        // FFMA only
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 1000;
    for (int mode = 0; mode < 3; ++mode)
        for (int threads : {128, 256, 512, 1024}) {
            if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
            if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
            if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double per = double(h[0]) / (iters * 16.0) / (threads / 128.0);
            printf("mode %d threads %4d: %.2f cycles per warp-instr-group per SMSP (total %lld)\n", mode, threads, per, h[0]);
        }
    return 0;
}
