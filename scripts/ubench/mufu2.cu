// micro-benchmark: issue rate of the exp2 flavours a softmax can use, per SM sub-partition
//   mode 0  ex2.approx.ftz.f32          (MUFU.EX2, one result per lane)
//   mode 1  ex2.approx.ftz.bf16x2       (two bf16 results per lane)
//   mode 2  ex2.approx.f16x2            (two half results per lane)
//   mode 3  tanh.approx.bf16x2 (for reference: another packed MUFU op)
// Reports cycles per warp instruction per SMSP with 1..8 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2_bf16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2_f16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned tanh_bf16x2(unsigned x) { unsigned y; asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
template <int MODE>
__global__ void k(unsigned* out, long long* cyc, int iters) {
    unsigned a[16];
    for (int i = 0; i < 16; ++i) a[i] = MODE == 0 ? __float_as_uint(-0.5f - threadIdx.x * 0.001f - i) : 0xbc00bc00u + threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = __float_as_uint(ex2(__uint_as_float(a[i])));
            if (MODE == 1) a[i] = ex2_bf16x2(a[i]);
            if (MODE == 2) a[i] = ex2_f16x2(a[i]);
            if (MODE == 3) a[i] = tanh_bf16x2(a[i]);
        }
    }
    long long t1 = clock64();
    unsigned s = 0; for (int i = 0; i < 16; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    unsigned* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 1000;
    for (int mode = 0; mode < 4; ++mode)
        for (int threads : {128, 256, 512, 1024}) {
            if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
            if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
            if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
            if (mode == 3) k<3><<<148, threads>>>(out, cyc, iters);
            cudaDeviceSynchronize();
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double per = double(h[0]) / (iters * 16.0) / (threads / 128.0);
            printf("mode %d threads %4d: %.2f cycles per warp instruction per SMSP (total %lld)\n", mode, threads, per, h[0]);
        }
    return 0;
}
