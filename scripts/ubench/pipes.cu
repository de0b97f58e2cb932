// micro-benchmark: issue cost (cycles per warp instruction per SM sub-partition) of the instructions the
// Snake / attention inner loops are built from: MUFU.{EX2,SIN}, F2FP packs, PRMT-based bf16 packs,
// FFMA forms, FFMA2, HMMA.16816 (legacy warp MMA).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/pipes scripts/ubench/pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[16];
    uint32_t u[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i * 0.37f, u[i] = threadIdx.x * 77 + i;
    float c4[4] = {0.f, 0.f, 0.f, 0.f};
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 1) asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(a[i]));       // FMUL (1/2pi) + MUFU.SIN
            if (MODE == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[i]));
            if (MODE == 3) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(__uint_as_float(u[(i + 1) & 15])));
            if (MODE == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])), "f"(__uint_as_float(u[(i + 1) & 15])));
            if (MODE == 5) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(u[(i + 3) & 15]), "r"(u[(i + 1) & 15]));
            if (MODE == 6) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 5) & 15]), "f"(a[(i + 9) & 15]));   // 3 registers
            if (MODE == 7) asm volatile("fma.rn.f32 %0, %0, 0f3F800347, 0fBF000000;" : "+f"(a[i]));                          // immediates
            if (MODE == 8) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 5) & 15]));
            if (MODE == 9) {
                if (i < 8) {
                    asm volatile("{.reg .b64 x, y, z;\n\tmov.b64 x, {%0, %1};\n\tmov.b64 y, {%2, %3};\n\tmov.b64 z, {%4, %5};\n\t"
                                 "fma.rn.f32x2 x, x, y, z;\n\tmov.b64 {%0, %1}, x;}"
                                 : "+f"(a[2 * i]), "+f"(a[2 * i + 1])
                                 : "f"(a[(2 * i + 4) & 15]), "f"(a[(2 * i + 5) & 15]), "f"(a[(2 * i + 8) & 15]), "f"(a[(2 * i + 9) & 15]));
                }
            }
            if (MODE == 10) {
                if (i < 8)   // independent accumulators: 2 rotating
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(a[(i & 3) * 4]), "+f"(a[(i & 3) * 4 + 1]), "+f"(a[(i & 3) * 4 + 2]), "+f"(a[(i & 3) * 4 + 3])
                                 : "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]));
            }
            if (MODE == 11) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 5) & 15]), "r"(u[(i + 9) & 15]));
            if (MODE == 12) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 5) & 15]));
            if (MODE == 13) asm volatile("{.reg .f16 h, g;\n\tmov.b32 {h, g}, %1;\n\tcvt.f32.f16 %0, h;}" : "=f"(a[i]) : "r"(__float_as_uint(a[i])));
            if (MODE == 14) asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 15) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 16) asm volatile("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(*reinterpret_cast<unsigned short*>(&u[i])) : "f"(a[i]), "f"(a[(i + 1) & 15]));
            if (MODE == 17) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(u[(i + 5) & 15]), "r"(u[(i + 9) & 15]));
        }
    }
    long long t1 = clock64();
    float s = c4[0];
    for (int i = 0; i < 16; ++i) s += a[i] + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc, int per_iter) {
    const int iters = 2000;
    for (int threads : {128, 512, 1024}) {
        k<MODE><<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double per = double(h[0]) / (double(iters) * per_iter) / (threads / 128.0);
        printf("%-34s threads %4d: %6.2f cycles per warp instruction per SMSP\n", name, threads, per);
    }
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    run<0>("MUFU.EX2", out, cyc, 16);
    run<1>("sin.approx (FMUL + MUFU.SIN)", out, cyc, 16);
    run<14>("cos.approx (FMUL + MUFU.COS)", out, cyc, 16);
    run<2>("MUFU.RCP", out, cyc, 16);
    run<15>("MUFU.TANH", out, cyc, 16);
    run<3>("F2FP.F16.F32.PACK_AB", out, cyc, 16);
    run<4>("F2FP.BF16.F32.PACK_AB", out, cyc, 16);
    run<16>("F2FP e4m3x2", out, cyc, 16);
    run<13>("F2F half -> float (HADD2.F32)", out, cyc, 16);
    run<5>("PRMT", out, cyc, 16);
    run<11>("LOP3", out, cyc, 16);
    run<12>("IADD", out, cyc, 16);
    run<6>("FFMA 3-register", out, cyc, 16);
    run<7>("FFMA immediates", out, cyc, 16);
    run<8>("FMUL", out, cyc, 16);
    run<17>("HFMA2", out, cyc, 16);
    run<9>("FFMA2 (fma.rn.f32x2)", out, cyc, 8);
    run<10>("HMMA.16816.F32 (mma.sync)", out, cyc, 8);
    return 0;
}
