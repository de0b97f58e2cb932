// Memory-bound kernels of the sampler: AdaLN / RMSNorm / LayerNorm modulation, CFG combine +
// Euler update, layout changes at the API boundary, reflect halo, timestep features.
// All are vectorised, coalesced HBM streaming kernels; reductions use warp shuffles.
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace svc {

// ------------------------------------------------------------------------------------------
// svc_norm_mod: one warp per row.  Reference: AdaptiveLayerNorm / RMSNorm / FinalLayer norm
// (modules/diffusion_transformer.py:30-48,274-285,401-403; v2 modules/v2/dit_model.py:20-54).
// ------------------------------------------------------------------------------------------
template <typename TO, int MAXV>
__global__ void __launch_bounds__(256) norm_mod_kernel(
    const float* __restrict__ x, long long x_bstride, long long x_rstride,
    const float* __restrict__ gamma, const float* __restrict__ mul, const float* __restrict__ add,
    float eps, int mode, TO* __restrict__ out, long long o_bstride, long long o_rstride, int B, int T,
    int D, TO* __restrict__ raw, int raw_f16) {
    const int warp_lin = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_lin >= B * T) return;
#ifdef SVC_NORM_FORWARD
    const int warp = warp_lin;
#else
    // rows from the END of the tensor first: the GEMM that produced x wrote it front to back, so its tail is what the
    // 126 MB L2 still holds; and the rows written last here (the front) are the ones the next GEMM reads first
    const int warp = B * T - 1 - warp_lin;
#endif
    const int b = warp / T, t = warp % T;
    const float* xr = x + static_cast<long long>(b) * x_bstride + static_cast<long long>(t) * x_rstride;
    float4 v[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < D) {
            v[i] = __ldg(reinterpret_cast<const float4*>(xr + c));
            s1 += v[i].x + v[i].y + v[i].z + v[i].w;
            s2 += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
        }
    }
    float mean = 0.f, rstd;
    if (mode == 0) {
        s2 = warp_sum(s2);
        rstd = rsqrtf(s2 / D + eps);
    } else {
        s1 = warp_sum(s1);
        mean = s1 / D;
        float var = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < D) {
                const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean,
                            a3 = v[i].w - mean;
                var += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
            }
        }
        var = warp_sum(var);
        rstd = rsqrtf(var / D + eps);
    }
    TO* orow = out + static_cast<long long>(b) * o_bstride + static_cast<long long>(t) * o_rstride;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < D) {
            float y[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd,
                          (v[i].w - mean) * rstd};
            if (gamma != nullptr) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
                y[0] *= g.x, y[1] *= g.y, y[2] *= g.z, y[3] *= g.w;
            }
            if (mul != nullptr) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(mul + c));
                y[0] *= g.x, y[1] *= g.y, y[2] *= g.z, y[3] *= g.w;
            }
            if (add != nullptr) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(add + c));
                y[0] += g.x, y[1] += g.y, y[2] += g.z, y[3] += g.w;
            }
            if constexpr (sizeof(TO) == 4) {
                *reinterpret_cast<float4*>(orow + c) = make_float4(y[0], y[1], y[2], y[3]);
            } else {
                uint2 q;
                q.x = pack2<TO>(y[0], y[1]);
                q.y = pack2<TO>(y[2], y[3]);
                *reinterpret_cast<uint2*>(orow + c) = q;
            }
            if (raw != nullptr) {      // un-normalised copy of the row in the operand dtype (same layout as out)
                TO* rrow = raw + static_cast<long long>(b) * o_bstride + static_cast<long long>(t) * o_rstride;
                if constexpr (sizeof(TO) == 4) {
                    *reinterpret_cast<float4*>(rrow + c) = v[i];
                } else {
                    uint2 q;             // the raw copy may use the other 16-bit format than the normalised output
                    q.x = pack_op16_rt(v[i].x, v[i].y, raw_f16);
                    q.y = pack_op16_rt(v[i].z, v[i].w, raw_f16);
                    *reinterpret_cast<uint2*>(rrow + c) = q;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// svc_cfg_euler.  Reference: modules/flow_matching.py:98-110, modules/v2/cfm.py:86-130.
// ------------------------------------------------------------------------------------------
template <typename TO>
__global__ void __launch_bounds__(256) cfg_euler_kernel(
    float* __restrict__ x, const float* __restrict__ v, int n_branch, float c0, float c1, float c2,
    float dt, int B, int T, int C, int prompt_len, const int* __restrict__ x_lens,
    TO* __restrict__ x_op, bool write_op) {
    const long long n4 = static_cast<long long>(B) * T * C / 4;
    const long long per_branch = static_cast<long long>(B) * T * C;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long e = i * 4;
        const int t = static_cast<int>((e / C) % T);
        const int b = static_cast<int>(e / (static_cast<long long>(C) * T));
        float4 xv = *reinterpret_cast<const float4*>(x + e);
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(v + e));
        float4 d = make_float4(c0 * v0.x, c0 * v0.y, c0 * v0.z, c0 * v0.w);
        if (n_branch > 1) {
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(v + per_branch + e));
            d.x = fmaf(c1, v1.x, d.x), d.y = fmaf(c1, v1.y, d.y), d.z = fmaf(c1, v1.z, d.z),
            d.w = fmaf(c1, v1.w, d.w);
        }
        if (n_branch > 2) {
            const float4 v2 = __ldg(reinterpret_cast<const float4*>(v + 2 * per_branch + e));
            d.x = fmaf(c2, v2.x, d.x), d.y = fmaf(c2, v2.y, d.y), d.z = fmaf(c2, v2.z, d.z),
            d.w = fmaf(c2, v2.w, d.w);
        }
        xv.x = fmaf(dt, d.x, xv.x), xv.y = fmaf(dt, d.y, xv.y), xv.z = fmaf(dt, d.z, xv.z),
        xv.w = fmaf(dt, d.w, xv.w);
        const bool dead = t < prompt_len || (x_lens != nullptr && t >= x_lens[b]);
        if (dead) xv = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(x + e) = xv;
        if (write_op) {
            if constexpr (sizeof(TO) == 4) {
                *reinterpret_cast<float4*>(x_op + e) = xv;
            } else {
                uint2 q;
                q.x = pack2<TO>(xv.x, xv.y);
                q.y = pack2<TO>(xv.z, xv.w);
                *reinterpret_cast<uint2*>(x_op + e) = q;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// layout changes
// ------------------------------------------------------------------------------------------
template <typename TO>
__global__ void __launch_bounds__(256) bct_to_btc_kernel(const float* __restrict__ in, TO* __restrict__ out,
                                                         long long o_bstride, long long o_rstride,
                                                         int C, int T, int zero_from, int zero_to) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per pass
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + tx;
        tile[i][tx] = (c < C && t < T) ? in[(static_cast<long long>(b) * C + c) * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + tx;
        if (t < T && c < C) {
            float v = tile[tx][i];
            if (t >= zero_from && t < zero_to) v = 0.f;
            out[static_cast<long long>(b) * o_bstride + static_cast<long long>(t) * o_rstride + c] =
                from_f32<TO>(v);
        }
    }
}

__global__ void __launch_bounds__(256) btc_to_bct_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         int T, int C) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + tx;
        tile[i][tx] = (t < T && c < C) ? in[(static_cast<long long>(b) * T + t) * C + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + tx;
        if (c < C && t < T) out[(static_cast<long long>(b) * C + c) * T + t] = tile[tx][i];
    }
}

template <typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ in, TO* __restrict__ out,
                                                   long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        out[i] = from_f32<TO>(in[i]);
}

// Reflect halo.  Reference: modules/encodec.py:96-113,212-228 (pad1d 'reflect' in SConv1d).
template <typename TO>
__global__ void reflect_halo_kernel(TO* __restrict__ buf, long long bstride, long long rstride, int T,
                                    int C, int pad, const int* __restrict__ lens) {
    const int b = blockIdx.y;
    const int i = blockIdx.x;  // 0..2*pad-1
    int len = lens != nullptr ? lens[b] : T;
    len = max(2, min(len, T));
    TO* base = buf + static_cast<long long>(b) * bstride;
    // source = a BODY row, index reflected inside [0, len-1] (never a halo row another block of this launch
    // is writing, whatever len is; for len > pad this is exactly torch's reflect padding)
    int dst, j;
    if (i < pad) {
        dst = pad - 1 - i;          // left halo row
        j = i + 1;                  // body row i+1
        if (j > len - 1) j = 2 * (len - 1) - j;
    } else {
        const int k = i - pad;
        dst = pad + len + k;        // right halo row
        j = len - 2 - k;            // body row len-2-k
        if (j < 0) j = -j;
    }
    j = min(max(j, 0), len - 1);
    const int src = pad + j;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
        base[static_cast<long long>(dst) * rstride + c] = base[static_cast<long long>(src) * rstride + c];
}

// Reference: modules/diffusion_transformer.py:341-359.
__global__ void timestep_embedding_kernel(const float* __restrict__ t, const float* __restrict__ freqs,
                                          float* __restrict__ out, int n, int half) {
    const int i = blockIdx.x;
    if (i >= n) return;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        const float a = 1000.0f * t[i] * freqs[j];
        out[i * 2 * half + j] = cosf(a);
        out[i * 2 * half + half + j] = sinf(a);
    }
}

__global__ void set_rows_kernel(const float* __restrict__ src, long long src_bstride,
                                float* __restrict__ dst, long long dst_bstride, int D) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < D; c += blockDim.x)
        dst[static_cast<long long>(b) * dst_bstride + c] = src[static_cast<long long>(b) * src_bstride + c];
}

static inline bool aligned16(const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; }

}  // namespace svc

using namespace svc;

static int norm_mod_impl(const float* x, long long x_bstride, long long x_rstride,
                            const float* gamma, const float* mul, const float* add, float eps,
                            int mode, void* out, long long o_bstride, long long o_rstride, int B,
                            int T, int D, int out_dtype, void* stream, void* raw_out, int raw_dtype) {
    if (D % 4 != 0 || D > 2048 || B < 1 || T < 1) {
        svc_set_error("svc_norm_mod: D must be a multiple of 4 and <= 2048");
        return SVC_ERR_ARG;
    }
    const int esz = out_dtype == SVC_F32 ? 4 : 2;
    if (out_dtype != SVC_F32 && out_dtype != SVC_BF16 && out_dtype != SVC_F16) {
        svc_set_error("svc_norm_mod: bad out_dtype");
        return SVC_ERR_ARG;
    }
    if (!aligned16(x) || (x_bstride % 4) || (x_rstride % 4) || !aligned16(out) ||
        (o_bstride * esz) % 8 || (o_rstride * esz) % 8 || (gamma && !aligned16(gamma)) ||
        (mul && !aligned16(mul)) || (add && !aligned16(add))) {
        svc_set_error("svc_norm_mod: misaligned pointer or stride");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long rows = static_cast<long long>(B) * T;
    const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
#define LAUNCH_NORM(TO, MAXV)                                                                   \
    norm_mod_kernel<TO, MAXV><<<blocks, 256, 0, st>>>(x, x_bstride, x_rstride, gamma, mul, add, \
                                                      eps, mode, static_cast<TO*>(out),         \
                                                      o_bstride, o_rstride, B, T, D,             \
                                                      static_cast<TO*>(raw_out), raw_dtype == SVC_F16)
    if (out_dtype == SVC_F32) {
        if (D <= 512) LAUNCH_NORM(float, 4);
        else if (D <= 1024) LAUNCH_NORM(float, 8);
        else LAUNCH_NORM(float, 16);
    } else if (out_dtype == SVC_F16) {
        if (D <= 512) LAUNCH_NORM(__half, 4);
        else if (D <= 1024) LAUNCH_NORM(__half, 8);
        else LAUNCH_NORM(__half, 16);
    } else {
        if (D <= 512) LAUNCH_NORM(__nv_bfloat16, 4);
        else if (D <= 1024) LAUNCH_NORM(__nv_bfloat16, 8);
        else LAUNCH_NORM(__nv_bfloat16, 16);
    }
#undef LAUNCH_NORM
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_norm_mod(const float* x, long long x_bstride, long long x_rstride, const float* gamma,
                            const float* mul, const float* add, float eps, int mode, void* out,
                            long long o_bstride, long long o_rstride, int B, int T, int D, int out_dtype,
                            void* stream) {
    return norm_mod_impl(x, x_bstride, x_rstride, gamma, mul, add, eps, mode, out, o_bstride, o_rstride, B, T, D,
                         out_dtype, stream, nullptr, out_dtype);
}

extern "C" int svc_norm_mod_copy(const float* x, long long x_bstride, long long x_rstride, const float* gamma,
                                 const float* mul, const float* add, float eps, int mode, void* out,
                                 void* raw_out, long long o_bstride, long long o_rstride, int B, int T, int D,
                                 int out_dtype, int raw_dtype, void* stream) {
    if ((out_dtype == SVC_F32) != (raw_dtype == SVC_F32)) {
        svc_set_error("svc_norm_mod_copy: raw_dtype must be out_dtype or the other 16-bit format");
        return SVC_ERR_ARG;
    }
    return norm_mod_impl(x, x_bstride, x_rstride, gamma, mul, add, eps, mode, out, o_bstride, o_rstride, B, T, D,
                         out_dtype, stream, raw_out, raw_dtype);
}

extern "C" int svc_cfg_euler(float* x, const float* v, int n_branch, float c0, float c1, float c2,
                             float dt, int B, int T, int C, int prompt_len, const int* x_lens,
                             void* x_op, int op_dtype, void* stream) {
    if (C % 4 != 0 || n_branch < 1 || n_branch > 3 || !aligned16(x) || !aligned16(v) ||
        (x_op && !aligned16(x_op))) {
        svc_set_error("svc_cfg_euler: C must be a multiple of 4, 1..3 branches, aligned pointers");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n4 = static_cast<long long>(B) * T * C / 4;
    const unsigned blocks = static_cast<unsigned>(std::min<long long>((n4 + 255) / 256, kNumSMs * 16));
    if (op_dtype == SVC_F32)
        cfg_euler_kernel<float><<<blocks, 256, 0, st>>>(x, v, n_branch, c0, c1, c2, dt, B, T, C,
                                                        prompt_len, x_lens, static_cast<float*>(x_op),
                                                        x_op != nullptr);
    else if (op_dtype == SVC_F16)
        cfg_euler_kernel<__half><<<blocks, 256, 0, st>>>(x, v, n_branch, c0, c1, c2, dt, B, T, C, prompt_len,
                                                         x_lens, static_cast<__half*>(x_op), x_op != nullptr);
    else
        cfg_euler_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
            x, v, n_branch, c0, c1, c2, dt, B, T, C, prompt_len, x_lens,
            static_cast<__nv_bfloat16*>(x_op), x_op != nullptr);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_bct_to_btc(const float* in, void* out, long long o_bstride, long long o_rstride,
                              int B, int C, int T, int zero_from, int zero_to, int out_dtype,
                              void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((T + 31) / 32, (C + 31) / 32, B);
    if (out_dtype == SVC_F32)
        bct_to_btc_kernel<float><<<grid, 256, 0, st>>>(in, static_cast<float*>(out), o_bstride,
                                                       o_rstride, C, T, zero_from, zero_to);
    else if (out_dtype == SVC_F16)
        bct_to_btc_kernel<__half><<<grid, 256, 0, st>>>(in, static_cast<__half*>(out), o_bstride, o_rstride, C, T,
                                                        zero_from, zero_to);
    else
        bct_to_btc_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
            in, static_cast<__nv_bfloat16*>(out), o_bstride, o_rstride, C, T, zero_from, zero_to);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_btc_to_bct(const float* in, float* out, int B, int T, int C, void* stream) {
    dim3 grid((T + 31) / 32, (C + 31) / 32, B);
    btc_to_bct_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, T, C);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_cast(const float* in, void* out, long long n, int out_dtype, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n <= 0) return SVC_OK;
    const unsigned blocks = static_cast<unsigned>(std::min<long long>((n + 255) / 256, kNumSMs * 16));
    if (out_dtype == SVC_F32)
        cast_kernel<float><<<blocks, 256, 0, st>>>(in, static_cast<float*>(out), n);
    else if (out_dtype == SVC_F16)
        cast_kernel<__half><<<blocks, 256, 0, st>>>(in, static_cast<__half*>(out), n);
    else
        cast_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(in, static_cast<__nv_bfloat16*>(out), n);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

// out[s][n][k] = W[n][k] * g[k] * mul[s][k]  (folded RMS norm: the norm's per-column factors into the weight columns)
template <typename TO>
__global__ void __launch_bounds__(256) scale_cols_kernel(const float* __restrict__ W, long long w_rstride,
                                                         const float* __restrict__ g, const float* __restrict__ mul,
                                                         long long mul_stride, TO* __restrict__ out, int N, int K) {
    const int s = blockIdx.y;
    const long long per = static_cast<long long>(N) * K;
    const float* ms = mul != nullptr ? mul + s * mul_stride : nullptr;
    TO* o = out + s * per;
    for (long long i = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4; i < per;
         i += static_cast<long long>(gridDim.x) * blockDim.x * 4) {
        const int n = static_cast<int>(i / K), k = static_cast<int>(i - static_cast<long long>(n) * K);
        float4 w = __ldg(reinterpret_cast<const float4*>(W + n * w_rstride + k));
        if (g != nullptr) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(g + k));
            w.x *= q.x, w.y *= q.y, w.z *= q.z, w.w *= q.w;
        }
        if (ms != nullptr) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(ms + k));
            w.x *= q.x, w.y *= q.y, w.z *= q.z, w.w *= q.w;
        }
        if constexpr (sizeof(TO) == 4) {
            *reinterpret_cast<float4*>(o + i) = w;
        } else {
            uint2 q;
            q.x = pack2<TO>(w.x, w.y);
            q.y = pack2<TO>(w.z, w.w);
            *reinterpret_cast<uint2*>(o + i) = q;
        }
    }
}

extern "C" int svc_scale_cols(const float* W, long long w_rstride, const float* g, const float* mul, long long mul_stride,
                              void* out, int out_dtype, int S, int N, int K, void* stream) {
    if (S < 1 || N < 1 || K < 4 || K % 4 != 0 || w_rstride % 4 != 0 || (mul != nullptr && mul_stride % 4 != 0) ||
        reinterpret_cast<uintptr_t>(W) % 16 != 0 || reinterpret_cast<uintptr_t>(out) % 16 != 0 ||
        (g != nullptr && reinterpret_cast<uintptr_t>(g) % 16 != 0) ||
        (mul != nullptr && reinterpret_cast<uintptr_t>(mul) % 16 != 0)) {
        svc_set_error("svc_scale_cols: K % 4 == 0 and 16-byte aligned pointers / strides required");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long per = static_cast<long long>(N) * K;
    dim3 grid(static_cast<unsigned>(std::min<long long>((per / 4 + 255) / 256, kNumSMs * 8)), S);
    if (out_dtype == SVC_F32)
        scale_cols_kernel<float><<<grid, 256, 0, st>>>(W, w_rstride, g, mul, mul_stride, static_cast<float*>(out), N, K);
    else if (out_dtype == SVC_F16)
        scale_cols_kernel<__half><<<grid, 256, 0, st>>>(W, w_rstride, g, mul, mul_stride, static_cast<__half*>(out), N, K);
    else
        scale_cols_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(W, w_rstride, g, mul, mul_stride,
                                                               static_cast<__nv_bfloat16*>(out), N, K);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_reflect_halo(void* buf, long long bstride, long long rstride, int B, int T, int C,
                                int pad, const int* lens, int dtype, void* stream) {
    if (pad < 1 || T < pad + 1) {
        svc_set_error("svc_reflect_halo: need T > pad >= 1");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(2 * pad, B);
    if (dtype == SVC_F32)
        reflect_halo_kernel<float><<<grid, 128, 0, st>>>(static_cast<float*>(buf), bstride, rstride, T,
                                                         C, pad, lens);
    else      // a row copy: the 16-bit instantiation serves bf16 and fp16 alike
        reflect_halo_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(static_cast<__nv_bfloat16*>(buf),
                                                                 bstride, rstride, T, C, pad, lens);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_timestep_embedding(const float* t, const float* freqs, float* out, int n, int half,
                                      void* stream) {
    if (n < 1) return SVC_OK;
    timestep_embedding_kernel<<<n, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, freqs, out, n, half);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_set_rows(const float* src, long long src_bstride, float* dst, long long dst_bstride,
                            int B, int D, void* stream) {
    if (B < 1) return SVC_OK;
    set_rows_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(src, src_bstride, dst,
                                                                     dst_bstride, D);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

// ------------------------------------------------------------------------------------------
// InterpolateRegulator pieces (modules/length_regulator.py:90-141)
// ------------------------------------------------------------------------------------------
// out[b, t, :] = src[b, idx[t], :] (+ add_vec) (+ emb[q[b, eidx[t]], :])  -- F.interpolate(mode='nearest')
// as a row gather (:115), `x + f0_mask` (:122) or `x + interpolate(f0_embedding(quantized_f0))` (:125-129)
template <typename TO>
__global__ void __launch_bounds__(128) interp_rows_kernel(
    const float* __restrict__ src, long long sb, long long sr, const int* __restrict__ idx,
    const float* __restrict__ add_vec, const float* __restrict__ emb, const int* __restrict__ q,
    long long qb, const int* __restrict__ eidx, TO* __restrict__ out, long long ob, long long orow, int D) {
    const int t = blockIdx.x, b = blockIdx.y;
    const float* s = src + b * sb + static_cast<long long>(idx[t]) * sr;
    const float* e = emb != nullptr ? emb + static_cast<long long>(q[b * qb + eidx[t]]) * D : nullptr;
    TO* o = out + b * ob + static_cast<long long>(t) * orow;
    for (int c = threadIdx.x; c < D; c += 128) {
        float v = s[c];
        if (add_vec != nullptr) v += add_vec[c];
        if (e != nullptr) v += e[c];
        o[c] = from_f32<TO>(v);
    }
}

// GroupNorm(1 group) statistics of one sample = all T*C values (nn.GroupNorm(1, C), :51): fp64 sums
__global__ void __launch_bounds__(256) gn1_stats_kernel(const float* __restrict__ x, long long bstride,
                                                        long long rstride, int T, int C,
                                                        double* __restrict__ stats) {
    const int b = blockIdx.y;
    const float* xb = x + b * bstride;
    double s = 0.0, ss = 0.0;
    const long long n = static_cast<long long>(T) * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const int t = static_cast<int>(i / C), c = static_cast<int>(i - static_cast<long long>(t) * C);
        const double v = xb[static_cast<long long>(t) * rstride + c];
        s += v;
        ss += v * v;
    }
    __shared__ double sh[2][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        ss += __shfl_down_sync(0xffffffffu, ss, o);
    }
    if ((threadIdx.x & 31) == 0) sh[0][threadIdx.x >> 5] = s, sh[1][threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c2 = 0.0;
        for (int w = 0; w < 8; ++w) a += sh[0][w], c2 += sh[1][w];
        atomicAdd(&stats[2 * b], a);
        atomicAdd(&stats[2 * b + 1], c2);
    }
}

// y = Mish((x - mean) * rstd * gamma[c] + beta[c]); Mish(y) = y * tanh(softplus(y)) (nn.Mish, :52)
template <typename TO, bool PRECISE>
__global__ void __launch_bounds__(256) gn1_mish_kernel(const float* __restrict__ x, long long bstride,
                                                       long long rstride, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps,
                                                       const double* __restrict__ stats, TO* __restrict__ out,
                                                       long long ob, long long orow, int T, int C) {
    const int b = blockIdx.y;
    const double n = static_cast<double>(T) * C;
    const double mean_d = stats[2 * b] / n;
    const double var_d = fmax(stats[2 * b + 1] / n - mean_d * mean_d, 0.0);
    const float mean = static_cast<float>(mean_d);
    const float rstd = static_cast<float>(1.0 / sqrt(var_d + static_cast<double>(eps)));
    const long long total = static_cast<long long>(T) * C;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
        const int t = static_cast<int>(i / C), c = static_cast<int>(i - static_cast<long long>(t) * C);
        const float v = x[b * bstride + static_cast<long long>(t) * rstride + c];
        const float y = (v - mean) * rstd * gamma[c] + beta[c];
        const float sp = y > 20.0f ? y : (PRECISE ? log1pf(expf(y)) : __logf(1.0f + __expf(y)));
        const float m = y * (PRECISE ? tanhf(sp) : tanhf(sp));
        out[b * ob + static_cast<long long>(t) * orow + c] = from_f32<TO>(m);
    }
}

// x[b, t, :] = 0 for t >= lens[b]   (`out * mask`, :140)
__global__ void __launch_bounds__(128) mask_rows_kernel(float* __restrict__ x, long long bstride,
                                                        long long rstride, const int* __restrict__ lens,
                                                        int D) {
    const int t = blockIdx.x, b = blockIdx.y;
    if (t < lens[b]) return;
    float* o = x + b * bstride + static_cast<long long>(t) * rstride;
    for (int c = threadIdx.x; c < D; c += 128) o[c] = 0.f;
}

extern "C" int svc_interp_rows(const float* src, long long src_bstride, long long src_rstride,
                               const int* idx, const float* add_vec, const float* emb, const int* emb_q,
                               long long q_bstride, const int* emb_idx, void* out, long long out_bstride,
                               long long out_rstride, int B, int Tout, int D, int dtype, void* stream) {
    if (B < 1 || Tout < 1 || D < 1 || src == nullptr || idx == nullptr || out == nullptr ||
        (emb != nullptr && (emb_q == nullptr || emb_idx == nullptr))) {
        svc_set_error("svc_interp_rows: bad arguments");
        return SVC_ERR_ARG;
    }
    dim3 grid(Tout, B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == SVC_F32)
        interp_rows_kernel<float><<<grid, 128, 0, st>>>(src, src_bstride, src_rstride, idx, add_vec, emb,
                                                        emb_q, q_bstride, emb_idx, static_cast<float*>(out),
                                                        out_bstride, out_rstride, D);
    else if (dtype == SVC_F16)
        interp_rows_kernel<__half><<<grid, 128, 0, st>>>(src, src_bstride, src_rstride, idx, add_vec, emb, emb_q,
                                                         q_bstride, emb_idx, static_cast<__half*>(out), out_bstride,
                                                         out_rstride, D);
    else
        interp_rows_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(
            src, src_bstride, src_rstride, idx, add_vec, emb, emb_q, q_bstride, emb_idx,
            static_cast<__nv_bfloat16*>(out), out_bstride, out_rstride, D);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_groupnorm1_mish(const float* x, long long bstride, long long rstride, const float* gamma,
                                   const float* beta, float eps, double* stats_ws, void* out,
                                   long long out_bstride, long long out_rstride, int B, int T, int C,
                                   int dtype, int precise, void* stream) {
    if (B < 1 || T < 1 || C < 1 || x == nullptr || gamma == nullptr || beta == nullptr ||
        stats_ws == nullptr || out == nullptr) {
        svc_set_error("svc_groupnorm1_mish: bad arguments");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * B, st);
    const long long n = static_cast<long long>(T) * C;
    const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 592));
    dim3 grid(blocks, B);
    gn1_stats_kernel<<<grid, 256, 0, st>>>(x, bstride, rstride, T, C, stats_ws);
    if (dtype == SVC_F32) {
        if (precise)
            gn1_mish_kernel<float, true><<<grid, 256, 0, st>>>(x, bstride, rstride, gamma, beta, eps, stats_ws,
                                                              static_cast<float*>(out), out_bstride,
                                                              out_rstride, T, C);
        else
            gn1_mish_kernel<float, false><<<grid, 256, 0, st>>>(x, bstride, rstride, gamma, beta, eps, stats_ws,
                                                               static_cast<float*>(out), out_bstride,
                                                               out_rstride, T, C);
    } else if (dtype == SVC_F16) {
        if (precise)
            gn1_mish_kernel<__half, true><<<grid, 256, 0, st>>>(x, bstride, rstride, gamma, beta, eps, stats_ws,
                                                               static_cast<__half*>(out), out_bstride, out_rstride,
                                                               T, C);
        else
            gn1_mish_kernel<__half, false><<<grid, 256, 0, st>>>(x, bstride, rstride, gamma, beta, eps, stats_ws,
                                                                static_cast<__half*>(out), out_bstride, out_rstride,
                                                                T, C);
    } else {
        if (precise)
            gn1_mish_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(
                x, bstride, rstride, gamma, beta, eps, stats_ws, static_cast<__nv_bfloat16*>(out), out_bstride,
                out_rstride, T, C);
        else
            gn1_mish_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(
                x, bstride, rstride, gamma, beta, eps, stats_ws, static_cast<__nv_bfloat16*>(out), out_bstride,
                out_rstride, T, C);
    }
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_mask_rows(float* x, long long bstride, long long rstride, const int* lens, int B, int T,
                             int D, void* stream) {
    if (B < 1 || T < 1 || D < 1 || x == nullptr || lens == nullptr) {
        svc_set_error("svc_mask_rows: bad arguments");
        return SVC_ERR_ARG;
    }
    mask_rows_kernel<<<dim3(T, B), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, bstride, rstride, lens, D);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

// ------------------------------------------------------------------------------------------
// Mel front-end pieces (modules/audio.py:45-82).  The STFT itself is a segmented GEMM over
// the padded audio viewed as rows of `hop` samples against the windowed DFT matrix.
// ------------------------------------------------------------------------------------------
// out[b, i] = y[b, reflect(i - pad)] for i < L + 2*pad, 0 for the tail up to out_len  (F.pad reflect, :58-61)
__global__ void __launch_bounds__(256) reflect_pad1d_kernel(const float* __restrict__ y, long long ybs, int L,
                                                            int pad, float* __restrict__ out, long long obs,
                                                            long long out_len) {
    const int b = blockIdx.y;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < out_len; i += gridDim.x * 256LL) {
        float v = 0.f;
        if (i < static_cast<long long>(L) + 2 * pad) {
            long long j = i - pad;
            if (j < 0) j = -j;
            if (j >= L) j = 2LL * (L - 1) - j;
            v = y[b * ybs + j];
        }
        out[b * obs + i] = v;
    }
}

// mag[m, k] = sqrt(re[m,k]^2 + im[m,k]^2 + eps) with spec rows = [re_0..re_{nb-1} | im_0..im_{nb-1} | pad]  (:78)
__global__ void __launch_bounds__(256) stft_mag_kernel(const float* __restrict__ spec, long long srs, int nb,
                                                       float eps, float* __restrict__ mag, long long mrs,
                                                       long long total) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
        const long long m = i / nb;
        const int k = static_cast<int>(i - m * nb);
        const float re = spec[m * srs + k], im = spec[m * srs + nb + k];
        mag[m * mrs + k] = sqrtf(re * re + im * im + eps);
    }
}

// x = log(max(x, clip))   (dynamic_range_compression_torch, :24-25)
__global__ void __launch_bounds__(256) log_clamp_kernel(float* __restrict__ x, long long n, float clip) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL)
        x[i] = logf(fmaxf(x[i], clip));
}

extern "C" int svc_reflect_pad1d(const float* y, long long y_bstride, int B, int L, int pad, float* out,
                                 long long out_bstride, long long out_len, void* stream) {
    if (B < 1 || L < 2 || pad < 0 || pad >= L || y == nullptr || out == nullptr ||
        out_len < static_cast<long long>(L) + 2 * pad) {
        svc_set_error("svc_reflect_pad1d: need 0 <= pad < L and out_len >= L + 2*pad");
        return SVC_ERR_ARG;
    }
    const int blocks = static_cast<int>(std::min<long long>((out_len + 255) / 256, 1184));
    reflect_pad1d_kernel<<<dim3(blocks, B), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        y, y_bstride, L, pad, out, out_bstride, out_len);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_stft_mag(const float* spec, long long spec_rstride, long long rows, int n_bins, float eps,
                            float* mag, long long mag_rstride, void* stream) {
    if (rows < 1 || n_bins < 1 || spec == nullptr || mag == nullptr) {
        svc_set_error("svc_stft_mag: bad arguments");
        return SVC_ERR_ARG;
    }
    const long long total = rows * n_bins;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 2368));
    stft_mag_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(spec, spec_rstride, n_bins, eps, mag,
                                                                          mag_rstride, total);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_log_clamp(float* x, long long n, float clip, void* stream) {
    if (n < 0 || (n > 0 && x == nullptr)) {
        svc_set_error("svc_log_clamp: bad arguments");
        return SVC_ERR_ARG;
    }
    if (n == 0) return SVC_OK;
    const int blocks = static_cast<int>(std::min<long long>((n + 255) / 256, 2368));
    log_clamp_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, clip);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

// ------------------------------------------------------------------------------------------
// Overlap / cos^2 crossfade stitching of vocoded chunks (inference.py:343-350,505-527;
// seed_vc_wrapper.py:190-285): every chunk but the last drops its final `ov` samples, every chunk
// but the first blends its first `ov` samples with the previous chunk's dropped tail.  The blend is
// done in fp64 and rounded to fp32 once, exactly as numpy does for float32 * float64 arrays.
// ------------------------------------------------------------------------------------------
__global__ void crossfade_stitch_kernel(const float* __restrict__ waves, long long wstride,
                                        const int* __restrict__ lens, const long long* __restrict__ offs,
                                        int n, int ov, const double* __restrict__ fade_in,
                                        const double* __restrict__ fade_out, float* __restrict__ out,
                                        long long total) {
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        int k = 0;
        while (k + 1 < n && idx >= offs[k + 1]) ++k;
        const long long i = idx - offs[k];
        float v = waves[k * wstride + i];
        if (k > 0 && i < ov) {
            const float t = waves[(k - 1) * wstride + lens[k - 1] - ov + i];
            v = static_cast<float>(__dadd_rn(__dmul_rn(static_cast<double>(v), fade_in[i]),
                                             __dmul_rn(static_cast<double>(t), fade_out[i])));
        }
        out[idx] = v;
    }
}

extern "C" int svc_crossfade_stitch(const float* waves, long long wave_stride, const int* lens,
                                    const long long* offs, int n_chunks, int overlap,
                                    const double* fade_in, const double* fade_out, float* out,
                                    long long total, void* stream) {
    if (n_chunks < 1 || overlap < 0 || total < 0 || waves == nullptr || lens == nullptr || offs == nullptr ||
        out == nullptr || (overlap > 0 && n_chunks > 1 && (fade_in == nullptr || fade_out == nullptr))) {
        svc_set_error("svc_crossfade_stitch: bad arguments");
        return SVC_ERR_ARG;
    }
    if (total == 0) return SVC_OK;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
    crossfade_stitch_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        waves, wave_stride, lens, offs, n_chunks, overlap, fade_in, fade_out, out, total);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

// ------------------------------------------------------------------------------------------
// Streaming SOLA stitch (real-time-gui.py:1103-1137): per stream, find the offset in [0, search] that
// maximises the normalised cross-correlation between the new block's head and the kept tail of the
// previous block, cut there, crossfade the first `sb` samples with the kept tail, keep the next tail.
// One CTA per stream; the window and the tail sit in shared memory.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sola_stitch_kernel(const float* __restrict__ infer, long long ibs,
                                                          float* __restrict__ sola_buf, long long bbs,
                                                          const float* __restrict__ fade_in,
                                                          const float* __restrict__ fade_out,
                                                          float* __restrict__ out, long long obs,
                                                          int* __restrict__ offsets, int sb, int search,
                                                          int block) {
    extern __shared__ float sm[];
    float* xs = sm;                    // sb + search samples of the new block's head
    float* bs = xs + sb + search;      // sb samples kept from the previous block
    __shared__ float best_v[256];
    __shared__ int best_o[256];
    const int b = blockIdx.x;
    const float* x = infer + b * ibs;
    float* buf = sola_buf + b * bbs;
    for (int i = threadIdx.x; i < sb + search; i += 256) xs[i] = x[i];
    for (int i = threadIdx.x; i < sb; i += 256) bs[i] = buf[i];
    __syncthreads();
    // argmax with torch.argmax semantics: NaN ranks above every number (the first NaN wins), ties go to the
    // smallest offset.  Threads without a candidate carry offset INT_MAX and lose every comparison; the final
    // offset is clamped to [0, search] so a non-finite input can never index outside the window.
    auto better = [](float v2, int o2, float v1, int o1) {
        const bool n2 = v2 != v2, n1 = v1 != v1;
        if (n2 != n1) return n2;
        if (n2 || v2 == v1) return o2 < o1;
        return v2 > v1;
    };
    float bv = -INFINITY;
    int bo = 0x7fffffff;
    for (int o = threadIdx.x; o <= search; o += 256) {
        float nom = 0.f, den = 0.f;
        for (int i = 0; i < sb; ++i) {
            const float v = xs[o + i];
            nom = fmaf(v, bs[i], nom);
            den = fmaf(v, v, den);
        }
        const float c = nom / sqrtf(den + 1e-8f);
        if (bo == 0x7fffffff || better(c, o, bv, bo)) bv = c, bo = o;
    }
    best_v[threadIdx.x] = bv;
    best_o[threadIdx.x] = bo;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const float v2 = best_v[threadIdx.x + s];
            const int o2 = best_o[threadIdx.x + s];
            if (o2 != 0x7fffffff &&
                (best_o[threadIdx.x] == 0x7fffffff || better(v2, o2, best_v[threadIdx.x], best_o[threadIdx.x])))
                best_v[threadIdx.x] = v2, best_o[threadIdx.x] = o2;
        }
        __syncthreads();
    }
    const int off = min(max(best_o[0], 0), search);
    if (threadIdx.x == 0) offsets[b] = off;
    // faded(i) = x[off+i]*fade_in[i] + buf[i]*fade_out[i] (two fp32 roundings, like the reference's
    // in-place `*=` then `+=`), raw(i) = x[off+i]
    auto val = [&](int i) {
        const float v = x[off + i];
        return i < sb ? __fadd_rn(__fmul_rn(v, fade_in[i]), __fmul_rn(bs[i], fade_out[i])) : v;
    };
    for (int i = threadIdx.x; i < block; i += 256) out[b * obs + i] = val(i);
    for (int i = threadIdx.x; i < sb; i += 256) buf[i] = val(block + i);
}

extern "C" int svc_sola_stitch(const float* infer, long long infer_bstride, int infer_len, float* sola_buf,
                               long long buf_bstride, const float* fade_in, const float* fade_out, float* out,
                               long long out_bstride, int* offsets, int B, int sola_buffer_frame,
                               int sola_search_frame, int block_frame, void* stream) {
    const int sb = sola_buffer_frame, search = sola_search_frame;
    if (B < 1 || sb < 1 || search < 0 || block_frame < 1 || infer == nullptr || sola_buf == nullptr ||
        fade_in == nullptr || fade_out == nullptr || out == nullptr || offsets == nullptr ||
        infer_len < search + block_frame + sb) {
        svc_set_error("svc_sola_stitch: need infer_len >= search + block + sola_buffer_frame");
        return SVC_ERR_ARG;
    }
    const int smem = (2 * sb + search) * 4;
    if (smem > 200 * 1024) {
        svc_set_error("svc_sola_stitch: window too large for shared memory");
        return SVC_ERR_ARG;
    }
    static int attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
        cudaFuncSetAttribute(sola_stitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_smem = smem;
    }
    sola_stitch_kernel<<<B, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        infer, infer_bstride, sola_buf, buf_bstride, fade_in, fade_out, out, out_bstride, offsets, sb, search,
        block_frame);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

// ------------------------------------------------------------------------------------------
// error plumbing shared by all translation units
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void svc_set_error(const char* msg) {
    strncpy(g_err, msg, sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
}
extern "C" const char* svc_last_error(void) { return g_err; }
extern "C" int svc_version(void) { return 101; }
