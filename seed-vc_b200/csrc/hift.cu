// HiFT vocoder pieces that are not dense contractions (reference: modules/hifigan/generator.py).
//   svc_unary         leaky-ReLU / ELU / Snake / |x| on frames-major activations, fp32 in -> operand dtype out
//   svc_hift_source   F0 -> harmonic source: nearest upsample, 9 phase accumulators, uv / noise mix, Linear + tanh
//                     (SineGen.forward :208-243, SourceModuleHnNSF.forward :262-279, _f02source :366-370)
//   svc_hift_stft     16-point Hann STFT of the source, hop 4, centre / reflect (_stft :372-378)
//   svc_hift_istft    exp / sin heads -> polar -> windowed 16-point irfft -> overlap-add / envelope -> clamp
//                     (_istft :380-385, forward :426-435)
// The convolutions of the generator run on svc_gemm.  All four kernels are HBM-streaming, one thread per output
// element (or frame), coalesced along the contiguous axis.
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace svc {

template <typename TO>
__device__ __forceinline__ void store_elem(TO* p, float v) {
    *p = from_f32<TO>(v);
}

// ------------------------------------------------------------------------------------------ unary
template <bool PRECISE>
__device__ __forceinline__ float unary_fn(float v, int kind, float slope, float a) {
    if (kind == 0) return v > 0.f ? v : v * slope;                                   // F.leaky_relu
    if (kind == 1) return v > 0.f ? v : (PRECISE ? expm1f(v) : __expf(v) - 1.0f);    // nn.ELU(alpha = 1)
    if (kind == 2) {                                                                 // Snake, linear-scale alpha (:79-90)
        const float s = PRECISE ? sinf(v * a) : __sinf(v * a);
        return v + (1.0f / (a + 1e-9f)) * s * s;
    }
    return fabsf(v);
}

// grid (row blocks, B): a block walks rows t = blockIdx.x, + gridDim.x, ...; threads cover the row 4 channels at a
// time (float4 in, 8 / 16-byte packed out), so every access is a contiguous row segment and nothing is divided.
template <typename TO, bool PRECISE, bool VEC>
__global__ void __launch_bounds__(256) unary_kernel(const float* __restrict__ x, long long xbs, long long xrs,
                                                    TO* __restrict__ out, long long obs, long long ors, int T, int C,
                                                    int kind, float slope, const float* __restrict__ alpha) {
    const int b = blockIdx.y;
    const int lanes = VEC ? C / 4 : C;                // work items per row
    const int rows_per_pass = 256 / lanes > 0 ? 256 / lanes : 1;
    const int sub = threadIdx.x / lanes, li = threadIdx.x % lanes;
    for (int t0 = blockIdx.x * rows_per_pass; t0 < T; t0 += gridDim.x * rows_per_pass) {
        const int t = t0 + sub;
        if (sub >= rows_per_pass || t >= T) continue;
        for (int i = li; i < lanes; i += (lanes < 256 ? lanes : 256)) {
            if constexpr (VEC) {
                const int c = 4 * i;
                const float4 v = *reinterpret_cast<const float4*>(x + b * xbs + t * xrs + c);
                float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kind == 2) a4 = __ldg(reinterpret_cast<const float4*>(alpha + c));
                const float y0 = unary_fn<PRECISE>(v.x, kind, slope, a4.x), y1 = unary_fn<PRECISE>(v.y, kind, slope, a4.y);
                const float y2 = unary_fn<PRECISE>(v.z, kind, slope, a4.z), y3 = unary_fn<PRECISE>(v.w, kind, slope, a4.w);
                TO* o = out + b * obs + t * ors + c;
                if constexpr (sizeof(TO) == 4) {
                    *reinterpret_cast<float4*>(o) = make_float4(y0, y1, y2, y3);
                } else {
                    uint2 q;
                    q.x = pack2<TO>(y0, y1);
                    q.y = pack2<TO>(y2, y3);
                    *reinterpret_cast<uint2*>(o) = q;
                }
            } else {
                const float v = x[b * xbs + t * xrs + i];
                store_elem<TO>(out + b * obs + t * ors + i, unary_fn<PRECISE>(v, kind, slope, kind == 2 ? __ldg(alpha + i) : 0.f));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ source
// prefix[b][h][q] = sum_{q' < q} scale * v[b][h][q'] in double, v = fl32(fl32(f0 * (h + 1)) / sr): the running phase
// (in cycles) at the start of frame q.  torch.cumsum on the CPU accumulates fp32 inputs in double and rounds each
// prefix to fp32; F0 is piecewise constant per frame after nn.Upsample(nearest), so the prefix at sample
// n = scale * q + r is prefix[q] + (r + 1) * v[q] in closed form (the same value to ~1e-13).
__global__ void hift_prefix_kernel(const float* __restrict__ f0, long long f0_bs, double* __restrict__ prefix, int Tm,
                                   int H, int scale, float sr) {
    const int b = blockIdx.x, h = threadIdx.x;
    if (h >= H) return;
    double acc = 0.0;
    double* p = prefix + (static_cast<long long>(b) * H + h) * Tm;
    for (int q = 0; q < Tm; ++q) {
        p[q] = acc;
        const float v = __fdiv_rn(__fmul_rn(f0[b * f0_bs + q], static_cast<float>(h + 1)), sr);
        acc += static_cast<double>(scale) * static_cast<double>(v);
    }
}

constexpr int kMaxHarm = 16;

__global__ void __launch_bounds__(256) hift_source_kernel(
    const float* __restrict__ f0, long long f0_bs, const double* __restrict__ prefix, const float* __restrict__ phase,
    const float* __restrict__ noise, const float* __restrict__ lin_w, float lin_b, float* __restrict__ out,
    long long out_bs, int Tm, int H, int scale, float sr, float sine_amp, float noise_std, float voiced_thr) {
    __shared__ float s_w[kMaxHarm], s_ph[kMaxHarm];
    const int b = blockIdx.y;
    if (threadIdx.x < H) {
        s_w[threadIdx.x] = lin_w[threadIdx.x];
        s_ph[threadIdx.x] = threadIdx.x == 0 ? 0.f : phase[b * H + threadIdx.x];    // phase_vec[:, 0, :] = 0
    }
    __syncthreads();
    const long long L = static_cast<long long>(Tm) * scale;
    const long long n = blockIdx.x * 256LL + threadIdx.x;
    if (n >= L) return;
    const int q = static_cast<int>(n / scale), r = static_cast<int>(n - static_cast<long long>(q) * scale);
    const float f = f0[b * f0_bs + q];
    const float uv = f > voiced_thr ? 1.f : 0.f;
    const float noise_amp = __fadd_rn(__fmul_rn(uv, noise_std), __fdiv_rn(__fmul_rn(1.f - uv, sine_amp), 3.f));
    const float two_pi = 6.283185307179586f;
    float acc = lin_b;
    for (int h = 0; h < H; ++h) {
        const float v = __fdiv_rn(__fmul_rn(f, static_cast<float>(h + 1)), sr);
        const double c = prefix[(static_cast<long long>(b) * H + h) * Tm + q] +
                         static_cast<double>(r + 1) * static_cast<double>(v);
        const float cf = static_cast<float>(c);
        const float frac = cf - floorf(cf);                                          // % 1
        const float theta = __fmul_rn(two_pi, frac);
        float s = __fmul_rn(sine_amp, sinf(__fadd_rn(theta, s_ph[h])));
        const float nz = noise != nullptr ? noise[(static_cast<long long>(b) * H + h) * L + n] : 0.f;
        s = __fadd_rn(__fmul_rn(s, uv), __fmul_rn(noise_amp, nz));
        acc = fmaf(s, s_w[h], acc);
    }
    out[b * out_bs + n] = tanhf(acc);
}

// ------------------------------------------------------------------------------------------ STFT / iSTFT
__constant__ float c_cos16[16] = {1.f, 0.9238795325f, 0.7071067812f, 0.3826834324f, 0.f, -0.3826834324f,
                                  -0.7071067812f, -0.9238795325f, -1.f, -0.9238795325f, -0.7071067812f,
                                  -0.3826834324f, 0.f, 0.3826834324f, 0.7071067812f, 0.9238795325f};
__constant__ float c_sin16[16] = {0.f, 0.3826834324f, 0.7071067812f, 0.9238795325f, 1.f, 0.9238795325f,
                                  0.7071067812f, 0.3826834324f, 0.f, -0.3826834324f, -0.7071067812f,
                                  -0.9238795325f, -1.f, -0.9238795325f, -0.7071067812f, -0.3826834324f};
// hann(16, periodic): 0.5 - 0.5 cos(2 pi n / 16)
__device__ __forceinline__ float hann16(int n) { return 0.5f - 0.5f * c_cos16[n & 15]; }

// one thread per (batch, frame): 9 complex bins -> channels [re0..re8, im0..im8, zero pad]
template <typename TO>
__global__ void __launch_bounds__(128) hift_stft_kernel(const float* __restrict__ s, long long s_bs, TO* __restrict__ out,
                                                        long long obs, long long ors, int L, int TT, int rows, int Cpad) {
    const int b = blockIdx.y;
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= rows) return;
    TO* o = out + b * obs + t * ors;
    if (t >= TT) {                         // rows added so the buffer regroups by 8 (zero = the conv's padding)
        for (int c = 0; c < Cpad; ++c) store_elem<TO>(o + c, 0.f);
        return;
    }
    float xw[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        int n = 4 * t - 8 + j;
        if (n < 0) n = -n;                 // reflect (centre = True)
        if (n >= L) n = 2 * (L - 1) - n;
        xw[j] = s[b * s_bs + n] * hann16(j);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            re = fmaf(xw[j], c_cos16[(k * j) & 15], re);
            im = fmaf(xw[j], -c_sin16[(k * j) & 15], im);
        }
        store_elem<TO>(o + k, re);
        store_elem<TO>(o + 9 + k, im);
    }
    for (int c = 18; c < Cpad; ++c) store_elem<TO>(o + c, 0.f);
}

// block = 128 consecutive frames (+3 frames of halo on the left): phase 1 one thread per frame -> 16 windowed
// time samples in shared memory; phase 2 one thread per output sample: overlap-add of up to 4 frames / envelope.
constexpr int kIstftFrames = 128;
template <bool PRECISE>
__global__ void __launch_bounds__(256) hift_istft_kernel(const float* __restrict__ x, long long xbs, long long xrs,
                                                         float* __restrict__ wav, long long wbs, int TT, int L,
                                                         float clip_mag, float limit) {
    __shared__ float fr[kIstftFrames + 3][17];
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * kIstftFrames - 3;          // first frame staged by this block
    for (int i = threadIdx.x; i < kIstftFrames + 3; i += 256) {
        const int t = t0 + i;
        if (t < 0 || t >= TT) {
#pragma unroll
            for (int m = 0; m < 16; ++m) fr[i][m] = 0.f;
            continue;
        }
        const float* xp = x + b * xbs + t * xrs;
        float re[9], im[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float mag = fminf(PRECISE ? expf(xp[k]) : __expf(xp[k]), clip_mag);
            const float ph = PRECISE ? sinf(xp[9 + k]) : __sinf(xp[9 + k]);
            float sn, cs;
            if (PRECISE) sincosf(ph, &sn, &cs);
            else __sincosf(ph, &sn, &cs);
            re[k] = mag * cs;
            im[k] = mag * sn;
        }
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            // irfft: (1/16) [Re X0 + (-1)^m Re X8 + 2 sum_{k=1..7} (Re Xk cos(2 pi k m / 16) - Im Xk sin(2 pi k m / 16))]
            float acc = re[0] + ((m & 1) ? -re[8] : re[8]);
#pragma unroll
            for (int k = 1; k < 8; ++k)
                acc += 2.0f * (re[k] * c_cos16[(k * m) & 15] - im[k] * c_sin16[(k * m) & 15]);
            fr[i][m] = acc * (1.0f / 16.0f) * hann16(m);
        }
    }
    __syncthreads();
    // output samples of this block: padded positions n' in [4 * (t0 + 3), 4 * (t0 + 3 + kIstftFrames))
    for (int i = threadIdx.x; i < 4 * kIstftFrames; i += 256) {
        const long long np = 4LL * (t0 + 3) + i;           // position in the un-trimmed signal
        const long long n = np - 8;                        // after trimming n_fft / 2
        if (n < 0 || n >= L) continue;
        const int tq = static_cast<int>(np >> 2);          // last frame covering np
        float acc = 0.f, env = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int t = tq - d;
            const int m = static_cast<int>(np - 4LL * t);  // 0..15
            if (t >= 0 && t < TT) {
                acc += fr[t - t0][m];
                const float w = hann16(m);
                env += w * w;
            }
        }
        const float y = acc / env;
        wav[b * wbs + n] = fminf(fmaxf(y, -limit), limit);
    }
}

}  // namespace svc

using namespace svc;

extern "C" int svc_unary(const float* x, long long x_bstride, long long x_rstride, void* out, long long o_bstride,
                         long long o_rstride, int B, int T, int C, int kind, float slope, const float* alpha,
                         int out_dtype, int precise, void* stream) {
    if (B < 1 || T < 1 || C < 1 || x == nullptr || out == nullptr || kind < 0 || kind > 3 ||
        (kind == 2 && alpha == nullptr) || B > 65535) {
        svc_set_error("svc_unary: bad arguments");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int esz = out_dtype == SVC_F32 ? 4 : 2;
    const bool vec = C % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && x_bstride % 4 == 0 && x_rstride % 4 == 0 &&
                     reinterpret_cast<uintptr_t>(out) % 16 == 0 && (o_bstride * esz) % 16 == 0 &&
                     (o_rstride * esz) % (esz == 4 ? 16 : 8) == 0 &&
                     (alpha == nullptr || reinterpret_cast<uintptr_t>(alpha) % 16 == 0);
    const int lanes = vec ? C / 4 : C;
    const int rows_per_pass = lanes >= 256 ? 1 : 256 / lanes;
    dim3 grid(static_cast<unsigned>(std::min<long long>((T + rows_per_pass - 1) / rows_per_pass, 148LL * 16)), B);
#define UNARY_V(TO, P, V)                                                                                          \
    unary_kernel<TO, P, V><<<grid, 256, 0, st>>>(x, x_bstride, x_rstride, static_cast<TO*>(out), o_bstride, o_rstride, \
                                                 T, C, kind, slope, alpha)
#define UNARY(TO)                                                  \
    do {                                                           \
        if (precise) { if (vec) UNARY_V(TO, true, true); else UNARY_V(TO, true, false); }    \
        else { if (vec) UNARY_V(TO, false, true); else UNARY_V(TO, false, false); }          \
    } while (0)
    if (out_dtype == SVC_F32) UNARY(float);
    else if (out_dtype == SVC_F16) UNARY(__half);
    else if (out_dtype == SVC_BF16) UNARY(__nv_bfloat16);
    else {
        svc_set_error("svc_unary: bad out_dtype");
        return SVC_ERR_ARG;
    }
#undef UNARY
#undef UNARY_V
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_hift_source(const float* f0, long long f0_bstride, const float* phase, const float* noise,
                               const float* lin_w, float lin_b, double* prefix_ws, float* out, long long out_bstride,
                               int B, int Tm, int H, int scale, float sampling_rate, float sine_amp, float noise_std,
                               float voiced_threshold, void* stream) {
    if (B < 1 || Tm < 1 || H < 1 || H > kMaxHarm || scale < 1 || f0 == nullptr || phase == nullptr ||
        lin_w == nullptr || prefix_ws == nullptr || out == nullptr || B > 65535) {
        svc_set_error("svc_hift_source: bad arguments (H <= 16)");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    hift_prefix_kernel<<<B, 32, 0, st>>>(f0, f0_bstride, prefix_ws, Tm, H, scale, sampling_rate);
    const long long L = static_cast<long long>(Tm) * scale;
    dim3 grid(static_cast<unsigned>((L + 255) / 256), B);
    hift_source_kernel<<<grid, 256, 0, st>>>(f0, f0_bstride, prefix_ws, phase, noise, lin_w, lin_b, out, out_bstride,
                                             Tm, H, scale, sampling_rate, sine_amp, noise_std, voiced_threshold);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_hift_stft(const float* s, long long s_bstride, void* out, long long o_bstride, long long o_rstride,
                             int B, int L, int rows, int Cpad, int out_dtype, void* stream) {
    if (B < 1 || L < 9 || (L % 4) != 0 || Cpad < 18 || s == nullptr || out == nullptr || B > 65535) {
        svc_set_error("svc_hift_stft: need L % 4 == 0, L >= 9 (reflect 8), Cpad >= 18");
        return SVC_ERR_ARG;
    }
    const int TT = L / 4 + 1;
    if (rows < TT) {
        svc_set_error("svc_hift_stft: rows < L / 4 + 1");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((rows + 127) / 128, B);
    if (out_dtype == SVC_F32)
        hift_stft_kernel<float><<<grid, 128, 0, st>>>(s, s_bstride, static_cast<float*>(out), o_bstride, o_rstride, L, TT,
                                                      rows, Cpad);
    else if (out_dtype == SVC_F16)
        hift_stft_kernel<__half><<<grid, 128, 0, st>>>(s, s_bstride, static_cast<__half*>(out), o_bstride, o_rstride, L,
                                                       TT, rows, Cpad);
    else
        hift_stft_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(s, s_bstride, static_cast<__nv_bfloat16*>(out), o_bstride,
                                                              o_rstride, L, TT, rows, Cpad);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_hift_istft(const float* x, long long x_bstride, long long x_rstride, float* wav, long long wav_bstride,
                              int B, int TT, float clip_mag, float audio_limit, int precise, void* stream) {
    if (B < 1 || TT < 2 || x == nullptr || wav == nullptr || B > 65535) {
        svc_set_error("svc_hift_istft: bad arguments");
        return SVC_ERR_ARG;
    }
    const int L = 4 * (TT - 1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // blocks cover padded positions [0, 4 * TT + 12); block k emits positions [4 * 128 k, 4 * 128 (k + 1))
    dim3 grid((TT + 3 + kIstftFrames - 1) / kIstftFrames, B);
    if (precise)
        hift_istft_kernel<true><<<grid, 256, 0, st>>>(x, x_bstride, x_rstride, wav, wav_bstride, TT, L, clip_mag,
                                                      audio_limit);
    else
        hift_istft_kernel<false><<<grid, 256, 0, st>>>(x, x_bstride, x_rstride, wav, wav_bstride, TT, L, clip_mag,
                                                       audio_limit);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}
