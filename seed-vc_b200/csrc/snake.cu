// Anti-aliased Snake / SnakeBeta on frames-major (B, L, C) activations, written anew for
// sm_100a (the reference's own op, anti_alias_activation_cuda.cu:44-179, is not reused).
//
// Maths (reference: alias_free_activation/torch/{act,resample,filter}.py,
// activations.py:107-119), h = 12-tap Kaiser-sinc low-pass:
//   u[m] = 2 * sum_{k, (m+5-k) even} h[k] * x[clamp((m+5-k)/2, 0, L-1)]      m in [0, 2L)
//   u[m] = u[m] + inv_b * sin(a * u[m])^2
//   y[n] = sum_{k=0..11} h[k] * u[clamp(2n+k-5, 0, 2L-1)]                     n in [0, L)
// For interior outputs, u[2n+5] and u[2n+6] both read x[n .. n+5], so one thread walks along
// time with a 6-deep x window and a 12-deep u window in registers: 1 shared-memory load, 24 FMA
// and 2 sin per output.  A block stages a (TL + 10) x CT tile (replicate-clamped rows) in
// shared memory with coalesced loads; lanes map to channels, so global and shared accesses
// are contiguous and conflict-free (row stride CT+1 spreads the time-chunks of narrow tiles
// over the banks).
#include <type_traits>

#include "common.cuh"

namespace svc {

__constant__ float c_h12[12] = {
    0.0020289648f, 0.0093894657f, -0.0255434588f, -0.0576573834f, 0.1285725832f, 0.4432097971f,
    0.4432097971f, 0.1285725832f, -0.0576573834f, -0.0255434588f, 0.0093894657f, 0.0020289648f};

// the same taps doubled: the upsampler's gain of 2 (resample.py:36) folded into the FIR (exact: a power of two)
__constant__ float c_h12x2[12] = {
    2 * 0.0020289648f, 2 * 0.0093894657f, 2 * -0.0255434588f, 2 * -0.0576573834f, 2 * 0.1285725832f, 2 * 0.4432097971f,
    2 * 0.4432097971f, 2 * 0.1285725832f, 2 * -0.0576573834f, 2 * -0.0255434588f, 2 * 0.0093894657f, 2 * 0.0020289648f};

template <bool PRECISE>
__device__ __forceinline__ float snake_fn(float u, float a, float inv_b) {
    const float s = PRECISE ? sinf(u * a) : __sinf(u * a);
    return fmaf(inv_b * s, s, u);
}

// u at (clamped) upsampled index m, reading x through `xat(l)` with l already clamped
template <bool PRECISE, typename F>
__device__ __forceinline__ float up_point(int m, int L, float a, float inv_b, F xat) {
    m = min(max(m, 0), 2 * L - 1);
    const int q = m >> 1;
    float acc = 0.f;
    if (m & 1) {
#pragma unroll
        for (int j = 0; j < 6; ++j) acc = fmaf(c_h12[2 * j], xat(min(max(q + 3 - j, 0), L - 1)), acc);
    } else {
#pragma unroll
        for (int j = 0; j < 6; ++j) acc = fmaf(c_h12[2 * j + 1], xat(min(max(q + 2 - j, 0), L - 1)), acc);
    }
    return snake_fn<PRECISE>(2.0f * acc, a, inv_b);
}

template <typename TI, typename TO, int CT, int LPT, bool PRECISE>
__global__ void __launch_bounds__(256) snake_aa_kernel(const TI* __restrict__ x, TO* __restrict__ out,
                                                       const float* __restrict__ a_p,
                                                       const float* __restrict__ invb_p, int L, int C) {
    constexpr int NCHUNK = 256 / CT;
    constexpr int TL = NCHUNK * LPT;
    constexpr int ROWS = TL + 10;
    constexpr int STRIDE = CT + 1;
    __shared__ float xs[ROWS * STRIDE];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CT;
    const int l0 = blockIdx.x * TL;
    const TI* xb = x + static_cast<long long>(b) * L * C;
    // stage rows l0-5 .. l0+TL+4, clamped to [0, L-1] (replicate padding of the upsampler)
    for (int i = threadIdx.x; i < ROWS * CT; i += 256) {
        const int r = i / CT, c = i % CT;
        const int l = min(max(l0 - 5 + r, 0), L - 1);
        float v = 0.f;
        if (c0 + c < C) v = to_f32<TI>(xb[static_cast<long long>(l) * C + c0 + c]);
        xs[r * STRIDE + c] = v;
    }
    __syncthreads();
    const int c = threadIdx.x % CT;
    const int chunk = threadIdx.x / CT;
    if (c0 + c >= C) return;
    const int n0 = l0 + chunk * LPT;
    if (n0 >= L) return;
    const float a = __ldg(a_p + c0 + c), inv_b = __ldg(invb_p + c0 + c);
    TO* ob = out + static_cast<long long>(b) * L * C + c0 + c;
    const float* col = xs + c;
    // local row of global index l is (l - l0 + 5)
    const int n_last = min(n0 + LPT, L) - 1;
    const bool interior = (2 * n0 - 5 >= 0) && (2 * n_last + 6 <= 2 * L - 1) && (n_last == n0 + LPT - 1);
    if (interior) {
        float xw[6];   // x[n .. n+5] for the current n
        float uw[12];  // u[2n-5 .. 2n+6]
        const int r0 = n0 - l0 + 5;
        // fill u[2n0-5 .. 2n0+4]: pairs (u[2q'+1], u[2q'+2]) come from x[q'-2 .. q'+3]
#pragma unroll
        for (int pz = 0; pz < 5; ++pz) {
            // pair index pz covers u[2(n0-3+pz)+1], u[2(n0-3+pz)+2]; window x[n0-5+pz .. n0+pz]
            float e = 0.f, o = 0.f;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const float xv = col[(r0 - 5 + pz + 5 - j) * STRIDE];
                e = fmaf(c_h12[2 * j], xv, e);       // odd index  m = 2q+1, q = n0-3+pz
                o = fmaf(c_h12[2 * j + 1], xv, o);   // even index m = 2q+2
            }
            uw[2 * pz] = snake_fn<PRECISE>(2.0f * e, a, inv_b);
            uw[2 * pz + 1] = snake_fn<PRECISE>(2.0f * o, a, inv_b);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) xw[j + 1] = col[(r0 + j) * STRIDE];  // x[n0 .. n0+4] -> slots 1..5
#pragma unroll
        for (int i = 0; i < LPT; ++i) {
            // shift in x[n+5]
#pragma unroll
            for (int j = 0; j < 5; ++j) xw[j] = xw[j + 1];
            xw[5] = col[(r0 + i + 5) * STRIDE];
            float e = 0.f, o = 0.f;  // u[2n+5] (odd, q=n+2), u[2n+6] (even, q=n+3): x[n+5-j]
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                e = fmaf(c_h12[2 * j], xw[5 - j], e);
                o = fmaf(c_h12[2 * j + 1], xw[5 - j], o);
            }
            uw[10] = snake_fn<PRECISE>(2.0f * e, a, inv_b);
            uw[11] = snake_fn<PRECISE>(2.0f * o, a, inv_b);
            float y = 0.f;
#pragma unroll
            for (int k = 0; k < 12; ++k) y = fmaf(c_h12[k], uw[k], y);
            ob[static_cast<long long>(n0 + i) * C] = from_f32<TO>(y);
#pragma unroll
            for (int k = 0; k < 10; ++k) uw[k] = uw[k + 2];
        }
    } else {
        auto xat = [&](int l) { return col[(l - l0 + 5) * STRIDE]; };
        for (int n = n0; n <= n_last; ++n) {
            float y = 0.f;
#pragma unroll
            for (int k = 0; k < 12; ++k)
                y = fmaf(c_h12[k], up_point<PRECISE>(2 * n + k - 5, L, a, inv_b, xat), y);
            ob[static_cast<long long>(n) * C] = from_f32<TO>(y);
        }
    }
}

// v2: two adjacent channels per thread with packed fp32x2 arithmetic (FFMA2 halves the FIR
// instruction count), float4 / 8-byte global accesses.  Same maths and tiling idea as above.
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }

template <bool PRECISE>
__device__ __forceinline__ float2 snake_fn2(float2 u, float2 a, float2 inv_b) {
    const float sx = PRECISE ? sinf(u.x * a.x) : __sinf(u.x * a.x);
    const float sy = PRECISE ? sinf(u.y * a.y) : __sinf(u.y * a.y);
    return make_float2(fmaf(inv_b.x * sx, sx, u.x), fmaf(inv_b.y * sy, sy, u.y));
}

template <typename TI>
__device__ __forceinline__ float4 load4(const TI* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&q.x);
    const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&q.y);
    return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}

template <>
__device__ __forceinline__ float4 load4<__half>(const __half* p) {
    const uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&q.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&q.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <typename TI, typename TO, int CT, int LPT, bool PRECISE>
__global__ void __launch_bounds__(256) snake_aa2_kernel(const TI* __restrict__ x, TO* __restrict__ out,
                                                        const float* __restrict__ a_p,
                                                        const float* __restrict__ invb_p, int L, int C) {
    constexpr int LANES = CT / 2;
    constexpr int NCHUNK = 256 / LANES;
    constexpr int TL = NCHUNK * LPT;
    constexpr int ROWS = TL + 10;
    constexpr int STRIDE = CT + 2;          // even (float2 alignment), spreads chunks over banks
    extern __shared__ float xs2[];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CT;
    const int l0 = blockIdx.x * TL;
    const TI* xb = x + static_cast<long long>(b) * L * C;
    constexpr int V4 = CT / 4;
    for (int i = threadIdx.x; i < ROWS * V4; i += 256) {
        const int r = i / V4, c4 = (i % V4) * 4;
        const int l = min(max(l0 - 5 + r, 0), L - 1);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + c4 < C) v = load4<TI>(xb + static_cast<long long>(l) * C + c0 + c4);
        float* d = xs2 + r * STRIDE + c4;
        *reinterpret_cast<float2*>(d) = make_float2(v.x, v.y);
        *reinterpret_cast<float2*>(d + 2) = make_float2(v.z, v.w);
    }
    __syncthreads();
    const int lane2 = threadIdx.x % LANES;
    const int chunk = threadIdx.x / LANES;
    const int c = c0 + 2 * lane2;
    if (chunk >= NCHUNK || c >= C) return;
    const int n0 = l0 + chunk * LPT;
    if (n0 >= L) return;
    const float2 a = __ldg(reinterpret_cast<const float2*>(a_p + c));
    const float2 inv_b = __ldg(reinterpret_cast<const float2*>(invb_p + c));
    TO* ob = out + static_cast<long long>(b) * L * C + c;
    const float* col = xs2 + 2 * lane2;
    auto ld = [&](int row) { return *reinterpret_cast<const float2*>(col + row * STRIDE); };
    auto st = [&](int n, float2 y) {
        if constexpr (sizeof(TO) == 4) {
            *reinterpret_cast<float2*>(ob + static_cast<long long>(n) * C) = y;
        } else {
            *reinterpret_cast<uint32_t*>(ob + static_cast<long long>(n) * C) = pack2<TO>(y.x, y.y);
        }
    };
    const int n_last = min(n0 + LPT, L) - 1;
    const bool interior = (2 * n0 - 5 >= 0) && (2 * n_last + 6 <= 2 * L - 1) && (n_last == n0 + LPT - 1);
    if (interior) {
        float2 xw[6], uw[12];
        const int r0 = n0 - l0 + 5;
#pragma unroll
        for (int pz = 0; pz < 5; ++pz) {
            float2 e = splat2(0.f), o = splat2(0.f);
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const float2 xv = ld(r0 + pz - j);
                e = __ffma2_rn(splat2(c_h12x2[2 * j]), xv, e);
                o = __ffma2_rn(splat2(c_h12x2[2 * j + 1]), xv, o);
            }
            uw[2 * pz] = snake_fn2<PRECISE>(e, a, inv_b);
            uw[2 * pz + 1] = snake_fn2<PRECISE>(o, a, inv_b);
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) xw[j + 1] = ld(r0 + j);
#pragma unroll
        for (int i = 0; i < LPT; ++i) {
#pragma unroll
            for (int j = 0; j < 5; ++j) xw[j] = xw[j + 1];
            xw[5] = ld(r0 + i + 5);
            float2 e = splat2(0.f), o = splat2(0.f);
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                e = __ffma2_rn(splat2(c_h12x2[2 * j]), xw[5 - j], e);
                o = __ffma2_rn(splat2(c_h12x2[2 * j + 1]), xw[5 - j], o);
            }
            uw[10] = snake_fn2<PRECISE>(e, a, inv_b);
            uw[11] = snake_fn2<PRECISE>(o, a, inv_b);
            float2 y = splat2(0.f);
#pragma unroll
            for (int k = 0; k < 12; ++k) y = __ffma2_rn(splat2(c_h12[k]), uw[k], y);
            st(n0 + i, y);
#pragma unroll
            for (int k = 0; k < 10; ++k) uw[k] = uw[k + 2];
        }
    } else {
        for (int n = n0; n <= n_last; ++n) {
            float y0 = 0.f, y1 = 0.f;
            auto xat0 = [&](int l) { return col[(l - l0 + 5) * STRIDE]; };
            auto xat1 = [&](int l) { return col[(l - l0 + 5) * STRIDE + 1]; };
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                y0 = fmaf(c_h12[k], up_point<PRECISE>(2 * n + k - 5, L, a.x, inv_b.x, xat0), y0);
                y1 = fmaf(c_h12[k], up_point<PRECISE>(2 * n + k - 5, L, a.y, inv_b.y, xat1), y1);
            }
            st(n, make_float2(y0, y1));
        }
    }
}

// v3 (16-bit output, non-PRECISE): both FIRs on the tensor cores as banded-Toeplitz warp MMAs, so the FMA
// pipe is left with the Snake nonlinearity only (the v2 kernel above is FMA-pipe bound at 0.41 of HBM peak).
// Phases of the 2x signal: uo[q] = u[2q+1] = 2 sum_j h[2j] x[q+3-j], ue[q] = u[2q] = 2 sum_j h[2j+1] x[q+2-j],
// y[n] = sum_j h[2j] uo[n-3+j] + sum_j h[2j+1] ue[n-2+j]   (j = 0..5; same maths as the header comment).
// One warp owns 16 channels and walks along time in tiles of 8 frames.  The DATA is the A operand
// (M = channels, K = 16 frames), the constant Toeplitz band is the B operand (K = 16 frames in, N = 8
// frames out; 13 of the 16 K slots carry taps) and lives in registers for the whole kernel:
//   u tile j (q in [N0-3+8j, +8)), both phases, from ONE A fragment = x frames [N0-6+8j, +16)
//       (ldmatrix.trans from the fp16 (frames, channels) smem tile);
//   Snake in fp32 on the accumulator fragment; the m16n8 accumulator layout IS the A-fragment layout, so
//       two consecutive u tiles pack (cvt.f16x2) straight into the A operand of the down-FIR MMA;
//   y tile (n in [N0+8t, +8)) = uo tiles (t, t+1) x Bd_o + ue tiles (t, t+1) x Bd_e.
// The taps are split hi + lo in fp16 (two MMAs per product) so the filter response is exact to 2^-22;
// x and u enter the MMAs as fp16 (2^-11 relative; the output is rounded to a 16-bit operand anyway).
// Sequence ends: x rows are staged replicate-clamped; u values outside [0, 2L) are replaced by u[0] /
// u[2L-1] (post-Snake, like the reference's replicate pad of the activated signal) with warp shuffles.
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Toeplitz entry B[k][n] of the four band matrices (kind 0: up odd, 1: up even, 2: down odd, 3: down even)
__device__ __forceinline__ float snake_band(int kind, int k, int n) {
    int j;
    switch (kind) {
        case 0: j = n + 6 - k; return (j >= 0 && j <= 5) ? c_h12x2[2 * j] : 0.f;
        case 1: j = n + 5 - k; return (j >= 0 && j <= 5) ? c_h12x2[2 * j + 1] : 0.f;
        case 2: j = k - n; return (j >= 0 && j <= 5) ? c_h12[2 * j] : 0.f;
        default: j = k - n - 1; return (j >= 0 && j <= 5) ? c_h12[2 * j + 1] : 0.f;
    }
}

// packed bf16 pair -> packed IEEE half pair
__device__ __forceinline__ uint32_t bf162_to_f162(uint32_t v) {
    const __nv_bfloat162 p = *reinterpret_cast<const __nv_bfloat162*>(&v);
    return pack_f16(__low2float(p), __high2float(p));
}

// (64 registers per thread: 4 resident blocks of 256 threads, 5 of 192)
template <typename TI, typename TO, int CT, int NS, int SPAN, bool SPLIT>
__global__ void __launch_bounds__(32 * ((CT + 15) / 16) * NS, 65536 / (64 * 32 * ((CT + 15) / 16) * NS))
snake_mma_kernel(const TI* __restrict__ x, TO* __restrict__ out, const float* __restrict__ a_p,
                 const float* __restrict__ invb_p, int L, int C) {
    constexpr int MT = (CT + 15) / 16;          // 16-channel MMA row tiles per block
    constexpr int CTP = MT * 16;                // staged channels (zero beyond CT)
    constexpr int NTHR = 32 * MT * NS;
    constexpr int TL = NS * SPAN;               // output frames per block
    constexpr int XROWS = TL + 16;              // x frames l0-6 .. l0+TL+9
    constexpr int XSTR = CTP * 2 + 16;          // bytes; +16 keeps the 8 ldmatrix rows on distinct banks
    constexpr int OSTR = 48;                    // per-warp output staging: 8 frames x 32 B, conflict-free stride
    extern __shared__ __align__(16) unsigned char sm3[];
    unsigned char* xs = sm3;
    unsigned char* os = sm3 + XROWS * XSTR;
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CT;
    const int l0 = blockIdx.x * TL;
    const TI* xb = x + static_cast<long long>(b) * L * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = warp % MT, ns = warp / MT;
    const int g = lane >> 2, t = lane & 3;
    const int N0 = l0 + ns * SPAN;
    float a_l, ib_l, a_h, ib_h;
    uint32_t bh[4][2], bl[4][2];
    // per-thread constants; called between the first batch of tile loads and its use, so the arithmetic and the
    // parameter loads overlap the DRAM round trip
    auto setup = [&]() {
        const int cl = c0 + mt * 16 + g, ch = cl + 8;
        const bool vl = mt * 16 + g < CT && cl < C, vh = mt * 16 + g + 8 < CT && ch < C;
        a_l = vl ? __ldg(a_p + cl) : 0.f, ib_l = vl ? __ldg(invb_p + cl) : 0.f;
        a_h = vh ? __ldg(a_p + ch) : 0.f, ib_h = vh ? __ldg(invb_p + ch) : 0.f;
        // constant band fragments: b0 = B[2t..2t+1][g], b1 = B[2t+8..2t+9][g]; hi / lo halves of every tap
#pragma unroll
        for (int kind = 0; kind < 4; ++kind)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float v0 = snake_band(kind, 2 * t + 8 * r, g), v1 = snake_band(kind, 2 * t + 8 * r + 1, g);
                const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
                bh[kind][r] = static_cast<uint32_t>(__half_as_ushort(h0)) | (static_cast<uint32_t>(__half_as_ushort(h1)) << 16);
                bl[kind][r] = pack_f16(v0 - __half2float(h0), v1 - __half2float(h1));
            }
    };
    if constexpr (sizeof(TI) == 4) {
        // stage: global fp32 (frames, C) -> fp16 (frames, CTP), rows clamped to the sequence.  Every thread owns one
        // 4-channel column chunk; loads are issued in batches (the whole tile in two rounds of latency)
        constexpr int V4 = CTP / 4, RPI = NTHR / V4, NIT = (XROWS + RPI - 1) / RPI;
        constexpr int BATCH = NIT <= 18 ? (NIT + 1) / 2 : 8;
        static_assert(NTHR % V4 == 0, "tile shape");
        const int c4 = (threadIdx.x % V4) * 4, rb = threadIdx.x / V4;
        const bool cv = c4 < CT && c0 + c4 < C;
        const TI* xc = xb + c0 + c4;
#pragma unroll
        for (int i0 = 0; i0 < NIT; i0 += BATCH) {
            float4 v[BATCH];
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                const int r = rb + (i0 + k) * RPI;
                const int l = min(max(l0 - 6 + r, 0), L - 1);
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i0 + k < NIT && r < XROWS && cv) v[k] = __ldg(reinterpret_cast<const float4*>(xc + static_cast<long long>(l) * C));
            }
            if (i0 == 0) setup();
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                const int r = rb + (i0 + k) * RPI;
                if (i0 + k < NIT && r < XROWS)
                    *reinterpret_cast<uint2*>(xs + r * XSTR + c4 * 2) = make_uint2(pack_f16(v[k].x, v[k].y), pack_f16(v[k].z, v[k].w));
            }
        }
    } else {
        // 16-bit input: 8-channel (16-byte) chunks, the whole tile in flight at once; a half input is copied as it is
        constexpr int V8 = CTP / 8, RPI = NTHR / V8, NIT = (XROWS + RPI - 1) / RPI;
        constexpr int BATCH = NIT <= 10 ? NIT : (NIT + 1) / 2;
        static_assert(NTHR % V8 == 0, "tile shape");
        const int c8 = (threadIdx.x % V8) * 8, rb = threadIdx.x / V8;
        const bool cv = c8 < CT && c0 + c8 < C;
        const TI* xc = xb + c0 + c8;
#pragma unroll
        for (int i0 = 0; i0 < NIT; i0 += BATCH) {
            uint4 v[BATCH];
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                const int r = rb + (i0 + k) * RPI;
                const int l = min(max(l0 - 6 + r, 0), L - 1);
                v[k] = make_uint4(0u, 0u, 0u, 0u);
                if (i0 + k < NIT && r < XROWS && cv) v[k] = __ldg(reinterpret_cast<const uint4*>(xc + static_cast<long long>(l) * C));
            }
            if (i0 == 0) setup();
#pragma unroll
            for (int k = 0; k < BATCH; ++k) {
                const int r = rb + (i0 + k) * RPI;
                if constexpr (std::is_same<TI, __nv_bfloat16>::value) {
                    v[k].x = bf162_to_f162(v[k].x); v[k].y = bf162_to_f162(v[k].y);
                    v[k].z = bf162_to_f162(v[k].z); v[k].w = bf162_to_f162(v[k].w);
                }
                if (i0 + k < NIT && r < XROWS) *reinterpret_cast<uint4*>(xs + r * XSTR + c8 * 2) = v[k];
            }
        }
    }
    __syncthreads();
    if (N0 >= L) return;                        // no block-level barrier below this line
    // ldmatrix.x4.trans row address of this lane: matrices (ch 0-7 | 8-15) x (frames 0-7 | 8-15)
    const int lm = lane >> 3, lr = lane & 7;
    uint32_t xaddr = smem_u32(xs) + (ns * SPAN + lr + (lm >> 1) * 8) * XSTR + (mt * 16 + (lm & 1) * 8) * 2;
    // per-warp output staging, double buffered: stmatrix.x2.trans rows in, one 8-byte chunk per lane out
    const uint32_t obase = smem_u32(os) + warp * (2 * 8 * OSTR);
    const uint32_t oaddr_st = obase + (lane & 7) * OSTR + ((lane >> 3) & 1) * 16;
    const uint32_t oaddr_ld = obase + g * OSTR + t * 8;
    // copy-out: lane = (frame g, channels 4t .. 4t+3)
    TO* op = out + (static_cast<long long>(b) * L + N0 + g) * C + c0 + mt * 16 + t * 4;
    const long long ostep = 8LL * C;
    const bool ovalid = mt * 16 + t * 4 < CT && c0 + mt * 16 + t * 4 < C;

    const float2 a2_l = make_float2(a_l, a_l), ib2_l = make_float2(ib_l, ib_l);
    const float2 a2_h = make_float2(a_h, a_h), ib2_h = make_float2(ib_h, ib_h);
    float vLl = 0.f, vLh = 0.f;                 // u[2L-1] of this thread's two channels, once seen
    uint32_t po[2], pe[2];                      // previous u tile, packed: [0] = channel g, [1] = channel g+8

    // one u tile (both phases) from the A fragment at xaddr; EDGE: a sequence end may lie inside (qs = q of column 0)
    auto u_tile = [&](auto edge, int qs, uint32_t (&no)[2], uint32_t (&ne)[2]) {
        constexpr bool EDGE = decltype(edge)::value;
        uint32_t xa[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(xa[0]), "=r"(xa[1]), "=r"(xa[2]), "=r"(xa[3])
                     : "r"(xaddr));
        xaddr += 8 * XSTR;
        float uo[4] = {0.f, 0.f, 0.f, 0.f}, ue[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(uo, xa, bh[0][0], bh[0][1]);
        mma16816(ue, xa, bh[1][0], bh[1][1]);
        if constexpr (SPLIT) {
            mma16816(uo, xa, bl[0][0], bl[0][1]);
            mma16816(ue, xa, bl[1][0], bl[1][1]);
        }
#ifdef SVC_SNAKE_SCALAR
        uo[0] = snake_fn<false>(uo[0], a_l, ib_l); uo[1] = snake_fn<false>(uo[1], a_l, ib_l);
        uo[2] = snake_fn<false>(uo[2], a_h, ib_h); uo[3] = snake_fn<false>(uo[3], a_h, ib_h);
        ue[0] = snake_fn<false>(ue[0], a_l, ib_l); ue[1] = snake_fn<false>(ue[1], a_l, ib_l);
        ue[2] = snake_fn<false>(ue[2], a_h, ib_h); ue[3] = snake_fn<false>(ue[3], a_h, ib_h);
#else
        // packed f32x2 arithmetic around the two scalar sin.approx of a pair (both values of a pair belong to the
        // same channel): 7 issue slots per pair instead of 10 - the loop is short of issue slots, not of pipe cycles
        auto snake2 = [](float& x, float& y, float2 a2, float2 ib2) {
            const float2 u = make_float2(x, y);
            const float2 t = __fmul2_rn(u, a2);
            const float2 sn = make_float2(__sinf(t.x), __sinf(t.y));
            const float2 r = __ffma2_rn(__fmul2_rn(ib2, sn), sn, u);
            x = r.x, y = r.y;
        };
        snake2(uo[0], uo[1], a2_l, ib2_l); snake2(uo[2], uo[3], a2_h, ib2_h);
        snake2(ue[0], ue[1], a2_l, ib2_l); snake2(ue[2], ue[3], a2_h, ib2_h);
#endif
        if constexpr (EDGE) {
            if (qs < 0 || qs + 8 >= L) {        // warp-uniform: frame 0 / frame L-1 or beyond inside this tile
                if (qs < 0) {                   // q < 0 -> u[0] = ue[q = 0], column -qs
                    const int n0c = -qs, src = (lane & ~3) | (n0c >> 1);
                    const float sl = __shfl_sync(0xffffffffu, (n0c & 1) ? ue[1] : ue[0], src);
                    const float sh = __shfl_sync(0xffffffffu, (n0c & 1) ? ue[3] : ue[2], src);
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (qs + 2 * t + e < 0) { uo[e] = sl; ue[e] = sl; uo[2 + e] = sh; ue[2 + e] = sh; }
                }
                if (qs <= L - 1 && L - 1 < qs + 8) {   // this tile holds q = L-1: u[2L-1] = uo[L-1]
                    const int nLc = L - 1 - qs, src = (lane & ~3) | (nLc >> 1);
                    vLl = __shfl_sync(0xffffffffu, (nLc & 1) ? uo[1] : uo[0], src);
                    vLh = __shfl_sync(0xffffffffu, (nLc & 1) ? uo[3] : uo[2], src);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (qs + 2 * t + e >= L) { uo[e] = vLl; ue[e] = vLl; uo[2 + e] = vLh; ue[2 + e] = vLh; }
            }
        }
        no[0] = pack_f16(uo[0], uo[1]); no[1] = pack_f16(uo[2], uo[3]);
        ne[0] = pack_f16(ue[0], ue[1]); ne[1] = pack_f16(ue[2], ue[3]);
    };
    // y tile from the previous and the new u tile; (channel, frame) fragment -> (frame, channel) rows through the
    // warp's staging tile, then 8-byte stores: one full 32-byte sector per frame and warp
    auto y_tile = [&](const uint32_t (&no)[2], const uint32_t (&ne)[2], int buf, bool store) {
        const uint32_t ao[4] = {po[0], po[1], no[0], no[1]}, ae[4] = {pe[0], pe[1], ne[0], ne[1]};
        float y[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(y, ao, bh[2][0], bh[2][1]);
        mma16816(y, ae, bh[3][0], bh[3][1]);
        if constexpr (SPLIT) {
            mma16816(y, ao, bl[2][0], bl[2][1]);
            mma16816(y, ae, bl[3][0], bl[3][1]);
        }
        po[0] = no[0]; po[1] = no[1]; pe[0] = ne[0]; pe[1] = ne[1];
        const uint32_t y0 = pack2<TO>(y[0], y[1]), y1 = pack2<TO>(y[2], y[3]);
        asm volatile("stmatrix.sync.aligned.m8n8.x2.trans.shared.b16 [%0], {%1,%2};" ::"r"(oaddr_st + buf * (8 * OSTR)),
                     "r"(y0), "r"(y1)
                     : "memory");
        __syncwarp();
        uint2 v;
        asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(oaddr_ld + buf * (8 * OSTR)) : "memory");
        if (store) *reinterpret_cast<uint2*>(op) = v;
        op += ostep;
    };

    // interior span: every u tile of the span lies inside [0, L) and every output frame exists
    if (N0 >= 3 && N0 + SPAN + 5 <= L) {
        u_tile(std::false_type{}, 0, po, pe);
#pragma unroll 4
        for (int tt = 0; tt < SPAN / 8; ++tt) {
            uint32_t no[2], ne[2];
            u_tile(std::false_type{}, 0, no, ne);
            y_tile(no, ne, tt & 1, ovalid);
        }
    } else {
        int qs = N0 - 3;
        u_tile(std::true_type{}, qs, po, pe);
        for (int tt = 0; tt < SPAN / 8; ++tt) {
            const int n0 = N0 + 8 * tt;
            if (n0 >= L) break;
            qs += 8;
            uint32_t no[2], ne[2];
            u_tile(std::true_type{}, qs, no, ne);
            y_tile(no, ne, tt & 1, ovalid && n0 + g < L);
        }
    }
}

constexpr int kPostTL = 128;    // output samples per block of snake_conv_post_kernel

// activation_post + conv_post (C -> 1, k taps, zero padding) + clamp / tanh.
// Reference: modules/bigvgan/bigvgan.py:377-384.
// One block = TL output samples.  Four phases through shared memory so every upsampled Snake value
// (6 FMA + 1 sin) is computed once: x tile -> u (2x rate) -> y (down FIR) -> conv over k taps x C.
// Threads are (row = tid / 32, channel = tid % 32): no integer division by the runtime C.
template <bool PRECISE>
__global__ void __launch_bounds__(256) snake_conv_post_kernel(
    const float* __restrict__ x, const float* __restrict__ a_p, const float* __restrict__ invb_p,
    const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out, int L, int C,
    int ksize, int use_tanh) {
    constexpr int TL = kPostTL;
    extern __shared__ float sm[];
    const int half = ksize / 2;
    const int CS = C + 1;                     // odd row stride: conflict-free column walks
    const int YR = TL + 2 * half;             // activation rows needed by the conv
    const int UR = 2 * YR + 10;               // upsampled rows needed by the down FIR
    const int XR = YR + 13;                   // x rows needed by the up FIR
    float* xs = sm;                           // XR * CS
    float* us = xs + XR * CS;                 // UR * CS
    float* ys = us + UR * CS;                 // YR * CS
    float* ws = ys + YR * CS;                 // ksize * C
    const int b = blockIdx.y;
    const int l0 = blockIdx.x * TL;
    const int lbase = l0 - half - 6;          // x row 0
    const int m0 = 2 * (l0 - half) - 5;       // u row 0
    const int tc = threadIdx.x & 31, tr = threadIdx.x >> 5;
    const bool interior = lbase >= 0 && lbase + XR <= L;
    const float* xb = x + static_cast<long long>(b) * L * C;
    for (int r = tr; r < XR; r += 8) {
        const int l = min(max(lbase + r, 0), L - 1);
        for (int c = tc; c < C; c += 32) xs[r * CS + c] = xb[static_cast<long long>(l) * C + c];
    }
    for (int i = threadIdx.x; i < ksize * C; i += 256) ws[i] = w[i];
    __syncthreads();
    for (int c = tc; c < C; c += 32) {
        const float a = a_p[c], inv_b = invb_p[c];
        for (int i = tr; i < UR; i += 8) {
            const int m = min(max(m0 + i, 0), 2 * L - 1);
            const int q = m >> 1;
            // odd m: taps h[0,2,..] on x[q+3-j]; even m: taps h[1,3,..] on x[q+2-j]  (j = 0..5)
            const int odd = m & 1;
            const float* xp = xs + (q + 2 + odd - lbase) * CS + c;
            float acc = 0.f;
            if (interior) {                   // no sequence end inside this block: taps are in range
#pragma unroll
                for (int j = 0; j < 6; ++j) acc = fmaf(odd ? c_h12[2 * j] : c_h12[2 * j + 1], xp[-j * CS], acc);
            } else {
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int l = min(max(q + 2 + odd - j, 0), L - 1) - (q + 2 + odd);   // clamp at the ends
                    acc = fmaf(odd ? c_h12[2 * j] : c_h12[2 * j + 1], xp[l * CS], acc);
                }
            }
            us[i * CS + c] = snake_fn<PRECISE>(2.0f * acc, a, inv_b);
        }
    }
    __syncthreads();
    for (int c = tc; c < C; c += 32) {
        for (int r = tr; r < YR; r += 8) {
            const int n = l0 - half + r;
            float y = 0.f;
            if (n >= 0 && n < L) {
                const float* up = us + (2 * r) * CS + c;
#pragma unroll
                for (int k = 0; k < 12; ++k) y = fmaf(c_h12[k], up[k * CS], y);
            }
            ys[r * CS + c] = y;
        }
    }
    __syncthreads();
    const int n = l0 + threadIdx.x;
    if (threadIdx.x < TL && n < L) {
        float acc = bias != nullptr ? bias[0] : 0.f;
        for (int k = 0; k < ksize; ++k) {
            const float* yp = ys + (threadIdx.x + k) * CS;
            const float* wp = ws + k * C;
            for (int c = 0; c < C; ++c) acc = fmaf(wp[c], yp[c], acc);
        }
        acc = use_tanh ? tanhf(acc) : fminf(fmaxf(acc, -1.0f), 1.0f);
        out[static_cast<long long>(b) * L + n] = acc;
    }
}

// conv_post on an already activated 16-bit (B, L, C) tensor (the tensor-core Snake's output): C -> 1 channel,
// ksize taps, zero padding, clamp / tanh.  One thread per output sample; the block stages its (TL + ksize - 1) x C
// rows with 16-byte loads and every thread walks its ksize rows with 16-byte shared-memory reads (row stride
// C * 2 = 48 bytes for C = 24: conflict-free per quarter warp).  Reference: modules/bigvgan/bigvgan.py:380-384.
template <typename TI>
__global__ void __launch_bounds__(256) conv_post_kernel(const TI* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias, float* __restrict__ out, int L,
                                                        int C, int ksize, int use_tanh) {
    constexpr int TL = 256;
    extern __shared__ __align__(16) unsigned char cps[];
    const int half = ksize / 2;
    const int rows = TL + 2 * half;
    const int rb = C * 2;                                  // bytes per row, multiple of 16 (C % 8 == 0)
    float* ws = reinterpret_cast<float*>(cps + ((rows * rb + 15) & ~15));
    const int b = blockIdx.y;
    const int l0 = blockIdx.x * TL;
    const unsigned char* xb = reinterpret_cast<const unsigned char*>(x + static_cast<long long>(b) * L * C);
    const int chunks = rb / 16;
    for (int i = threadIdx.x; i < rows * chunks; i += 256) {
        const int r = i / chunks, c = i - r * chunks;
        const int l = l0 - half + r;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);              // zero padding of the conv
        if (l >= 0 && l < L) v = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(l) * rb) + c);
        *reinterpret_cast<uint4*>(cps + r * rb + c * 16) = v;
    }
    for (int i = threadIdx.x; i < ksize * C; i += 256) ws[i] = w[i];
    __syncthreads();
    const int n = l0 + threadIdx.x;
    if (n >= L) return;
    float acc = bias != nullptr ? bias[0] : 0.f;
    for (int k = 0; k < ksize; ++k) {
        const uint4* row = reinterpret_cast<const uint4*>(cps + (threadIdx.x + k) * rb);
        const float* wk = ws + k * C;
        for (int c = 0; c < chunks; ++c) {
            const uint4 q = row[c];
            const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float lo, hi;
                if constexpr (std::is_same<TI, __half>::value) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[j]));
                    lo = f.x, hi = f.y;
                } else {
                    lo = __uint_as_float(u[j] << 16), hi = __uint_as_float(u[j] & 0xffff0000u);
                }
                acc = fmaf(wk[c * 8 + 2 * j], lo, acc);
                acc = fmaf(wk[c * 8 + 2 * j + 1], hi, acc);
            }
        }
    }
    acc = use_tanh ? tanhf(acc) : fminf(fmaxf(acc, -1.0f), 1.0f);
    out[static_cast<long long>(b) * L + n] = acc;
}

template <typename TI, typename TO, int CT, int LPT, bool PRECISE>
static void launch_snake2(const TI* xi, TO* o, const float* a, const float* inv_b, int B, int L, int C,
                          cudaStream_t st) {
    constexpr int LANES = CT / 2, NCHUNK = 256 / LANES, TL = NCHUNK * LPT;
    constexpr int SMEM = (TL + 10) * (CT + 2) * 4;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(snake_aa2_kernel<TI, TO, CT, LPT, PRECISE>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        attr = true;
    }
    dim3 grid((L + TL - 1) / TL, (C + CT - 1) / CT, B);
    snake_aa2_kernel<TI, TO, CT, LPT, PRECISE><<<grid, 256, SMEM, st>>>(xi, o, a, inv_b, L, C);
}

#ifndef SVC_SNAKE_NS        // time spans per block / frames per span of the 64-channel tile (experiment builds override)
#define SVC_SNAKE_NS 2
#define SVC_SNAKE_SPAN 128
#endif
#ifdef SVC_SNAKE_NOSPLIT
constexpr bool kSnakeSplitTaps = false;     // experiment build: single fp16 tap (filter response exact to 2^-11 only)
#else
constexpr bool kSnakeSplitTaps = true;
#endif

template <typename TI, typename TO, int CT, int NS, int SPAN>
static int launch_snake3(const TI* xi, TO* o, const float* a, const float* inv_b, int B, int L, int C,
                         cudaStream_t st) {
    constexpr int MT = (CT + 15) / 16, TL = NS * SPAN;
    constexpr int SMEM = (TL + 16) * (MT * 32 + 16) + MT * NS * 2 * 8 * 48;
    auto kern = snake_mma_kernel<TI, TO, CT, NS, SPAN, kSnakeSplitTaps>;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        attr = true;
    }
    dim3 grid((L + TL - 1) / TL, (C + CT - 1) / CT, B);
    kern<<<grid, 32 * MT * NS, SMEM, st>>>(xi, o, a, inv_b, L, C);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

template <typename TI, typename TO, bool PRECISE>
static int launch_snake(const void* x, void* out, const float* a, const float* inv_b, int B, int L,
                        int C, cudaStream_t st) {
    const TI* xi = static_cast<const TI*>(x);
    TO* o = static_cast<TO*>(out);
    const bool al = reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 8 == 0 &&
                    reinterpret_cast<uintptr_t>(a) % 8 == 0 && reinterpret_cast<uintptr_t>(inv_b) % 8 == 0;
    if constexpr (!PRECISE && sizeof(TO) == 2) {
        // 16-bit operand output: FIRs on the tensor cores (snake_mma_kernel)
        if (al && C % 8 == 0) {
            if (C % 64 == 0) return launch_snake3<TI, TO, 64, SVC_SNAKE_NS, SVC_SNAKE_SPAN>(xi, o, a, inv_b, B, L, C, st);
            if (C % 48 == 0) return launch_snake3<TI, TO, 48, 2, 128>(xi, o, a, inv_b, B, L, C, st);
            if (C % 32 == 0) return launch_snake3<TI, TO, 32, 4, 64>(xi, o, a, inv_b, B, L, C, st);
            if (C % 24 == 0) return launch_snake3<TI, TO, 24, 4, 64>(xi, o, a, inv_b, B, L, C, st);
            if (C % 16 == 0) return launch_snake3<TI, TO, 16, 8, 32>(xi, o, a, inv_b, B, L, C, st);
        }
    }
    if (al && C % 64 == 0) {
        launch_snake2<TI, TO, 64, 16, PRECISE>(xi, o, a, inv_b, B, L, C, st);
    } else if (al && C % 32 == 0) {
        launch_snake2<TI, TO, 32, 16, PRECISE>(xi, o, a, inv_b, B, L, C, st);
    } else if (al && C % 48 == 0) {
        launch_snake2<TI, TO, 48, 16, PRECISE>(xi, o, a, inv_b, B, L, C, st);
    } else if (al && C % 24 == 0) {
        launch_snake2<TI, TO, 24, 16, PRECISE>(xi, o, a, inv_b, B, L, C, st);
    } else if (C % 32 == 0 || C > 64) {
        constexpr int CT = 32, LPT = 16, TL = (256 / CT) * LPT;
        dim3 grid((L + TL - 1) / TL, (C + CT - 1) / CT, B);
        snake_aa_kernel<TI, TO, CT, LPT, PRECISE><<<grid, 256, 0, st>>>(xi, o, a, inv_b, L, C);
    } else if (C % 16 == 0) {
        constexpr int CT = 16, LPT = 16, TL = (256 / CT) * LPT;
        dim3 grid((L + TL - 1) / TL, (C + CT - 1) / CT, B);
        snake_aa_kernel<TI, TO, CT, LPT, PRECISE><<<grid, 256, 0, st>>>(xi, o, a, inv_b, L, C);
    } else {
        constexpr int CT = 8, LPT = 8, TL = (256 / CT) * LPT;
        dim3 grid((L + TL - 1) / TL, (C + CT - 1) / CT, B);
        snake_aa_kernel<TI, TO, CT, LPT, PRECISE><<<grid, 256, 0, st>>>(xi, o, a, inv_b, L, C);
    }
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

}  // namespace svc

using namespace svc;

extern "C" int svc_snake_aa(const void* x, int x_dtype, void* out, int out_dtype, const float* a,
                            const float* inv_b, int B, int L, int C, int precise, void* stream) {
    if (B < 1 || L < 1 || C < 1 || B > 65535) {
        svc_set_error("svc_snake_aa: bad shape");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define SNAKE_DISPATCH(TI, TO) return launch_snake<TI, TO, false>(x, out, a, inv_b, B, L, C, st)
    if (precise) {   // libm sinf: the fp32 parity mode only
        if (x_dtype == SVC_F32 && out_dtype == SVC_F32) return launch_snake<float, float, true>(x, out, a, inv_b, B, L, C, st);
        svc_set_error("svc_snake_aa: precise = 1 needs fp32 input and output");
        return SVC_ERR_UNSUPPORTED;
    }
    if (x_dtype == SVC_F32 && out_dtype == SVC_F32) { SNAKE_DISPATCH(float, float); }
    if (x_dtype == SVC_F32 && out_dtype == SVC_BF16) { SNAKE_DISPATCH(float, __nv_bfloat16); }
    if (x_dtype == SVC_BF16 && out_dtype == SVC_BF16) { SNAKE_DISPATCH(__nv_bfloat16, __nv_bfloat16); }
    if (x_dtype == SVC_BF16 && out_dtype == SVC_F32) { SNAKE_DISPATCH(__nv_bfloat16, float); }
    if (x_dtype == SVC_F32 && out_dtype == SVC_F16) { SNAKE_DISPATCH(float, __half); }
    if (x_dtype == SVC_F16 && out_dtype == SVC_F16) { SNAKE_DISPATCH(__half, __half); }
    if (x_dtype == SVC_F16 && out_dtype == SVC_BF16) { SNAKE_DISPATCH(__half, __nv_bfloat16); }
    if (x_dtype == SVC_F16 && out_dtype == SVC_F32) { SNAKE_DISPATCH(__half, float); }
#undef SNAKE_DISPATCH
    svc_set_error("svc_snake_aa: unsupported dtype");
    return SVC_ERR_UNSUPPORTED;
}

extern "C" int svc_conv_post(const void* act, int act_dtype, const float* w, const float* bias, float* out, int B, int L,
                             int C, int ksize, int use_tanh, void* stream) {
    if (B < 1 || L < 1 || C < 8 || C % 8 != 0 || C > 256 || ksize < 1 || ksize > 15 || (ksize % 2) == 0 || B > 65535 ||
        (act_dtype != SVC_F16 && act_dtype != SVC_BF16) || reinterpret_cast<uintptr_t>(act) % 16 != 0) {
        svc_set_error("svc_conv_post: 16-bit (B, L, C) input, C % 8 == 0, C <= 256, odd ksize <= 15, 16-byte aligned");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rows = 256 + 2 * (ksize / 2);
    const int smem = ((rows * C * 2 + 15) & ~15) + ksize * C * 4;
    dim3 grid((L + 255) / 256, B);
    if (act_dtype == SVC_F16) {
        static bool attr = false;
        if (!attr) cudaFuncSetAttribute(conv_post_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024), attr = true;
        conv_post_kernel<__half><<<grid, 256, smem, st>>>(static_cast<const __half*>(act), w, bias, out, L, C, ksize, use_tanh);
    } else {
        static bool attr = false;
        if (!attr) cudaFuncSetAttribute(conv_post_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024), attr = true;
        conv_post_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(static_cast<const __nv_bfloat16*>(act), w, bias, out, L, C,
                                                                 ksize, use_tanh);
    }
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

extern "C" int svc_snake_conv_post(const float* x, const float* a, const float* inv_b, const float* w,
                                   const float* bias, float* out, int B, int L, int C, int ksize,
                                   int use_tanh, int precise, void* stream) {
    if (B < 1 || L < 1 || C < 1 || C > 64 || ksize < 1 || ksize > 15 || (ksize % 2) == 0) {
        svc_set_error("svc_snake_conv_post: C <= 64 and odd ksize <= 15 required");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int half = ksize / 2;
    const int yr = kPostTL + 2 * half;
    const int smem = ((yr + 13) + (2 * yr + 10) + yr) * (C + 1) * 4 + ksize * C * 4;
    if (smem > 200 * 1024) {
        svc_set_error("svc_snake_conv_post: C * ksize too large for the shared-memory tile");
        return SVC_ERR_ARG;
    }
    dim3 grid((L + kPostTL - 1) / kPostTL, B);
    if (precise) {
        cudaFuncSetAttribute(snake_conv_post_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        snake_conv_post_kernel<true><<<grid, 256, smem, st>>>(x, a, inv_b, w, bias, out, L, C, ksize, use_tanh);
    } else {
        cudaFuncSetAttribute(snake_conv_post_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        snake_conv_post_kernel<false><<<grid, 256, smem, st>>>(x, a, inv_b, w, bias, out, L, C, ksize, use_tanh);
    }
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}
