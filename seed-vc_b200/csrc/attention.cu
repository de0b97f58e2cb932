// svc_attention: non-causal softmax attention, head_dim 64, keys masked by kv_len[b].
// Replaces F.scaled_dot_product_attention with the (B,1,T,T) bool key mask
// (reference: modules/diffusion_transformer.py:255,518-520).  q and k arrive already rotated
// (RoPE) and q pre-scaled by 1/sqrt(64) from the wqkv GEMM epilogue.
//
// bf16: flash-style tcgen05 kernel.  One CTA = 128 queries of one (batch, head).
//   warp 0  : TMA producer (Q once, then K_j / V_j tiles of 128 keys through mbarrier rings)
//   warp 1  : MMA issuer.  S_j = Q K_j^T -> TMEM (2 buffers of 128 fp32 columns),
//             O_j = P_j V_j -> TMEM (2 buffers of 64 columns, V consumed MN-major)
//   warps 2-5: one query row per thread: online softmax on S_j (tcgen05.ld), P_j -> bf16 ->
//             128B-swizzled smem for the PV MMA, running O in registers rescaled per block.
// fp32: FFMA kernel, one query per thread ("fp32 mode").
#include "common.cuh"

namespace svc {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per block
constexpr int AT_HD = 64;
constexpr int AT_KST = 2;    // K / V ring depth

struct alignas(64) AttnTcParams {
    CUtensorMap qmap, kmap, vmap;  // (H*64, T, B) bf16 views, box {64, 128, 1}
    __nv_bfloat16* out;
    long long o_bstride, o_rstride;
    const int* kv_len;
    int T, H;
};

struct AttnSmem {
    static constexpr int TILE = AT_BM * AT_HD * 2;          // 16 KB
    static constexpr int Q_OFF = 0;
    static constexpr int K_OFF = Q_OFF + TILE;
    static constexpr int V_OFF = K_OFF + AT_KST * TILE;
    static constexpr int P_OFF = V_OFF + AT_KST * TILE;     // 2 buffers x (2 tiles of 128x64)
    static constexpr int BAR_OFF = P_OFF + 2 * 2 * TILE;
    static constexpr int TOTAL = BAR_OFF + 512 + 1024;
};

// descriptor for a MN-major (rows = K index, 64 contiguous MN elements per row) SW128 tile
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t addr) {
    return umma_desc_sw128(addr, 1024, 1024);
}

__global__ void __launch_bounds__(192, 1) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
    using S = AttnSmem;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* q_full = bars;                // 1
    uint64_t* k_full = bars + 1;            // AT_KST
    uint64_t* k_empty = k_full + AT_KST;
    uint64_t* v_full = k_empty + AT_KST;
    uint64_t* v_empty = v_full + AT_KST;
    uint64_t* s_full = v_empty + AT_KST;    // 2
    uint64_t* s_empty = s_full + 2;
    uint64_t* p_full = s_empty + 2;
    uint64_t* p_empty = p_full + 2;
    uint64_t* o_full = p_empty + 2;
    uint64_t* o_empty = o_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * AT_BM;
    const int h = blockIdx.y;
    const int b = blockIdx.z;
    int kv_len = p.kv_len != nullptr ? p.kv_len[b] : p.T;
    kv_len = max(1, min(kv_len, p.T));
    const int n_blocks = (kv_len + AT_BN - 1) / AT_BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.qmap);
        tma_prefetch_desc(&p.kmap);
        tma_prefetch_desc(&p.vmap);
        mbar_init(q_full, 1);
        for (int i = 0; i < AT_KST; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 128);
            mbar_init(&p_full[i], 128);
            mbar_init(&p_empty[i], 1);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 128);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;          // 2 x 128 columns
    const uint32_t tmem_O = tmem_base + 256;    // 2 x 64 columns

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, S::TILE);
            tma_load_3d(smem + S::Q_OFF, &p.qmap, q_full, h * AT_HD, q0, b);
            for (int j = 0; j < n_blocks; ++j) {
                const int st = j % AT_KST;
                const uint32_t ph = (j / AT_KST) & 1;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], S::TILE);
                tma_load_3d(smem + S::K_OFF + st * S::TILE, &p.kmap, &k_full[st], h * AT_HD,
                            j * AT_BN, b);
                mbar_wait(&v_empty[st], ph ^ 1);
                mbar_expect_tx(&v_full[st], S::TILE);
                tma_load_3d(smem + S::V_OFF + st * S::TILE, &p.vmap, &v_full[st], h * AT_HD,
                            j * AT_BN, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN, 0, 0);
            const uint32_t idesc_o = umma_idesc_bf16(AT_BM, AT_HD, 0, 1);   // B (=V) MN-major
            const uint32_t sq = smem_u32(smem + S::Q_OFF);
            auto issue_s = [&](int j) {
                const int st = j % AT_KST, sb = j & 1;
                mbar_wait(&k_full[st], (j / AT_KST) & 1);
                mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t sk = smem_u32(smem + S::K_OFF + st * S::TILE);
#pragma unroll
                for (int k = 0; k < AT_HD / 16; ++k)
                    tc_mma_f16(tmem_S + sb * AT_BN, umma_desc_sw128(sq + k * 32, 0, 1024),
                               umma_desc_sw128(sk + k * 32, 0, 1024), idesc_s, k != 0);
                tc_commit(&k_empty[st]);
                tc_commit(&s_full[sb]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_blocks; ++j) {
                if (j + 1 < n_blocks) issue_s(j + 1);
                const int st = j % AT_KST, pb = j & 1;
                mbar_wait(&v_full[st], (j / AT_KST) & 1);
                mbar_wait(&o_empty[pb], ((j >> 1) & 1) ^ 1);
                mbar_wait(&p_full[pb], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t sp = smem_u32(smem + S::P_OFF + pb * 2 * S::TILE);
                const uint32_t sv = smem_u32(smem + S::V_OFF + st * S::TILE);
#pragma unroll
                for (int k = 0; k < AT_BN / 16; ++k) {
                    const uint64_t da = umma_desc_sw128(sp + (k >> 2) * S::TILE + (k & 3) * 32, 0, 1024);
                    const uint64_t db = umma_desc_mn_sw128(sv + k * 2048);
                    tc_mma_f16(tmem_O + pb * AT_HD, da, db, idesc_o, k != 0);
                }
                tc_commit(&v_empty[st]);
                tc_commit(&p_empty[pb]);
                tc_commit(&o_full[pb]);
            }
        }
    } else {
        // ===================== softmax / output: one query row per thread =====================
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
        constexpr float kLog2e = 1.4426950408889634f;
        float o_acc[AT_HD];
#pragma unroll
        for (int d = 0; d < AT_HD; ++d) o_acc[d] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

        for (int j = 0; j < n_blocks; ++j) {
            const int sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            const int kbase = j * AT_BN;
            // pass 1: row max
            float m_blk = -INFINITY;
#pragma unroll
            for (int c = 0; c < AT_BN; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_S + sb * AT_BN + lane_addr + c, r);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float s = (kbase + c + i < kv_len) ? __uint_as_float(r[i]) : -INFINITY;
                    m_blk = fmaxf(m_blk, s);
                }
            }
            const float m_new = fmaxf(m_run, m_blk);       // finite: block has >= 1 valid key
            const float alpha = exp2f((m_run - m_new) * kLog2e);
            const float mscaled = m_new * kLog2e;
            // P buffer must be free (PV of block j-2 retired)
            mbar_wait(&p_empty[sb], ((j >> 1) & 1) ^ 1);
            uint8_t* pbuf = smem + S::P_OFF + sb * 2 * S::TILE;
            float l_blk = 0.f;
#pragma unroll
            for (int c = 0; c < AT_BN; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_S + sb * AT_BN + lane_addr + c, r);
                tc_wait_ld();
                uint32_t packed[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = (kbase + c + i < kv_len)
                                   ? exp2f(fmaf(__uint_as_float(r[i]), kLog2e, -mscaled)) : 0.f;
                    float p1 = (kbase + c + i + 1 < kv_len)
                                   ? exp2f(fmaf(__uint_as_float(r[i + 1]), kLog2e, -mscaled)) : 0.f;
                    // sum what the MMA will see (bf16-rounded), like flash kernels do not; keep fp32
                    l_blk += p0 + p1;
                    packed[i >> 1] = pack_bf16(p0, p1);
                }
                // 32 keys = 4 chunks of 16 B; tile = c / 64, chunk index within the 128 B row
                uint8_t* tile = pbuf + (c >> 6) * S::TILE + row * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = ((c & 63) >> 3) + q;
                    uint4 val = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2],
                                           packed[4 * q + 3]);
                    *reinterpret_cast<uint4*>(tile + ((chunk ^ (row & 7)) << 4)) = val;
                }
            }
            tc_fence_before();
            mbar_arrive(&s_empty[sb]);
            fence_proxy_async_smem();
            mbar_arrive(&p_full[sb]);
            l_run = l_run * alpha + l_blk;
            m_run = m_new;
            // fold in the previous block's PV product (had a whole iteration to finish)
            if (j > 0) {
                const int ob = (j - 1) & 1;
                mbar_wait(&o_full[ob], ((j - 1) >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < AT_HD; c += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_O + ob * AT_HD + lane_addr + c, r);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        o_acc[c + i] = fmaf(o_acc[c + i], alpha_prev, __uint_as_float(r[i]));
                }
                tc_fence_before();
                mbar_arrive(&o_empty[ob]);
            }
            alpha_prev = alpha;
        }
        {
            const int j = n_blocks - 1;
            const int ob = j & 1;
            mbar_wait(&o_full[ob], (j >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < AT_HD; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_O + ob * AT_HD + lane_addr + c, r);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    o_acc[c + i] = fmaf(o_acc[c + i], alpha_prev, __uint_as_float(r[i]));
            }
        }
        const int t = q0 + row;
        if (t < p.T) {
            const float inv_l = 1.0f / l_run;
            __nv_bfloat16* o = p.out + static_cast<long long>(b) * p.o_bstride +
                               static_cast<long long>(t) * p.o_rstride + h * AT_HD;
#pragma unroll
            for (int d = 0; d < AT_HD; d += 8) {
                uint4 q;
                q.x = pack_bf16(o_acc[d] * inv_l, o_acc[d + 1] * inv_l);
                q.y = pack_bf16(o_acc[d + 2] * inv_l, o_acc[d + 3] * inv_l);
                q.z = pack_bf16(o_acc[d + 4] * inv_l, o_acc[d + 5] * inv_l);
                q.w = pack_bf16(o_acc[d + 6] * inv_l, o_acc[d + 7] * inv_l);
                *reinterpret_cast<uint4*>(o + d) = q;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// fp32 / debug path: one query per thread, K/V tiles of 32 keys staged in shared memory
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) attention_simt_kernel(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, long long bstride,
    long long rstride, T* __restrict__ out, long long o_bstride, long long o_rstride, int Tn,
    const int* __restrict__ kv_len_p) {
    constexpr int KT = 32;
    __shared__ float ks[KT][AT_HD];
    __shared__ float vs[KT][AT_HD];
    const int h = blockIdx.y, b = blockIdx.z;
    const int t = blockIdx.x * 128 + threadIdx.x;
    int kv_len = kv_len_p != nullptr ? kv_len_p[b] : Tn;
    kv_len = max(1, min(kv_len, Tn));
    const long long base = static_cast<long long>(b) * bstride + h * AT_HD;
    float qr[AT_HD], o[AT_HD];
    const int tq = min(t, Tn - 1);
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) {
        qr[d] = to_f32<T>(q[base + static_cast<long long>(tq) * rstride + d]);
        o[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < kv_len; k0 += KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < KT * AT_HD; i += 128) {
            const int r = i / AT_HD, d = i % AT_HD;
            const int kk = k0 + r;
            float kv = 0.f, vv = 0.f;
            if (kk < kv_len) {
                kv = to_f32<T>(k[base + static_cast<long long>(kk) * rstride + d]);
                vv = to_f32<T>(v[base + static_cast<long long>(kk) * rstride + d]);
            }
            ks[r][d] = kv;
            vs[r][d] = vv;
        }
        __syncthreads();
        const int nk = min(KT, kv_len - k0);
        for (int r = 0; r < nk; ++r) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < AT_HD; ++d) s = fmaf(qr[d], ks[r][d], s);
            const float m_new = fmaxf(m, s);
            const float a = __expf(m - m_new);
            const float pz = __expf(s - m_new);
            l = l * a + pz;
#pragma unroll
            for (int d = 0; d < AT_HD; ++d) o[d] = fmaf(o[d], a, pz * vs[r][d]);
            m = m_new;
        }
    }
    if (t < Tn) {
        const float inv = 1.0f / l;
        T* op = out + static_cast<long long>(b) * o_bstride + static_cast<long long>(t) * o_rstride +
                h * AT_HD;
#pragma unroll
        for (int d = 0; d < AT_HD; ++d) op[d] = from_f32<T>(o[d] * inv);
    }
}

bool encode_bf16_map(CUtensorMap* map, const void* ptr, int K, long long rows, long long rstride,
                     long long batches, long long bstride, int box_rows);

}  // namespace svc

extern "C" int svc_attention(const void* q, const void* k, const void* v, long long qkv_bstride,
                             long long qkv_rstride, void* out, long long out_bstride,
                             long long out_rstride, int B, int T, int H, const int* kv_len,
                             int dtype, int backend, void* stream) {
    using namespace svc;
    if (B < 1 || T < 1 || H < 1 || q == nullptr || k == nullptr || v == nullptr || out == nullptr) {
        svc_set_error("svc_attention: bad arguments");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == SVC_BF16 && backend == SVC_BACKEND_AUTO) {
        AttnTcParams p;
        if (!encode_bf16_map(&p.qmap, q, H * AT_HD, T, qkv_rstride, B, qkv_bstride, AT_BM) ||
            !encode_bf16_map(&p.kmap, k, H * AT_HD, T, qkv_rstride, B, qkv_bstride, AT_BN) ||
            !encode_bf16_map(&p.vmap, v, H * AT_HD, T, qkv_rstride, B, qkv_bstride, AT_BN)) {
            svc_set_error("svc_attention: tensor map encode failed (alignment)");
            return SVC_ERR_ARG;
        }
        if (reinterpret_cast<uintptr_t>(out) % 16 || (out_bstride * 2) % 16 || (out_rstride * 2) % 16) {
            svc_set_error("svc_attention: out must be 16-byte aligned");
            return SVC_ERR_ARG;
        }
        p.out = static_cast<__nv_bfloat16*>(out);
        p.o_bstride = out_bstride;
        p.o_rstride = out_rstride;
        p.kv_len = kv_len;
        p.T = T;
        p.H = H;
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 AttnSmem::TOTAL);
            attr_set = true;
        }
        dim3 grid((T + AT_BM - 1) / AT_BM, H, B);
        attention_tc_kernel<<<grid, 192, AttnSmem::TOTAL, st>>>(p);
        SVC_CHECK_LAUNCH();
        return SVC_OK;
    }
    dim3 grid((T + 127) / 128, H, B);
    if (dtype == SVC_F32)
        attention_simt_kernel<float><<<grid, 128, 0, st>>>(
            static_cast<const float*>(q), static_cast<const float*>(k), static_cast<const float*>(v),
            qkv_bstride, qkv_rstride, static_cast<float*>(out), out_bstride, out_rstride, T, kv_len);
    else
        attention_simt_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(
            static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
            static_cast<const __nv_bfloat16*>(v), qkv_bstride, qkv_rstride,
            static_cast<__nv_bfloat16*>(out), out_bstride, out_rstride, T, kv_len);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}
