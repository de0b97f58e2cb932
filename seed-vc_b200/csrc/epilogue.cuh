// Fused GEMM epilogue shared by the tcgen05 and the SIMT mainloops (see svc_gemm in
// include/seedvc_b200.h for the contract).  One call handles CH consecutive accumulator
// columns of one output row.
#pragma once
#include "common.cuh"

namespace svc {

struct EpiParams {
    int N, N_out, act;
    const float* bias;
    const float* rowbias;
    long long rowbias_bstride;
    const float* rope_tab;
    const float* rope_tab_t;   // pair-major (32, rope_ld, 2) copy, or nullptr
    int rope_ld;
    int rope_cols, rope_pos0, q_cols;
    int rope_mod;              // > 0: rows are (batch, frame) flattened, position = rope_pos0 + row % rope_mod
    float q_scale;
    const float* gate;
    long long gate_bstride;
    const float* res;
    long long res_bstride, res_rstride;
    float alpha;
    int accumulate;
    float* out_f32;
    long long of_bstride, of_rstride;
    void* out_op;
    long long oo_bstride, oo_rstride;
    int op_is_f32;
    int op_is_f16;             // 16-bit operand format: 1 = IEEE half, 0 = bf16
    int vec_ok;  // every pointer / stride involved is 16-byte aligned
    float* row_ss_out;         // folded RMS norm: per-row sums of squares out / in (tensor-core direct epilogues only)
    const float* row_ss_in;
    float rs_inv_dim, rs_eps;
};

inline EpiParams make_epi_params(const svc_gemm_desc& d) {
    EpiParams e;
    const bool pair = d.act == SVC_ACT_SWIGLU_PAIR || d.act == SVC_ACT_TANH_SIG_PAIR;
    e.N = d.N;
    e.N_out = pair ? d.N / 2 : d.N;
    e.act = d.act;
    e.bias = d.bias;
    e.rowbias = d.rowbias;
    e.rowbias_bstride = d.rowbias_bstride;
    e.rope_tab = d.rope_tab;
    e.rope_tab_t = d.rope_tab_t;
    e.rope_ld = d.rope_ld;
    e.rope_cols = d.rope_cols;
    e.rope_pos0 = d.rope_pos0;
    e.rope_mod = 0;
    e.row_ss_out = d.row_ss_out;
    e.row_ss_in = d.row_ss_in;
    e.rs_inv_dim = d.rs_inv_dim;
    e.rs_eps = d.rs_eps;
    e.q_cols = d.q_cols;
    e.q_scale = d.q_scale;
    e.gate = d.gate;
    e.gate_bstride = d.gate_bstride;
    e.res = d.res;
    e.res_bstride = d.res_bstride;
    e.res_rstride = d.res_rstride;
    e.alpha = d.alpha;
    e.accumulate = d.accumulate;
    e.out_f32 = d.out_f32;
    e.of_bstride = d.of_bstride;
    e.of_rstride = d.of_rstride;
    e.out_op = d.out_op;
    e.oo_bstride = d.oo_bstride;
    e.oo_rstride = d.oo_rstride;
    e.op_is_f32 = d.dtype == SVC_F32;
    e.op_is_f16 = (d.out_op_dtype_p1 > 0 ? d.out_op_dtype_p1 - 1 : d.dtype) == SVC_F16;
    auto al = [](const void* p, long long s0, long long s1, int esz) {
        return p == nullptr || ((reinterpret_cast<uintptr_t>(p) % 16 == 0) &&
                                ((s0 * esz) % 16 == 0) && ((s1 * esz) % 16 == 0));
    };
    const int osz = e.op_is_f32 ? 4 : 2;
    e.vec_ok = al(d.res, d.res_bstride, d.res_rstride, 4) &&
               al(d.out_f32, d.of_bstride, d.of_rstride, 4) &&
               al(d.out_op, d.oo_bstride, d.oo_rstride, osz) && (e.N_out % 8 == 0) &&
               al(d.bias, 0, 0, 4) && al(d.rowbias, d.rowbias_bstride, 0, 4) &&
               al(d.gate, d.gate_bstride, 0, 4) && al(d.rope_tab, 0, 0, 4);
    return e;
}

// CH accumulator columns [n0, n0+CH) of row (b, t).  n0 is a multiple of CH, CH % 8 == 0 or
// CH == 4 (SIMT micro-tile).
template <int CH>
__device__ __forceinline__ void epilogue_chunk(const EpiParams& e, int b, int t, int n0,
                                               float (&v)[CH]) {
    // ---- bias ---------------------------------------------------------------------------
    if (e.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < CH; ++j)
            if (n0 + j < e.N) v[j] += __ldg(e.bias + n0 + j);
    }
    if (e.rowbias != nullptr) {
        const float* rb = e.rowbias + static_cast<long long>(b) * e.rowbias_bstride;
#pragma unroll
        for (int j = 0; j < CH; ++j)
            if (n0 + j < e.N) v[j] += __ldg(rb + n0 + j);
    }
    // ---- activation ---------------------------------------------------------------------
    constexpr int CO_MAX = CH;
    int co = CH, c0 = n0;
    if (e.act == SVC_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = siluf_(v[j]);
    } else if (e.act == SVC_ACT_SWIGLU_PAIR) {
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) v[j] = siluf_(v[2 * j]) * v[2 * j + 1];
        co = CH / 2;
        c0 = n0 / 2;
    } else if (e.act == SVC_ACT_TANH_SIG_PAIR) {
#pragma unroll
        for (int j = 0; j < CH / 2; ++j) v[j] = tanhf(v[2 * j]) * sigmoidf_(v[2 * j + 1]);
        co = CH / 2;
        c0 = n0 / 2;
    } else if (e.act == SVC_ACT_ROPE) {
        if (n0 < e.rope_cols) {
            const float* tab = e.rope_tab + static_cast<long long>(e.rope_pos0 + t) * 64;
#pragma unroll
            for (int j = 0; j < CH / 2; ++j) {
                const int i = ((n0 + 2 * j) & 63) >> 1;
                const float2 cs = __ldg(reinterpret_cast<const float2*>(tab) + i);
                const float x0 = v[2 * j], x1 = v[2 * j + 1];
                v[2 * j] = x0 * cs.x - x1 * cs.y;
                v[2 * j + 1] = x1 * cs.x + x0 * cs.y;
            }
        }
        if (n0 < e.q_cols) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] *= e.q_scale;
        }
    }
    // ---- gate, residual, alpha, accumulate ----------------------------------------------
    const bool full = (c0 + co <= e.N_out);
    if (e.gate != nullptr) {
        const float* g = e.gate + static_cast<long long>(b) * e.gate_bstride + c0;
#pragma unroll
        for (int j = 0; j < CO_MAX; ++j)
            if (j < co && c0 + j < e.N_out) v[j] *= __ldg(g + j);
    }
    const bool vec = e.vec_ok && full && (co % 4 == 0) && (c0 % 4 == 0);
    if (e.res != nullptr) {
        const float* r = e.res + static_cast<long long>(b) * e.res_bstride +
                         static_cast<long long>(t) * e.res_rstride + c0;
        if (vec) {
#pragma unroll
            for (int j = 0; j < CO_MAX; j += 4)
                if (j < co) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(r + j));
                    v[j] += q.x, v[j + 1] += q.y, v[j + 2] += q.z, v[j + 3] += q.w;
                }
        } else {
#pragma unroll
            for (int j = 0; j < CO_MAX; ++j)
                if (j < co && c0 + j < e.N_out) v[j] += __ldg(r + j);
        }
    }
    if (e.alpha != 1.0f) {
#pragma unroll
        for (int j = 0; j < CO_MAX; ++j) v[j] *= e.alpha;
    }
    if (e.out_f32 != nullptr) {
        float* o = e.out_f32 + static_cast<long long>(b) * e.of_bstride +
                   static_cast<long long>(t) * e.of_rstride + c0;
        if (vec) {
#pragma unroll
            for (int j = 0; j < CO_MAX; j += 4)
                if (j < co) {
                    if (e.accumulate) {
                        const float4 q = *reinterpret_cast<const float4*>(o + j);
                        v[j] += q.x, v[j + 1] += q.y, v[j + 2] += q.z, v[j + 3] += q.w;
                    }
                    *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
        } else {
#pragma unroll
            for (int j = 0; j < CO_MAX; ++j)
                if (j < co && c0 + j < e.N_out) {
                    if (e.accumulate) v[j] += o[j];
                    o[j] = v[j];
                }
        }
    }
    if (e.out_op != nullptr) {
        if (e.op_is_f32) {
            float* o = static_cast<float*>(e.out_op) + static_cast<long long>(b) * e.oo_bstride +
                       static_cast<long long>(t) * e.oo_rstride + c0;
            if (vec) {
#pragma unroll
                for (int j = 0; j < CO_MAX; j += 4)
                    if (j < co)
                        *reinterpret_cast<float4*>(o + j) =
                            make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < CO_MAX; ++j)
                    if (j < co && c0 + j < e.N_out) o[j] = v[j];
            }
        } else {
            uint16_t* o = static_cast<uint16_t*>(e.out_op) +
                          static_cast<long long>(b) * e.oo_bstride +
                          static_cast<long long>(t) * e.oo_rstride + c0;
            const int f16 = e.op_is_f16;
            if (vec && (co % 8 == 0) && (c0 % 8 == 0)) {
#pragma unroll
                for (int j = 0; j < CO_MAX; j += 8)
                    if (j < co) {
                        uint4 q;
                        q.x = pack_op16_rt(v[j], v[j + 1], f16);
                        q.y = pack_op16_rt(v[j + 2], v[j + 3], f16);
                        q.z = pack_op16_rt(v[j + 4], v[j + 5], f16);
                        q.w = pack_op16_rt(v[j + 6], v[j + 7], f16);
                        *reinterpret_cast<uint4*>(o + j) = q;
                    }
            } else {
#pragma unroll
                for (int j = 0; j < CO_MAX; ++j)
                    if (j < co && c0 + j < e.N_out) o[j] = cvt_op16_rt(v[j], f16);
            }
        }
    }
}

}  // namespace svc
