// svc_dit_step: one estimator call of the Euler loop as a C entry point (see include/seedvc_b200.h).
// Pure host code: it only sequences the kernels of this library on the given stream.  The Python engine
// (seed-vc_b200/dit_engine.py:DiTEngine.step) calls this for every step; the launch sequence is the one that
// file documents and that tests/test_host_logic.py checks on the CPU with emulated ops.
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace {

struct Seg {
    const void* a;
    long long bs, rs;
    int rows, shift;
    const void* w;
    long long wrs;
    int K;
};

struct Gemm {
    svc_gemm_desc d;
    Gemm(int dtype, int B, int T, int N) {
        memset(&d, 0, sizeof(d));
        d.dtype = dtype, d.B = B, d.T = T, d.N = N;
        d.alpha = 1.0f;
    }
    Gemm& seg(const Seg& s) {
        const int i = d.n_seg++;
        d.a_ptr[i] = s.a, d.a_bstride[i] = s.bs, d.a_rstride[i] = s.rs, d.a_rows[i] = s.rows, d.a_shift[i] = s.shift;
        d.w_ptr[i] = s.w, d.w_rstride[i] = s.wrs, d.K[i] = s.K;
        return *this;
    }
    Gemm& bias(const float* b) { d.bias = b; return *this; }
    Gemm& rowbias(const float* b, long long bs) { d.rowbias = b, d.rowbias_bstride = bs; return *this; }
    Gemm& gate(const float* g, long long bs) { d.gate = g, d.gate_bstride = bs; return *this; }
    Gemm& act(int a) { d.act = a; return *this; }
    Gemm& res(const float* r, long long bs, long long rs) { d.res = r, d.res_bstride = bs, d.res_rstride = rs; return *this; }
    Gemm& out_f32(float* o, long long bs, long long rs) { d.out_f32 = o, d.of_bstride = bs, d.of_rstride = rs; return *this; }
    Gemm& out_op(void* o, long long bs, long long rs, int dtype) {
        d.out_op = o, d.oo_bstride = bs, d.oo_rstride = rs;
        if (dtype != d.dtype) d.out_op_dtype_p1 = 1 + dtype;
        return *this;
    }
    int run(void* stream) { return svc_gemm(&d, SVC_BACKEND_AUTO, stream); }
};

inline size_t esize(int dtype) { return dtype == SVC_F32 ? 4 : 2; }
inline const char* at(const void* p, long long elems, int dtype) {
    return static_cast<const char*>(p) + elems * static_cast<long long>(esize(dtype));
}
inline char* at(void* p, long long elems, int dtype) {
    return static_cast<char*>(p) + elems * static_cast<long long>(esize(dtype));
}

#define RUN(x)                 \
    do {                       \
        const int _rc = (x);   \
        if (_rc != SVC_OK) return _rc; \
    } while (0)

}  // namespace

extern "C" int svc_dit_step(const svc_dit_weights* w, const svc_dit_state* st, int s, const void* x_op, void* stream) {
    if (w == nullptr || st == nullptr || x_op == nullptr || s < 0 || s >= st->n_steps || w->L < 1 ||
        w->L > SVC_MAX_LAYERS || st->n_branch < 1 || st->n_branch > SVC_MAX_BRANCH || (w->version != 1 && w->version != 2) ||
        (w->head == 1 && (w->wn_layers < 1 || w->wn_layers > SVC_MAX_WN_LAYERS))) {
        svc_set_error("svc_dit_step: bad arguments");
        return SVC_ERR_ARG;
    }
    const int B = st->B, T = st->T, nb = st->n_branch, R = nb * B;
    const int D = w->D, C = w->C, L = w->L, H = w->H, I = w->I;
    const int ntok = (w->time_as_token ? 1 : 0) + (w->style_as_token ? 1 : 0);
    const int Tq = T + ntok;
    const int od = w->op_dtype, sd = w->stream_dtype;
    const long long hD = static_cast<long long>(Tq) * D;        // batch stride of (R, Tq, D) buffers
    const float* ada = st->ada + static_cast<long long>(s) * st->n_ada;
    float* h = st->h;
    float* hb = h + static_cast<long long>(ntok) * D;            // frame rows of the hidden state

    // ---- input assembly: x columns of cond_x_merge_linear + the hoisted constants --------------------------
    for (int k = 0; k < nb; ++k) {
        Gemm g(sd, B, T, D);
        g.seg({x_op, static_cast<long long>(T) * C, C, T, 0, w->merge_wx, w->merge_w_rstride, C});
        if (st->const_kind[k] == 0) g.res(st->const_ptr[k], static_cast<long long>(T) * D, D);
        else g.bias(st->const_ptr[k]);
        g.out_f32(hb + static_cast<long long>(k) * B * hD, hD, D);
        RUN(g.run(stream));
    }
    if (w->time_as_token)
        RUN(svc_set_rows(st->t1 + static_cast<long long>(s) * D, 0, h, hD, R, D, stream));
    if (w->style_as_token) {
        const int r = w->time_as_token ? 1 : 0;
        for (int k = 0; k < nb; ++k) {
            float* dst = h + static_cast<long long>(k) * B * hD + static_cast<long long>(r) * D;
            if (st->branch_style[k]) RUN(svc_set_rows(st->style_tok, D, dst, hD, B, D, stream));
            else RUN(svc_set_rows(st->style_tok_null, 0, dst, hD, B, D, stream));
        }
    }

    // ---- transformer ---------------------------------------------------------------------------------------
    const int n_skip = w->uvit ? L / 2 : 0;
    int skip_stack[SVC_MAX_LAYERS / 2];
    int n_stack = 0, next_buf = 0;
    void* pending_raw = nullptr;       // skip tensor of the previous (emit) layer, written by this layer's first norm
    for (int i = 0; i < L; ++i) {
        const bool emit = w->uvit && i < L / 2, recv = w->uvit && i > L / 2;
        const bool next_recv = w->uvit && (i + 1) > L / 2 && (i + 1) < L;
        if (recv) {
            if (n_stack < 1 || w->skip_w[i] == nullptr) {
                svc_set_error("svc_dit_step: receive layer without a skip tensor / skip_in_linear");
                return SVC_ERR_ARG;
            }
            const void* skip = st->skips[skip_stack[--n_stack]];
            Gemm g(sd, R, Tq, D);
            g.seg({st->h_op, hD, D, Tq, 0, w->skip_w[i], 2LL * D, D});
            g.seg({skip, hD, D, Tq, 0, at(w->skip_w[i], D, sd), 2LL * D, D});
            g.bias(w->skip_b[i]).out_f32(h, hD, D);
            RUN(g.run(stream));
        }
        // ---- attention
        const float *mul = nullptr, *add = nullptr, *gate_a = nullptr, *gate_m = nullptr;
        const float *mul_f = nullptr, *add_f = nullptr;
        if (w->version == 1) {
            if (w->ada_attn[i] >= 0) mul = ada + w->ada_attn[i], add = mul + D;
            if (w->ada_ffn[i] >= 0) mul_f = ada + w->ada_ffn[i], add_f = mul_f + D;
        } else {                       // shift, 1+scale, gate, shift, 1+scale, gate
            const float* a = ada + w->ada_attn[i];
            add = a, mul = a + D, gate_a = a + 2 * D, add_f = a + 3 * D, mul_f = a + 4 * D, gate_m = a + 5 * D;
        }
        if (pending_raw != nullptr)
            RUN(svc_norm_mod_copy(h, hD, D, w->g_attn[i], mul, add, 1e-5f, 0, st->xn, pending_raw, hD, D, R, Tq, D, od, sd,
                                  stream));
        else
            RUN(svc_norm_mod(h, hD, D, w->g_attn[i], mul, add, 1e-5f, 0, st->xn, hD, D, R, Tq, D, od, stream));
        pending_raw = nullptr;
        {
            Gemm g(od, R, Tq, 3 * D);
            g.seg({st->xn, hD, D, Tq, 0, w->wqkv[i], D, D}).act(SVC_ACT_ROPE);
            g.d.rope_tab = w->rope_tab, g.d.rope_cols = 2 * D, g.d.rope_pos0 = 0, g.d.q_cols = D, g.d.q_scale = 0.125f;
            g.d.rope_tab_t = w->rope_tab_t, g.d.rope_ld = w->rope_ld;
            g.out_op(st->qkv, 3 * hD, 3LL * D, od);
            RUN(g.run(stream));
        }
        RUN(svc_attention(st->qkv, at(st->qkv, D, od), at(st->qkv, 2LL * D, od), 3 * hD, 3LL * D, st->att, hD, D, R, Tq, H,
                          st->kv_len, od, SVC_BACKEND_AUTO, stream));
        {
            Gemm g(od, R, Tq, D);
            g.seg({st->att, hD, D, Tq, 0, w->wo[i], D, D});
            if (gate_a != nullptr) g.gate(gate_a, 0);
            g.res(h, hD, D).out_f32(h, hD, D);
            RUN(g.run(stream));
        }
        // ---- feed-forward
        RUN(svc_norm_mod(h, hD, D, w->g_ffn[i], mul_f, add_f, 1e-5f, 0, st->xn, hD, D, R, Tq, D, od, stream));
        {
            Gemm g(od, R, Tq, 2 * I);
            g.seg({st->xn, hD, D, Tq, 0, w->w13[i], D, D}).act(SVC_ACT_SWIGLU_PAIR);
            g.out_op(st->ff, static_cast<long long>(Tq) * I, I, od);
            RUN(g.run(stream));
        }
        void* out_op = nullptr;
        if (emit) {
            if (next_buf >= n_skip) {
                svc_set_error("svc_dit_step: out of skip buffers");
                return SVC_ERR_ARG;
            }
            void* buf = st->skips[next_buf];
            skip_stack[n_stack++] = next_buf++;
            // v1: the skip copy falls out of the next layer's attention norm (it reads h anyway) unless that layer
            // starts with skip_in_linear
            if (w->version == 1 && (i + 1) < L && !next_recv) pending_raw = buf;
            else out_op = buf;
        } else if (next_recv) {
            out_op = st->h_op;
        }
        {
            Gemm g(od, R, Tq, D);
            g.seg({st->ff, static_cast<long long>(Tq) * I, I, Tq, 0, w->w2[i], I, I});
            if (gate_m != nullptr) g.gate(gate_m, 0);
            g.res(h, hD, D).out_f32(h, hD, D);
            if (out_op != nullptr) g.out_op(out_op, hD, D, sd);
            RUN(g.run(stream));
        }
    }
    // ---- final norm (uses c even when time is a token, diffusion_transformer.py:142) -----------------------------
    {
        const float* a = ada + w->ada_final;
        RUN(svc_norm_mod(h, hD, D, w->g_final, a, a + D, 1e-5f, 0, st->xn_f, hD, D, R, Tq, D, sd, stream));
    }
    const void* xf = at(st->xn_f, static_cast<long long>(ntok) * D, sd);    // frame rows, batch stride hD
    const void* xr = xf;
    long long xr_bs = hD;
    const long long tD = static_cast<long long>(T) * D;
    if (w->long_skip) {                // skip_linear(cat[x_res, x]) without the concat (:524-525)
        for (int k = 0; k < nb; ++k) {
            Gemm g(sd, B, T, D);
            g.seg({at(xf, static_cast<long long>(k) * B * hD, sd), hD, D, T, 0, w->lskip_w, static_cast<long long>(D + C), D});
            g.seg({x_op, static_cast<long long>(T) * C, C, T, 0, at(w->lskip_w, D, sd), static_cast<long long>(D + C), C});
            g.bias(w->lskip_b).out_op(at(st->x_res, static_cast<long long>(k) * B * tD, sd), tD, D, sd);
            RUN(g.run(stream));
        }
        xr = st->x_res, xr_bs = tD;
    }
    float* v = st->v;
    const long long tC = static_cast<long long>(T) * C;
    if (w->head == 0) {
        Gemm g0(sd, R, T, D);
        g0.seg({xr, xr_bs, D, T, 0, w->mlp0_w, D, D}).bias(w->mlp0_b).act(SVC_ACT_SILU).out_op(st->y, tD, D, sd);
        RUN(g0.run(stream));
        Gemm g1(sd, R, T, C);
        g1.seg({st->y, tD, D, T, 0, w->mlp2_w, D, D}).bias(w->mlp2_b).out_f32(v, tC, C);
        return g1.run(stream);
    }
    // ---- WaveNet head (wavenet.py:138-166 with SConv1d reflect padding, encodec.py:212-228) -----------------------
    const int Dw = w->Dw, nl = w->wn_layers, ks = w->wn_kernel, pad = (ks - 1) / 2;
    const long long tW = static_cast<long long>(T) * Dw, pW = static_cast<long long>(T + 2 * pad) * Dw;
    void* body = at(st->xw_op, static_cast<long long>(pad) * Dw, sd);
    {
        Gemm g(sd, R, T, Dw);
        g.seg({xr, xr_bs, D, T, 0, w->conv1_w, D, D}).bias(w->conv1_b).out_f32(st->xw, tW, Dw).out_op(body, pW, Dw, sd);
        RUN(g.run(stream));
    }
    RUN(svc_reflect_halo(st->xw_op, pW, Dw, R, T, Dw, pad, st->wn_lens, sd, stream));
    const float* g_all = st->wn_g + static_cast<long long>(s) * nl * 2 * Dw;
    const long long aW = static_cast<long long>(T) * nl * Dw;
    for (int l = 0; l < nl; ++l) {
        void* acts_l = at(st->acts, static_cast<long long>(l) * Dw, sd);
        Gemm g(sd, R, T, 2 * Dw);
        for (int j = 0; j < ks; ++j)
            g.seg({st->xw_op, pW, Dw, T + 2 * pad, j, at(w->wn_in_w[l], static_cast<long long>(j) * 2 * Dw * Dw, sd), Dw, Dw});
        g.rowbias(g_all + static_cast<long long>(l) * 2 * Dw, 0).act(SVC_ACT_TANH_SIG_PAIR);
        g.out_op(acts_l, aW, static_cast<long long>(nl) * Dw, sd);
        RUN(g.run(stream));
        if (l < nl - 1) {
            Gemm r(sd, R, T, Dw);
            r.seg({acts_l, aW, static_cast<long long>(nl) * Dw, T, 0, w->wn_rs_w[l], Dw, Dw}).bias(w->wn_rs_b[l]);
            r.res(st->xw, tW, Dw).out_f32(st->xw, tW, Dw).out_op(body, pW, Dw, sd);
            RUN(r.run(stream));
            RUN(svc_reflect_halo(st->xw_op, pW, Dw, R, T, Dw, pad, st->wn_lens, sd, stream));
        }
    }
    {   // output = sum_l skip_l + res_projection(x_res): one GEMM, K = nl * Dw + D
        Gemm g(sd, R, T, Dw);
        g.seg({st->acts, aW, static_cast<long long>(nl) * Dw, T, 0, w->wn_skip_w, static_cast<long long>(nl) * Dw, nl * Dw});
        g.seg({xr, xr_bs, D, T, 0, w->resp_w, D, D});
        g.bias(w->wn_skip_b).out_f32(st->wn_out, tW, Dw);
        RUN(g.run(stream));
    }
    {
        const float* a = ada + w->ada_fl;          // shift, 1 + scale
        RUN(svc_norm_mod(st->wn_out, tW, Dw, nullptr, a + Dw, a, 1e-6f, 1, st->ln, tW, Dw, R, T, Dw, sd, stream));
    }
    {
        Gemm g(sd, R, T, Dw);
        g.seg({st->ln, tW, Dw, T, 0, w->fl_w, Dw, Dw}).bias(w->fl_b).out_op(st->y, tW, Dw, sd);
        RUN(g.run(stream));
    }
    Gemm g(sd, R, T, C);
    g.seg({st->y, tW, Dw, T, 0, w->conv2_w, Dw, Dw}).bias(w->conv2_b).out_f32(v, tC, C);
    return g.run(stream);
}


// ------------------------------------------------------------------------------------------------------------
// svc_bigvgan_forward
// ------------------------------------------------------------------------------------------------------------
namespace {

struct VocLayout {
    long long mel_op, pre, xs, xt, y, nxt, act, nxt_op[2], total;
};

// workspace carve-up: M = B * Tm * max_i(L_i / Tm * O_i) elements per stage tensor
VocLayout voc_layout(const svc_bigvgan_weights* w, int B, int Tm) {
    const long long es = w->op_dtype == SVC_F32 ? 4 : 2;
    long long per_frame = 0, up = 1;
    for (int i = 0; i < w->n_stages; ++i) {
        up *= w->stages[i].u;
        per_frame = std::max(per_frame, up * w->stages[i].O);
    }
    const long long M = static_cast<long long>(B) * Tm * per_frame;
    auto al = [](long long v) { return (v + 255) & ~255LL; };
    VocLayout l;
    long long off = 0;
    l.mel_op = off, off += al(static_cast<long long>(B) * Tm * w->n_mels * es);
    l.pre = off, off += al(static_cast<long long>(B) * Tm * w->c0 * es);
    l.xs = off, off += al(M * 4);
    l.xt = off, off += al(M * 4);
    l.y = off, off += al(M * 4);
    l.nxt = off, off += al(M * 4);
    l.act = off, off += al(M * es);
    l.nxt_op[0] = off, off += al(M * es);
    l.nxt_op[1] = off, off += al(M * es);
    l.total = off;
    return l;
}

int voc_conv(const svc_conv_plan& c, int dtype, const void* act, int B, long long L, int O, const float* res, float alpha,
             int accumulate, float* out_f32, void* out_op, int out_op_dtype, void* stream) {
    const int f = c.f;
    const int N = O * f;
    const long long rows = L / f;
    Gemm g(dtype, B, static_cast<int>(rows), N);
    for (int i = 0; i < c.n_taps; ++i)
        g.seg({act, rows * N, N, static_cast<int>(rows), c.shifts[i], at(c.w, static_cast<long long>(i) * N * N, dtype), N, N});
    g.bias(c.b);
    if (res != nullptr) g.res(res, rows * N, N);
    g.d.alpha = alpha, g.d.accumulate = accumulate;
    if (out_f32 != nullptr) g.out_f32(out_f32, rows * N, N);
    if (out_op != nullptr) g.out_op(out_op, rows * N, N, out_op_dtype);
    return g.run(stream);
}

}  // namespace

extern "C" long long svc_bigvgan_workspace_bytes(const svc_bigvgan_weights* w, int B, int Tm) {
    if (w == nullptr || B < 1 || Tm < 1 || w->n_stages < 1 || w->n_stages > SVC_MAX_STAGES) return -1;
    return voc_layout(w, B, Tm).total;
}

extern "C" int svc_bigvgan_forward(const svc_bigvgan_weights* w, const float* mel, void* workspace, float* out, int B,
                                   int Tm, void* stream) {
    if (w == nullptr || mel == nullptr || workspace == nullptr || out == nullptr || B < 1 || Tm < 1 || w->n_stages < 1 ||
        w->n_stages > SVC_MAX_STAGES || w->n_kernels < 1 || w->n_kernels > 3 || w->n_dil < 1 || w->n_dil > 3) {
        svc_set_error("svc_bigvgan_forward: bad arguments");
        return SVC_ERR_ARG;
    }
    const int od = w->op_dtype;
    const VocLayout lay = voc_layout(w, B, Tm);
    char* ws = static_cast<char*>(workspace);
    void* mel_op = ws + lay.mel_op;
    RUN(svc_bct_to_btc(mel, mel_op, static_cast<long long>(Tm) * w->n_mels, w->n_mels, B, w->n_mels, Tm, 0, 0, od, stream));
    void* cur_op = ws + lay.pre;
    {
        Gemm g(od, B, Tm, w->c0);
        for (int j = 0; j < 7; ++j)
            g.seg({mel_op, static_cast<long long>(Tm) * w->n_mels, w->n_mels, Tm, j - 3,
                   at(w->pre_w, static_cast<long long>(j) * w->c0 * w->n_mels, od), w->n_mels, w->n_mels});
        g.bias(w->pre_b).out_op(cur_op, static_cast<long long>(Tm) * w->c0, w->c0, od);
        RUN(g.run(stream));
    }
    float* xs = reinterpret_cast<float*>(ws + lay.xs);
    // conv1's output feeds only the second Snake of the pair, whose tensor-core FIRs read it as IEEE half:
    // in the 16-bit modes it is written once, as half, by the conv's epilogue (no fp32 round trip)
    void* xt = ws + lay.xt;
    const int xd = od == SVC_F32 ? SVC_F32 : SVC_F16;
    float* y = reinterpret_cast<float*>(ws + lay.y);
    float* nxt = reinterpret_cast<float*>(ws + lay.nxt);
    void* act = ws + lay.act;
    long long L = Tm;
    int I = w->c0;
    const int nk = w->n_kernels;
    for (int si = 0; si < w->n_stages; ++si) {
        const svc_bigvgan_stage& st = w->stages[si];
        const int u = st.u, O = st.O;
        {   // ConvTranspose1d as a polyphase GEMM with N = u * O
            Gemm g(od, B, static_cast<int>(L), u * O);
            for (int di = 0; di < st.n_delta; ++di)
                g.seg({cur_op, L * I, I, static_cast<int>(L), st.deltas[di],
                       at(st.up_w, static_cast<long long>(di) * u * O * I, od), I, I});
            g.bias(st.up_b).out_f32(xs, L * u * O, static_cast<long long>(u) * O);
            RUN(g.run(stream));
        }
        L *= u;
        const bool last_stage = si == w->n_stages - 1;
        void* nxt_op = last_stage ? nullptr : ws + lay.nxt_op[si & 1];
        for (int j = 0; j < nk; ++j) {
            const float* src = xs;
            for (int l = 0; l < w->n_dil; ++l) {
                const svc_amp_pair& pr = st.pairs[j][l];
                RUN(svc_snake_aa(src, SVC_F32, act, od, pr.a1, pr.inv_b1, B, static_cast<int>(L), O, w->precise, stream));
                if (xd == SVC_F32) {
                    RUN(voc_conv(pr.c1, od, act, B, L, O, nullptr, 1.0f, 0, static_cast<float*>(xt), nullptr, od, stream));
                } else {
                    RUN(voc_conv(pr.c1, od, act, B, L, O, nullptr, 1.0f, 0, nullptr, xt, xd, stream));
                }
                RUN(svc_snake_aa(xt, xd, act, od, pr.a2, pr.inv_b2, B, static_cast<int>(L), O, w->precise, stream));
                if (l < w->n_dil - 1) {
                    RUN(voc_conv(pr.c2, od, act, B, L, O, src, 1.0f, 0, y, nullptr, od, stream));
                    src = y;
                } else {    // last pair: residual, then (r0 + r1 + r2) / 3 accumulated in place
                    RUN(voc_conv(pr.c2, od, act, B, L, O, src, 1.0f / nk, j > 0, nxt,
                                 (j == nk - 1) ? nxt_op : nullptr, od, stream));
                }
            }
        }
        cur_op = nxt_op;
        I = O;
    }
    if (od != SVC_F32 && I % 8 == 0) {   // activation_post on the tensor cores (16-bit out), conv_post + clamp behind it
        RUN(svc_snake_aa(nxt, SVC_F32, act, od, w->post_a, w->post_inv_b, B, static_cast<int>(L), I, 0, stream));
        return svc_conv_post(act, od, w->post_w, w->post_b, out, B, static_cast<int>(L), I, w->post_k, w->use_tanh, stream);
    }
    return svc_snake_conv_post(nxt, w->post_a, w->post_inv_b, w->post_w, w->post_b, out, B, static_cast<int>(L), I,
                               w->post_k, w->use_tanh, w->precise, stream);
}
