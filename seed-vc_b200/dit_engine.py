"""Host-side orchestration of the DiT velocity estimator on the sm_100a kernels.

This is the part of the hot path the reference runs as ~1000 small PyTorch ops per
Euler step.  Here one solve is:

``begin``  - once per ``solve_euler`` call: everything that does not depend on ``x``
             (SURVEY App. B-3): ``cond_projection(mu)``, the prompt / content / style
             columns of ``cond_x_merge_linear``, the null-branch constants, the style
             token, and - for *all* Euler steps at once - the timestep MLPs, every
             AdaLN projection, the WaveNet ``cond_layer`` and the FinalLayer modulation.
``step``   - per Euler step: the ``x`` columns of the merge GEMM (K = n_mels), the
             transformer layers, the head; returns the velocity of every CFG branch.

Activations are frames-major ``(rows, T', D)``; residual streams are fp32, GEMM operands
are in the operand dtype (bf16 on tensor cores / fp32 in fp32 mode).  All arithmetic is
in ``libseedvc_b200.so``; torch only allocates.

Reference: modules/diffusion_transformer.py:77-147 (Transformer), :150-191 (block),
:194-260 (attention), :263-271 (FFN), :30-48/:274-285 (AdaLN/RMSNorm), :388-405 (FinalLayer),
:486-537 (DiT.forward); modules/wavenet.py:138-166; modules/encodec.py:212-228;
v2: modules/v2/dit_model.py:82-143, modules/v2/dit_wrapper.py:114-152.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

from ._lib import ACT_NONE, ACT_ROPE, ACT_SILU, ACT_SWIGLU_PAIR, ACT_TANH_SIG_PAIR


@dataclass
class DiTSpec:
    """Static description of one estimator (v1 or v2)."""
    version: int          # 1 or 2
    D: int
    H: int
    L: int
    C: int                # mel bins
    content_dim: int
    style_dim: int
    time_as_token: bool
    style_as_token: bool
    style_in_merge: bool  # v1 style_condition and not style_as_token
    uvit: bool
    long_skip: bool
    head: str             # "mlp" | "wavenet"
    Dw: int = 0
    wn_layers: int = 0
    wn_kernel: int = 5
    prefix: str = "estimator."

    @property
    def I(self):
        n = int(2 * 4 * self.D / 3)
        return n if n % 256 == 0 else n + 256 - (n % 256)

    @property
    def ntok(self):
        return int(self.time_as_token) + int(self.style_as_token)


def _fold_wn(sd, prefix):
    """w = g * v / ||v|| over all dims but 0 (torch weight_norm dim=0)."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"].float()
    v, g = sd[prefix + ".weight_v"].float(), sd[prefix + ".weight_g"].float()
    n = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return g * v / n


def _interleave_rows(a, b):
    """rows a0,b0,a1,b1,... (pair activations read adjacent accumulator columns)."""
    return torch.stack([a, b], dim=1).reshape(a.shape[0] * 2, *a.shape[1:])


def rope_table(n_pos, head_dim=64, base=10000.0, bf16_round=False):
    """cos/sin table exactly as the reference builds it
    (modules/diffusion_transformer.py:288-297; v2 rounds it to bf16)."""
    freqs = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(n_pos), freqs)
    fc = torch.polar(torch.ones_like(ang), ang)
    tab = torch.stack([fc.real, fc.imag], dim=-1)
    if bf16_round:
        tab = tab.to(torch.bfloat16).float()
    return tab.contiguous()


class DiTEngine:
    def __init__(self, spec: DiTSpec, ops):
        self.spec = spec
        self.ops = ops
        self.w = None
        self.rope = None
        self.max_pos = 0

    # ------------------------------------------------------------------ weights
    def load_weights(self, sd, device):
        """Fold weight-norm, fuse / interleave / split and cast the reference state_dict."""
        self._cw = None
        sp, ops = self.spec, self.ops
        p = sp.prefix
        od = ops.op_dtype
        sdt = ops.stream_dtype        # operands that carry a residual stream (see Ops.stream_dtype)
        D, L, C = sp.D, sp.L, sp.C

        def dev(t, dtype=None):
            return t.detach().to(device=device, dtype=dtype or torch.float32).contiguous()

        w = {}
        layers = []
        # v1 RMS norms folded into the GEMMs around them (see fold_begin): needs the fp32 masters of wqkv / w13
        self.fold = bool(getattr(ops, "fold_norms", False)) and sp.version == 1
        ada_w, ada_b = [], []          # stacked AdaLN projections, fp32
        ada_index = {}

        def add_ada(name, weight, bias, one_plus_rows=()):
            bias = bias.clone().float()
            for lo, hi in one_plus_rows:      # "(1 + scale)" folded into the bias
                bias[lo:hi] += 1.0
            ada_index[name] = (sum(x.shape[0] for x in ada_w), weight.shape[0])
            ada_w.append(weight.float())
            ada_b.append(bias)

        for i in range(L):
            lp = f"{p}transformer.layers.{i}."
            lw = {
                "wqkv": dev(sd[lp + "attention.wqkv.weight"], od),
                "wo": dev(sd[lp + "attention.wo.weight"], od),
                "w13": dev(_interleave_rows(sd[lp + "feed_forward.w1.weight"].float(),
                                            sd[lp + "feed_forward.w3.weight"].float()), od),
                "w2": dev(sd[lp + "feed_forward.w2.weight"], od),
            }
            if self.fold:
                lw["wqkv_f32"] = dev(sd[lp + "attention.wqkv.weight"])
                lw["w13_f32"] = dev(_interleave_rows(sd[lp + "feed_forward.w1.weight"].float(),
                                                     sd[lp + "feed_forward.w3.weight"].float()))
            if sp.version == 1:
                lw["g_attn"] = dev(sd[lp + "attention_norm.norm.weight"])
                lw["g_ffn"] = dev(sd[lp + "ffn_norm.norm.weight"])
                if not sp.time_as_token:
                    add_ada(f"attn{i}", sd[lp + "attention_norm.project_layer.weight"],
                            sd[lp + "attention_norm.project_layer.bias"])
                    add_ada(f"ffn{i}", sd[lp + "ffn_norm.project_layer.weight"],
                            sd[lp + "ffn_norm.project_layer.bias"])
                if sp.uvit and i > L // 2:
                    sk = dev(sd[lp + "skip_in_linear.weight"], sdt)
                    lw["skip_w"] = sk
                    lw["skip_b"] = dev(sd[lp + "skip_in_linear.bias"])
            else:
                lw["g_attn"] = dev(sd[lp + "attention_norm.norm.weight"])
                lw["g_ffn"] = dev(sd[lp + "ffn_norm.weight"])
                # chunks: shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
                add_ada(f"blk{i}", sd[lp + "attention_norm.linear.weight"],
                        sd[lp + "attention_norm.linear.bias"],
                        one_plus_rows=((D, 2 * D), (4 * D, 5 * D)))
            layers.append(lw)
        w["layers"] = layers
        w["g_final"] = dev(sd[p + "transformer.norm.norm.weight"])
        if sp.version == 1:
            add_ada("final", sd[p + "transformer.norm.project_layer.weight"],
                    sd[p + "transformer.norm.project_layer.bias"])
        else:  # chunks: scale, shift (dit_model.py:50-53)
            add_ada("final", sd[p + "transformer.norm.linear.weight"],
                    sd[p + "transformer.norm.linear.bias"], one_plus_rows=((0, D),))
        if sp.head == "wavenet":  # FinalLayer: shift, scale = chunk(Linear(SiLU(t1)))
            Dw = sp.Dw
            add_ada("fl", sd[p + "final_layer.adaLN_modulation.1.weight"],
                    sd[p + "final_layer.adaLN_modulation.1.bias"], one_plus_rows=((Dw, 2 * Dw),))
        w["ada_w"] = dev(torch.cat(ada_w, 0))
        w["ada_b"] = dev(torch.cat(ada_b, 0))
        w["ada_index"] = ada_index
        # which rows of the stacked projection take SiLU(t1) instead of t1
        w["ada_silu"] = sp.version == 2

        for name in ("t_embedder",) + (("t_embedder2",) if sp.head == "wavenet" else ()):
            w[name] = {
                "w0": dev(sd[f"{p}{name}.mlp.0.weight"]), "b0": dev(sd[f"{p}{name}.mlp.0.bias"]),
                "w2": dev(sd[f"{p}{name}.mlp.2.weight"]), "b2": dev(sd[f"{p}{name}.mlp.2.bias"]),
            }
        half = 128
        w["freqs"] = dev(torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half))

        w["cond_w"] = dev(sd[p + "cond_projection.weight"], sdt)
        w["cond_b"] = dev(sd[p + "cond_projection.bias"])
        mw = sd[p + "cond_x_merge_linear.weight"].float()
        w["merge_w"] = dev(mw, sdt)                      # column blocks are taken as views
        w["merge_b"] = dev(sd[p + "cond_x_merge_linear.bias"])
        # null-branch constant of the content columns: W_c @ b_cond (fp32, exact)
        Wc = mw[:, 2 * C:2 * C + D]
        w["merge_null_c"] = dev(Wc @ sd[p + "cond_projection.bias"].float())
        if sp.style_in_merge:
            w["merge_ws"] = dev(mw[:, 2 * C + D:2 * C + D + sp.style_dim])   # fp32 (B rows only)
        if sp.style_as_token or sp.version == 2:
            w["style_in_w"] = dev(sd[p + "style_in.weight"])
            w["style_in_b"] = dev(sd[p + "style_in.bias"])
        if sp.long_skip:
            w["lskip_w"] = dev(sd[p + "skip_linear.weight"], sdt)
            w["lskip_b"] = dev(sd[p + "skip_linear.bias"])
        if sp.head == "mlp":
            w["mlp0_w"] = dev(sd[p + "final_mlp.0.weight"], sdt)
            w["mlp0_b"] = dev(sd[p + "final_mlp.0.bias"])
            w["mlp2_w"] = dev(sd[p + "final_mlp.2.weight"], sdt)
            w["mlp2_b"] = dev(sd[p + "final_mlp.2.bias"])
        else:
            Dw, nl, ks = sp.Dw, sp.wn_layers, sp.wn_kernel
            w["conv1_w"] = dev(sd[p + "conv1.weight"], sdt)
            w["conv1_b"] = dev(sd[p + "conv1.bias"])
            w["resp_w"] = dev(sd[p + "res_projection.weight"], sdt)
            w["resp_b"] = dev(sd[p + "res_projection.bias"])
            w["conv2_w"] = dev(sd[p + "conv2.weight"].float().reshape(C, Dw), sdt)
            w["conv2_b"] = dev(sd[p + "conv2.bias"])
            w["fl_w"] = dev(_fold_wn(sd, p + "final_layer.linear"), sdt)
            w["fl_b"] = dev(sd[p + "final_layer.linear.bias"])
            cw = _fold_wn(sd, p + "wavenet.cond_layer.conv.conv").reshape(2 * Dw * nl, Dw)
            cb = sd[p + "wavenet.cond_layer.conv.conv.bias"].float()
            cws, cbs, wn = [], [], []
            for l in range(nl):
                wi = _fold_wn(sd, f"{p}wavenet.in_layers.{l}.conv.conv")      # (2Dw, Dw, k)
                bi = sd[f"{p}wavenet.in_layers.{l}.conv.conv.bias"].float()
                wi = _interleave_rows(wi[:Dw], wi[Dw:])                        # tanh/sigmoid pairs
                bi = _interleave_rows(bi[:Dw], bi[Dw:])
                cl = cw[l * 2 * Dw:(l + 1) * 2 * Dw]
                cbl = cb[l * 2 * Dw:(l + 1) * 2 * Dw]
                cws.append(_interleave_rows(cl[:Dw], cl[Dw:]))
                cbs.append(_interleave_rows(cbl[:Dw], cbl[Dw:]) + bi)           # conv bias folded in
                wr = _fold_wn(sd, f"{p}wavenet.res_skip_layers.{l}.conv.conv")
                wr = wr.reshape(wr.shape[0], Dw)
                br = sd[f"{p}wavenet.res_skip_layers.{l}.conv.conv.bias"].float()
                wn.append({
                    "in_w": dev(wi.permute(2, 0, 1), sdt),                     # (k, 2Dw, Dw)
                    "rs_w": dev(wr, sdt), "rs_b": dev(br),
                    "rs_b_res": dev(br[:Dw]), "rs_b_skip": dev(br[Dw:]),
                })
            w["wn"] = wn
            w["wn_cond_w"] = dev(torch.cat(cws, 0))                            # (nl*2Dw, Dw) fp32
            w["wn_cond_b"] = dev(torch.cat(cbs, 0))
            # skip path of all layers as ONE contraction over the concatenated gated activations:
            # output = sum_l acts_l @ W_skip_l^T (+ res_projection as a second segment)
            skip_w, skip_b = [], sd[p + "res_projection.bias"].float().clone()
            for l in range(nl):
                wr = _fold_wn(sd, f"{p}wavenet.res_skip_layers.{l}.conv.conv")
                wr = wr.reshape(wr.shape[0], Dw)
                br = sd[f"{p}wavenet.res_skip_layers.{l}.conv.conv.bias"].float()
                skip_w.append(wr[Dw:] if l < nl - 1 else wr)
                skip_b += br[Dw:] if l < nl - 1 else br
            w["wn_skip_w"] = dev(torch.cat(skip_w, 1), sdt)                    # (Dw, nl*Dw)
            w["wn_skip_b"] = dev(skip_b)
        self.w = w

    def setup_rope(self, n_pos, device):
        if self.rope is None or self.max_pos < n_pos:
            self.rope = rope_table(n_pos, bf16_round=self.spec.version == 2).to(device)
            self.rope_t = self.rope.permute(1, 0, 2).contiguous()      # pair-major copy for the row-layout epilogue
            self.max_pos = n_pos
            self._cw = None

    # ------------------------------------------------------------------ C structs of svc_dit_step
    def _c_weights(self):
        """svc_dit_weights over the prepared tensors (kept alive by self.w)."""
        if getattr(self, "_cw", None) is not None:
            return self._cw
        from . import _lib
        sp, w, ops = self.spec, self.w, self.ops
        c = _lib.DitWeights()
        c.version, c.D, c.H, c.L, c.C, c.I = sp.version, sp.D, sp.H, sp.L, sp.C, sp.I
        c.time_as_token, c.style_as_token = int(sp.time_as_token), int(sp.style_as_token)
        c.uvit, c.long_skip = int(sp.uvit), int(sp.long_skip)
        c.head = 1 if sp.head == "wavenet" else 0
        c.Dw, c.wn_layers, c.wn_kernel = sp.Dw, sp.wn_layers, sp.wn_kernel
        c.op_dtype, c.stream_dtype = ops._code(ops.op_dtype), ops._code(ops.stream_dtype)
        idx = w["ada_index"]
        for i, lw in enumerate(w["layers"]):
            c.wqkv[i], c.wo[i], c.w13[i], c.w2[i] = (lw[k].data_ptr() for k in ("wqkv", "wo", "w13", "w2"))
            c.g_attn[i], c.g_ffn[i] = lw["g_attn"].data_ptr(), lw["g_ffn"].data_ptr()
            if "skip_w" in lw:
                c.skip_w[i], c.skip_b[i] = lw["skip_w"].data_ptr(), lw["skip_b"].data_ptr()
            if sp.version == 1:
                c.ada_attn[i] = idx[f"attn{i}"][0] if f"attn{i}" in idx else -1
                c.ada_ffn[i] = idx[f"ffn{i}"][0] if f"ffn{i}" in idx else -1
            else:
                c.ada_attn[i], c.ada_ffn[i] = idx[f"blk{i}"][0], -1
        c.g_final, c.ada_final = w["g_final"].data_ptr(), idx["final"][0]
        c.merge_wx, c.merge_w_rstride = w["merge_w"].data_ptr(), w["merge_w"].stride(0)
        if sp.long_skip:
            c.lskip_w, c.lskip_b = w["lskip_w"].data_ptr(), w["lskip_b"].data_ptr()
        if sp.head == "mlp":
            c.mlp0_w, c.mlp0_b = w["mlp0_w"].data_ptr(), w["mlp0_b"].data_ptr()
            c.mlp2_w, c.mlp2_b = w["mlp2_w"].data_ptr(), w["mlp2_b"].data_ptr()
        else:
            c.conv1_w, c.conv1_b, c.resp_w = w["conv1_w"].data_ptr(), w["conv1_b"].data_ptr(), w["resp_w"].data_ptr()
            c.conv2_w, c.conv2_b = w["conv2_w"].data_ptr(), w["conv2_b"].data_ptr()
            c.fl_w, c.fl_b = w["fl_w"].data_ptr(), w["fl_b"].data_ptr()
            for l, wl in enumerate(w["wn"]):
                c.wn_in_w[l], c.wn_rs_w[l] = wl["in_w"].data_ptr(), wl["rs_w"].data_ptr()
                c.wn_rs_b[l] = wl["rs_b_res"].data_ptr()
            c.wn_skip_w, c.wn_skip_b = w["wn_skip_w"].data_ptr(), w["wn_skip_b"].data_ptr()
            c.ada_fl = idx["fl"][0]
        c.rope_tab, c.rope_tab_t, c.rope_ld = self.rope.data_ptr(), self.rope_t.data_ptr(), self.rope.shape[0]
        self._cw = c
        return c

    def _c_state(self, st):
        from . import _lib
        sp = self.spec
        c = _lib.DitState()
        c.B, c.T, c.n_branch, c.n_steps, c.n_ada = st["B"], st["T"], st["nb"], st["N"], st["ada"].shape[1]
        c.ada, c.t1 = st["ada"].data_ptr(), st["t1"].data_ptr()
        if "wn_g" in st:
            c.wn_g = st["wn_g"].data_ptr()
        for k, (kind, val) in enumerate(st["consts"]):
            c.const_kind[k], c.const_ptr[k] = (0 if kind == "mat" else 1), val.data_ptr()
        if sp.style_as_token:
            c.style_tok, c.style_tok_null = st["style_tok"].data_ptr(), st["style_tok_null"].data_ptr()
            for k, (use_p, use_s, use_m) in enumerate(st["branches"]):
                c.branch_style[k] = int(bool(use_s))
        c.kv_len = st["kv_len"].data_ptr()
        for name in ("h", "xn", "xn_f", "qkv", "att", "ff", "h_op", "v", "x_res", "y", "xw", "xw_op", "acts",
                     "wn_out", "ln", "wn_lens"):
            if name in st:
                setattr(c, name, st[name].data_ptr())
        for k, t in enumerate(st["skips"]):
            c.skips[k] = t.data_ptr()
        return c

    # ------------------------------------------------------------------ per-solve precompute
    def begin(self, branches, prompt_op, mu, style, x_lens, t_values):
        """branches: list of (use_prompt, use_style, use_mu) flags, one per CFG branch.

        prompt_op: (B, T, C) operand dtype (zeros outside the prompt); mu: (B, T, cd) fp32;
        style: (B, style_dim) fp32; x_lens: (B,) int; t_values: (N,) fp32 CPU tensor.
        """
        sp, ops, w = self.spec, self.ops, self.w
        dev = mu.device
        B, T, _ = mu.shape
        D, C, L = sp.D, sp.C, sp.L
        nb = len(branches)
        ntok = sp.ntok
        Tq = T + ntok
        N = int(t_values.numel())
        self.setup_rope(max(Tq, 1), dev)
        st = {"B": B, "T": T, "Tq": Tq, "nb": nb, "N": N, "branches": branches}
        f32 = torch.float32

        # ---- time conditioning for every step at once (rows = steps) -------------------
        tdev = t_values.to(device=dev, dtype=f32).contiguous()

        def t_embed(name):
            tw = w[name]
            Dh = tw["w0"].shape[0]
            feat = torch.empty(N, 256, dtype=f32, device=dev)
            ops.timestep_embedding(tdev, w["freqs"], feat)
            h0 = torch.empty(1, N, Dh, dtype=f32, device=dev)
            ops.gemm([(feat.view(1, N, 256), 0, tw["w0"])], Dh, B=1, T=N, bias=tw["b0"],
                     act=ACT_SILU, out_f32=h0, f32=True)
            t1 = torch.empty(1, N, Dh, dtype=f32, device=dev)
            ops.gemm([(h0, 0, tw["w2"])], Dh, B=1, T=N, bias=tw["b2"], out_f32=t1, f32=True)
            t1s = torch.empty(1, N, Dh, dtype=f32, device=dev)
            ops.gemm([(h0, 0, tw["w2"])], Dh, B=1, T=N, bias=tw["b2"], act=ACT_SILU, out_f32=t1s,
                     f32=True)
            return t1, t1s

        t1, t1s = t_embed("t_embedder")
        st["t1"] = t1[0]                                   # (N, D)
        n_ada = w["ada_w"].shape[0]
        ada = torch.empty(1, N, n_ada, dtype=f32, device=dev)
        if sp.version == 2:
            ops.gemm([(t1s, 0, w["ada_w"])], n_ada, B=1, T=N, bias=w["ada_b"], out_f32=ada, f32=True)
        else:
            # v1: AdaLN projections take t1; the FinalLayer modulation takes SiLU(t1)
            if sp.head == "wavenet":
                lo, n_fl = w["ada_index"]["fl"]
                ops.gemm([(t1, 0, w["ada_w"][:lo])], lo, B=1, T=N, bias=w["ada_b"][:lo],
                         out_f32=ada[:, :, :lo], f32=True)
                ops.gemm([(t1s, 0, w["ada_w"][lo:])], n_fl, B=1, T=N, bias=w["ada_b"][lo:],
                         out_f32=ada[:, :, lo:], f32=True)
            else:
                ops.gemm([(t1, 0, w["ada_w"])], n_ada, B=1, T=N, bias=w["ada_b"], out_f32=ada,
                         f32=True)
        st["ada"] = ada[0]                                 # (N, n_ada)
        self.st_ada = ada[0]
        if sp.head == "wavenet":
            t2, _ = t_embed("t_embedder2")
            ng = w["wn_cond_w"].shape[0]
            g = torch.empty(1, N, ng, dtype=f32, device=dev)
            ops.gemm([(t2, 0, w["wn_cond_w"])], ng, B=1, T=N, bias=w["wn_cond_b"], out_f32=g, f32=True)
            st["wn_g"] = g[0]                              # (N, nl*2Dw), conv bias included

        # ---- step-invariant part of cond_x_merge_linear ----------------------------------
        sdt = ops.stream_dtype
        mu_op = ops.empty(B, T, sp.content_dim, device=dev, dtype=sdt)
        ops.cast(mu.contiguous(), mu_op)
        cond_op = ops.empty(B, T, D, device=dev, dtype=sdt)
        ops.gemm([(mu_op, 0, w["cond_w"])], D, B=B, T=T, bias=w["cond_b"], out_op=cond_op)
        mw = w["merge_w"]
        Wx, Wp, Wc = mw[:, :C], mw[:, C:2 * C], mw[:, 2 * C:2 * C + D]
        st["Wx"] = Wx
        style_rb = None
        if sp.style_in_merge:
            style_rb = torch.empty(1, B, D, dtype=f32, device=dev)   # style @ Ws^T + b_merge
            ops.gemm([(style.contiguous().view(1, B, sp.style_dim), 0, w["merge_ws"])], D, B=1, T=B,
                     bias=w["merge_b"], out_f32=style_rb, f32=True)
        null_vec = (w["merge_null_c"] + w["merge_b"]).contiguous()   # tiny host-side prep
        consts = []
        cache = {}
        for (use_p, use_s, use_m) in branches:
            key = (use_p, use_s and sp.style_in_merge, use_m)
            if key in cache:
                consts.append(cache[key])
                continue
            segs = []
            if use_p:
                segs.append((prompt_op, 0, Wp))
            if use_m:
                segs.append((cond_op, 0, Wc))
            if not segs:
                entry = ("vec", null_vec if not key[1] else None)
                if key[1]:   # style without prompt/content: not produced by any reference branch
                    raise NotImplementedError("style-only CFG branch")
            else:
                Hk = torch.empty(B, T, D, dtype=f32, device=dev)
                rb, bias = None, w["merge_b"]
                if key[1]:
                    rb, bias = style_rb[0], None
                if not use_m:   # content columns see cond_projection(0) = b_cond
                    bias = (w["merge_null_c"] + (w["merge_b"] if bias is not None else 0)).contiguous()
                ops.gemm(segs, D, B=B, T=T, bias=bias, rowbias=rb, out_f32=Hk)
                entry = ("mat", Hk)
            cache[key] = entry
            consts.append(entry)
        st["consts"] = consts

        # ---- tokens ------------------------------------------------------------------------
        if sp.style_as_token:
            tok = torch.empty(1, B, D, dtype=f32, device=dev)
            ops.gemm([(style.contiguous().view(1, B, sp.style_dim), 0, w["style_in_w"])], D, B=1, T=B,
                     bias=w["style_in_b"], out_f32=tok, f32=True)
            st["style_tok"] = tok[0]                            # (B, D)
            st["style_tok_null"] = w["style_in_b"].view(1, D)   # style_in(0)
        kv = (x_lens.to(device=dev, dtype=torch.int32) + ntok).repeat(nb).contiguous()
        st["kv_len"] = kv
        st["x_lens"] = x_lens.to(device=dev, dtype=torch.int32).contiguous()

        # ---- work buffers ------------------------------------------------------------------
        R = nb * B
        od = ops.op_dtype
        st["h"] = torch.empty(R, Tq, D, dtype=f32, device=dev)
        st["xn"] = torch.empty(R, Tq, D, dtype=od, device=dev)
        st["qkv"] = torch.empty(R, Tq, 3 * D, dtype=od, device=dev)
        st["att"] = torch.empty(R, Tq, D, dtype=od, device=dev)
        st["ff"] = torch.empty(R, Tq, sp.I, dtype=od, device=dev)
        st["h_op"] = torch.empty(R, Tq, D, dtype=sdt, device=dev)
        # output of the final norm: input of the head chain (stream dtype); aliases xn when the dtypes agree
        st["xn_f"] = st["xn"] if sdt == od else torch.empty(R, Tq, D, dtype=sdt, device=dev)
        n_skip = L // 2 if sp.uvit else 0
        st["skips"] = [torch.empty(R, Tq, D, dtype=sdt, device=dev) for _ in range(n_skip)]
        st["v"] = torch.empty(R, T, C, dtype=f32, device=dev)
        if sp.long_skip:
            st["x_res"] = torch.empty(R, T, D, dtype=sdt, device=dev)
        if sp.head == "mlp":
            st["y"] = torch.empty(R, T, D, dtype=sdt, device=dev)
        else:
            Dw, pad = sp.Dw, (sp.wn_kernel - 1) // 2
            st["xw"] = torch.empty(R, T, Dw, dtype=f32, device=dev)
            st["xw_op"] = torch.zeros(R, T + 2 * pad, Dw, dtype=sdt, device=dev)
            st["acts"] = torch.empty(R, T, sp.wn_layers * Dw, dtype=sdt, device=dev)
            st["wn_out"] = torch.empty(R, T, Dw, dtype=f32, device=dev)
            st["ln"] = torch.empty(R, T, Dw, dtype=sdt, device=dev)
            st["y"] = torch.empty(R, T, Dw, dtype=sdt, device=dev)
            st["wn_lens"] = st["x_lens"].repeat(nb).contiguous()
        if self.fold:
            self._fold_begin(st, N, R, Tq, dev)
        self.st = st
        st["c_state"] = self._c_state(st) if hasattr(ops, "dit_step") else None
        return st

    # ------------------------------------------------------------------ folded RMS norms (v1)
    def _copy_dtype(self, i):
        """dtype of the operand copy of h that layer i's wqkv reads: the previous layer's w2 writes it; when that
        layer is a U-ViT emit layer the copy IS the skip tensor (stream dtype)."""
        sp = self.spec
        emit_prev = sp.uvit and (i - 1) < sp.L // 2
        return self.ops.stream_dtype if emit_prev else self.ops.op_dtype

    def _fold_begin(self, st, N, R, Tq, dev):
        """AdaptiveLayerNorm over RMSNorm feeding a Linear (diffusion_transformer.py:30-48, 173-191):
            (h * r * g * w_s + b_s) W^T  =  r * (h (W * g * w_s)^T) + b_s W^T,   r = rsqrt(mean(h^2) + eps) per row,
        so per Euler step s and layer the weight W'_s = W * (g * w_s) and the bias b_s W^T are prepared here (fp32
        masters, one rounding), the GEMM that produces h also writes a 16-bit copy and the row sums of squares
        (row_ss_out) and wqkv / w13 scale their accumulator rows (row_scale).  Removes both norm passes of a
        layer; layer 0's attention norm (h assembled from several launches) and the final norm stay kernels."""
        sp, ops, w = self.spec, self.ops, self.w
        D, L = sp.D, sp.L
        f32 = torch.float32
        S = 1 if sp.time_as_token else N
        st["fold_S"] = S
        st["ss"] = torch.zeros(R * Tq, 8, dtype=f32, device=dev)          # SVC_SS_SLOTS
        st["hn"] = torch.empty(R, Tq, D, dtype=ops.op_dtype, device=dev)   # copy of h for wqkv (non-skip layers)
        st["hn2"] = torch.empty(R, Tq, D, dtype=ops.op_dtype, device=dev)  # copy of h for w13
        fw = []
        for i in range(L):
            lw = w["layers"][i]
            e = {}
            for side, gname, wname, n_out in (("q", "g_attn", "wqkv_f32", 3 * D), ("f", "g_ffn", "w13_f32", 2 * sp.I)):
                if side == "q" and i == 0:
                    continue
                mul = add = None
                if not sp.time_as_token:
                    a = self._ada_all(("attn" if side == "q" else "ffn") + str(i))     # (N, 2D): weight, bias
                    mul, add = a[:, :D], a[:, D:]
                dt = self._copy_dtype(i) if side == "q" else ops.op_dtype
                Wf = torch.empty(S, n_out, D, dtype=dt, device=dev)
                ops.scale_cols(lw[wname], lw[gname], mul, Wf)
                rb = torch.zeros(1, S, n_out, dtype=f32, device=dev)
                if add is not None:
                    ops.gemm([(add.unsqueeze(0), 0, lw[wname])], n_out, B=1, T=S, out_f32=rb, f32=True)
                e["w" + side], e["b" + side] = Wf, rb[0]
            fw.append(e)
        st["fold_w"] = fw

    def _ada_all(self, name):
        lo, n = self.w["ada_index"][name]
        return self.st_ada[:, lo:lo + n]

    def launches_per_step(self):
        """Kernel launches of one estimator call (for the gpu_launches bookkeeping of the C path)."""
        sp, st = self.spec, self.st
        nb, L = st["nb"], sp.L
        n = nb + int(sp.time_as_token) + (nb if sp.style_as_token else 0)
        n += L * 7 + (L - L // 2 - 1 if sp.uvit else 0) + 1
        if self.fold:
            n -= 2 * L - 1
        n += nb if sp.long_skip else 0
        n += 2 if sp.head == "mlp" else (2 + sp.wn_layers + 2 * (sp.wn_layers - 1) + 1 + 3)
        return n

    # ------------------------------------------------------------------ one estimator call
    def _ada(self, s, name):
        lo, n = self.w["ada_index"][name]
        return self.st["ada"][s, lo:lo + n]

    def _layers_folded(self, s, rope, emit, recv):
        """The transformer layers with the RMS norms folded into the GEMMs (see _fold_begin)."""
        sp, ops, w, st = self.spec, self.ops, self.w, self.st
        D, L, H = sp.D, sp.L, sp.H
        R, Tq = st["nb"] * st["B"], st["Tq"]
        h, xn, qkv, att, ff = st["h"], st["xn"], st["qkv"], st["att"], st["ff"]
        ss, hn, hn2 = st["ss"], st["hn"], st["hn2"]
        fs = 0 if st["fold_S"] == 1 else s
        skips, skip_bufs = [], list(st["skips"])
        a_in = None                     # operand copy of h for this layer's wqkv (None: layer 0, norm kernel)
        for i in range(L):
            lw, fw = w["layers"][i], st["fold_w"][i]
            if i in recv:
                skip = skips.pop()
                ops.gemm([(st["h_op"], 0, lw["skip_w"][:, :D]), (skip, 0, lw["skip_w"][:, D:])], D,
                         B=R, T=Tq, bias=lw["skip_b"], out_f32=h, out_op=hn, row_ss_out=ss)
                a_in = hn
            if a_in is None:
                if sp.time_as_token:
                    ops.norm_mod(h, xn, gamma=lw["g_attn"])
                else:
                    a = self._ada(s, f"attn{i}")
                    ops.norm_mod(h, xn, gamma=lw["g_attn"], mul=a[:D], add=a[D:])
                ops.gemm([(xn, 0, lw["wqkv"])], 3 * D, B=R, T=Tq, act=ACT_ROPE, rope=rope, out_op=qkv)
            else:
                ops.gemm([(a_in, 0, fw["wq"][fs])], 3 * D, B=R, T=Tq, bias=fw["bq"][fs], act=ACT_ROPE, rope=rope,
                         out_op=qkv, row_scale=(ss, D, 1e-5))
            ops.attention(qkv, att, H, st["kv_len"])
            ops.gemm([(att, 0, lw["wo"])], D, B=R, T=Tq, res=h, out_f32=h, out_op=hn2, row_ss_out=ss)
            ops.gemm([(hn2, 0, fw["wf"][fs])], 2 * sp.I, B=R, T=Tq, bias=fw["bf"][fs], act=ACT_SWIGLU_PAIR,
                     out_op=ff, row_scale=(ss, D, 1e-5))
            # w2 also writes the operand copy its consumer reads: the skip tensor (emit layers: next wqkv reads
            # the skip buffer itself), h_op (skip_in_linear of a receive layer comes next), or hn
            nxt_recv = (i + 1) in recv
            if i in emit:
                copy = skip_bufs.pop(0)
                skips.append(copy)
            elif nxt_recv:
                copy = st["h_op"]
            else:
                copy = hn if i + 1 < L else None
            direct = copy is not None and not nxt_recv          # next layer's wqkv consumes (copy, ss) directly
            ops.gemm([(ff, 0, lw["w2"])], D, B=R, T=Tq, res=h, out_f32=h, out_op=copy,
                     row_ss_out=ss if direct else None)
            assert not (i in emit and nxt_recv)
            a_in = copy if direct else None
            assert a_in is None or a_in.dtype == self._copy_dtype(i + 1)

    def step(self, s, x_op):
        """Velocity of every branch at step ``s``: (nb*B, T, C) fp32.  x_op: (B, T, C).

        On the CUDA library this is ONE call of the graph-level C entry point ``svc_dit_step`` (csrc/graph.cu),
        which issues exactly the launch sequence written out below; the Python sequence runs when per-launch
        profiling is on (bench.py's kernel breakdown) and on the emulated ops of the CPU host-logic tests."""
        if self.st.get("c_state") is not None and self.ops.profile is None and not self.fold:
            assert x_op.is_contiguous() and x_op.dtype == self.ops.stream_dtype
            self.ops.dit_step(self._c_weights(), self.st["c_state"], s, x_op, self.launches_per_step())
            return self.st["v"]
        sp, ops, w, st = self.spec, self.ops, self.w, self.st
        B, T, Tq, nb = st["B"], st["T"], st["Tq"], st["nb"]
        D, C, L, H = sp.D, sp.C, sp.L, sp.H
        ntok = sp.ntok
        R = nb * B
        h, xn, qkv, att, ff = st["h"], st["xn"], st["qkv"], st["att"], st["ff"]
        hb = h[:, ntok:, :]                                   # frame rows of the hidden state

        # ---- input assembly: x columns of cond_x_merge_linear + hoisted constants ----------
        for k, (kind, val) in enumerate(st["consts"]):
            out = hb[k * B:(k + 1) * B]
            if kind == "mat":
                ops.gemm([(x_op, 0, st["Wx"])], D, B=B, T=T, res=val, out_f32=out)
            else:
                ops.gemm([(x_op, 0, st["Wx"])], D, B=B, T=T, bias=val, out_f32=out)
        if sp.time_as_token:
            ops.set_rows(st["t1"][s:s + 1], h[:, 0, :])
        if sp.style_as_token:
            r = int(sp.time_as_token)
            for k, (use_p, use_s, use_m) in enumerate(st["branches"]):
                src = st["style_tok"] if use_s else st["style_tok_null"]
                ops.set_rows(src, h[k * B:(k + 1) * B, r, :])

        rope = (self.rope, 2 * D, 0, D, 0.125)
        emit = set(range(L // 2)) if sp.uvit else set()
        recv = set(range(L // 2 + 1, L)) if sp.uvit else set()
        skips = []
        skip_bufs = list(st["skips"])
        pending_raw = None        # skip tensor of the previous (emit) layer: written by this layer's first norm
        if self.fold:
            self._layers_folded(s, rope, emit, recv)
        for i in range(L if not self.fold else 0):
            lw = w["layers"][i]
            if i in recv:
                skip = skips.pop()
                ops.gemm([(st["h_op"], 0, lw["skip_w"][:, :D]), (skip, 0, lw["skip_w"][:, D:])], D,
                         B=R, T=Tq, bias=lw["skip_b"], out_f32=h)
            # ---- attention --------------------------------------------------------------
            if sp.version == 1:
                # the operand copy of an emit layer's output (the U-ViT skip tensor) falls out of this norm's
                # read of h, so that layer's w2 GEMM needs only its fp32 output
                if sp.time_as_token:
                    ops.norm_mod(h, xn, gamma=lw["g_attn"], raw_out=pending_raw)
                else:
                    a = self._ada(s, f"attn{i}")
                    ops.norm_mod(h, xn, gamma=lw["g_attn"], mul=a[:D], add=a[D:], raw_out=pending_raw)
                pending_raw = None
                gate_a = gate_m = None
            else:
                a = self._ada(s, f"blk{i}")     # shift, 1+scale, gate, shift, 1+scale, gate
                ops.norm_mod(h, xn, gamma=lw["g_attn"], mul=a[D:2 * D], add=a[:D])
                gate_a = a[2 * D:3 * D].view(1, D).expand(R, D)
                gate_m = a[5 * D:6 * D].view(1, D).expand(R, D)
            ops.gemm([(xn, 0, lw["wqkv"])], 3 * D, B=R, T=Tq, act=ACT_ROPE, rope=rope, out_op=qkv)
            ops.attention(qkv, att, H, st["kv_len"])
            ops.gemm([(att, 0, lw["wo"])], D, B=R, T=Tq, gate=gate_a, res=h, out_f32=h)
            # ---- feed-forward -----------------------------------------------------------
            if sp.version == 1:
                if sp.time_as_token:
                    ops.norm_mod(h, xn, gamma=lw["g_ffn"])
                else:
                    a = self._ada(s, f"ffn{i}")
                    ops.norm_mod(h, xn, gamma=lw["g_ffn"], mul=a[:D], add=a[D:])
            else:
                ops.norm_mod(h, xn, gamma=lw["g_ffn"], mul=a[4 * D:5 * D], add=a[3 * D:4 * D])
            ops.gemm([(xn, 0, lw["w13"])], 2 * sp.I, B=R, T=Tq, act=ACT_SWIGLU_PAIR, out_op=ff)
            out_op = None
            if i in emit and sp.version == 1 and (i + 1) < L and (i + 1) not in recv:
                pending_raw = skip_bufs.pop(0)       # filled by layer i+1's attention norm
                skips.append(pending_raw)
            elif i in emit:
                out_op = skip_bufs.pop(0)
                skips.append(out_op)
            elif (i + 1) in recv:
                out_op = st["h_op"]
            ops.gemm([(ff, 0, lw["w2"])], D, B=R, T=Tq, gate=gate_m, res=h, out_f32=h, out_op=out_op)
        # ---- final norm (uses c even when time is a token, diffusion_transformer.py:142) ----
        a = self._ada(s, "final")         # v1: w, b of the AdaLN projection; v2: (1 + scale), shift
        ops.norm_mod(h, st["xn_f"], gamma=w["g_final"], mul=a[:D], add=a[D:])
        xf = st["xn_f"][:, ntok:, :]
        v = st["v"]
        if sp.long_skip:     # skip_linear(cat[x_res, x]) without the concat (:524-525)
            lw_, lb_ = w["lskip_w"], w["lskip_b"]
            x_res = st["x_res"]
            for k in range(nb):
                sl = slice(k * B, (k + 1) * B)
                ops.gemm([(xf[sl], 0, lw_[:, :D]), (x_op, 0, lw_[:, D:D + C])], D, B=B, T=T, bias=lb_,
                         out_op=x_res[sl])
            xr = x_res
        else:
            xr = xf
        if sp.head == "mlp":
            y = st["y"]
            ops.gemm([(xr, 0, w["mlp0_w"])], D, B=R, T=T, bias=w["mlp0_b"], act=ACT_SILU, out_op=y)
            ops.gemm([(y, 0, w["mlp2_w"])], C, B=R, T=T, bias=w["mlp2_b"], out_f32=v)
            return v
        # ---- WaveNet head -------------------------------------------------------------------
        Dw, nl, ks = sp.Dw, sp.wn_layers, sp.wn_kernel
        pad = (ks - 1) // 2
        xw, xw_op, acts, wn_out = st["xw"], st["xw_op"], st["acts"], st["wn_out"]
        body = xw_op[:, pad:pad + T, :]
        ops.gemm([(xr, 0, w["conv1_w"])], Dw, B=R, T=T, bias=w["conv1_b"], out_f32=xw, out_op=body)
        ops.reflect_halo(xw_op, T, pad, st["wn_lens"])
        g_all = st["wn_g"][s]
        for l in range(nl):
            wl = w["wn"][l]
            g_l = g_all[l * 2 * Dw:(l + 1) * 2 * Dw].view(1, 2 * Dw).expand(R, 2 * Dw)
            acts_l = acts[:, :, l * Dw:(l + 1) * Dw]
            segs = [(xw_op, j, wl["in_w"][j]) for j in range(ks)]
            ops.gemm(segs, 2 * Dw, B=R, T=T, rowbias=g_l, act=ACT_TANH_SIG_PAIR, out_op=acts_l)
            if l < nl - 1:
                ops.gemm([(acts_l, 0, wl["rs_w"][:Dw])], Dw, B=R, T=T, bias=wl["rs_b_res"],
                         res=xw, out_f32=xw, out_op=body)
                ops.reflect_halo(xw_op, T, pad, st["wn_lens"])
        # output = sum_l skip_l + res_projection(x_res): one GEMM, K = nl*Dw + D, no accumulate passes
        ops.gemm([(acts, 0, w["wn_skip_w"]), (xr, 0, w["resp_w"])], Dw, B=R, T=T, bias=w["wn_skip_b"],
                 out_f32=wn_out)
        a = self._ada(s, "fl")                                   # shift, 1+scale
        ops.norm_mod(wn_out, st["ln"], mul=a[Dw:], add=a[:Dw], eps=1e-6, mode=1)
        ops.gemm([(st["ln"], 0, w["fl_w"])], Dw, B=R, T=T, bias=w["fl_b"], out_op=st["y"])
        ops.gemm([(st["y"], 0, w["conv2_w"])], C, B=R, T=T, bias=w["conv2_b"], out_f32=v)
        return v
