"""BigVGAN-v2 generator on the sm_100a kernels, API-compatible with the reference.

Drop-in surface: ``BigVGAN(h, use_cuda_kernel=False)``, ``from_pretrained``, ``remove_weight_norm``,
``forward(mel (B, n_mels, Tm)) -> (B, 1, Tm * hop)`` (reference: modules/bigvgan/bigvgan.py:266-491).

Execution plan per call (frames-major activations, fp32 residual stream, operand-dtype GEMM inputs):
  conv_pre                      7-tap segmented GEMM                           bigvgan.py:285-287,362
  per stage i:
    ups[i] ConvTranspose1d      polyphase: 3-tap GEMM with N = stride * C_out   :300-316,367
    3 x AMPBlock1               6 x (svc_snake_aa -> k-tap dilated GEMM), residual and the
                                (r0 + r1 + r2) / 3 average fused in GEMM epilogues   :132-141,369-375
  activation_post + conv_post + clamp   one fused kernel                        :377-384
"""
from __future__ import annotations

import json
import os

import torch
from torch import nn

from . import synth
from .configs import AttrDict, to_attr
from .flow_matching import _Params
from .ops import Ops


class _Act(nn.Module):
    """Activation1d(SnakeBeta): parameters + the two FIR buffers of the reference."""

    def __init__(self, ch, filt):
        super().__init__()
        self.act = _Params(alpha=(ch,), beta=(ch,))
        self.upsample = nn.Module()
        self.upsample.register_buffer("filter", filt.clone())
        self.downsample = nn.Module()
        self.downsample.lowpass = nn.Module()
        self.downsample.lowpass.register_buffer("filter", filt.clone())


class _AMPBlock(nn.Module):
    def __init__(self, ch, k, n_dil, filt):
        super().__init__()
        self.convs1 = nn.ModuleList([_Params(weight=(ch, ch, k), bias=(ch,)) for _ in range(n_dil)])
        self.convs2 = nn.ModuleList([_Params(weight=(ch, ch, k), bias=(ch,)) for _ in range(n_dil)])
        self.activations = nn.ModuleList([_Act(ch, filt) for _ in range(2 * n_dil)])


def kaiser_sinc_filter12():
    """Same taps as the kernel's constant table (filter.py:30-62 with 0.25 / 0.3 / 12)."""
    import math
    ks, cutoff, half_width = 12, 0.25, 0.3
    half = ks // 2
    A = 2.285 * (half - 1) * math.pi * 4 * half_width + 7.95
    beta = 0.1102 * (A - 8.7)
    win = torch.kaiser_window(ks, beta=beta, periodic=False)
    time = torch.arange(-half, half) + 0.5
    f = 2 * cutoff * win * torch.sinc(2 * cutoff * time)
    return (f / f.sum()).view(1, 1, ks)


def conv_plan(weight, bias, dilation, od, length_divisor_ok=True):
    """Tap list and weights for a 'same' Conv1d as a segmented GEMM.

    Narrow layers (C_out < 96) are regrouped by *time-to-depth*: f consecutive frames are read
    as one row of f*C channels - the (B, L, C) buffer IS a (B, L/f, f*C) buffer - and the kernel
    becomes a block-structured (f*C x f*C) weight per super-tap:
        y[f*q + p_out] = sum_k w_k x[f*q + p_out + (k - half) d]
                       = sum_{s'} W'[s'][p_out-block, p_in-block] x'[q + s'],  m = p_out + (k-half) d,
                         s' = floor(m / f), p_in = m - f s'.
    Tiles get f x wider (N = f*C fills the 128-row MMA), their number drops f x, and for d = 1 the
    k taps collapse into ~k/f super-taps.  Zero padding is unchanged (rows outside [0, L/f)).
    Chosen only when it lowers the number of 64-wide K blocks issued per output frame.
    Returns dict(f, shifts, w (n_taps, f*O, f*I) operand dtype, b (f*O,) fp32).
    """
    O, I, k = weight.shape
    half = (k - 1) // 2
    w = weight.detach().float()
    b = bias.detach().float()
    best = {"f": 1, "shifts": [(t - half) * dilation for t in range(k)],
            "w": w.permute(2, 0, 1).contiguous().to(od), "b": b.contiguous(), "k": k}
    if not length_divisor_ok or O != I or O >= 96:
        return best
    cost = {1: k * ((I + 63) // 64)}
    cand = None
    for f in (2, 4):
        if O * f > 128:
            continue
        smin = min((p + (t - half) * dilation) // f for p in range(f) for t in range(k))
        smax = max((p + (t - half) * dilation) // f for p in range(f) for t in range(k))
        n_super = smax - smin + 1
        if n_super > 16:
            continue
        c = n_super * ((f * I + 63) // 64) / f          # K blocks per original frame-tile
        if c <= min(cost.values()):
            cost[f] = c
            cand = (f, smin, smax)
    if cand is None:
        return best
    f, smin, smax = cand
    wp = torch.zeros(smax - smin + 1, f * O, f * I, device=w.device)
    for p_out in range(f):
        for t in range(k):
            m = p_out + (t - half) * dilation
            sp, p_in = m // f, m % f
            wp[sp - smin, p_out * O:(p_out + 1) * O, p_in * I:(p_in + 1) * I] += w[:, :, t]
    return {"f": f, "shifts": list(range(smin, smax + 1)), "w": wp.to(od).contiguous(),
            "b": b.repeat(f).contiguous(), "k": k}


def polyphase_plan(wt, u):
    """ConvTranspose1d(I -> O, kernel k, stride u, padding (k - u) // 2) as a few-tap GEMM with N = u * O:
        out[u*q + r] = sum_delta x[q + delta] . w[:, :, r + pad - u*delta].
    wt: (I, O, k) folded weight.  Returns (deltas, poly (n_delta, u*O, I) fp32)."""
    I, O, k = wt.shape
    pad = (k - u) // 2
    dmin = min(-((-(r + pad - k + 1)) // u) for r in range(u))   # ceil((r+pad-k+1)/u)
    dmax = max((r + pad) // u for r in range(u))
    deltas = list(range(dmin, dmax + 1))
    poly = torch.zeros(len(deltas), u * O, I, device=wt.device)
    for di, dl in enumerate(deltas):
        for r in range(u):
            kk = r + pad - u * dl
            if 0 <= kk < k:
                poly[di, r * O:(r + 1) * O, :] = wt[:, :, kk].t()
    return deltas, poly


class BigVGAN(nn.Module):
    def __init__(self, h, use_cuda_kernel: bool = False, mode: str = "bf16"):
        super().__init__()
        self.h = h if isinstance(h, AttrDict) else to_attr(dict(h))
        self.h["use_cuda_kernel"] = use_cuda_kernel      # accepted; the sm_100a kernels always run
        h = self.h
        if h.resblock != "1":
            raise NotImplementedError("only AMPBlock1 (resblock '1'); AMPBlock2.forward returns None "
                                      "in the reference (SURVEY App. D-7)")
        if h.activation not in ("snake", "snakebeta"):
            raise NotImplementedError("activation incorrectly specified")
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        filt = kaiser_sinc_filter12()
        c0 = h.upsample_initial_channel
        self.conv_pre = _Params(weight=(c0, h.num_mels, 7), bias=(c0,))
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        ch = c0
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            self.ups.append(nn.ModuleList([_Params(weight=(c0 // 2 ** i, c0 // 2 ** (i + 1), k),
                                                   bias=(c0 // 2 ** (i + 1),))]))
            ch = c0 // 2 ** (i + 1)
            for ks, dil in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(_AMPBlock(ch, ks, len(dil), filt))
        self.activation_post = _Act(ch, filt)
        self.use_bias_at_final = h.get("use_bias_at_final", True)
        self.conv_post = _Params(weight=(1, ch, 7), **({"bias": (1,)} if self.use_bias_at_final else {}))
        self.use_tanh_at_final = h.get("use_tanh_at_final", True)
        synth.fill_parameters_(self, seed=0)
        self.mode = mode
        self._w = None
        self._w_key = None
        self._register_load_state_dict_pre_hook(self._fold_weight_norm_hook)

    # -- reference API ---------------------------------------------------------------------
    def remove_weight_norm(self):
        """Weights are always held folded; checkpoints with weight_g / weight_v are folded on load."""
        return None

    @staticmethod
    def _fold_weight_norm_hook(state_dict, prefix, *args):
        for k in [k for k in state_dict if k.endswith("weight_g")]:
            base = k[: -len("weight_g")]
            g, v = state_dict.pop(k), state_dict.pop(base + "weight_v")
            n = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
            state_dict[base + "weight"] = g * v / n        # norm over all dims but 0 (App. a16)

    @classmethod
    def from_pretrained(cls, model_id, use_cuda_kernel=False, mode="bf16", **ignored):
        """Local-directory loader (config.json + bigvgan_generator.pt with key 'generator'),
        the layout the reference's save_pretrained writes (bigvgan.py:403-411).  Hub download is
        out of scope (no network); extra HF kwargs are accepted and ignored (App. D-9)."""
        if not os.path.isdir(model_id):
            raise FileNotFoundError(f"{model_id}: from_pretrained needs a local directory")
        with open(os.path.join(model_id, "config.json")) as f:
            h = to_attr(json.load(f))
        model = cls(h, use_cuda_kernel=use_cuda_kernel, mode=mode)
        ck = torch.load(os.path.join(model_id, "bigvgan_generator.pt"), map_location="cpu")
        model.load_state_dict(ck["generator"])
        return model

    def set_mode(self, mode):
        if mode != self.mode:
            self.mode, self._w = mode, None

    # -- kernel-side weights ----------------------------------------------------------------
    def _prepare(self):
        dev = self.conv_pre.weight.device
        if dev.type != "cuda":
            raise RuntimeError("seedvc_b200.BigVGAN runs on CUDA only: call .to('cuda') first "
                               "(there is no CPU fallback)")
        key = (str(dev), self.mode, tuple(p._version for p in self.parameters()),
               tuple(p.data_ptr() for p in self.parameters()))
        if self._w is None or key != self._w_key:
            # every GEMM operand of the vocoder is an un-normalised stream value (there is no norm layer in BigVGAN),
            # so "bf16" mode runs it on IEEE-half operands like the other stream-carrying operands of that mode
            # (Ops.stream_dtype): waveform rel-L2 1.2e-3 instead of 9.7e-3 of a 1e-2 budget for +1.6 % vocoder time
            self._w, self._w_key = self._build_weights(Ops("fp16" if self.mode == "bf16" else self.mode)), key
        return self._w

    def _build_weights(self, ops):
        """Fold / permute / cast the parameters into what the kernels read."""
        od = ops.op_dtype
        h = self.h
        logscale = bool(h.get("snake_logscale", False))
        is_beta = h.activation == "snakebeta"

        def conv_w(p):                       # (O, I, k) -> (k, O, I) operand dtype
            return p.weight.detach().float().permute(2, 0, 1).contiguous().to(od)

        def f32(t):
            return t.detach().float().contiguous()

        def snake(a):
            alpha = a.act.alpha.detach().float()
            beta = a.act.beta.detach().float() if is_beta else alpha
            if logscale:
                alpha, beta = torch.exp(alpha), torch.exp(beta)
            return alpha.contiguous(), (1.0 / (beta + 1e-9)).contiguous()

        w = {"ops": ops, "pre_w": conv_w(self.conv_pre), "pre_b": f32(self.conv_pre.bias)}
        stages = []
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            wt = self.ups[i][0].weight.detach().float()          # (I, O, k)
            O = wt.shape[1]
            deltas, poly = polyphase_plan(wt, u)
            st = {"u": u, "O": O, "deltas": deltas, "up_w": poly.to(od).contiguous(),
                  "up_b": f32(self.ups[i][0].bias).repeat(u).contiguous(), "blocks": []}
            for j in range(self.num_kernels):
                rb = self.resblocks[i * self.num_kernels + j]
                ks = h.resblock_kernel_sizes[j]
                pairs = []
                for l, d in enumerate(h.resblock_dilation_sizes[j]):
                    pairs.append({
                        "a1": snake(rb.activations[2 * l]), "a2": snake(rb.activations[2 * l + 1]),
                        "c1": conv_plan(rb.convs1[l].weight, rb.convs1[l].bias, d, od),
                        "c2": conv_plan(rb.convs2[l].weight, rb.convs2[l].bias, 1, od),
                    })
                st["blocks"].append(pairs)
            stages.append(st)
        w["stages"] = stages
        w["post_a"] = snake(self.activation_post)
        w["post_w"] = f32(self.conv_post.weight[0].t())           # (k, C)
        w["post_b"] = f32(self.conv_post.bias) if self.use_bias_at_final else None
        w["c"] = self._c_weights(w, ops) if hasattr(ops, "bigvgan_forward") else None
        return w

    def _c_weights(self, w, ops):
        """svc_bigvgan_weights over the prepared tensors (kept alive by ``w``)."""
        from . import _lib
        c = _lib.BigVGANWeights()
        h = self.h
        c.n_mels, c.c0, c.n_stages = h.num_mels, h.upsample_initial_channel, len(w["stages"])
        c.n_kernels, c.n_dil = self.num_kernels, len(h.resblock_dilation_sizes[0])
        if any(len(d) != c.n_dil for d in h.resblock_dilation_sizes) or c.n_kernels > 3 or c.n_dil > 3:
            return None                                        # shapes outside the C struct: Python sequence
        c.op_dtype, c.precise = ops.op_code, ops.precise
        c.pre_w, c.pre_b = w["pre_w"].data_ptr(), w["pre_b"].data_ptr()

        def plan(dst, p):
            dst.f, dst.n_taps, dst.k = p["f"], len(p["shifts"]), p["k"]
            for i, sh in enumerate(p["shifts"]):
                dst.shifts[i] = sh
            dst.w, dst.b = p["w"].data_ptr(), p["b"].data_ptr()

        for si, st in enumerate(w["stages"]):
            cs = c.stages[si]
            cs.u, cs.O, cs.n_delta = st["u"], st["O"], len(st["deltas"])
            for i, d in enumerate(st["deltas"]):
                cs.deltas[i] = d
            cs.up_w, cs.up_b = st["up_w"].data_ptr(), st["up_b"].data_ptr()
            for j, pairs in enumerate(st["blocks"]):
                for l, pr in enumerate(pairs):
                    cp = cs.pairs[j][l]
                    cp.a1, cp.inv_b1 = pr["a1"][0].data_ptr(), pr["a1"][1].data_ptr()
                    cp.a2, cp.inv_b2 = pr["a2"][0].data_ptr(), pr["a2"][1].data_ptr()
                    plan(cp.c1, pr["c1"])
                    plan(cp.c2, pr["c2"])
        c.post_a, c.post_inv_b = w["post_a"][0].data_ptr(), w["post_a"][1].data_ptr()
        c.post_w, c.post_k = w["post_w"].data_ptr(), w["post_w"].shape[0]
        c.post_b = w["post_b"].data_ptr() if w["post_b"] is not None else None
        c.use_tanh = int(bool(self.use_tanh_at_final))
        return c

    def launches_per_call(self):
        n_st = len(self.h.upsample_rates)
        split_post = self.mode != "fp32" and (self.h.upsample_initial_channel // (2 ** n_st)) % 8 == 0
        return 2 + n_st * (1 + self.num_kernels * len(self.h.resblock_dilation_sizes[0]) * 4) + 1 + int(split_post)

    @torch.no_grad()
    def forward(self, x):
        w = self._prepare()
        ops: Ops = w["ops"]
        od = ops.op_dtype
        dev = x.device
        B, n_mels, Tm = x.shape
        f32 = torch.float32
        if w.get("c") is not None and ops.profile is None:
            # ONE call of the graph-level C entry point (csrc/graph.cu) issues the launch sequence written out below
            mel = x.float().contiguous()
            up = 1
            for u in self.h.upsample_rates:
                up *= u
            out = torch.empty(B, Tm * up, dtype=f32, device=dev)
            ops.bigvgan_forward(w["c"], mel, out, B, Tm, self.launches_per_call())
            return out.view(B, 1, Tm * up)
        mel_op = ops.empty(B, Tm, n_mels, device=dev)
        ops.bct_to_btc(x.float().contiguous(), mel_op)
        c0 = self.h.upsample_initial_channel
        cur_op = ops.empty(B, Tm, c0, device=dev)
        ops.gemm([(mel_op, j - 3, w["pre_w"][j]) for j in range(7)], c0, B=B, T=Tm, bias=w["pre_b"],
                 out_op=cur_op)
        L = Tm
        cur = None
        nk = self.num_kernels
        for si, st in enumerate(w["stages"]):
            u, O = st["u"], st["O"]
            xs = torch.empty(B, L * u, O, dtype=f32, device=dev)
            ops.gemm([(cur_op, dl, st["up_w"][di]) for di, dl in enumerate(st["deltas"])], u * O,
                     B=B, T=L, bias=st["up_b"], out_f32=xs.view(B, L, u * O))
            L = L * u
            act = torch.empty(B, L, O, dtype=od, device=dev)
            # conv1's output feeds only the pair's second Snake, whose tensor-core FIRs read IEEE half: in the
            # 16-bit modes the conv epilogue writes it once, as half (no fp32 round trip)
            xt = torch.empty(B, L, O, dtype=f32 if od == f32 else torch.float16, device=dev)
            y = torch.empty(B, L, O, dtype=f32, device=dev)
            nxt = torch.empty(B, L, O, dtype=f32, device=dev)
            last_stage = si == len(w["stages"]) - 1
            nxt_op = None if last_stage else torch.empty(B, L, O, dtype=od, device=dev)
            for j, pairs in enumerate(st["blocks"]):
                src = xs
                for l, pr in enumerate(pairs):
                    c1, c2 = pr["c1"], pr["c2"]

                    def segs_of(plan, a):
                        f = plan["f"]
                        av = a.view(B, L // f, O * f)
                        return [(av, sh, plan["w"][i]) for i, sh in enumerate(plan["shifts"])], f

                    ops.snake(src, act, *pr["a1"])
                    sg, f = segs_of(c1, act)
                    fl = 2.0 * B * L * O * O            # algorithmic FLOPs per tap of the original conv
                    xo = {"out_f32" if od == f32 else "out_op": xt.view(B, L // f, O * f)}
                    ops.gemm(sg, O * f, B=B, T=L // f, bias=c1["b"], algo_flops=fl * c1["k"], **xo)
                    ops.snake(xt, act, *pr["a2"])
                    sg, f = segs_of(c2, act)
                    vw = (B, L // f, O * f)
                    if l < len(pairs) - 1:
                        ops.gemm(sg, O * f, B=B, T=L // f, bias=c2["b"], res=src.view(vw), out_f32=y.view(vw),
                                 algo_flops=fl * c2["k"])
                        src = y
                    else:   # last pair: residual, then (r0 + r1 + r2) / 3 accumulated in place
                        ops.gemm(sg, O * f, B=B, T=L // f, bias=c2["b"], res=src.view(vw), alpha=1.0 / nk,
                                 accumulate=j > 0, out_f32=nxt.view(vw), algo_flops=fl * c2["k"],
                                 out_op=nxt_op.view(vw) if (j == nk - 1 and nxt_op is not None) else None)
            cur, cur_op = nxt, nxt_op
        out = torch.empty(B, L, dtype=f32, device=dev)
        if od != f32 and cur.shape[2] % 8 == 0:
            # activation_post on the tensor cores (16-bit out), conv_post + clamp behind it
            ops.snake(cur, act, *w["post_a"])
            ops.conv_post(act, w["post_w"], w["post_b"], out, self.use_tanh_at_final)
        else:
            ops.snake_conv_post(cur, *w["post_a"], w["post_w"], w["post_b"], out, self.use_tanh_at_final)
        return out.view(B, 1, L)
