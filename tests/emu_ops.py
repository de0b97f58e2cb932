"""TEST INFRASTRUCTURE ONLY - a torch-on-CPU stand-in for ``seedvc_b200.ops.Ops``.

Implements the documented semantics of every C-ABI entry point (include/seedvc_b200.h) with
plain fp32 torch ops, so the *host logic* of the package (weight folding, hoisting, polyphase
transposed conv, CFG branch layouts, token handling ...) can be checked against the oracle
without a GPU.  The product never imports this; on a GPU the same call sequence goes to
``libseedvc_b200.so``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

ACT_NONE, ACT_SILU, ACT_SWIGLU_PAIR, ACT_TANH_SIG_PAIR, ACT_ROPE = 0, 1, 2, 3, 4


class EmuOps:
    def __init__(self, mode="fp32", fold_norms=False):
        self.mode = "fp32"
        self.fold_norms = fold_norms      # exercise the engine's folded-norm orchestration on the emulated ops
        self.op_dtype = torch.float32
        self.stream_dtype = torch.float32
        self.launches = 0

    def empty(self, *shape, dtype=None, device="cpu"):
        return torch.full(shape, float("nan"), dtype=dtype or torch.float32, device=device)

    def zeros(self, *shape, dtype=None, device="cpu"):
        return torch.zeros(*shape, dtype=dtype or torch.float32, device=device)

    def gemm(self, segs, N, *, B, T, bias=None, rowbias=None, act=ACT_NONE, rope=None, gate=None,
             res=None, alpha=1.0, accumulate=False, out_f32=None, out_op=None, f32=False,
             algo_flops=None, row_ss_out=None, row_scale=None):
        self.launches += 1
        dev = segs[0][0].device
        acc = torch.zeros(B, T, N, device=dev)
        for A, shift, W in segs:
            assert A.shape[0] == B and W.shape == (N, A.shape[2])
            rows = A.shape[1]
            idx = torch.arange(T, device=dev) + shift
            ok = (idx >= 0) & (idx < rows)
            a = torch.zeros(B, T, A.shape[2], device=dev)
            a[:, ok] = A[:, idx[ok]].float()
            acc += a @ W.float().t()
        v = acc
        if row_scale is not None:      # folded RMS norm: accumulator rows scaled by 1 / rms of the producer's rows
            ss, dim, eps = row_scale
            v = v * torch.rsqrt(ss.sum(-1).view(B, T, 1) / dim + eps)
        if bias is not None:
            v = v + bias.view(1, 1, N)
        if rowbias is not None:
            v = v + rowbias.reshape(B, 1, N)
        if act == ACT_SILU:
            v = F.silu(v)
        elif act == ACT_SWIGLU_PAIR:
            v = F.silu(v[..., 0::2]) * v[..., 1::2]
        elif act == ACT_TANH_SIG_PAIR:
            v = torch.tanh(v[..., 0::2]) * torch.sigmoid(v[..., 1::2])
        elif act == ACT_ROPE:
            tab, rope_cols, pos0, q_cols, q_scale = rope
            v = v.clone()
            r = v[..., :rope_cols].reshape(B, T, rope_cols // 64, 32, 2)
            c = tab[pos0:pos0 + T, :, 0].view(1, T, 1, 32)
            s = tab[pos0:pos0 + T, :, 1].view(1, T, 1, 32)
            rot = torch.stack([r[..., 0] * c - r[..., 1] * s, r[..., 1] * c + r[..., 0] * s], -1)
            v[..., :rope_cols] = rot.reshape(B, T, rope_cols)
            v[..., :q_cols] *= q_scale
        if gate is not None:
            v = v * gate.reshape(B, 1, -1)
        if res is not None:
            v = v + res
        v = v * alpha
        if accumulate:
            v = v + out_f32
        if out_f32 is not None:
            out_f32.copy_(v)
        if out_op is not None:
            out_op.copy_(v)
        if row_ss_out is not None:     # all of the row's sum of squares in slot 0 (consumers add the slots)
            row_ss_out.zero_()
            row_ss_out[:, 0] = (v * v).sum(-1).reshape(B * T)

    def attention(self, qkv, out, H, kv_len):
        self.launches += 1
        B, T, W = qkv.shape
        D = W // 3
        q, k, v = qkv.float().split([D, D, D], dim=-1)
        q = q.view(B, T, H, 64).transpose(1, 2)
        k = k.view(B, T, H, 64).transpose(1, 2)
        v = v.view(B, T, H, 64).transpose(1, 2)
        s = q @ k.transpose(-1, -2)                 # scale already folded into q
        ok = torch.arange(T, device=qkv.device)[None, :] < kv_len[:, None]
        s = s.masked_fill(~ok[:, None, None, :], float("-inf"))
        y = torch.softmax(s, -1) @ v
        out.copy_(y.transpose(1, 2).reshape(B, T, D))

    def norm_mod(self, x, out, *, gamma=None, mul=None, add=None, eps=1e-5, mode=0, raw_out=None):
        self.launches += 1
        if raw_out is not None:
            raw_out.copy_(x)
        if mode == 0:
            y = x * torch.rsqrt(torch.mean(x * x, -1, keepdim=True) + eps)
        else:
            y = F.layer_norm(x, (x.shape[-1],), eps=eps)
        for g in (gamma, mul):
            if g is not None:
                y = y * g
        if add is not None:
            y = y + add
        out.copy_(y)

    @staticmethod
    def _snake(x, a, inv_b):
        """x: (B, L, C) -> anti-aliased snake with prepared a / inv_b."""
        h = torch.tensor([0.0020289648, 0.0093894657, -0.0255434588, -0.0576573834, 0.1285725832,
                          0.4432097971, 0.4432097971, 0.1285725832, -0.0576573834, -0.0255434588,
                          0.0093894657, 0.0020289648], device=x.device)
        xc = x.float().transpose(1, 2)
        C = xc.shape[1]
        xp = F.pad(xc, (5, 5), mode="replicate")
        u = 2.0 * F.conv_transpose1d(xp, h.view(1, 1, 12).expand(C, -1, -1), stride=2, groups=C)
        u = u[..., 15:-15]
        u = u + inv_b.view(1, C, 1) * torch.sin(u * a.view(1, C, 1)) ** 2
        up = F.pad(u, (5, 6), mode="replicate")
        y = F.conv1d(up, h.view(1, 1, 12).expand(C, -1, -1), stride=2, groups=C)
        return y.transpose(1, 2)

    def snake(self, x, out, a, inv_b):
        self.launches += 1
        out.copy_(self._snake(x, a, inv_b))

    def conv_post(self, act, w, bias, out, use_tanh):
        self.launches += 1
        y = act.float().transpose(1, 2)                               # (B, C, L)
        k = w.shape[0]
        o = F.conv1d(y, w.t().unsqueeze(0), bias, padding=k // 2)[:, 0]
        out.copy_(torch.tanh(o) if use_tanh else o.clamp(-1.0, 1.0))

    def snake_conv_post(self, x, a, inv_b, w, bias, out, use_tanh):
        self.launches += 1
        y = self._snake(x, a, inv_b).transpose(1, 2)                  # (B, C, L)
        k = w.shape[0]
        o = F.conv1d(y, w.t().unsqueeze(0), bias, padding=k // 2)[:, 0]
        out.copy_(torch.tanh(o) if use_tanh else o.clamp(-1, 1))

    def cfg_euler(self, x, v, coefs, dt, prompt_len, x_lens=None, x_op=None):
        self.launches += 1
        B, T, C = x.shape
        d = sum(c * v[i * B:(i + 1) * B] for i, c in enumerate(coefs))
        x += dt * d
        x[:, :prompt_len] = 0
        if x_lens is not None:
            for b in range(B):
                x[b, int(x_lens[b]):] = 0
        if x_op is not None:
            x_op.copy_(x)

    def bct_to_btc(self, inp, out, zero_from=0, zero_to=0):
        self.launches += 1
        v = inp.transpose(1, 2).clone()
        v[:, zero_from:zero_to] = 0
        out.copy_(v)

    def btc_to_bct(self, inp, out):
        self.launches += 1
        out.copy_(inp.transpose(1, 2))

    def cast(self, inp, out):
        self.launches += 1
        out.copy_(inp.view(out.shape))

    def scale_cols(self, W, g, mul, out):
        self.launches += 1
        w = W.unsqueeze(0)
        if g is not None:
            w = w * g.view(1, 1, -1)
        if mul is not None:
            w = w * mul.unsqueeze(1)
        out.copy_(w.expand(out.shape))

    def reflect_halo(self, buf, T, pad, lens=None):
        self.launches += 1
        B = buf.shape[0]
        for b in range(B):
            n = T if lens is None else max(2, min(int(lens[b]), T))
            for i in range(pad):
                buf[b, pad - 1 - i] = buf[b, pad + i + 1]
                buf[b, pad + n + i] = buf[b, pad + n - 2 - i]

    def timestep_embedding(self, t, freqs, out):
        self.launches += 1
        a = 1000.0 * t[:, None] * freqs[None]
        out.copy_(torch.cat([torch.cos(a), torch.sin(a)], -1))

    def set_rows(self, src, dst):
        self.launches += 1
        dst.copy_(src.expand_as(dst))

    # ---------------------------------------------------------------- HiFT pieces (svc_unary, svc_hift_*)
    def unary(self, x, out, kind, slope=0.0, alpha=None):
        self.launches += 1
        if kind == 0:
            y = F.leaky_relu(x, slope)
        elif kind == 1:
            y = F.elu(x)
        elif kind == 2:
            a = alpha.view(1, 1, -1)
            y = x + (1.0 / (a + 1e-9)) * torch.sin(x * a) ** 2
        else:
            y = x.abs()
        out.copy_(y)

    def hift_source(self, f0, phase, noise, lin_w, lin_b, out, scale, sr, sine_amp, noise_std, voiced_thr):
        self.launches += 1
        B, Tm = f0.shape
        H = phase.shape[1]
        f0u = f0.repeat_interleave(scale, dim=1)[:, None, :]
        F_mat = torch.cat([f0u * (i + 1) / sr for i in range(H)], dim=1)
        cum = torch.cumsum(F_mat.double(), dim=-1).float()        # ATen's CPU cumsum accumulates fp32 in double
        theta = 2 * 3.141592653589793 * (cum % 1)
        ph = phase.clone().view(B, H, 1)
        ph[:, 0] = 0
        sine = sine_amp * torch.sin(theta + ph)
        uv = (f0u > voiced_thr).float()
        amp = uv * noise_std + (1 - uv) * sine_amp / 3
        sine = sine * uv + (amp * noise if noise is not None else 0)
        out.copy_(torch.tanh((sine * lin_w.view(1, H, 1)).sum(1) + lin_b))

    def hift_stft(self, s, out):
        self.launches += 1
        spec = torch.stft(s, 16, 4, 16, window=torch.hann_window(16, periodic=True, device=s.device), return_complex=True)
        TT = spec.shape[-1]
        out.zero_()
        out[:, :TT, 0:9] = spec.real.transpose(1, 2)
        out[:, :TT, 9:18] = spec.imag.transpose(1, 2)

    def hift_istft(self, x, wav, clip_mag=1e2, audio_limit=0.99):
        self.launches += 1
        mag = torch.clip(torch.exp(x[..., 0:9]), max=clip_mag).transpose(1, 2)
        ph = torch.sin(x[..., 9:18]).transpose(1, 2)
        y = torch.istft(torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)), 16, 4, 16,
                        window=torch.hann_window(16, periodic=True, device=x.device))
        wav.copy_(y.clamp(-audio_limit, audio_limit))
