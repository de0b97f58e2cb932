"""InterpolateRegulator - the producer of ``mu`` for the sampler, on the same kernels.

Mirror of the reference class (modules/length_regulator.py:28-141): same constructor arguments, the
same parameter names and shapes (``content_in_proj``, ``model.{0,3,6,9}`` Conv1d k=3, ``model.{1,4,7,10}``
GroupNorm(1), ``model.12`` 1x1 conv, ``embedding``, ``mask_token``, ``f0_embedding``, ``f0_mask``), the same
``forward(x, ylens, n_quantizers, f0) -> (out * mask, olens, None, None, None)``.

Compute: ``content_in_proj`` = segmented GEMM; ``F.interpolate(mode='nearest')`` = a row gather
(``svc_interp_rows``, index table built with ATen's float32 formula); every Conv1d(k=3) = a 3-segment GEMM
whose zero padding is the TMA out-of-bounds fill; GroupNorm(1)+Mish = ``svc_groupnorm1_mish``; the final
1x1 conv = GEMM; ``* mask`` = ``svc_mask_rows``.  The continuous (``is_discrete=False``), non-VQ
configuration every released v1 preset uses is implemented; the discrete / VQ branches raise.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .ops import Ops

f0_max, f0_min = 1100.0, 50.0
f0_mel_min = 1127 * np.log(1 + f0_min / 700)
f0_mel_max = 1127 * np.log(1 + f0_max / 700)


def f0_to_coarse(f0, f0_bin):
    """modules/length_regulator.py:15-26 (index preparation, tiny)."""
    f0_mel = 1127 * (1 + f0 / 700).log()
    a = (f0_bin - 2) / (f0_mel_max - f0_mel_min)
    b = f0_mel_min * a - 1.
    f0_mel = torch.where(f0_mel > 0, f0_mel * a - b, f0_mel)
    f0_coarse = torch.round(f0_mel).long()
    f0_coarse = f0_coarse * (f0_coarse > 0)
    f0_coarse = f0_coarse + ((f0_coarse < 1) * 1)
    f0_coarse = f0_coarse * (f0_coarse < f0_bin)
    f0_coarse = f0_coarse + ((f0_coarse >= f0_bin) * (f0_bin - 1))
    return f0_coarse


def nearest_index(n_in: int, n_out: int) -> np.ndarray:
    """Source index of every output position of F.interpolate(mode='nearest') (ATen ``nearest_idx``:
    identity / halving fast paths, else ``min(floorf(dst * (float)n_in / n_out), n_in - 1)`` in fp32)."""
    dst = np.arange(n_out)
    if n_out == n_in:
        return dst.astype(np.int32)
    if n_out == 2 * n_in:
        return (dst >> 1).astype(np.int32)
    scale = np.float32(n_in) / np.float32(n_out)
    return np.minimum(np.floor(dst.astype(np.float32) * scale).astype(np.int64), n_in - 1).astype(np.int32)


class InterpolateRegulator(nn.Module):
    def __init__(self, channels, sampling_ratios, is_discrete=False, in_channels=None, vector_quantize=False,
                 codebook_size=1024, out_channels=None, groups=1, n_codebooks=1, quantizer_dropout=0.0,
                 f0_condition=False, n_f0_bins=512, mode: str = "bf16", v2: bool = False):
        super().__init__()
        if vector_quantize or n_codebooks != 1:
            raise NotImplementedError("vector-quantised / multi-codebook InterpolateRegulator is not built")
        if groups != 1:
            raise NotImplementedError("GroupNorm with one group only (the reference default)")
        self.sampling_ratios = sampling_ratios
        out_channels = out_channels or channels
        model = nn.ModuleList([])
        self.interpolate = len(sampling_ratios) > 0
        for _ in sampling_ratios:                     # parameter containers with the reference's names
            model.extend([nn.Conv1d(channels, channels, 3, 1, 1), nn.GroupNorm(groups, channels), nn.Mish()])
        # v2 (modules/v2/length_regulator.py:53-55) drops the 1x1 conv when the widths agree
        model.append(nn.Identity() if (v2 and channels == out_channels) else nn.Conv1d(channels, out_channels, 1, 1))
        self.model = nn.Sequential(*model)
        self.v2 = v2
        self.embedding = nn.Embedding(codebook_size, channels)
        self.is_discrete = is_discrete
        self.mask_token = nn.Parameter(torch.zeros(1, channels))
        self.n_codebooks = n_codebooks
        self.quantizer_dropout = quantizer_dropout
        self.f0_condition = bool(f0_condition)
        if f0_condition:
            self.f0_embedding = nn.Embedding(n_f0_bins, channels)
            self.n_f0_bins = n_f0_bins
            self.f0_mask = nn.Parameter(torch.zeros(1, channels))
        if not is_discrete:
            self.content_in_proj = nn.Linear(in_channels, channels)
        self.channels, self.out_channels = channels, out_channels
        self.mode = mode
        self._prep = None

    def set_mode(self, mode):
        self.mode, self._prep = mode, None

    def _prepare(self):
        dev = self.mask_token.device
        # like DiT.engine() / BigVGAN._prepare(): any parameter update (in-place write, a parent module's
        # load_state_dict, .to()) changes _version or data_ptr and rebuilds the kernel-side weights
        key = (self.mode, str(dev), tuple(p._version for p in self.parameters()),
               tuple(p.data_ptr() for p in self.parameters()))
        if self._prep is not None and self._prep["key"] == key:
            return self._prep
        ops = Ops(self.mode)
        od = ops.op_dtype
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        w = {"key": key, "ops": ops, "blocks": []}
        if self.is_discrete:
            w["emb"] = f(self.embedding.weight)
            w["zero_row"] = torch.zeros(1, 1, self.channels, dtype=torch.float32, device=dev)
        else:
            w["proj_w"], w["proj_b"] = f(self.content_in_proj.weight).to(od), f(self.content_in_proj.bias)
        mods = list(self.model)
        for i in range(0, len(mods) - 1, 3):
            conv, gn = mods[i], mods[i + 1]
            w["blocks"].append(dict(w=f(conv.weight).permute(2, 0, 1).contiguous().to(od), b=f(conv.bias),
                                    g=f(gn.weight), beta=f(gn.bias), eps=gn.eps))
        last = mods[-1]
        if isinstance(last, nn.Conv1d):
            w["out_w"] = f(last.weight)[:, :, 0].contiguous().to(od)
            w["out_b"] = f(last.bias)
        if self.f0_condition:
            w["f0_emb"] = f(self.f0_embedding.weight)
            w["f0_mask"] = f(self.f0_mask).view(-1)
        self._prep = w
        return w

    def load_state_dict(self, *a, **k):
        self._prep = None
        return super().load_state_dict(*a, **k)

    @torch.no_grad()
    def forward(self, x, ylens=None, n_quantizers=None, f0=None):
        if x.device.type != "cuda":
            raise RuntimeError("seedvc_b200 InterpolateRegulator runs on a CUDA (sm_100a) device only")
        w = self._prepare()
        ops, od = w["ops"], w["ops"].op_dtype
        dev = x.device
        D = self.channels
        if self.is_discrete:                         # token ids: (B, T) or (B, n_codebooks, T) -> first book
            tok = (x if x.dim() == 2 else x[:, 0]).to(torch.int32).contiguous()
            B, Tin = tok.shape
        else:
            B, Tin, Cin = x.shape
            xin = x.to(torch.float32).contiguous()
            x_op = xin if od == torch.float32 else ops.empty(B, Tin, Cin, device=dev)
            if od != torch.float32:
                ops.cast(xin, x_op)
            h0 = torch.empty(B, Tin, D, dtype=torch.float32, device=dev)
            ops.gemm([(x_op, 0, w["proj_w"])], D, B=B, T=Tin, bias=w["proj_b"], out_f32=h0)
        if self.interpolate:
            Tout = int(ylens.max())
            idx = torch.from_numpy(nearest_index(Tin, Tout)).to(dev)
            olens = ylens
        else:                                        # length_regulator.py:116-119
            Tout = Tin
            idx = torch.arange(Tin, dtype=torch.int32, device=dev)
            olens = ylens if self.v2 else ylens.clamp(max=Tin).long()
        bare = not w["blocks"] and "out_w" not in w       # no conv stack at all (v2 ar_length_regulator)
        cur = torch.empty(B, Tout, D, dtype=torch.float32, device=dev) if bare else ops.empty(B, Tout, D, device=dev)
        if self.is_discrete:
            if self.f0_condition:
                raise NotImplementedError("discrete tokens together with F0 conditioning (no preset uses it)")
            # embedding lookup + nearest interpolation in one gather: zero source row + embedding add
            ops.interp_rows(w["zero_row"].expand(B, 1, D), torch.zeros(Tout, dtype=torch.int32, device=dev), cur,
                            emb=w["emb"], emb_q=tok, emb_idx=idx)
        else:
            kw = {}
            if self.f0_condition:
                if f0 is None:
                    kw["add_vec"] = w["f0_mask"]
                else:
                    q = f0_to_coarse(f0.to(dev), self.n_f0_bins).clamp(0, self.n_f0_bins - 1).to(torch.int32).contiguous()
                    kw.update(emb=w["f0_emb"], emb_q=q,
                              emb_idx=torch.from_numpy(nearest_index(q.shape[1], Tout)).to(dev))
            ops.interp_rows(h0, idx, cur, **kw)
        y = torch.empty(B, Tout, D, dtype=torch.float32, device=dev)
        y_pre, last_blk = None, None
        for blk in w["blocks"]:
            ops.gemm([(cur, s - 1, blk["w"][s]) for s in range(3)], D, B=B, T=Tout, bias=blk["b"], out_f32=y)
            ops.groupnorm1_mish(y, blk["g"], blk["beta"], cur, eps=blk["eps"])
            y_pre, last_blk = y, blk
        if "out_w" in w:
            out = torch.empty(B, Tout, self.out_channels, dtype=torch.float32, device=dev)
            ops.gemm([(cur, 0, w["out_w"])], self.out_channels, B=B, T=Tout, bias=w["out_b"], out_f32=out)
        elif bare:
            out = cur
        else:                                        # Identity tail: last block's activation, in fp32
            out = y
            ops.groupnorm1_mish(y_pre, last_blk["g"], last_blk["beta"], out, eps=last_blk["eps"])
        if self.interpolate or not self.v2:          # v2 without interpolation applies no mask (:90-93,:108)
            ops.mask_rows(out, olens.to(device=dev, dtype=torch.int32).contiguous())
        if self.v2:
            return out, olens                        # modules/v2/length_regulator.py:110
        return out, olens, None, None, None
