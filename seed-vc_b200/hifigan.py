"""HiFT vocoder (neural source filter + iSTFTNet) on the sm_100a kernels, API-compatible with the reference.

Drop-in surface: ``HiFTGenerator(**configs/hifigan.yml:hift, f0_predictor=ConvRNNF0Predictor(...))``,
``forward(x (B, 80, Tm), f0=None) -> (B, 256 * Tm)``, ``inference``, ``remove_weight_norm``; parameters under the
reference's names (weight_g / weight_v kept, as in the released ``hift.pt``), so ``load_state_dict`` of that
checkpoint works (reference: modules/hifigan/generator.py:282-454, modules/hifigan/f0_predictor.py:19-55,
inference.py:114-122, real-time-gui.py:214-222).

Execution plan per call (frames-major activations, fp32 residual streams, operand-dtype GEMM inputs):
  F0 predictor      5 x (3-tap segmented GEMM -> svc_unary ELU), Linear 512 -> 1 in fp32, |.|      f0_predictor.py:52-55
  source            svc_hift_source (upsample x256, 9 harmonics, uv / noise mix, Linear + tanh)     generator.py:366-370
  source STFT       svc_hift_stft -> (B, 64 Tm + 8, 24) operand buffer                               :372-378,391-392
  conv_pre          7-tap GEMM, leaky-ReLU by svc_unary                                              :394,396
  per stage i       ups[i] as a 3-tap polyphase GEMM with N = 8 * C_out; source_downs[i] (stride-8 conv = 3 taps
                    over the regrouped (B, rows / 8, 8 * 24) view, or 1x1) + source ResBlock whose last GEMM adds
                    into x; 3 ResBlocks (Snake by svc_unary, k-tap dilated GEMMs, residual and the 1/3 average
                    fused in the epilogues)                                                          :395-421
  head              leaky-ReLU, conv_post (N padded 18 -> 24), svc_hift_istft                        :424-435
The two random draws of ``SineGen`` (uniform phase per harmonic, Gaussian noise per sample, generator.py:222-236)
can be injected with the keyword-only ``phase`` / ``noise`` arguments (parity tests); otherwise they are drawn
with ``torch.rand`` / ``torch.randn`` on the device like the reference does.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import synth
from .bigvgan import conv_plan, polyphase_plan
from .dit_engine import _fold_wn
from .flow_matching import _Params
from .ops import Ops

S_CPAD = 24          # channels of the source-STFT operand buffer (18 used; 48-byte rows keep TMA strides aligned)


def _wn_conv(o, i, k):
    return _Params(bias=(o,), weight_g=(o, 1, 1), weight_v=(o, i, k))


class _ResBlock(nn.Module):
    def __init__(self, ch, k, n_dil):
        super().__init__()
        self.convs1 = nn.ModuleList([_wn_conv(ch, ch, k) for _ in range(n_dil)])
        self.convs2 = nn.ModuleList([_wn_conv(ch, ch, k) for _ in range(n_dil)])
        self.activations1 = nn.ModuleList([_Params(alpha=(ch,)) for _ in range(n_dil)])
        self.activations2 = nn.ModuleList([_Params(alpha=(ch,)) for _ in range(n_dil)])


class ConvRNNF0Predictor(nn.Module):
    """Parameter holder + forward of the reference's F0 predictor (f0_predictor.py:19-55)."""

    def __init__(self, num_class: int = 1, in_channels: int = 80, cond_channels: int = 512):
        super().__init__()
        if num_class != 1:
            raise NotImplementedError("ConvRNNF0Predictor: num_class 1 only (configs/hifigan.yml)")
        mods = []
        for i in range(5):
            mods += [_wn_conv(cond_channels, in_channels if i == 0 else cond_channels, 3), nn.Identity()]
        self.condnet = nn.ModuleList(mods)
        self.classifier = _Params(weight=(1, cond_channels), bias=(1,))
        self.cond_channels = cond_channels


class HiFTGenerator(nn.Module):
    def __init__(self, in_channels=80, base_channels=512, nb_harmonics=8, sampling_rate=22050, nsf_alpha=0.1,
                 nsf_sigma=0.003, nsf_voiced_threshold=10, upsample_rates=(8, 8), upsample_kernel_sizes=(16, 16),
                 istft_params=None, resblock_kernel_sizes=(3, 7, 11),
                 resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), source_resblock_kernel_sizes=(7, 11),
                 source_resblock_dilation_sizes=((1, 3, 5), (1, 3, 5)), lrelu_slope=0.1, audio_limit=0.99,
                 f0_predictor=None, mode: str = "bf16"):
        super().__init__()
        istft_params = dict(istft_params or {"n_fft": 16, "hop_len": 4})
        if istft_params["n_fft"] != 16 or istft_params["hop_len"] != 4:
            raise NotImplementedError("HiFT kernels are written for the released iSTFT head (n_fft 16, hop 4)")
        if len(upsample_rates) != 2 or len(source_resblock_kernel_sizes) != 2:
            raise NotImplementedError("HiFT: two upsampling stages (configs/hifigan.yml)")
        self.out_channels = 1
        self.nb_harmonics, self.sampling_rate = nb_harmonics, sampling_rate
        self.istft_params, self.lrelu_slope, self.audio_limit = istft_params, lrelu_slope, audio_limit
        self.nsf_alpha, self.nsf_sigma, self.nsf_voiced_threshold = nsf_alpha, nsf_sigma, nsf_voiced_threshold
        self.upsample_rates, self.upsample_kernel_sizes = list(upsample_rates), list(upsample_kernel_sizes)
        self.resblock_kernel_sizes = list(resblock_kernel_sizes)
        self.resblock_dilation_sizes = [list(d) for d in resblock_dilation_sizes]
        self.source_resblock_kernel_sizes = list(source_resblock_kernel_sizes)
        self.source_resblock_dilation_sizes = [list(d) for d in source_resblock_dilation_sizes]
        self.num_kernels, self.num_upsamples = len(resblock_kernel_sizes), len(upsample_rates)
        self.scale = int(math.prod(upsample_rates)) * istft_params["hop_len"]
        self.m_source = nn.Module()
        self.m_source.l_linear = _Params(weight=(1, nb_harmonics + 1), bias=(1,))
        self.conv_pre = _wn_conv(base_channels, in_channels, 7)
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(upsample_rates, upsample_kernel_sizes)):
            ci, co = base_channels // 2 ** i, base_channels // 2 ** (i + 1)
            self.ups.append(_Params(bias=(co,), weight_g=(ci, 1, 1), weight_v=(ci, co, k)))
        nf = istft_params["n_fft"] + 2
        down = [1] + list(upsample_rates)[::-1][:-1]
        cum = [math.prod(down[:i + 1]) for i in range(len(down))][::-1]         # [8, 1]
        self.source_rates = cum
        self.source_downs, self.source_resblocks = nn.ModuleList(), nn.ModuleList()
        for i, (u, k, d) in enumerate(zip(cum, source_resblock_kernel_sizes, source_resblock_dilation_sizes)):
            co = base_channels // 2 ** (i + 1)
            self.source_downs.append(_Params(weight=(co, nf, 1 if u == 1 else 2 * u), bias=(co,)))
            self.source_resblocks.append(_ResBlock(co, k, len(d)))
        self.resblocks = nn.ModuleList()
        for i in range(len(upsample_rates)):
            ch = base_channels // 2 ** (i + 1)
            for k, d in zip(resblock_kernel_sizes, resblock_dilation_sizes):
                self.resblocks.append(_ResBlock(ch, k, len(d)))
        self.conv_post = _wn_conv(nf, ch, 7)
        self.f0_predictor = f0_predictor
        synth.fill_parameters_(self, seed=0)
        self.mode = mode
        self._w = self._w_key = None

    # -- reference API ---------------------------------------------------------------------
    def remove_weight_norm(self):
        """Weights are folded when the kernel-side copies are built; nothing to do."""
        return None

    def set_mode(self, mode):
        if mode != self.mode:
            self.mode, self._w = mode, None

    @torch.inference_mode()
    def inference(self, mel, f0=None):
        return self.forward(mel, f0=f0)

    # -- kernel-side weights ----------------------------------------------------------------
    def _prepare(self):
        dev = self.conv_pre.bias.device
        if dev.type != "cuda":
            raise RuntimeError("seedvc_b200.HiFTGenerator runs on CUDA only: call .to('cuda') first "
                               "(there is no CPU fallback)")
        key = (str(dev), self.mode, tuple(p._version for p in self.parameters()),
               tuple(p.data_ptr() for p in self.parameters()))
        if self._w is None or key != self._w_key:
            self._w, self._w_key = self._build_weights(Ops(self.mode)), key
        return self._w

    def _build_weights(self, ops):
        # Every HiFT activation is an un-normalised signal whose absolute error reaches the waveform through
        # exp() / sin() heads, so the whole generator uses the stream dtype (IEEE half in both 16-bit modes; measured
        # waveform rel-L2 vs the reference: 2.9e-3 with half operands, 2.1e-2 with bf16 operands).
        od = sdt = ops.stream_dtype
        sd = {k: v.detach() for k, v in self.state_dict().items()}

        def f32(t):
            return t.detach().float().contiguous()

        def plan(prefix, dil, dtype=od):
            return conv_plan(_fold_wn(sd, prefix), sd[prefix + ".bias"], dil, dtype)

        def resblock(prefix, k, dils):
            return [dict(a1=f32(sd[f"{prefix}.activations1.{i}.alpha"]), a2=f32(sd[f"{prefix}.activations2.{i}.alpha"]),
                         c1=plan(f"{prefix}.convs1.{i}", d), c2=plan(f"{prefix}.convs2.{i}", 1))
                    for i, d in enumerate(dils)]

        w = {"ops": ops, "pre": plan("conv_pre", 1)}
        stages = []
        for i, u in enumerate(self.upsample_rates):
            wt = _fold_wn(sd, f"ups.{i}")                       # (I, O, k), norm per input channel
            O = wt.shape[1]
            deltas, poly = polyphase_plan(wt, u)
            st = dict(u=u, O=O, deltas=deltas, up_w=poly.to(od).contiguous(),
                      up_b=f32(sd[f"ups.{i}.bias"]).repeat(u).contiguous())
            # source_downs[i]: Conv1d(18 -> O, kernel 2r, stride r, padding r/2) over the (rows, 24) STFT buffer viewed
            # as (rows / r, r * 24): out[q] = sum_k W[:, :, k] s[r q + k - r/2]; m = k - r/2 -> super-row floor(m / r)
            r = self.source_rates[i]
            ws = sd[f"source_downs.{i}.weight"].float()         # (O, 18, k)
            kk = ws.shape[2]
            if r == 1:
                sw = torch.zeros(1, O, S_CPAD, device=ws.device)
                sw[0, :, :ws.shape[1]] = ws[:, :, 0]
                shifts = [0]
            else:
                pad = r // 2
                sps = sorted({(k - pad) // r for k in range(kk)})
                sw = torch.zeros(len(sps), O, r * S_CPAD, device=ws.device)
                for k in range(kk):
                    m = k - pad
                    sp, p_in = m // r, m % r
                    sw[sps.index(sp), :, p_in * S_CPAD:p_in * S_CPAD + ws.shape[1]] = ws[:, :, k]
                shifts = sps
            st.update(src_r=r, src_w=sw.to(od).contiguous(), src_shifts=shifts, src_b=f32(sd[f"source_downs.{i}.bias"]),
                      src_block=resblock(f"source_resblocks.{i}", self.source_resblock_kernel_sizes[i],
                                         self.source_resblock_dilation_sizes[i]),
                      blocks=[resblock(f"resblocks.{i * self.num_kernels + j}", self.resblock_kernel_sizes[j],
                                       self.resblock_dilation_sizes[j]) for j in range(self.num_kernels)])
            stages.append(st)
        w["stages"] = stages
        # conv_post: N = 18 padded to 24 output columns (zero rows) so the fp32 rows stay 16-byte aligned
        wp = _fold_wn(sd, "conv_post")
        nf = wp.shape[0]
        wpad = torch.zeros(S_CPAD, wp.shape[1], wp.shape[2], device=wp.device)
        wpad[:nf] = wp
        bpad = torch.zeros(S_CPAD, device=wp.device)
        bpad[:nf] = sd["conv_post.bias"].float()
        w["post"] = conv_plan(wpad, bpad, 1, od)
        w["lin_w"] = f32(sd["m_source.l_linear.weight"].reshape(-1))
        w["lin_b"] = float(sd["m_source.l_linear.bias"].reshape(-1)[0])
        if self.f0_predictor is not None:
            pf = "f0_predictor."
            # the predictor runs in the stream dtype (fp16 in both 16-bit modes): F0 feeds phase accumulators
            w["f0"] = dict(convs=[plan(f"{pf}condnet.{2 * i}", 1, sdt) for i in range(5)],
                           cls_w=f32(sd[pf + "classifier.weight"]), cls_b=f32(sd[pf + "classifier.bias"]))
        return w

    # -- forward ------------------------------------------------------------------------------
    @staticmethod
    def _conv(ops, plan, a, B, L, **kw):
        segs = [(a, sh, plan["w"][i]) for i, sh in enumerate(plan["shifts"])]
        ops.gemm(segs, plan["w"].shape[1], B=B, T=L, bias=plan["b"], **kw)

    def _resblock(self, ops, pairs, src, B, L, C, act, xt, tmp, last_kw):
        """ResBlock.forward (:151-158); the last conv writes through ``last_kw`` (epilogue-fused sums)."""
        for l, pr in enumerate(pairs):
            ops.unary(src, act, Ops.UNARY_SNAKE, alpha=pr["a1"])
            self._conv(ops, pr["c1"], act, B, L, out_f32=xt)
            ops.unary(xt, act, Ops.UNARY_SNAKE, alpha=pr["a2"])
            if l < len(pairs) - 1:
                self._conv(ops, pr["c2"], act, B, L, res=src, out_f32=tmp)
                src = tmp
            else:
                self._conv(ops, pr["c2"], act, B, L, res=src, **last_kw)

    def _predict_f0(self, w, ops, mel):
        B, _, Tm = mel.shape
        dev = mel.device
        pw = w["f0"]
        sdt = ops.stream_dtype
        a = torch.empty(B, Tm, mel.shape[1], dtype=sdt, device=dev)
        ops.bct_to_btc(mel, a)
        Cc = self.f0_predictor.cond_channels
        h = torch.empty(B, Tm, Cc, dtype=torch.float32, device=dev)
        for i, plan in enumerate(pw["convs"]):
            self._conv(ops, plan, a, B, Tm, out_f32=h)
            last = i == len(pw["convs"]) - 1
            a = torch.empty(B, Tm, Cc, dtype=torch.float32 if last else sdt, device=dev)
            ops.unary(h, a, Ops.UNARY_ELU)
        f0 = torch.empty(B, Tm, 1, dtype=torch.float32, device=dev)
        ops.gemm([(a, 0, pw["cls_w"])], 1, B=B, T=Tm, bias=pw["cls_b"], out_f32=f0, f32=True)
        out = torch.empty_like(f0)
        ops.unary(f0, out, Ops.UNARY_ABS)
        return out[:, :, 0]

    @torch.no_grad()
    def forward(self, x, f0=None, *, phase=None, noise=None):
        w = self._prepare()
        ops: Ops = w["ops"]
        od, f32 = ops.stream_dtype, torch.float32
        dev = x.device
        B, n_mels, Tm = x.shape
        mel = x.float().contiguous()
        if f0 is None:
            if self.f0_predictor is None:
                raise ValueError("HiFTGenerator: f0 is None and no f0_predictor was given")
            f0 = self._predict_f0(w, ops, mel)
        f0 = f0.float().contiguous()
        H = self.nb_harmonics + 1
        L = Tm * self.scale
        if phase is None:                                      # SineGen's own draws (generator.py:222-224,234-236)
            phase = (torch.rand(B, H, 1, device=dev) * 2 - 1) * math.pi
        if noise is None:
            noise = torch.randn(B, H, L, device=dev)
        s = torch.empty(B, L, dtype=f32, device=dev)
        ops.hift_source(f0, phase.reshape(B, H).float().contiguous(), noise.float().contiguous(), w["lin_w"],
                        w["lin_b"], s, self.scale, self.sampling_rate, self.nsf_alpha, self.nsf_sigma,
                        self.nsf_voiced_threshold)
        TT = L // 4 + 1
        r0 = self.source_rates[0]
        rows = r0 * (TT // r0 + 2)                             # regroupable by r0, >= one zero super-row at the end
        s_buf = torch.empty(B, rows, S_CPAD, dtype=od, device=dev)
        ops.hift_stft(s, s_buf)

        mel_op = torch.empty(B, Tm, n_mels, dtype=od, device=dev)
        ops.bct_to_btc(mel, mel_op)
        c0 = self.conv_pre.bias.numel()
        pre = torch.empty(B, Tm, c0, dtype=f32, device=dev)
        self._conv(ops, w["pre"], mel_op, B, Tm, out_f32=pre)
        cur = pre
        Lc = Tm
        nk = self.num_kernels
        for si, st in enumerate(w["stages"]):
            u, O = st["u"], st["O"]
            cur_op = torch.empty(cur.shape, dtype=od, device=dev)
            ops.unary(cur, cur_op, Ops.UNARY_LRELU, slope=self.lrelu_slope)
            last_stage = si == self.num_upsamples - 1
            Ln = Lc * u + (1 if last_stage else 0)             # ReflectionPad1d((1, 0)) on the last stage (:398-399)
            xs = torch.empty(B, Ln, O, dtype=f32, device=dev)
            body = xs[:, 1:, :] if last_stage else xs
            ops.gemm([(cur_op, dl, st["up_w"][di]) for di, dl in enumerate(st["deltas"])], u * O, B=B, T=Lc,
                     bias=st["up_b"], out_f32=body.as_strided((B, Lc, u * O), (body.stride(0), u * O, 1),
                                                              body.storage_offset()))
            if last_stage:
                ops.set_rows(xs[:, 2, :], xs[:, 0, :])         # reflect: padded[0] = x[1]
            act = torch.empty(B, Ln, O, dtype=od, device=dev)
            xt = torch.empty(B, Ln, O, dtype=f32, device=dev)
            tmp = torch.empty(B, Ln, O, dtype=f32, device=dev)
            # ---- fusion: x = x + source_resblocks[i](source_downs[i](s_stft))  (:402-404)
            r = st["src_r"]
            si_buf = torch.empty(B, Ln, O, dtype=f32, device=dev)
            if r == 1:
                a_view = s_buf[:, :TT, :]
            else:
                a_view = s_buf.view(B, rows // r, r * S_CPAD)
            assert Ln == (TT if r == 1 else (TT + 2 * (r // 2) - 2 * r) // r + 1)
            ops.gemm([(a_view, sh, st["src_w"][i]) for i, sh in enumerate(st["src_shifts"])], O, B=B, T=Ln,
                     bias=st["src_b"], out_f32=si_buf)
            self._resblock(ops, st["src_block"], si_buf, B, Ln, O, act, xt, tmp,
                           dict(accumulate=True, out_f32=xs))
            # ---- (r0 + r1 + r2) / 3  (:406-412)
            nxt = torch.empty(B, Ln, O, dtype=f32, device=dev)
            for j, pairs in enumerate(st["blocks"]):
                self._resblock(ops, pairs, xs, B, Ln, O, act, xt, tmp,
                               dict(alpha=1.0 / nk, accumulate=j > 0, out_f32=nxt))
            cur, Lc = nxt, Ln
        cur_op = torch.empty(cur.shape, dtype=od, device=dev)
        ops.unary(cur, cur_op, Ops.UNARY_LRELU, slope=0.01)    # F.leaky_relu default slope (:424)
        post = torch.empty(B, Lc, S_CPAD, dtype=f32, device=dev)
        self._conv(ops, w["post"], cur_op, B, Lc, out_f32=post)
        wav = torch.empty(B, 4 * (Lc - 1), dtype=f32, device=dev)
        ops.hift_istft(post, wav, clip_mag=1e2, audio_limit=self.audio_limit)
        return wav
