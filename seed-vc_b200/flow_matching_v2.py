"""v2 sampler + estimator (reference: modules/v2/cfm.py, modules/v2/dit_wrapper.py,
modules/v2/dit_model.py), same kernels as v1.

Differences from v1 handled here / in ``DiTEngine``: adaLN-Zero style 6-way modulation with
gates (dit_model.py:20-36,130-143), final ``RMSNorm*(1+scale)+shift`` (:38-54), style and time
as prepended tokens, bf16-rounded RoPE table, cosine time grid
(cfm.py:47-48) and the five CFG branch layouts (cfm.py:77-125).
"""
from __future__ import annotations

import torch
from torch import nn

from . import synth
from .dit_engine import DiTEngine, DiTSpec, rope_table
from .flow_matching import _Params, _TimestepEmbedder, _find_multiple, _linear, _wn_linear
from .ops import Ops


class _AdaLN6(nn.Module):
    def __init__(self, d, n):
        super().__init__()
        self.linear = _linear(n * d, d)
        self.norm = _Params(weight=(d,))


class _BlockV2(nn.Module):
    def __init__(self, d, inter):
        super().__init__()
        self.attention = nn.Module()
        self.attention.wqkv = _linear(3 * d, d, bias=False)
        self.attention.wo = _linear(d, d, bias=False)
        self.feed_forward = nn.Module()
        self.feed_forward.w1 = _linear(inter, d, bias=False)
        self.feed_forward.w3 = _linear(inter, d, bias=False)
        self.feed_forward.w2 = _linear(d, inter, bias=False)
        self.ffn_norm = _Params(weight=(d,))
        self.attention_norm = _AdaLN6(d, 6)


class _TransformerV2(nn.Module):
    def __init__(self, d, inter, depth, block_size):
        super().__init__()
        self.layers = nn.ModuleList([_BlockV2(d, inter) for _ in range(depth)])
        self.norm = _AdaLN6(d, 2)
        # persistent buffer in the reference state_dict (bf16); the 64 MB causal_mask is not kept
        self.register_buffer("freqs_cis", rope_table(block_size, bf16_round=True).to(torch.bfloat16))


class _TimestepEmbedderV2(nn.Module):
    def __init__(self, d, freq_dim=256):
        super().__init__()
        self.mlp = nn.ModuleList([_linear(d, freq_dim), nn.Identity(), _linear(d, d)])


class DiT(nn.Module):
    """Reference: modules/v2/dit_wrapper.py:58-152."""

    def __init__(self, time_as_token, style_as_token, uvit_skip_connection, block_size, depth,
                 num_heads, hidden_dim, in_channels, content_dim, style_encoder_dim,
                 class_dropout_prob, dropout_rate, attn_dropout_rate, mode: str = "bf16"):
        super().__init__()
        self.time_as_token, self.style_as_token = bool(time_as_token), bool(style_as_token)
        self.uvit_skip_connection = bool(uvit_skip_connection)   # the v2 Transformer never uses it
        if hidden_dim // num_heads != 64 or hidden_dim % num_heads:
            raise NotImplementedError("seedvc_b200 attention kernel supports head_dim 64 only")
        D, C = hidden_dim, in_channels
        inter = _find_multiple(int(2 * 4 * D / 3), 256)
        self.transformer = _TransformerV2(D, inter, depth, block_size)
        self.in_channels = self.out_channels = C
        self.num_heads = num_heads
        self.x_embedder = _wn_linear(D, C)
        self.content_dim = content_dim
        self.cond_projection = _linear(D, content_dim)
        self.t_embedder = _TimestepEmbedderV2(D)
        self.final_mlp = nn.ModuleList([_linear(D, D), nn.Identity(), _linear(C, D)])
        self.class_dropout_prob = class_dropout_prob
        self.cond_x_merge_linear = _linear(D, D + 2 * C)
        self.style_in = _linear(D, style_encoder_dim)
        synth.fill_parameters_(self, seed=0, prefix="estimator.")
        self.spec = DiTSpec(version=2, D=D, H=num_heads, L=depth, C=C, content_dim=content_dim,
                            style_dim=style_encoder_dim, time_as_token=self.time_as_token,
                            style_as_token=self.style_as_token, style_in_merge=False, uvit=False,
                            long_skip=False, head="mlp", prefix="")
        self.mode = mode
        self._engine = None
        self._engine_key = None

    def setup_caches(self, max_batch_size, max_seq_length):
        pass    # the v2 reference builds its caches in the constructor

    def set_mode(self, mode):
        if mode != self.mode:
            self.mode, self._engine = mode, None

    def engine(self) -> DiTEngine:
        dev = self.cond_projection.weight.device
        if dev.type != "cuda":
            raise RuntimeError("seedvc_b200.DiTv2 runs on CUDA only (no CPU fallback)")
        key = (str(dev), self.mode, tuple(p._version for p in self.parameters()),
               tuple(p.data_ptr() for p in self.parameters()))
        if self._engine is None or key != self._engine_key:
            eng = DiTEngine(self.spec, Ops(self.mode))
            eng.load_weights(self.state_dict(), dev)
            self._engine, self._engine_key = eng, key
        return self._engine

    @torch.no_grad()
    def forward(self, x, prompt_x, x_lens, t, style, cond):
        eng = self.engine()
        ops = eng.ops
        N, C, T = x.shape
        dev = x.device
        if not bool((t == t[0]).all()):
            raise NotImplementedError("rows of one estimator call must share the timestep")
        prompt_op = ops.empty(N, T, C, device=dev, dtype=ops.stream_dtype)
        ops.bct_to_btc(prompt_x.float().contiguous(), prompt_op)
        x_op = ops.empty(N, T, C, device=dev, dtype=ops.stream_dtype)
        ops.bct_to_btc(x.float().contiguous(), x_op)
        eng.begin([(True, True, True)], prompt_op, cond.float(), style.float(), x_lens.to(dev),
                  t[:1].detach().float().cpu())
        v = eng.step(0, x_op)
        out = torch.empty(N, C, T, dtype=torch.float32, device=dev)
        ops.btc_to_bct(v, out)
        return out


class CFM(nn.Module):
    """Reference: modules/v2/cfm.py:4-132."""

    def __init__(self, estimator: nn.Module):
        super().__init__()
        self.sigma_min = 1e-6
        self.estimator = estimator
        self.in_channels = estimator.in_channels

    def set_mode(self, mode):
        self.estimator.set_mode(mode)

    @torch.inference_mode()
    def inference(self, mu, x_lens, prompt, style, n_timesteps=10, temperature=1.0,
                  inference_cfg_rate=[0.5, 0.5], random_voice=False):
        B, T = mu.size(0), mu.size(1)
        z = torch.randn([B, self.in_channels, T], device=mu.device) * temperature
        t_span = torch.linspace(0, 1, n_timesteps + 1, device=mu.device)
        t_span = t_span + (-1) * (torch.cos(torch.pi / 2 * t_span) - 1 + t_span)
        return self.solve_euler(z, x_lens, prompt, mu, style, t_span, inference_cfg_rate, random_voice)

    @torch.no_grad()
    def solve_euler(self, x, x_lens, prompt, mu, style, t_span, inference_cfg_rate=[0.5, 0.5],
                    random_voice=False, *, step_hook=None):
        eng = self.estimator.engine()
        ops = eng.ops
        dev = x.device
        B, C, T = x.shape
        Tp = min(int(prompt.size(-1)), T)
        w0, w1 = float(inference_cfg_rate[0]), float(inference_cfg_rate[1])
        # time grid as the reference walks it (cfm.py:69,126-129), fp32 on the host
        ts = t_span.detach().float().cpu()
        t, dt = ts[0].clone(), ts[1] - ts[0]
        t_vals, dts = [], []
        for step in range(1, len(ts)):
            t_vals.append(t.clone())
            dts.append(float(dt))
            t = t + dt
            if step < len(ts) - 1:
                dt = ts[step + 1] - t
        full, txt, null = (True, True, True), (False, False, True), (False, False, False)
        if random_voice:
            branches, coefs = [txt, null], [1.0 + w0, -w0]
        elif w0 == 0 and w1 == 0:
            branches, coefs = [full], [1.0]
        elif w0 == 0:
            branches, coefs = [full, txt], [1.0 + w1, -w1]
        elif w1 == 0:
            branches, coefs = [full, null], [1.0 + w0, -w0]
        else:
            branches, coefs = [full, txt, null], [1.0 + w0 + w1, -w1, -w0]
        xs = torch.empty(B, T, C, dtype=torch.float32, device=dev)
        ops.bct_to_btc(x.float().contiguous(), xs, zero_from=0, zero_to=Tp)
        x_op = ops.empty(B, T, C, device=dev, dtype=ops.stream_dtype)
        ops.bct_to_btc(x.float().contiguous(), x_op, zero_from=0, zero_to=Tp)
        prompt_op = ops.zeros(B, T, C, device=dev, dtype=ops.stream_dtype)
        if Tp > 0:
            ops.bct_to_btc(prompt[..., :Tp].float().contiguous(), prompt_op[:, :Tp, :])
        st = eng.begin(branches, prompt_op, mu.float(), style.float(), x_lens.to(dev),
                       torch.stack(t_vals))
        for s in range(len(dts)):
            v = eng.step(s, x_op)
            if step_hook is not None:      # parity tests: the CFG-combined velocity of this step, (B, T, C)
                step_hook(s, sum(c * v[k * B:(k + 1) * B] for k, c in enumerate(coefs)))
            ops.cfg_euler(xs, v, coefs, dts[s], Tp, st["x_lens"], x_op)
        out = torch.empty(B, C, T, dtype=torch.float32, device=dev)
        ops.btc_to_bct(xs, out)
        return out


CFMv2, DiTv2 = CFM, DiT
