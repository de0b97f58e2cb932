"""v1 conditional-flow-matching sampler + DiT estimator, API-compatible with the reference.

Drop-in surface:
  ``CFM(args)``                         modules/flow_matching.py:159-167
  ``CFM.inference(mu, x_lens, prompt, style, f0, n_timesteps, temperature, inference_cfg_rate)``
                                        modules/flow_matching.py:30-53
  ``CFM.solve_euler(x, x_lens, prompt, mu, style, f0, t_span, inference_cfg_rate)``   :55-112
  ``CFM.estimator.setup_caches(max_batch_size, max_seq_length)``   diffusion_transformer.py:484
  ``DiT.forward(x, prompt_x, x_lens, t, style, cond, mask_content)``                  :486-537
The modules hold parameters under the reference's names and shapes so
``load_state_dict`` of a reference checkpoint works; the compute happens in
``libseedvc_b200.so`` through ``DiTEngine``.

Unlike the reference (batch 1 only under CFG) a batch of utterances is
supported; the result for each utterance equals the reference's batch-1 result with
``x_lens[b]`` as its length.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import synth
from .dit_engine import DiTEngine, DiTSpec
from .ops import Ops


def _find_multiple(n, k):
    return n if n % k == 0 else n + k - (n % k)


class _Params(nn.Module):
    """A bag of parameters with given names/shapes (no compute of its own)."""

    def __init__(self, **shapes):
        super().__init__()
        for k, s in shapes.items():
            self.register_parameter(k, nn.Parameter(torch.empty(*s), requires_grad=False))


def _linear(o, i, bias=True):
    return _Params(weight=(o, i), **({"bias": (o,)} if bias else {}))


def _wn_linear(o, i):
    return _Params(bias=(o,), weight_g=(o, 1), weight_v=(o, i))


class _AdaLN(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.project_layer = _linear(2 * d, d)
        self.norm = _Params(weight=(d,))


class _Block(nn.Module):
    def __init__(self, d, inter, uvit):
        super().__init__()
        self.attention = nn.Module()
        self.attention.wqkv = _linear(3 * d, d, bias=False)
        self.attention.wo = _linear(d, d, bias=False)
        self.feed_forward = nn.Module()
        self.feed_forward.w1 = _linear(inter, d, bias=False)
        self.feed_forward.w3 = _linear(inter, d, bias=False)
        self.feed_forward.w2 = _linear(d, inter, bias=False)
        self.ffn_norm = _AdaLN(d)
        self.attention_norm = _AdaLN(d)
        if uvit:
            self.skip_in_linear = _linear(d, 2 * d)


class _TimestepEmbedder(nn.Module):
    def __init__(self, d, freq_dim=256):
        super().__init__()
        self.mlp = nn.ModuleList([_linear(d, freq_dim), nn.Identity(), _linear(d, d)])
        half = freq_dim // 2
        self.register_buffer("freqs", torch.exp(
            -math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half))


class _SConv(nn.Module):
    """Parameter holder mirroring SConv1d -> NormConv1d -> weight-normed Conv1d."""

    def __init__(self, i, o, k):
        super().__init__()
        self.conv = nn.Module()
        self.conv.conv = _Params(bias=(o,), weight_g=(o, 1, 1), weight_v=(o, i, k))


class _WN(nn.Module):
    def __init__(self, hidden, k, n_layers):
        super().__init__()
        self.in_layers = nn.ModuleList([_SConv(hidden, 2 * hidden, k) for _ in range(n_layers)])
        self.res_skip_layers = nn.ModuleList(
            [_SConv(hidden, 2 * hidden if i < n_layers - 1 else hidden, 1) for i in range(n_layers)])
        self.cond_layer = _SConv(hidden, 2 * hidden * n_layers, 1)


class _Transformer(nn.Module):
    def __init__(self, d, inter, depth, uvit):
        super().__init__()
        self.layers = nn.ModuleList([_Block(d, inter, uvit) for _ in range(depth)])
        self.norm = _AdaLN(d)


def _get(o, name, default):
    return getattr(o, name) if hasattr(o, name) else default


class DiT(nn.Module):
    """Velocity estimator (reference: modules/diffusion_transformer.py:407-537)."""

    def __init__(self, args, mode: str = "bf16"):
        super().__init__()
        d = args.DiT
        self.time_as_token = bool(_get(d, "time_as_token", False))
        self.style_as_token = bool(_get(d, "style_as_token", False))
        self.uvit_skip_connection = bool(_get(d, "uvit_skip_connection", False))
        D, C = d.hidden_dim, d.in_channels
        self.in_channels = self.out_channels = C
        self.num_heads = d.num_heads
        if D // d.num_heads != 64 or D % d.num_heads:
            raise NotImplementedError("seedvc_b200 attention kernel supports head_dim 64 only")
        if _get(d, "is_causal", False):
            raise NotImplementedError("is_causal DiT is not used by any released config")
        inter = _find_multiple(int(2 * 4 * D / 3), 256)
        self.transformer = _Transformer(D, inter, d.depth, self.uvit_skip_connection)
        self.x_embedder = _wn_linear(D, C)                        # unused by forward (App. D-8)
        self.content_type = d.content_type
        self.content_dim = d.content_dim
        self.cond_embedder = _Params(weight=(d.content_codebook_size, D))   # unused
        self.cond_projection = _linear(D, d.content_dim)
        self.is_causal = False
        self.t_embedder = _TimestepEmbedder(D)
        self.register_buffer("input_pos", torch.arange(16384))
        self.final_layer_type = d.final_layer_type
        style_dim = args.style_encoder.dim
        self.transformer_style_condition = bool(d.style_condition)
        if self.final_layer_type == "wavenet":
            wn = args.wavenet
            Dw = wn.hidden_dim
            self.t_embedder2 = _TimestepEmbedder(Dw)
            self.conv1 = _linear(Dw, D)
            self.conv2 = _Params(weight=(C, Dw, 1), bias=(C,))
            if wn.dilation_rate != 1:
                raise NotImplementedError("WaveNet head: only dilation_rate 1 (all released configs)")
            self.wavenet = _WN(Dw, wn.kernel_size, wn.num_layers)
            self.final_layer = nn.Module()
            self.final_layer.linear = _wn_linear(Dw, Dw)
            self.final_layer.adaLN_modulation = nn.ModuleList([nn.Identity(), _linear(2 * Dw, Dw)])
            self.res_projection = _linear(Dw, D)
        else:
            self.final_mlp = nn.ModuleList([_linear(D, D), nn.Identity(), _linear(C, D)])
        self.class_dropout_prob = d.class_dropout_prob
        self.content_mask_embedder = _Params(weight=(1, D))                 # unused
        self.long_skip_connection = bool(d.long_skip_connection)
        self.skip_linear = _linear(D, D + C)
        k_merge = D + 2 * C + style_dim * int(self.transformer_style_condition) * int(not self.style_as_token)
        self.cond_x_merge_linear = _linear(D, k_merge)
        if self.style_as_token:
            self.style_in = _linear(D, style_dim)
        synth.fill_parameters_(self, seed=0, prefix="estimator.")
        self.spec = DiTSpec(
            version=1, D=D, H=d.num_heads, L=d.depth, C=C, content_dim=d.content_dim,
            style_dim=style_dim, time_as_token=self.time_as_token,
            style_as_token=self.style_as_token,
            style_in_merge=self.transformer_style_condition and not self.style_as_token,
            uvit=self.uvit_skip_connection,
            long_skip=self.long_skip_connection,
            head=self.final_layer_type,
            Dw=args.wavenet.hidden_dim if self.final_layer_type == "wavenet" else 0,
            wn_layers=args.wavenet.num_layers if self.final_layer_type == "wavenet" else 0,
            wn_kernel=args.wavenet.kernel_size if self.final_layer_type == "wavenet" else 5,
            prefix="")
        self.mode = mode
        self._engine = None
        self._engine_key = None
        self.max_seq_length = -1

    # -- reference API ---------------------------------------------------------------------
    def setup_caches(self, max_batch_size, max_seq_length):
        """Reference: diffusion_transformer.py:484 -> Transformer.setup_caches (:90-110).
        Only the RoPE table is needed here (the 64 MB causal mask is never read)."""
        self.max_seq_length = _find_multiple(max_seq_length, 8)
        self.max_batch_size = max_batch_size
        if self._engine is not None:
            self._engine.setup_rope(self.max_seq_length, self._device())

    def _device(self):
        return self.cond_projection.weight.device

    def set_mode(self, mode):
        if mode != self.mode:
            self.mode = mode
            self._engine = None

    def engine(self) -> DiTEngine:
        """Kernel-side weights; rebuilt when parameters change (load_state_dict / .to())."""
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("seedvc_b200.DiT runs on CUDA only: call .to('cuda') first "
                               "(there is no CPU fallback)")
        key = (str(dev), self.mode, getattr(self, "fold_norms", False), tuple(p._version for p in self.parameters()),
               tuple(p.data_ptr() for p in self.parameters()))
        if self._engine is None or key != self._engine_key:
            eng = DiTEngine(self.spec, Ops(self.mode, fold_norms=getattr(self, "fold_norms", False)))
            eng.load_weights(self.state_dict(), dev)
            if self.max_seq_length > 0:
                eng.setup_rope(self.max_seq_length, dev)
            self._engine, self._engine_key = eng, key
        return self._engine

    @torch.no_grad()
    def forward(self, x, prompt_x, x_lens, t, style, cond, mask_content=False):
        """One estimator call with explicit (already stacked) inputs; all rows share t[0].

        x, prompt_x: (N, C, T); x_lens: (N,) or (1,); t: (N,); style: (N, 192); cond: (N, T, cd).
        Returns (N, C, T) fp32."""
        if isinstance(mask_content, torch.Tensor) and bool(mask_content.any()):
            raise NotImplementedError("mask_content (class dropout) is a training-time path")
        eng = self.engine()
        ops = eng.ops
        N, C, T = x.shape
        dev = x.device
        if not bool((t == t[0]).all()):
            raise NotImplementedError("rows of one estimator call must share the timestep")
        x_lens = x_lens.to(dev)
        if x_lens.numel() == 1 and N > 1:
            x_lens = x_lens.expand(N)
        prompt_op = ops.empty(N, T, C, device=dev, dtype=ops.stream_dtype)
        ops.bct_to_btc(prompt_x.float().contiguous(), prompt_op)
        x_op = ops.empty(N, T, C, device=dev, dtype=ops.stream_dtype)
        ops.bct_to_btc(x.float().contiguous(), x_op)
        eng.begin([(True, True, True)], prompt_op, cond.float(), style.float(), x_lens,
                  t[:1].detach().float().cpu())
        v = eng.step(0, x_op)
        out = torch.empty(N, C, T, dtype=torch.float32, device=dev)
        ops.btc_to_bct(v, out)
        return out


class BASECFM(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.sigma_min = 1e-6
        self.estimator = None
        self.in_channels = args.DiT.in_channels
        self.zero_prompt_speech_token = bool(_get(args.DiT, "zero_prompt_speech_token", False))

    @torch.inference_mode()
    def inference(self, mu, x_lens, prompt, style, f0, n_timesteps, temperature=1.0,
                  inference_cfg_rate=0.5):
        """Reference: modules/flow_matching.py:30-53."""
        B, T = mu.size(0), mu.size(1)
        z = torch.randn([B, self.in_channels, T], device=mu.device) * temperature
        t_span = torch.linspace(0, 1, n_timesteps + 1, device=mu.device)
        return self.solve_euler(z, x_lens, prompt, mu, style, f0, t_span, inference_cfg_rate)

    @torch.no_grad()
    def solve_euler(self, x, x_lens, prompt, mu, style, f0, t_span, inference_cfg_rate=0.5, *,
                    t_values_dev=None, step_hook=None):
        """Fixed-step Euler with batched CFG.  Reference: modules/flow_matching.py:55-112.

        ``t_values_dev`` (keyword only, used by graphs.GraphedConversion): the per-step times already on
        the device, so that no host-to-device copy happens inside a CUDA-graph capture.
        ``step_hook(s, v)`` (keyword only, parity tests): called after every Euler step with the CFG-combined
        velocity ``v`` (B, T, C) fp32 that step integrated (costs an extra torch op per step; product callers
        leave it None).

        x: (B, C, T) noise; prompt: (B, C, Tp); mu: (B, T, content_dim); style: (B, 192);
        ``f0`` is accepted and ignored like in the reference (App. D-3)."""
        eng = self.estimator.engine()
        ops = eng.ops
        dev = x.device
        B, C, T = x.shape
        Tp = min(int(prompt.size(-1)), T)
        x_lens = x_lens.to(dev)
        if self.zero_prompt_speech_token:
            mu = mu.clone()           # the reference writes into the caller's tensor
            mu[..., :int(prompt.size(-1))] = 0   # last dim, literally as flow_matching.py:80-81
        # time grid exactly as the reference accumulates it (fp32, on the host)
        ts = t_span.detach().float().cpu()
        t = ts[0].clone()
        t_vals, dts = [], []
        for step in range(1, len(ts)):
            dt = ts[step] - ts[step - 1]
            t_vals.append(t.clone())
            dts.append(float(dt))
            t = t + dt
        w = float(inference_cfg_rate)
        if w > 0:
            branches = [(True, True, True), (False, False, False)]
            coefs = [1.0 + w, -w]
        else:
            branches, coefs = [(True, True, True)], [1.0]
        # frames-major state: x fp32 master + operand copy; prompt region zeroed
        xs = torch.empty(B, T, C, dtype=torch.float32, device=dev)
        ops.bct_to_btc(x.float().contiguous(), xs, zero_from=0, zero_to=Tp)
        x_op = ops.empty(B, T, C, device=dev, dtype=ops.stream_dtype)
        ops.bct_to_btc(x.float().contiguous(), x_op, zero_from=0, zero_to=Tp)
        prompt_op = ops.zeros(B, T, C, device=dev, dtype=ops.stream_dtype)
        if Tp > 0:
            ops.bct_to_btc(prompt[..., :Tp].float().contiguous(), prompt_op[:, :Tp, :])
        st = eng.begin(branches, prompt_op, mu.float(), style.float(), x_lens,
                       torch.stack(t_vals) if t_values_dev is None else t_values_dev)
        for s in range(len(dts)):
            v = eng.step(s, x_op)
            if step_hook is not None:
                step_hook(s, sum(c * v[k * B:(k + 1) * B] for k, c in enumerate(coefs)))
            ops.cfg_euler(xs, v, coefs, dts[s], Tp, st["x_lens"], x_op)
        out = torch.empty(B, C, T, dtype=torch.float32, device=dev)
        ops.btc_to_bct(xs, out)
        return out


class CFM(BASECFM):
    def __init__(self, args, mode: str = "bf16"):
        super().__init__(args)
        if args.dit_type == "DiT":
            self.estimator = DiT(args, mode=mode)
        else:
            raise NotImplementedError(f"Unknown diffusion type {args.dit_type}")

    def set_mode(self, mode):
        self.estimator.set_mode(mode)
