"""Deterministic synthetic weights and inputs (no checkpoints or datasets offline).

``fill_parameters_`` writes every parameter of a module from a generator seeded by
the parameter's *name*, so the reference module (in ``oracle/gen_golden.py``) and
this package's module get bit-identical weights as long as their parameter names
and shapes agree - which is itself the drop-in contract
(reference loader: modules/commons.py:412-455).

The fill is chosen to keep activations O(1) through the whole path and to make
every per-channel parameter non-trivial (RMSNorm weights, Snake alpha/beta and
biases are all-ones / all-zeros after the reference constructors).
"""
from __future__ import annotations

import math
import zlib

import torch


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _randn(shape, g):
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32)


def synth_tensor(name: str, shape, seed: int = 0) -> torch.Tensor:
    """Value of the parameter called ``name`` with ``shape``."""
    g = _gen(name, seed)
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    if name == "conv_post.weight_g":            # HiFT: keeps exp(magnitude rows) O(0.2) so the +-0.99 clamp is rare
        return 0.35 + 0.3 * torch.rand(shape, generator=g)
    if leaf == "weight_g":                      # weight-norm gain == row norm of w
        return 0.7 + 0.6 * torch.rand(shape, generator=g)
    if leaf == "alpha" and (".activations1." in name or ".activations2." in name):
        return (1.0 + 0.25 * _randn(shape, g)).clamp_min(0.3)     # HiFT Snake: linear-scale alpha around 1
    if leaf in ("alpha", "beta"):               # Snake, log scale
        return 0.3 * _randn(shape, g)
    if leaf == "bias":
        b = 0.1 * _randn(shape, g)
        if name.endswith("project_layer.bias"):  # v1 AdaLN: first half multiplies the norm
            b[: shape[0] // 2] += 1.0
        if name == "conv_post.bias" and shape[0] == 18:   # HiFT: log-magnitude rows centred at -1.5
            b[:9] -= 1.5
        return b
    if len(shape) == 1:                          # RMSNorm weights
        return 1.0 + 0.1 * _randn(shape, g)
    if len(shape) == 2:                          # Linear (O, I) / Embedding
        return _randn(shape, g) / math.sqrt(shape[1])
    if len(shape) == 3:
        if ".ups." in name or name.startswith("ups."):   # ConvTranspose1d (I, O, k), stride k/2
            return _randn(shape, g) / math.sqrt(shape[0] * 2.0)
        fan_in = shape[1] * shape[2]                      # Conv1d (O, I, k)
        scale = 1.0
        if "resblocks" in name:
            scale = 0.6
        elif "conv_post" in name:
            scale = 0.08
        return scale * _randn(shape, g) / math.sqrt(fan_in)
    return _randn(shape, g)


@torch.no_grad()
def fill_parameters_(module: torch.nn.Module, seed: int = 0, prefix: str = "") -> None:
    for name, p in module.named_parameters():
        p.copy_(synth_tensor(prefix + name, p.shape, seed).to(p.dtype))


def synth_utterance(utt_id: int, T: int, Tp: int, n_mels: int, content_dim: int,
                    style_dim: int = 192):
    """Synthetic conditioning for one utterance.

    Returns ``mu (T, content_dim)``, ``prompt (n_mels, Tp)`` log-mel-like,
    ``style (style_dim,)`` and the injected noise ``z (n_mels, T)``.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(1000 + utt_id)
    mu = torch.randn(T, content_dim, generator=g)
    prompt = (-4.0 + 2.0 * torch.randn(n_mels, Tp, generator=g)).clamp_(-11.5, 2.5)
    style = torch.randn(style_dim, generator=g)
    z = torch.randn(n_mels, T, generator=g)
    return mu, prompt, style, z


def synth_batch(B: int, T: int, Tp: int, n_mels: int, content_dim: int,
                style_dim: int = 192, first_id: int = 0):
    parts = [synth_utterance(first_id + i, T, Tp, n_mels, content_dim, style_dim)
             for i in range(B)]
    mu = torch.stack([p[0] for p in parts])
    prompt = torch.stack([p[1] for p in parts])
    style = torch.stack([p[2] for p in parts])
    z = torch.stack([p[3] for p in parts])
    return mu, prompt, style, z


def synth_f0(B: int, Tm: int, seed: int = 3) -> torch.Tensor:
    """Frame-rate F0 track in Hz with voiced runs (80-400 Hz, slowly varying) and unvoiced gaps (0)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    base = 80.0 + 320.0 * torch.rand(B, 1, generator=g)
    wob = torch.cumsum(4.0 * torch.randn(B, Tm, generator=g), dim=1)
    f0 = (base + wob).clamp_(60.0, 500.0)
    seg = (torch.arange(Tm)[None, :] // 9 + torch.randint(0, 4, (B, 1), generator=g)) % 4
    return torch.where(seg == 3, torch.zeros_like(f0), f0)


def synth_hift_noise(B: int, H: int, L: int, seed: int = 5):
    """The two random draws of HiFT's SineGen, made reproducible: phase (B, H, 1) in [-pi, pi), noise (B, H, L)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    phase = (torch.rand(B, H, 1, generator=g) * 2 - 1) * math.pi
    noise = torch.randn(B, H, L, generator=g)
    return phase, noise


def synth_mel(B: int, n_mels: int, Tm: int, seed: int = 7) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return (-4.0 + 2.0 * torch.randn(B, n_mels, Tm, generator=g)).clamp_(-11.5, 2.5)


_BUFFER_LEAVES = ("freqs", "input_pos", "freqs_cis", "causal_mask", "filter")


def synth_state_dict(key_shapes: dict, seed: int = 0) -> dict:
    """State dict for a ``{name: shape}`` manifest; buffers are skipped (they are
    deterministic functions of the config, not weights)."""
    return {k: synth_tensor(k, s, seed) for k, s in key_shapes.items()
            if k.split(".")[-1] not in _BUFFER_LEAVES}
