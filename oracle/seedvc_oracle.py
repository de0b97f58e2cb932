"""TEST INFRASTRUCTURE ONLY - CPU restatement of the Seed-VC conversion hot path.

This file is the *oracle*: a plain fp32 PyTorch-on-CPU restatement of the
reference algorithm, written from the reference's maths (each function cites the
reference file:line it follows).  It must never be imported by the product
package ``seed-vc_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, as the
checker or as the timed CPU baseline.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the authoring container by ``oracle/gen_golden.py``
(which imports /root/reference) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every fixture.

All functions take a flat ``state_dict`` with the reference's parameter names.
Batched semantics: the reference's CFG path only runs at batch 1
(SURVEY.md App. D-1); a batch here is *defined* as the per-utterance batch-1
result, each utterance using its own ``x_lens[b]`` as its whole length.
"""
from __future__ import annotations

import math

import numpy as np

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------
def _wn_weight(sd, prefix):
    """w = g * v / ||v||, norm over all dims but 0 (torch weight_norm dim=0;
    reference: modules/diffusion_transformer.py:395,431, modules/wavenet.py:120-135)."""
    if prefix + ".weight" in sd:
        return sd[prefix + ".weight"]
    v, g = sd[prefix + ".weight_v"], sd[prefix + ".weight_g"]
    n = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return g * v / n


def _linear(sd, prefix, x):
    return F.linear(x, _wn_weight(sd, prefix), sd.get(prefix + ".bias"))


def rmsnorm(x, w, eps=1e-5):
    """modules/diffusion_transformer.py:274-285"""
    return x * torch.rsqrt(torch.mean(x * x, dim=-1, keepdim=True) + eps) * w


def timestep_embedding(t, dim=256, max_period=10000.0, scale=1000.0):
    """modules/diffusion_transformer.py:341-359 (v2: modules/v2/dit_wrapper.py:32-50)"""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = scale * t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def t_embedder(sd, prefix, t):
    """modules/diffusion_transformer.py:361-364"""
    h = _linear(sd, prefix + ".mlp.0", timestep_embedding(t))
    return _linear(sd, prefix + ".mlp.2", F.silu(h))


def rope_table(n_pos, head_dim=64, base=10000.0, bf16_round=False):
    """modules/diffusion_transformer.py:288-297.  v2 stores the table in bf16
    (modules/v2/dit_model.py:100-101,225-234)."""
    freqs = 1.0 / (base ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(n_pos), freqs)
    tab = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1)
    if bf16_round:
        tab = tab.to(torch.bfloat16).float()
    return tab  # (n_pos, hd/2, 2)


def apply_rope(x, tab):
    """Interleaved-pair rotation, modules/diffusion_transformer.py:300-312. x: (B,T,H,hd)."""
    xs = x.float().reshape(*x.shape[:-1], -1, 2)
    c = tab[None, : x.shape[1], None, :, 0]
    s = tab[None, : x.shape[1], None, :, 1]
    out = torch.stack([xs[..., 0] * c - xs[..., 1] * s, xs[..., 1] * c + xs[..., 0] * s], -1)
    return out.flatten(3)


def attention(sd, prefix, x, tab, n_head, kv_len):
    """modules/diffusion_transformer.py:222-260: wqkv, RoPE, masked SDPA, wo."""
    B, T, D = x.shape
    hd = D // n_head
    q, k, v = F.linear(x, sd[prefix + ".wqkv.weight"]).split([D, D, D], dim=-1)
    q = apply_rope(q.view(B, T, n_head, hd), tab).transpose(1, 2)
    k = apply_rope(k.view(B, T, n_head, hd), tab).transpose(1, 2)
    v = v.view(B, T, n_head, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    key_ok = torch.arange(T)[None, :] < kv_len[:, None]            # (B, T)
    s = s.masked_fill(~key_ok[:, None, None, :], float("-inf"))
    y = torch.softmax(s, dim=-1) @ v
    y = y.transpose(1, 2).reshape(B, T, D)
    return F.linear(y, sd[prefix + ".wo.weight"])


def feed_forward(sd, prefix, x):
    """modules/diffusion_transformer.py:263-271"""
    return F.linear(F.silu(F.linear(x, sd[prefix + ".w1.weight"])) *
                    F.linear(x, sd[prefix + ".w3.weight"]), sd[prefix + ".w2.weight"])


def adaln_v1(sd, prefix, x, c):
    """modules/diffusion_transformer.py:30-48: w*RMSNorm(x)+b, no SiLU, no '1+'."""
    n = rmsnorm(x, sd[prefix + ".norm.weight"])
    if c is None:
        return n
    D = x.shape[-1]
    w, b = torch.split(_linear(sd, prefix + ".project_layer", c), D, dim=-1)
    return w * n + b


# ----------------------------------------------------------------------------
# v1 estimator
# ----------------------------------------------------------------------------
def wavenet(sd, prefix, x, g, n_layers, hidden):
    """modules/wavenet.py:138-166 with SConv1d reflect padding
    (modules/encodec.py:212-228; the conv's own padding kwarg is swallowed,
    SURVEY section 8 a8).  x: (B, Dw, T); g: (B, Dw, 1); mask is all ones at
    batch-1 semantics."""
    out = torch.zeros_like(x)
    g = F.conv1d(g, _wn_weight(sd, prefix + ".cond_layer.conv.conv"),
                 sd[prefix + ".cond_layer.conv.conv.bias"])
    for i in range(n_layers):
        w = _wn_weight(sd, f"{prefix}.in_layers.{i}.conv.conv")
        k = w.shape[-1]
        pad_total = k - 1
        pr = pad_total // 2
        pl = pad_total - pr
        x_in = F.conv1d(F.pad(x, (pl, pr), mode="reflect"), w,
                        sd[f"{prefix}.in_layers.{i}.conv.conv.bias"])
        a = x_in + g[:, i * 2 * hidden:(i + 1) * 2 * hidden]
        acts = torch.tanh(a[:, :hidden]) * torch.sigmoid(a[:, hidden:])   # commons.py:131-138
        rs = F.conv1d(acts, _wn_weight(sd, f"{prefix}.res_skip_layers.{i}.conv.conv"),
                      sd[f"{prefix}.res_skip_layers.{i}.conv.conv.bias"])
        if i < n_layers - 1:
            x = x + rs[:, :hidden]
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out


def dit_v1_forward(sd, args, x, prompt_x, x_lens, t, style, cond, pfx="estimator."):
    """modules/diffusion_transformer.py:486-537 (inference branch).

    x, prompt_x: (N, C, T); t: (N,); style: (N, 192); cond: (N, T, content_dim).
    Every row uses its own full length T (batch-1 semantics)."""
    dit = args.DiT
    D, H, L, C = dit.hidden_dim, dit.num_heads, dit.depth, dit.in_channels
    tat = bool(getattr(dit, "time_as_token", False))
    sat = bool(getattr(dit, "style_as_token", False))
    uvit = bool(getattr(dit, "uvit_skip_connection", False))
    N, _, T = x.shape
    t1 = t_embedder(sd, pfx + "t_embedder", t)
    cond = _linear(sd, pfx + "cond_projection", cond)
    xt = x.transpose(1, 2)
    x_in = torch.cat([xt, prompt_x.transpose(1, 2), cond], dim=-1)
    if dit.style_condition and not sat:
        x_in = torch.cat([x_in, style[:, None, :].repeat(1, T, 1)], dim=-1)
    h = _linear(sd, pfx + "cond_x_merge_linear", x_in)
    if sat:
        h = torch.cat([_linear(sd, pfx + "style_in", style).unsqueeze(1), h], dim=1)
    if tat:
        h = torch.cat([t1.unsqueeze(1), h], dim=1)
    ntok = int(tat) + int(sat)
    Tq = T + ntok
    kv_len = x_lens + ntok
    tab = rope_table(Tq)
    c = t1.unsqueeze(1)
    c_layer = None if tat else c                                     # :184
    emit = [i for i in range(L) if i < L // 2] if uvit else []
    recv = [i for i in range(L) if i > L // 2] if uvit else []
    skips = []
    for i in range(L):
        lp = f"{pfx}transformer.layers.{i}"
        if i in recv:
            h = _linear(sd, lp + ".skip_in_linear", torch.cat([h, skips.pop(-1)], dim=-1))
        h = h + attention(sd, lp + ".attention", adaln_v1(sd, lp + ".attention_norm", h, c_layer),
                          tab, H, kv_len)
        h = h + feed_forward(sd, lp + ".feed_forward", adaln_v1(sd, lp + ".ffn_norm", h, c_layer))
        if i in emit:
            skips.append(h)
    h = adaln_v1(sd, pfx + "transformer.norm", h, c)                 # :142 uses c even with tokens
    h = h[:, ntok:]
    if dit.long_skip_connection:
        h = _linear(sd, pfx + "skip_linear", torch.cat([h, xt], dim=-1))
    if dit.final_layer_type == "wavenet":
        Dw = args.wavenet.hidden_dim
        y = _linear(sd, pfx + "conv1", h).transpose(1, 2)
        t2 = t_embedder(sd, pfx + "t_embedder2", t)
        y = wavenet(sd, pfx + "wavenet", y, t2.unsqueeze(2), args.wavenet.num_layers, Dw)
        y = y.transpose(1, 2) + _linear(sd, pfx + "res_projection", h)
        mod = _linear(sd, pfx + "final_layer.adaLN_modulation.1", F.silu(t1))   # :401-405
        shift, scale = mod.chunk(2, dim=1)
        y = F.layer_norm(y, (Dw,), eps=1e-6) * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)
        y = _linear(sd, pfx + "final_layer.linear", y).transpose(1, 2)
        return F.conv1d(y, sd[pfx + "conv2.weight"], sd[pfx + "conv2.bias"])
    y = _linear(sd, pfx + "final_mlp.2", F.silu(_linear(sd, pfx + "final_mlp.0", h)))
    return y.transpose(1, 2)


def solve_euler_v1(sd, args, z, x_lens, prompt, mu, style, t_span, cfg_rate, return_steps=False):
    """modules/flow_matching.py:55-112, applied per utterance (batch-1 semantics).

    z: (B, C, T) injected noise; prompt: (B, C, Tp); mu: (B, T, content_dim);
    style: (B, 192).  Returns (B, C, T) with rows t >= x_lens[b] zeroed."""
    B, C, T = z.shape
    outs, steps = [], []
    for b in range(B):
        Tb = int(x_lens[b])
        x = z[b:b + 1, :, :Tb].clone()
        Tp = min(prompt.shape[-1], Tb)
        prompt_x = torch.zeros_like(x)
        prompt_x[..., :Tp] = prompt[b:b + 1, :, :Tp]
        x[..., :Tp] = 0
        m = mu[b:b + 1, :Tb]
        s = style[b:b + 1]
        xl = torch.tensor([Tb])
        t = t_span[0]
        vs = []
        for step in range(1, len(t_span)):
            dt = t_span[step] - t_span[step - 1]
            if cfg_rate > 0:
                v2 = dit_v1_forward(
                    sd, args, torch.cat([x, x]), torch.cat([prompt_x, torch.zeros_like(prompt_x)]),
                    xl, torch.stack([t, t]), torch.cat([s, torch.zeros_like(s)]),
                    torch.cat([m, torch.zeros_like(m)]))
                v = (1.0 + cfg_rate) * v2[0:1] - cfg_rate * v2[1:2]
            else:
                v = dit_v1_forward(sd, args, x, prompt_x, xl, t.unsqueeze(0), s, m)
            vs.append(v)
            x = x + dt * v
            t = t + dt
            x[:, :, :Tp] = 0
        outs.append(F.pad(x, (0, T - Tb)))
        steps.append(vs)
    out = torch.cat(outs)
    return (out, steps) if return_steps else out


# ----------------------------------------------------------------------------
# v2 estimator and sampler
# ----------------------------------------------------------------------------
def dit_v2_forward(sd, kw, x, prompt_x, x_lens, t, style, cond, pfx=""):
    """modules/v2/dit_wrapper.py:114-152 + modules/v2/dit_model.py:109-143."""
    D, H, L, C = kw["hidden_dim"], kw["num_heads"], kw["depth"], kw["in_channels"]
    tat, sat = bool(kw["time_as_token"]), bool(kw["style_as_token"])
    N, _, T = x.shape
    t1 = t_embedder(sd, pfx + "t_embedder", t)
    cond = _linear(sd, pfx + "cond_projection", cond)
    x_in = torch.cat([x.transpose(1, 2), prompt_x.transpose(1, 2), cond], dim=-1)
    h = _linear(sd, pfx + "cond_x_merge_linear", x_in)
    st = _linear(sd, pfx + "style_in", style)
    if sat:
        h = torch.cat([st.unsqueeze(1), h], dim=1)
    if tat:
        h = torch.cat([t1.unsqueeze(1), h], dim=1)
    ntok = int(tat) + int(sat)
    kv_len = x_lens + ntok
    tab = rope_table(T + ntok, bf16_round=True)
    c = t1.unsqueeze(1)
    for i in range(L):
        lp = f"{pfx}transformer.layers.{i}"
        emb = _linear(sd, lp + ".attention_norm.linear", F.silu(c))
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = torch.chunk(emb, 6, dim=-1)
        n = rmsnorm(h, sd[lp + ".attention_norm.norm.weight"]) * (1 + sc_a) + sh_a
        h = h + g_a * attention(sd, lp + ".attention", n, tab, H, kv_len)
        n = rmsnorm(h, sd[lp + ".ffn_norm.weight"]) * (1 + sc_m) + sh_m
        h = h + g_m * feed_forward(sd, lp + ".feed_forward", n)
    emb = _linear(sd, pfx + "transformer.norm.linear", F.silu(c))
    scale, shift = torch.chunk(emb, 2, dim=-1)                       # dit_model.py:50-53
    h = rmsnorm(h, sd[pfx + "transformer.norm.norm.weight"]) * (1 + scale) + shift
    h = h[:, ntok:]
    y = _linear(sd, pfx + "final_mlp.2", F.silu(_linear(sd, pfx + "final_mlp.0", h)))
    return y.transpose(1, 2)


def v2_t_span(n_timesteps):
    """modules/v2/cfm.py:47-48"""
    t_span = torch.linspace(0, 1, n_timesteps + 1)
    return t_span + (-1) * (torch.cos(torch.pi / 2 * t_span) - 1 + t_span)


def solve_euler_v2(sd, kw, z, x_lens, prompt, mu, style, t_span, cfg_rate=(0.5, 0.5),
                   random_voice=False, pfx="estimator."):
    """modules/v2/cfm.py:50-132, per utterance."""
    B, C, T = z.shape
    w0, w1 = float(cfg_rate[0]), float(cfg_rate[1])
    outs = []
    for b in range(B):
        Tb = int(x_lens[b])
        x = z[b:b + 1, :, :Tb].clone()
        Tp = min(prompt.shape[-1], Tb)
        px = torch.zeros_like(x)
        px[..., :Tp] = prompt[b:b + 1, :, :Tp]
        x[..., :Tp] = 0
        m, s, xl = mu[b:b + 1, :Tb], style[b:b + 1], torch.tensor([Tb])
        zp, zs, zm = torch.zeros_like(px), torch.zeros_like(s), torch.zeros_like(m)
        t = t_span[0]
        dt = t_span[1] - t_span[0]

        def est(xs, ps, ss, ms):
            n = len(xs)
            return dit_v2_forward(sd, kw, torch.cat(xs), torch.cat(ps), xl.repeat(n),
                                  t.repeat(n), torch.cat(ss), torch.cat(ms), pfx=pfx)

        for step in range(1, len(t_span)):
            if random_voice:
                o = est([x, x], [zp, zp], [zs, zs], [m, zm])
                v = (1.0 + w0) * o[0:1] - w0 * o[1:2]
            elif w0 == 0 and w1 == 0:
                v = est([x], [px], [s], [m])
            elif w0 == 0:
                o = est([x, x], [px, zp], [s, zs], [m, m])
                v = (1.0 + w1) * o[0:1] - w1 * o[1:2]
            elif w1 == 0:
                o = est([x, x], [px, zp], [s, zs], [m, zm])
                v = (1.0 + w0) * o[0:1] - w0 * o[1:2]
            else:
                o = est([x, x, x], [px, zp, zp], [s, zs, zs], [m, m, zm])
                v = (1.0 + w0 + w1) * o[0:1] - w0 * o[2:3] - w1 * o[1:2]
            x = x + dt * v
            t = t + dt
            if step < len(t_span) - 1:
                dt = t_span[step + 1] - t
            x[:, :, :Tp] = 0
        outs.append(F.pad(x, (0, T - Tb)))
    return torch.cat(outs)


# ----------------------------------------------------------------------------
# BigVGAN
# ----------------------------------------------------------------------------
def kaiser_sinc_filter12():
    """modules/bigvgan/alias_free_activation/torch/filter.py:30-62 with
    cutoff 0.25, half_width 0.3, kernel_size 12 (resample.py:22-24, :47-52)."""
    ks, cutoff, half_width = 12, 0.25, 0.3
    half = ks // 2
    A = 2.285 * (half - 1) * math.pi * 4 * half_width + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50 else (
        0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A >= 21 else 0.0)
    win = torch.kaiser_window(ks, beta=beta, periodic=False)
    time = torch.arange(-half, half) + 0.5
    f = 2 * cutoff * win * torch.sinc(2 * cutoff * time)
    return f / f.sum()


def snake_aa(x, alpha_log, beta_log, h12=None):
    """Anti-aliased SnakeBeta, closed form of act.py:25-30 + resample.py:29-38 +
    filter.py:94-101 + activations.py:107-119.  x: (B, C, L)."""
    if h12 is None:
        h12 = kaiser_sinc_filter12()
    B, C, L = x.shape
    xp = F.pad(x, (5, 5), mode="replicate")
    u = 2.0 * F.conv_transpose1d(xp, h12.view(1, 1, 12).expand(C, -1, -1), stride=2, groups=C)
    u = u[..., 15:-15]
    a = torch.exp(alpha_log).view(1, C, 1)
    b = torch.exp(beta_log).view(1, C, 1)
    u = u + (1.0 / (b + 1e-9)) * torch.sin(u * a) ** 2
    up = F.pad(u, (5, 6), mode="replicate")
    return F.conv1d(up, h12.view(1, 1, 12).expand(C, -1, -1), stride=2, groups=C)


def bigvgan_forward(sd, h, mel):
    """modules/bigvgan/bigvgan.py:360-386 with AMPBlock1 (:132-141).
    ``sd`` holds folded weights (after remove_weight_norm) or weight_g/weight_v."""
    h12 = kaiser_sinc_filter12()
    nk = len(h.resblock_kernel_sizes)
    x = F.conv1d(mel, _wn_weight(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
        x = F.conv_transpose1d(x, _wn_weight_convT(sd, f"ups.{i}.0"), sd[f"ups.{i}.0.bias"],
                               stride=u, padding=(k - u) // 2)
        xs = None
        for j in range(nk):
            rb = f"resblocks.{i * nk + j}"
            ks = h.resblock_kernel_sizes[j]
            y = x
            for l, d in enumerate(h.resblock_dilation_sizes[j]):
                xt = snake_aa(y, sd[f"{rb}.activations.{2 * l}.act.alpha"],
                              sd[f"{rb}.activations.{2 * l}.act.beta"], h12)
                xt = F.conv1d(xt, _wn_weight(sd, f"{rb}.convs1.{l}"), sd[f"{rb}.convs1.{l}.bias"],
                              dilation=d, padding=d * (ks - 1) // 2)
                xt = snake_aa(xt, sd[f"{rb}.activations.{2 * l + 1}.act.alpha"],
                              sd[f"{rb}.activations.{2 * l + 1}.act.beta"], h12)
                xt = F.conv1d(xt, _wn_weight(sd, f"{rb}.convs2.{l}"), sd[f"{rb}.convs2.{l}.bias"],
                              padding=(ks - 1) // 2)
                y = xt + y
            xs = y if xs is None else xs + y
        x = xs / nk
    x = snake_aa(x, sd["activation_post.act.alpha"], sd["activation_post.act.beta"], h12)
    x = F.conv1d(x, _wn_weight(sd, "conv_post"), sd.get("conv_post.bias"), padding=3)
    if h.get("use_tanh_at_final", True):
        return torch.tanh(x)
    return torch.clamp(x, min=-1.0, max=1.0)


def _wn_weight_convT(sd, prefix):
    """ConvTranspose1d weight (I, O, k): weight_norm dim=0 is per *input* channel
    (SURVEY section 8 a16)."""
    return _wn_weight(sd, prefix)


# ---------------------------------------------------------------------------------------------
# Long-form chunk loop
# ---------------------------------------------------------------------------------------------
def crossfade(chunk1, chunk2, overlap):
    """inference.py:343-350 / seed_vc_wrapper.py:190-199 (numpy; chunk2 is modified in place)."""
    fade_out = np.cos(np.linspace(0, np.pi / 2, overlap)) ** 2
    fade_in = np.cos(np.linspace(np.pi / 2, 0, overlap)) ** 2
    if len(chunk2) < overlap:
        chunk2[:overlap] = chunk2[:overlap] * fade_in[:len(chunk2)] + (chunk1[-overlap:] * fade_out)[:len(chunk2)]
    else:
        chunk2[:overlap] = chunk2[:overlap] * fade_in + chunk1[-overlap:] * fade_out
    return chunk2


def chunk_plan(n_source_frames, n_prompt_frames, max_context_window, overlap_frame_len=16):
    """(start, length, is_last) of every window the reference loop visits (inference.py:470-476 and the
    ``processed_frames += vc_target.size(2) - overlap_frame_len`` updates at :512,:516,:522)."""
    window = max_context_window - n_prompt_frames
    plan, processed = [], 0
    while processed < n_source_frames:
        length = min(window, n_source_frames - processed)
        is_last = processed + window >= n_source_frames
        plan.append((processed, length, is_last))
        if is_last:
            break
        processed += length - overlap_frame_len
    return plan


def stitch_chunks(waves, overlap_wave_len):
    """waves: list of 1-D float32 arrays, one per window in loop order -> concatenated output
    (inference.py:505-527; seed_vc_wrapper.py:227-285 with stream_output=False)."""
    out, previous = [], None
    n = len(waves)
    for k, w in enumerate(waves):
        w = np.array(w, dtype=np.float32, copy=True)
        is_last = k == n - 1
        if k == 0:
            if is_last:
                out.append(w)
                break
            out.append(w[:-overlap_wave_len])
            previous = w[-overlap_wave_len:]
        elif is_last:
            out.append(crossfade(previous, w, overlap_wave_len))
        else:
            out.append(crossfade(previous, w[:-overlap_wave_len], overlap_wave_len))
            previous = w[-overlap_wave_len:]
    return np.concatenate(out)


# ---------------------------------------------------------------------------------------------
# InterpolateRegulator - modules/length_regulator.py:90-141, continuous / non-VQ branch
# ---------------------------------------------------------------------------------------------
def _f0_to_coarse(f0, f0_bin):
    """modules/length_regulator.py:9-26."""
    f0_mel_min = 1127 * np.log(1 + 50.0 / 700)
    f0_mel_max = 1127 * np.log(1 + 1100.0 / 700)
    f0_mel = 1127 * (1 + f0 / 700).log()
    a = (f0_bin - 2) / (f0_mel_max - f0_mel_min)
    b = f0_mel_min * a - 1.
    f0_mel = torch.where(f0_mel > 0, f0_mel * a - b, f0_mel)
    c = torch.round(f0_mel).long()
    c = c * (c > 0)
    c = c + ((c < 1) * 1)
    c = c * (c < f0_bin)
    c = c + ((c >= f0_bin) * (f0_bin - 1))
    return c


def interpolate_regulator(sd, x, ylens, n_blocks=4, f0=None, f0_condition=False, n_f0_bins=512):
    """sd: reference state_dict of InterpolateRegulator; x (B, Tin, in_channels); ylens (B,) long.
    Returns ``out * mask`` (B, max(ylens), out_channels) - length_regulator.py:112-141."""
    x = F.linear(x.float(), sd["content_in_proj.weight"], sd["content_in_proj.bias"])          # :111
    Tout = int(ylens.max())
    mask = (torch.arange(Tout)[None, :] < ylens[:, None]).unsqueeze(-1)                         # :113
    x = F.interpolate(x.transpose(1, 2).contiguous(), size=Tout, mode="nearest")               # :115
    if f0_condition:
        if f0 is None:
            x = x + sd["f0_mask"].unsqueeze(-1)                                                 # :122
        else:
            q = _f0_to_coarse(f0, n_f0_bins).clamp(0, n_f0_bins - 1).long()                     # :125-126
            e = F.embedding(q, sd["f0_embedding.weight"])
            x = x + F.interpolate(e.transpose(1, 2).contiguous(), size=Tout, mode="nearest")    # :127-129
    for i in range(n_blocks):                                                                   # :47-53
        x = F.conv1d(x, sd[f"model.{3 * i}.weight"], sd[f"model.{3 * i}.bias"], padding=1)
        x = F.group_norm(x, 1, sd[f"model.{3 * i + 1}.weight"], sd[f"model.{3 * i + 1}.bias"], 1e-5)
        x = F.mish(x)
    k = 3 * n_blocks
    x = F.conv1d(x, sd[f"model.{k}.weight"], sd[f"model.{k}.bias"])                             # :55-57
    return x.transpose(1, 2).contiguous() * mask                                                # :131,140


# ---------------------------------------------------------------------------------------------
# Mel front-end - modules/audio.py:45-82
# ---------------------------------------------------------------------------------------------
def slaney_mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney') -
    the third-party function modules/audio.py:4,54 calls.  librosa is NOT installed in the authoring
    container, so this restates librosa's published algorithm (Slaney's Auditory Toolbox mel scale:
    linear below 1 kHz at 200/3 Hz per mel, log above with step ln(6.4)/27; triangular filters between
    consecutive mel points; each filter scaled by 2 / bandwidth) and is unpinned by a librosa run."""
    fmax = sr / 2.0 if fmax is None else float(fmax)

    def hz2mel(f):
        f = float(f)
        return f / (200.0 / 3) if f < 1000.0 else 15.0 + math.log(f / 1000.0) / (math.log(6.4) / 27.0)

    def mel2hz(m):
        return (200.0 / 3) * m if m < 15.0 else 1000.0 * math.exp((math.log(6.4) / 27.0) * (m - 15.0))

    lo, hi = hz2mel(fmin), hz2mel(fmax)
    pts = [mel2hz(lo + (hi - lo) * i / (n_mels + 1)) for i in range(n_mels + 2)]
    nb = 1 + n_fft // 2
    fb = np.zeros((n_mels, nb), dtype=np.float64)
    for i in range(n_mels):
        left, centre, right = pts[i], pts[i + 1], pts[i + 2]
        for k in range(nb):
            f = (sr / 2.0) * k / (nb - 1)
            up = (f - left) / (centre - left)
            down = (right - f) / (right - centre)
            fb[i, k] = max(0.0, min(up, down)) * 2.0 / (right - left)
    return fb.astype(np.float32)


def mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False):
    """modules/audio.py:45-82 line by line (torch CPU), with slaney_mel_filterbank for librosa_mel_fn."""
    mel = torch.from_numpy(slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)).float()
    window = torch.hann_window(win_size)
    y = F.pad(y.unsqueeze(1), (int((n_fft - hop_size) / 2), int((n_fft - hop_size) / 2)), mode="reflect").squeeze(1)
    spec = torch.view_as_real(torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=window,
                                         center=center, pad_mode="reflect", normalized=False, onesided=True,
                                         return_complex=True))
    spec = torch.sqrt(spec.pow(2).sum(-1) + 1e-9)
    spec = torch.matmul(mel, spec)
    return torch.log(torch.clamp(spec, min=1e-5))


def interpolate_regulator_v2(sd, tokens, ylens, n_blocks=4):
    """modules/v2/length_regulator.py:74-110, discrete branch without F0: embedding -> nearest
    interpolation (or none when ``n_blocks == 0``, the ar_length_regulator) -> conv stack -> [1x1 conv]."""
    x = F.embedding(tokens if tokens.dim() == 2 else tokens[:, 0], sd["embedding.weight"])      # :76-79
    if n_blocks > 0:
        Tout = int(ylens.max())
        mask = (torch.arange(Tout)[None, :] < ylens[:, None]).unsqueeze(-1)                     # :85
        x = F.interpolate(x.transpose(1, 2).contiguous(), size=Tout, mode="nearest")           # :86
    else:
        x, mask = x.transpose(1, 2).contiguous(), None                                          # :88-89
    for i in range(n_blocks):
        x = F.conv1d(x, sd[f"model.{3 * i}.weight"], sd[f"model.{3 * i}.bias"], padding=1)
        x = F.group_norm(x, 1, sd[f"model.{3 * i + 1}.weight"], sd[f"model.{3 * i + 1}.bias"], 1e-5)
        x = F.mish(x)
    k = 3 * n_blocks
    if f"model.{k}.weight" in sd:                                                               # :53-55
        x = F.conv1d(x, sd[f"model.{k}.weight"], sd[f"model.{k}.bias"])
    out = x.transpose(1, 2).contiguous()
    return out * mask if mask is not None else out                                              # :107-108


def sola_step(infer_wav, sola_buffer, fade_in_window, fade_out_window, sola_buffer_frame, sola_search_frame,
              block_frame):
    """One stream, one tick of real-time-gui.py:1103-1137.  Returns (out block, new sola_buffer, offset)."""
    infer_wav = infer_wav.clone()
    conv_input = infer_wav[None, None, :sola_buffer_frame + sola_search_frame]                # :1104-1106
    cor_nom = F.conv1d(conv_input, sola_buffer[None, None, :])                                # :1108
    cor_den = torch.sqrt(F.conv1d(conv_input ** 2, torch.ones(1, 1, sola_buffer_frame)) + 1e-8)   # :1109-1115
    tensor = cor_nom[0, 0] / cor_den[0, 0]
    sola_offset = int(torch.argmax(tensor, dim=0).item()) if tensor.numel() > 1 else int(tensor.item())
    infer_wav = infer_wav[sola_offset:]                                                       # :1130
    infer_wav[:sola_buffer_frame] *= fade_in_window                                           # :1131
    infer_wav[:sola_buffer_frame] += sola_buffer * fade_out_window                            # :1132-1134
    new_buffer = infer_wav[block_frame:block_frame + sola_buffer_frame].clone()               # :1135-1137
    return infer_wav[:block_frame].clone(), new_buffer, sola_offset



# ---------------------------------------------------------------------------------------------
# HiFT vocoder: modules/hifigan/generator.py:282-454, f0_predictor.py:19-55,
# configs/hifigan.yml.  The two random draws of SineGen (:222-236) are INPUTS here so the same noise can be
# injected into every implementation: ``phase`` (B, nb_harmonics + 1, 1) and ``noise`` (B, nb_harmonics + 1, L)
# standard normal; the third draw (SourceModuleHnNSF.forward's noise branch, :277-278) never reaches the output.
# ---------------------------------------------------------------------------------------------
HIFT_CFG = dict(in_channels=80, base_channels=512, nb_harmonics=8, sampling_rate=22050, nsf_alpha=0.1,
                nsf_sigma=0.003, nsf_voiced_threshold=10, upsample_rates=[8, 8], upsample_kernel_sizes=[16, 16],
                n_fft=16, hop_len=4, resblock_kernel_sizes=[3, 7, 11],
                resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], source_resblock_kernel_sizes=[7, 11],
                source_resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5]], lrelu_slope=0.1, audio_limit=0.99)


def hift_f0_predictor(sd, mel, pfx="f0_predictor."):
    """ConvRNNF0Predictor.forward (f0_predictor.py:52-55): 5 x [weight-normed Conv1d k3 + ELU], Linear, abs."""
    x = mel
    for i in range(5):
        x = F.elu(F.conv1d(x, _wn_weight(sd, f"{pfx}condnet.{2 * i}"), sd[f"{pfx}condnet.{2 * i}.bias"], padding=1))
    return torch.abs(_linear(sd, pfx + "classifier", x.transpose(1, 2)).squeeze(-1))


def hift_source(sd, f0, phase, noise, cfg=HIFT_CFG):
    """_f02source (:366-370) -> SourceModuleHnNSF.forward (:262-279) -> SineGen.forward (:208-243).
    f0 (B, Tm) Hz -> harmonic source (B, L), L = Tm * prod(upsample_rates) * hop_len."""
    scale = int(np.prod(cfg["upsample_rates"])) * cfg["hop_len"]
    f0u = f0[:, None].repeat_interleave(scale, dim=-1)               # nn.Upsample(nearest), (B, 1, L)
    H = cfg["nb_harmonics"] + 1
    F_mat = torch.cat([f0u * (i + 1) / cfg["sampling_rate"] for i in range(H)], dim=1)     # (B, H, L)
    theta = 2 * np.pi * (torch.cumsum(F_mat, dim=-1) % 1)
    ph = phase.clone()
    ph[:, 0, :] = 0
    sine = cfg["nsf_alpha"] * torch.sin(theta + ph)
    uv = (f0u > cfg["nsf_voiced_threshold"]).float()
    noise_amp = uv * cfg["nsf_sigma"] + (1 - uv) * cfg["nsf_alpha"] / 3
    sine = sine * uv + noise_amp * noise
    merged = torch.tanh(_linear(sd, "m_source.l_linear", sine.transpose(1, 2)))            # (B, L, 1)
    return merged[..., 0]


def _hann_periodic(n):
    return torch.hann_window(n, periodic=True)


def hift_stft(x, n_fft=16, hop=4):
    """torch.stft(center=True, reflect, onesided) written out (:372-378): (B, L) -> real, imag (B, n_fft/2+1, L/hop+1)."""
    xp = F.pad(x[:, None], (n_fft // 2, n_fft // 2), mode="reflect")[:, 0]
    frames = xp.unfold(-1, n_fft, hop) * _hann_periodic(n_fft)       # (B, TT, n_fft)
    k = torch.arange(n_fft // 2 + 1, dtype=torch.float64)[:, None]
    n = torch.arange(n_fft, dtype=torch.float64)[None, :]
    ang = 2 * math.pi * k * n / n_fft
    re = frames @ torch.cos(ang).float().t()
    im = frames @ (-torch.sin(ang)).float().t()
    return re.transpose(1, 2), im.transpose(1, 2)


def hift_istft(mag, phase, n_fft=16, hop=4):
    """_istft (:380-385): clip, polar -> torch.istft(center=True) written out as windowed irfft + overlap-add /
    window envelope.  mag, phase (B, n_fft/2+1, TT) -> (B, hop * (TT - 1))."""
    mag = torch.clip(mag, max=1e2)
    re, im = mag * torch.cos(phase), mag * torch.sin(phase)
    B, nb, TT = re.shape
    k = torch.arange(nb, dtype=torch.float64)[:, None]
    n = torch.arange(n_fft, dtype=torch.float64)[None, :]
    ang = 2 * math.pi * k * n / n_fft
    wk = torch.full((nb, 1), 2.0, dtype=torch.float64)
    wk[0] = wk[-1] = 1.0                                             # c2r: DC / Nyquist once, their imag ignored
    C = (wk * torch.cos(ang) / n_fft).float()
    S = (-wk * torch.sin(ang) / n_fft).float()
    w = _hann_periodic(n_fft)
    fr = (re.transpose(1, 2) @ C + im.transpose(1, 2) @ S) * w       # (B, TT, n_fft)
    total = n_fft + hop * (TT - 1)
    y = torch.zeros(B, total)
    env = torch.zeros(total)
    for t in range(TT):
        y[:, t * hop:t * hop + n_fft] += fr[:, t]
        env[t * hop:t * hop + n_fft] += w * w
    half = n_fft // 2
    return y[:, half:total - half] / env[half:total - half]


def _hift_snake(x, alpha):
    """Snake (generator.py:79-90), linear-scale alpha."""
    a = alpha.view(1, -1, 1)
    return x + (1.0 / (a + 1e-9)) * torch.sin(x * a) ** 2


def _hift_resblock(sd, pfx, x, k, dils):
    """ResBlock.forward (:151-158)."""
    for i, d in enumerate(dils):
        xt = _hift_snake(x, sd[f"{pfx}.activations1.{i}.alpha"])
        xt = F.conv1d(xt, _wn_weight(sd, f"{pfx}.convs1.{i}"), sd[f"{pfx}.convs1.{i}.bias"], dilation=d,
                      padding=d * (k - 1) // 2)
        xt = _hift_snake(xt, sd[f"{pfx}.activations2.{i}.alpha"])
        xt = F.conv1d(xt, _wn_weight(sd, f"{pfx}.convs2.{i}"), sd[f"{pfx}.convs2.{i}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


def hift_forward(sd, mel, phase, noise, f0=None, cfg=HIFT_CFG):
    """HiFTGenerator.forward (:387-436).  mel (B, 80, Tm) -> waveform (B, 256 * Tm)."""
    if f0 is None:
        f0 = hift_f0_predictor(sd, mel)
    s = hift_source(sd, f0, phase, noise, cfg)
    re, im = hift_stft(s, cfg["n_fft"], cfg["hop_len"])
    s_stft = torch.cat([re, im], dim=1)
    nk, nu = len(cfg["resblock_kernel_sizes"]), len(cfg["upsample_rates"])
    down = [1] + cfg["upsample_rates"][::-1][:-1]
    down_cum = list(np.cumprod(down))[::-1]                          # [8, 1]
    x = F.conv1d(mel, _wn_weight(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    for i in range(nu):
        u, k = cfg["upsample_rates"][i], cfg["upsample_kernel_sizes"][i]
        x = F.leaky_relu(x, cfg["lrelu_slope"])
        x = F.conv_transpose1d(x, _wn_weight_convT(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)
        if i == nu - 1:
            x = F.pad(x, (1, 0), mode="reflect")
        r = int(down_cum[i])
        if r == 1:
            si = F.conv1d(s_stft, sd[f"source_downs.{i}.weight"], sd[f"source_downs.{i}.bias"])
        else:
            si = F.conv1d(s_stft, sd[f"source_downs.{i}.weight"], sd[f"source_downs.{i}.bias"], stride=r,
                          padding=r // 2)
        si = _hift_resblock(sd, f"source_resblocks.{i}", si, cfg["source_resblock_kernel_sizes"][i],
                            cfg["source_resblock_dilation_sizes"][i])
        x = x + si
        xs = None
        for j in range(nk):
            y = _hift_resblock(sd, f"resblocks.{i * nk + j}", x, cfg["resblock_kernel_sizes"][j],
                               cfg["resblock_dilation_sizes"][j])
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)                                              # default slope 0.01 (:424)
    x = F.conv1d(x, _wn_weight(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    nb = cfg["n_fft"] // 2 + 1
    mag = torch.exp(x[:, :nb])
    ph = torch.sin(x[:, nb:])
    y = hift_istft(mag, ph, cfg["n_fft"], cfg["hop_len"])
    return torch.clamp(y, -cfg["audio_limit"], cfg["audio_limit"])
