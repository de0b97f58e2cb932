"""Long-form conversion: window the source condition, convert every window in ONE batched call,
stitch the vocoded windows with the reference's cos^2 crossfade on the GPU.

Mirrors the reference's chunk loop (inference.py:470-527, seed_vc_wrapper.py:201-285,560-622):
windows of ``max_context_window - Tp`` source frames, advancing by ``window - overlap_frame_len``;
each window is converted with the prompt in front (``cat[prompt_condition, chunk_cond]``), the
prompt part of the mel is cut, the rest is vocoded, and consecutive waves overlap by
``overlap_frame_len * hop`` samples.  The reference converts the windows one after another and
crossfades on the host through ``.cpu().numpy()``; here the windows are rows of one ragged batch
(the sampler handles per-row lengths) and nothing leaves the device.
"""
from __future__ import annotations

import torch


def chunk_plan(n_source_frames: int, n_prompt_frames: int, max_context_window: int,
               overlap_frame_len: int = 16):
    """[(start, length, is_last)] over the source condition frames (inference.py:470-476,512-526)."""
    window = max_context_window - n_prompt_frames
    if window <= overlap_frame_len:
        raise ValueError("max_context_window must exceed the prompt length by more than the overlap")
    plan, processed = [], 0
    while processed < n_source_frames:
        length = min(window, n_source_frames - processed)
        is_last = processed + window >= n_source_frames
        plan.append((processed, length, is_last))
        if is_last:
            break
        processed += length - overlap_frame_len
    return plan


@torch.no_grad()
def convert_chunks(cfm, vocoder, cond, prompt_condition, mel2, style2, n_timesteps, inference_cfg_rate,
                   max_context_window, overlap_frame_len=16, hop=256, z=None, temperature=1.0):
    """cond (1, S, D) source condition, prompt_condition (1, Tp, D), mel2 (1, C, Tp), style2 (1, 192)
    -> waveform (1, n_samples) on the device.

    ``z`` (n_chunks, C, Tp + window) injects the sampler noise per window (tests); otherwise it is
    drawn like ``CFM.inference`` does (flow_matching.py:50)."""
    dev = cond.device
    S, D = cond.shape[1], cond.shape[2]
    Tp, C = prompt_condition.shape[1], mel2.shape[1]
    plan = chunk_plan(S, Tp, max_context_window, overlap_frame_len)
    n = len(plan)
    Tmax = Tp + max(length for _, length, _ in plan)
    mu = torch.zeros(n, Tmax, D, dtype=torch.float32, device=dev)
    lens = []
    for k, (start, length, _) in enumerate(plan):
        mu[k, :Tp] = prompt_condition[0]
        mu[k, Tp:Tp + length] = cond[0, start:start + length]
        lens.append(Tp + length)
    x_lens = torch.tensor(lens, dtype=torch.long, device=dev)
    if z is None:
        z = torch.randn(n, C, Tmax, device=dev) * temperature
    t_span = torch.linspace(0, 1, n_timesteps + 1, device=dev)
    mel = cfm.solve_euler(z.clone(), x_lens, mel2.expand(n, -1, -1).contiguous(), mu,
                          style2.expand(n, -1).contiguous(), None, t_span, inference_cfg_rate)
    # vocode the generated part of every window; rows of equal length share one vocoder batch
    ops = vocoder._prepare()["ops"] if hasattr(vocoder, "_prepare") else None
    full = [k for k, (_, length, _) in enumerate(plan) if length == plan[0][1]]
    waves = torch.zeros(n, (Tmax - Tp) * hop, dtype=torch.float32, device=dev)
    if full:
        w = vocoder(mel[full, :, Tp:Tp + plan[0][1]].contiguous())
        waves[full, :plan[0][1] * hop] = w[:, 0, :]
    for k, (_, length, _) in enumerate(plan):
        if k not in full:
            w = vocoder(mel[k:k + 1, :, Tp:Tp + length].contiguous())
            waves[k, :length * hop] = w[0, 0, :]
    wave_lens = [length * hop for _, length, _ in plan]
    if ops is None:
        from .ops import Ops
        ops = Ops("fp32")
    return ops.crossfade_stitch(waves, wave_lens, overlap_frame_len * hop)[None, :]
