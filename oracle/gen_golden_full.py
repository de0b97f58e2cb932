"""TEST INFRASTRUCTURE ONLY - goldens at the BENCHMARKED sizes, produced by the REAL reference.

Run in the authoring container (needs /root/reference; a few minutes of CPU):

    python oracle/gen_golden_full.py [case ...]

VERDICT r1 "weak 1/2": the round-1 fixtures stop at T = 200 / 3 Euler steps / 12 vocoder frames.  These
cases run the unmodified reference modules (modules/flow_matching.py:55-112, modules/v2/cfm.py:50-132,
modules/bigvgan/bigvgan.py:360-386) at the BASELINE.json shapes:

  full_small_T2580_n25   config 2 utterance: whisper-small-wavenet, T = 2580 (prompt 430), cfg 0.7, 25 steps
  full_small_T323_n25    config-5 frame counts (T = 323, prompt 258) with 25 steps
  full_tiny_T1291_n10    config 1: xlsr-tiny, T = 1291 (prompt 430), cfg 0.7, 10 steps
  full_base_T2580_n2     config 3 model: whisper-base-f0-44k (C = 128), T = 2580, cfg 0.7, 2 steps
  full_v2_T2580_n2       config 4: v2 small, 3-branch CFG [0.7, 0.7], T = 2580 (+2 tokens), cosine grid, 2 steps
  full_bigvgan22k_256 / full_bigvgan22k_2150 / full_bigvgan44k_256   vocoder at 256 / 2150 mel frames

Every sampler case stores, for EVERY Euler step, the CFG-combined velocity the reference integrates
(``v_steps[s]``; recovered from the estimator's outputs with a forward hook) and the final ``out``, both on a
strided subset of the generated frames (``frames``) so the fixtures stay small; vocoder cases store a strided
subset of the waveform samples (``idx``).  Weights / inputs are the seeded synthetic ones of
``seed-vc_b200/synth.py`` (by parameter name / utterance id), so fixtures hold outputs only.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
import seedvc_b200  # noqa: E402,F401
from seedvc_b200 import configs, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name: (kind, model, T, Tp, n_steps, cfg, frame stride of the stored subset)
SAMPLER_CASES = {
    "full_small_T2580_n25": ("v1", "whisper_small", 2580, 430, 25, 0.7, 8),
    "full_small_T323_n25": ("v1", "whisper_small", 323, 258, 25, 0.7, 1),
    "full_tiny_T1291_n10": ("v1", "xlsr_tiny", 1291, 430, 10, 0.7, 4),
    "full_base_T2580_n2": ("v1", "whisper_base", 2580, 430, 2, 0.7, 8),
    "full_v2_T2580_n2": ("v2", "v2_small", 2580, 430, 2, (0.7, 0.7), 8),
}
# name: (config, B, Tm, sample stride)
VOCODER_CASES = {
    "full_bigvgan22k_256": ("bigvgan_22k", 1, 256, 1),
    "full_bigvgan22k_2150": ("bigvgan_22k", 1, 2150, 4),
    "full_bigvgan44k_256": ("bigvgan_44k", 1, 256, 2),
}


def utt_id(name):
    """Utterance id (input seed) of a case: distinct from the ids the small fixtures use."""
    return 300 + sorted(SAMPLER_CASES).index(name)


def frames_of(T, Tp, fs):
    return np.arange(Tp, T, fs, dtype=np.int64)


def v2_grid(n_steps):
    t = torch.linspace(0, 1, n_steps + 1)
    return t + (-1) * (torch.cos(torch.pi / 2 * t) - 1 + t)          # modules/v2/cfm.py:48


def sampler_case(ns, name):
    from munch import Munch

    def munch(d):
        return Munch({k: munch(v) for k, v in d.items()}) if isinstance(d, dict) else d

    kind, model, T, Tp, n_steps, cfg, fs = SAMPLER_CASES[name]
    if kind == "v1":
        a = configs.v1_model_params(model)
        cfm = ns.CFM(munch(a)).eval()
        synth.fill_parameters_(cfm, seed=0)
        cfm.estimator.setup_caches(1, 8192)
        C, cd = a.DiT.in_channels, a.DiT.content_dim
        coefs = [1.0 + cfg, -cfg]
        t_span = torch.linspace(0, 1, n_steps + 1)
    else:
        kw = configs.v2_estimator_kwargs()
        est = ns.DiTv2(**kw).eval()
        synth.fill_parameters_(est, seed=0, prefix="estimator.")
        cfm = ns.CFMv2(est).eval()
        C, cd = kw["in_channels"], kw["content_dim"]
        w0, w1 = cfg
        coefs = [1.0 + w0 + w1, -w1, -w0]            # stacked order: txt+spk, txt, uncond (cfm.py:113-125)
        t_span = v2_grid(n_steps)
    mu, prompt, style, z = synth.synth_batch(1, T, Tp, C, cd, first_id=utt_id(name))
    fr = frames_of(T, Tp, fs)
    v_steps = []

    def hook(_m, _inp, out):
        v = sum(c * out[i:i + 1] for i, c in enumerate(coefs))
        v_steps.append(v[0][:, fr].clone().numpy())

    h = cfm.estimator.register_forward_hook(hook)
    t0 = time.time()
    with torch.no_grad():
        if kind == "v1":
            out = cfm.solve_euler(z.clone(), torch.tensor([T]), prompt, mu.clone(), style, None, t_span, cfg)
        else:
            out = cfm.solve_euler(z.clone(), torch.tensor([T]), prompt, mu.clone(), style, t_span, list(cfg), False)
    h.remove()
    assert len(v_steps) == n_steps
    assert float(out[:, :, :Tp].abs().max()) == 0.0
    meta = dict(kind=kind, model=model, T=T, Tp=Tp, n_steps=n_steps, cfg=cfg, frame_stride=fs, utt_id=utt_id(name))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), frames=fr, v_steps=np.stack(v_steps),
                        out=out[0][:, fr].numpy(), out_absmean=np.float64(out.double().abs().mean()),
                        meta=json.dumps(meta))
    print(f"{name}: {time.time() - t0:.1f}s  mean|out| {float(out.abs().mean()):.6f}  "
          f"|v| per step {[round(float(np.sqrt((v ** 2).mean())), 4) for v in v_steps[:3]]} ...")


def vocoder_case(ns, name):
    cfgname, B, Tm, ss = VOCODER_CASES[name]
    h = ns.BigVGANAttrDict(dict(configs.bigvgan_h(cfgname)))
    voc = ns.BigVGAN(h).eval()
    voc.remove_weight_norm()
    synth.fill_parameters_(voc, seed=0)
    mel = synth.synth_mel(B, h.num_mels, Tm, seed=23)
    t0 = time.time()
    with torch.no_grad():
        wav = voc(mel)
    idx = np.arange(0, wav.shape[-1], ss, dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), idx=idx, wav=wav[:, :, idx].numpy(),
                        rms=np.float64(wav.double().pow(2).mean().sqrt()),
                        meta=json.dumps(dict(config=cfgname, B=B, Tm=Tm, sample_stride=ss, mel_seed=23)))
    print(f"{name}: {time.time() - t0:.1f}s  rms {float(wav.pow(2).mean().sqrt()):.5f}  clamped "
          f"{float((wav.abs() >= 1).float().mean()):.4f}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_import.load()
    torch.manual_seed(0)
    want = sys.argv[1:] or list(SAMPLER_CASES) + list(VOCODER_CASES)
    for name in want:
        if name in SAMPLER_CASES:
            sampler_case(ns, name)
        else:
            vocoder_case(ns, name)


if __name__ == "__main__":
    main()
