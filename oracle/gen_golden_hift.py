"""TEST INFRASTRUCTURE ONLY - HiFT vocoder goldens from the REAL reference.

    python oracle/gen_golden_hift.py          (authoring container; needs /root/reference)

Runs modules/hifigan/generator.py:HiFTGenerator (with ConvRNNF0Predictor, configs/hifigan.yml values) on seeded
synthetic weights / mels / F0 tracks.  SineGen draws a uniform phase per harmonic and Gaussian noise per sample
(generator.py:222-236); for parity the same draws are INJECTED: ``Uniform.sample`` and ``torch.randn_like`` are
patched for the duration of the call to return ``synth.synth_hift_noise`` (the third draw, the unused noise branch
of SourceModuleHnNSF, gets zeros).  Stored: waveform, the F0 the predictor produced (case without F0), the source
signal and its STFT (intermediate pins for the oracle).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_import  # noqa: E402
import seedvc_b200  # noqa: E402,F401
from seedvc_b200 import synth  # noqa: E402
import seedvc_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
# name: (B, Tm, f0 given?, mel seed, f0 seed, noise seed)
CASES = {
    "hift_b2_t65": (2, 65, True, 31, 3, 5),        # one streaming block (config 5: 65 frames)
    "hift_b1_t300": (1, 300, True, 32, 4, 6),
    "hift_pred_t40": (1, 40, False, 33, 0, 7),     # F0 from the predictor
}


def build(ns):
    cfg = dict(orc.HIFT_CFG)
    kw = {k: v for k, v in cfg.items() if k not in ("n_fft", "hop_len")}
    kw["istft_params"] = {"n_fft": cfg["n_fft"], "hop_len": cfg["hop_len"]}
    gen = ns.HiFTGenerator(**kw, f0_predictor=ns.ConvRNNF0Predictor(num_class=1, in_channels=80, cond_channels=512))
    gen.eval()
    synth.fill_parameters_(gen, seed=0)
    return gen


def run_reference(ns, gen, mel, f0, phase, noise):
    mod = ns.hift_module
    calls = {"n": 0}

    class FakeUniform:
        def __init__(self, *a, **k):
            pass

        def sample(self, sample_shape=()):
            assert tuple(sample_shape) == tuple(phase.shape)
            return phase.clone()

    real_randn_like = torch.randn_like

    def fake_randn_like(t, *a, **k):
        calls["n"] += 1
        if calls["n"] == 1:
            assert t.shape == noise.shape, (t.shape, noise.shape)
            return noise.clone()
        return torch.zeros_like(t)              # SourceModuleHnNSF noise branch: unused by HiFTGenerator

    saved = mod.Uniform
    mod.Uniform = FakeUniform
    torch.randn_like = fake_randn_like
    try:
        with torch.no_grad():
            wav = gen(mel, f0=f0)
    finally:
        mod.Uniform = saved
        torch.randn_like = real_randn_like
    assert calls["n"] == 2
    return wav


def main():
    ns = ref_import.load()
    gen = build(ns)
    keys = {k: list(v.shape) for k, v in gen.state_dict().items()}
    man_path = os.path.join(GOLD, "manifest.json")
    man = json.load(open(man_path))
    man["keys_hift"] = keys
    json.dump(man, open(man_path, "w"))
    out, meta = {}, {}
    H = orc.HIFT_CFG["nb_harmonics"] + 1
    for name, (B, Tm, given, ms, fs, nsd) in CASES.items():
        mel = synth.synth_mel(B, 80, Tm, seed=ms)
        f0 = synth.synth_f0(B, Tm, seed=fs) if given else None
        phase, noise = synth.synth_hift_noise(B, H, Tm * 256, seed=nsd)
        wav = run_reference(ns, gen, mel, f0, phase, noise)
        with torch.no_grad():
            f0_pred = gen.f0_predictor(mel)
        out[name + "_wav"] = wav.numpy()
        out[name + "_f0pred"] = f0_pred.numpy()
        meta[name] = dict(B=B, Tm=Tm, f0_given=given, mel_seed=ms, f0_seed=fs, noise_seed=nsd)
        print(name, tuple(wav.shape), "rms", float(wav.pow(2).mean().sqrt()), "clamped",
              float((wav.abs() >= 0.99).float().mean()), "f0pred mean", float(f0_pred.mean()))
    np.savez_compressed(os.path.join(GOLD, "hift.npz"), meta=json.dumps(meta), **out)


if __name__ == "__main__":
    main()
