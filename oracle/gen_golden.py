"""TEST INFRASTRUCTURE ONLY - generate tests/golden/*.npz by running the REAL reference.

Run in the authoring container (needs /root/reference):

    python oracle/gen_golden.py

Each fixture records the case description (model, sizes, seeds) and the
reference's outputs.  Weights come from ``seed-vc_b200/synth.py`` (seeded by
parameter name) and inputs from ``synth_batch`` (seeded by utterance id), so the
fixtures hold only outputs.  ``tests/test_oracle_golden.py`` replays every case
through ``oracle/seedvc_oracle.py``; the GPU tests replay them through the CUDA
path.
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
import seedvc_b200  # noqa: E402  (root shim -> seed-vc_b200/)
from seedvc_b200 import configs, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (case name, model, scaled?, T, Tp, n_steps, cfg)
V1_CASES = [
    ("v1_small_scaled_cfg", "whisper_small", True, 70, 20, 3, 0.7),
    ("v1_small_scaled_nocfg", "whisper_small", True, 37, 0, 2, 0.0),
    ("v1_tiny_scaled_cfg", "xlsr_tiny", True, 65, 17, 2, 0.7),
    ("v1_base_scaled_cfg", "whisper_base", True, 50, 49, 2, 0.5),
    ("v1_tiny_full", "xlsr_tiny", False, 200, 60, 2, 0.7),
    ("v1_small_full", "whisper_small", False, 150, 40, 2, 0.7),
    ("v1_base_full", "whisper_base", False, 100, 30, 1, 0.7),
]
# (case name, T, Tp, n_steps, cfg pair, random_voice)
V2_CASES = [
    ("v2_small_3branch", 120, 30, 2, (0.7, 0.7), False),
    ("v2_small_spk_only", 60, 20, 1, (0.0, 0.7), False),
    ("v2_small_txt_only", 60, 20, 1, (0.7, 0.0), False),
    ("v2_small_nocfg", 60, 20, 1, (0.0, 0.0), False),
    ("v2_small_random_voice", 60, 20, 1, (0.7, 0.7), True),
]
BIGVGAN_CASES = [("bigvgan_22k_t12", "bigvgan_22k", 1, 12), ("bigvgan_22k_b2_t7", "bigvgan_22k", 2, 7),
                 ("bigvgan_44k_t6", "bigvgan_44k", 1, 6)]


def v1_args(model, scaled):
    a = configs.v1_model_params(model)
    return configs.scaled_down(a) if scaled else a


def main():
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_import.load()
    from munch import Munch

    def munch(d):
        return Munch({k: munch(v) for k, v in d.items()}) if isinstance(d, dict) else d

    manifest = {}
    torch.manual_seed(0)

    # --- anti-aliased Snake known-answer test --------------------
    act = ns.Activation1d(activation=ns.SnakeBeta(4, alpha_logscale=True))
    x = (torch.arange(40, dtype=torch.float32).reshape(1, 4, 10) / 10)
    with torch.no_grad():
        y0 = act(x)
        g = torch.Generator().manual_seed(5)
        act.act.alpha.copy_(0.3 * torch.randn(4, generator=g))
        act.act.beta.copy_(0.3 * torch.randn(4, generator=g))
        xr = torch.randn(2, 4, 33, generator=g)
        y1 = act(xr)
    np.savez(os.path.join(GOLD, "snake_kat.npz"), x0=x.numpy(), y0=y0.numpy(),
             alpha=act.act.alpha.detach().numpy(), beta=act.act.beta.detach().numpy(),
             x1=xr.numpy(), y1=y1.numpy(), filter=act.upsample.filter.reshape(-1).numpy())

    # --- v1 sampler ----------------------------------------------------------------
    for name, model, scaled, T, Tp, n_steps, cfg in V1_CASES:
        a = v1_args(model, scaled)
        cfm = ns.CFM(munch(a)).eval()
        synth.fill_parameters_(cfm, seed=0)
        cfm.estimator.setup_caches(1, 8192)
        keys = {k: list(v.shape) for k, v in cfm.state_dict().items()}
        manifest[f"keys_{model}{'_scaled' if scaled else ''}"] = keys
        mu, prompt, style, z = synth.synth_batch(1, T, Tp, a.DiT.in_channels, a.DiT.content_dim)
        t_span = torch.linspace(0, 1, n_steps + 1)
        with torch.no_grad():
            # first estimator call (velocity at t=0), cond branch only
            x0 = z.clone()
            px = torch.zeros_like(x0)
            px[..., :Tp] = prompt
            x0[..., :Tp] = 0
            v0 = cfm.estimator(x0, px, torch.tensor([T]), t_span[0:1], style, mu)
            out = cfm.solve_euler(z.clone(), torch.tensor([T]), prompt, mu.clone(), style, None,
                                  t_span, cfg)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), v0=v0.numpy(), out=out.numpy(),
                            meta=json.dumps(dict(model=model, scaled=scaled, T=T, Tp=Tp,
                                                 n_steps=n_steps, cfg=cfg)))
        print(name, "mean|out| =", float(out.abs().mean()), "mean|v0| =", float(v0.abs().mean()))
        del cfm

    # --- v2 sampler ----------------------------------------------------------------
    kw = configs.v2_estimator_kwargs()
    est = ns.DiTv2(**kw).eval()
    synth.fill_parameters_(est, seed=0, prefix="estimator.")
    cfm2 = ns.CFMv2(est).eval()
    manifest["keys_v2_small"] = {k: list(v.shape) for k, v in cfm2.state_dict().items()
                                 if "causal_mask" not in k}
    for name, T, Tp, n_steps, cfg, rv in V2_CASES:
        mu, prompt, style, z = synth.synth_batch(1, T, Tp, kw["in_channels"], kw["content_dim"])
        t_span = torch.linspace(0, 1, n_steps + 1)
        t_span = t_span + (-1) * (torch.cos(torch.pi / 2 * t_span) - 1 + t_span)
        with torch.no_grad():
            out = cfm2.solve_euler(z.clone(), torch.tensor([T]), prompt, mu.clone(), style, t_span,
                                   list(cfg), rv)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), out=out.numpy(),
                            meta=json.dumps(dict(T=T, Tp=Tp, n_steps=n_steps, cfg=list(cfg),
                                                 random_voice=rv)))
        print(name, "mean|out| =", float(out.abs().mean()))
    del cfm2, est

    # --- BigVGAN -------------------------------------------------------------------
    for name, cfgname, B, Tm in BIGVGAN_CASES:
        h = ns.BigVGANAttrDict(dict(configs.bigvgan_h(cfgname)))
        voc = ns.BigVGAN(h).eval()
        voc.remove_weight_norm()
        synth.fill_parameters_(voc, seed=0)
        manifest["keys_" + cfgname] = {k: list(v.shape) for k, v in voc.state_dict().items()}
        mel = synth.synth_mel(B, h.num_mels, Tm)
        with torch.no_grad():
            wav = voc(mel)
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), wav=wav.numpy(),
                            meta=json.dumps(dict(config=cfgname, B=B, Tm=Tm)))
        print(name, "rms =", float(wav.pow(2).mean().sqrt()), "clamped frac =",
              float((wav.abs() >= 1).float().mean()))
        del voc

    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
