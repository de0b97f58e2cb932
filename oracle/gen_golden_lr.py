"""TEST INFRASTRUCTURE ONLY - golden vectors for InterpolateRegulator from the REAL
reference class (modules/length_regulator.py), weights from synth.fill_parameters_ (seeded by name).

    python oracle/gen_golden_lr.py       (needs /root/reference; writes tests/golden/length_regulator.npz)
"""
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

# (name, constructor kwargs, B, Tin, ylens, f0: None | "given")
CASES = [
    ("small_b1", dict(channels=512, sampling_ratios=[1, 1, 1, 1], in_channels=768), 1, 60, [103], None),
    ("small_b2_ragged", dict(channels=512, sampling_ratios=[1, 1, 1, 1], in_channels=768), 2, 47, [81, 64], None),
    ("small_double", dict(channels=512, sampling_ratios=[1, 1, 1, 1], in_channels=768), 1, 40, [80], None),
    ("f0_none", dict(channels=512, sampling_ratios=[1, 1, 1, 1], in_channels=768, f0_condition=True,
                     n_f0_bins=256), 1, 50, [86], None),
    ("f0_given", dict(channels=512, sampling_ratios=[1, 1, 1, 1], in_channels=768, f0_condition=True,
                      n_f0_bins=256), 1, 50, [86], "given"),
]


# v2 (modules/v2/length_regulator.py): (name, kwargs, B, Tin, ylens)
CASES_V2 = [
    ("v2_cfm", dict(channels=512, is_discrete=True, codebook_size=2048, sampling_ratios=[1, 1, 1, 1],
                    f0_condition=False), 2, 40, [69, 52]),
    ("v2_out256", dict(channels=512, is_discrete=True, codebook_size=2048, sampling_ratios=[1, 1], out_channels=256,
                       f0_condition=False), 1, 33, [57]),
    ("v2_ar", dict(channels=768, is_discrete=True, codebook_size=32, sampling_ratios=[], f0_condition=False),
     2, 29, [29, 29]),
]


def tokens(name, B, Tin, codebook):
    g = torch.Generator().manual_seed(71 + [c[0] for c in CASES_V2].index(name))
    return torch.randint(0, codebook, (B, Tin), generator=g)


def inputs(name, B, Tin, Cin, Tf0=None):
    g = torch.Generator().manual_seed(31 + [c[0] for c in CASES].index(name))
    x = torch.randn(B, Tin, Cin, generator=g)
    f0 = None
    if Tf0:
        f0 = torch.rand(B, Tf0, generator=g) * 500 + 60
        f0[:, ::7] = 0.0            # unvoiced frames
    return x, f0


def main():
    ns = ref_import.load()
    out, meta = {}, {}
    for name, kw, B, Tin, ylens, f0mode in CASES:
        m = ns.InterpolateRegulator(**kw).eval()
        with torch.no_grad():
            synth.fill_parameters_(m, seed=0, prefix="length_regulator.")
        x, f0 = inputs(name, B, Tin, kw["in_channels"], Tf0=Tin + 3 if f0mode else None)
        with torch.no_grad():
            y, olens, *_ = m(x, ylens=torch.tensor(ylens), n_quantizers=3, f0=f0)
        out[name] = y.numpy()
        meta[name] = dict(kw=kw, B=B, Tin=Tin, ylens=ylens, f0=f0mode,
                          keys={k: list(v.shape) for k, v in m.state_dict().items()})
        print(name, tuple(y.shape), "mean|y| =", float(y.abs().mean()))
    for name, kw, B, Tin, ylens in CASES_V2:
        m = ns.InterpolateRegulatorV2(**kw).eval()
        with torch.no_grad():
            synth.fill_parameters_(m, seed=0, prefix="cfm_length_regulator.")
        tok = tokens(name, B, Tin, kw["codebook_size"])
        with torch.no_grad():
            y, olens = m(tok, ylens=torch.tensor(ylens), f0=None)
        out[name] = y.numpy()
        meta[name] = dict(kw=kw, B=B, Tin=Tin, ylens=ylens, v2=True,
                          keys={k: list(v.shape) for k, v in m.state_dict().items()})
        print(name, tuple(y.shape), "mean|y| =", float(y.abs().mean()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "length_regulator.npz"), meta=json.dumps(meta), **out)


if __name__ == "__main__":
    main()
