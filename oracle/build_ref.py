"""TEST / BASELINE INFRASTRUCTURE ONLY - stage the reference's own hot-path modules under ``oracle/_ref/``.

    python oracle/build_ref.py            (run in the authoring container; needs /root/reference)

The reference is a flat Python research repo (no setup.py / pyproject, so it cannot be pip-installed into
``baseline/_ref``); its hot path is ~20 Python files.  This recipe copies exactly those files, unmodified, into
``oracle/_ref/modules/`` so that ``bench.py --impl reference`` can time the REFERENCE'S OWN ``CFM.solve_euler`` and
``BigVGAN.forward`` on the GPU box's host cores (``/root/reference`` does not exist there).  ``oracle/_ref/`` is
git-ignored (the sources never enter this repo's history) but not gpurun-ignored, so it travels with the snapshot
like the built ``.so``.  Nothing under ``seed-vc_b200/`` imports it; ``oracle/ref_import.py`` is the only loader.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SEEDVC_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")

FILES = [
    "modules/flow_matching.py",                   # CFM.inference / solve_euler
    "modules/diffusion_transformer.py",           # DiT estimator
    "modules/wavenet.py", "modules/encodec.py", "modules/commons.py",
    "modules/length_regulator.py", "modules/audio.py",
    "modules/v2/cfm.py", "modules/v2/dit_wrapper.py", "modules/v2/dit_model.py", "modules/v2/length_regulator.py",
    "modules/bigvgan/bigvgan.py", "modules/bigvgan/activations.py", "modules/bigvgan/utils.py",
    "modules/bigvgan/env.py", "modules/bigvgan/meldataset.py", "modules/bigvgan/config.json",
    "modules/bigvgan/alias_free_activation/torch/__init__.py",
    "modules/bigvgan/alias_free_activation/torch/act.py",
    "modules/bigvgan/alias_free_activation/torch/filter.py",
    "modules/bigvgan/alias_free_activation/torch/resample.py",
    "modules/hifigan/generator.py", "modules/hifigan/f0_predictor.py",
]


def build(verbose=True) -> str | None:
    if not os.path.isdir(os.path.join(REF, "modules")):
        if verbose:
            print(f"build_ref: {REF} not present - keeping whatever is under {DST}")
        return DST if os.path.isdir(os.path.join(DST, "modules")) else None
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Unmodified copies of ArmingLou/seed-vc hot-path modules, staged by oracle/build_ref.py for the\n"
                "reference arm of bench.py. Git-ignored; not part of this repository's source.\n")
    if verbose:
        print(f"build_ref: staged {len(FILES)} reference files under {DST}")
    return DST


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
