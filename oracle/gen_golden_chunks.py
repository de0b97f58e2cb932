"""TEST INFRASTRUCTURE ONLY - golden vectors for the long-form chunk loop.

Runs the REAL reference code: the methods ``crossfade`` and ``_stream_wave_chunks`` are cut out of
/root/reference/seed_vc_wrapper.py with ``ast`` (the module itself cannot be imported here: it pulls
in librosa / torchaudio / pydub) and executed unmodified on seeded synthetic chunk waves, driven by
the same loop as ``convert_voice`` (seed_vc_wrapper.py:560-622) with ``stream_output=False``.

    python oracle/gen_golden_chunks.py      (needs /root/reference; writes tests/golden/stitch_kat.npz)
"""
import ast
import json
import os
import textwrap

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = "/root/reference/seed_vc_wrapper.py"

# (name, source frames S, prompt frames Tp, max_context_window, hop)
CASES = [("three_windows", 150, 20, 80, 4), ("short_last", 105, 20, 80, 4), ("single", 40, 20, 80, 4),
         ("two_windows", 100, 30, 90, 8), ("ragged_last", 106, 20, 80, 2)]


def load_reference_methods():
    tree = ast.parse(open(SRC).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SeedVCWrapper")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("crossfade", "_stream_wave_chunks")]
    assert len(fns) == 2
    mod = ast.Module(body=[ast.ClassDef(name="Ref", bases=[], keywords=[], body=fns, decorator_list=[])],
                     type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np, "torch": torch}
    exec(compile(mod, SRC, "exec"), ns)
    ref = ns["Ref"]()
    ref.overlap_frame_len = 16
    ref.bitrate = "320k"
    return ref


def main():
    ref = load_reference_methods()
    out = {}
    meta = {}
    for name, S, Tp, mcw, hop in CASES:
        g = torch.Generator().manual_seed(7 + [c[0] for c in CASES].index(name))
        overlap_wave_len = ref.overlap_frame_len * hop
        max_source_window = mcw - Tp
        processed, prev, chunks, plan, waves = 0, None, [], [], []
        result = None
        while processed < S:                                   # seed_vc_wrapper.py:560-566
            n_frames = min(max_source_window, S - processed)
            is_last = processed + max_source_window >= S
            vc_target = torch.zeros(1, 80, n_frames)
            vc_wave = torch.randn(1, n_frames * hop, generator=g)
            plan.append((processed, n_frames, bool(is_last)))
            waves.append(vc_wave[0].numpy().copy())
            processed, prev, stop, _, full = ref._stream_wave_chunks(
                vc_wave, processed, vc_target, overlap_wave_len, chunks, prev, is_last, False, 22050)
            if stop:
                result = full
                break
        if result is None:
            result = np.concatenate(chunks)
        out[name + "_out"] = result.astype(np.float32)
        for k, w in enumerate(waves):
            out[f"{name}_w{k}"] = w
        meta[name] = dict(S=S, Tp=Tp, max_context_window=mcw, hop=hop, plan=plan, overlap_frame_len=16)
        print(name, "windows", plan, "->", result.shape)
    # the reference's crossfade on a second chunk shorter than the overlap (its `len(chunk2) < overlap`
    # branch, which the chunk loop itself cannot reach)
    g = torch.Generator().manual_seed(99)
    c1 = torch.randn(64, generator=g).numpy()
    c2 = torch.randn(40, generator=g).numpy()
    out["xf_c1"], out["xf_c2"] = c1.copy(), c2.copy()
    out["xf_out"] = ref.crossfade(c1.copy(), c2.copy(), 64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "stitch_kat.npz"), meta=json.dumps(meta), **out)


if __name__ == "__main__":
    main()
