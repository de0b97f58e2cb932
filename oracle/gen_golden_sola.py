"""TEST INFRASTRUCTURE ONLY - golden vectors for the streaming SOLA stitch from the REAL reference lines.

The arithmetic lives inline in ``audio_callback`` of /root/reference/real-time-gui.py (a GUI method that
cannot be imported or called here).  The exact source lines from ``# SOLA algorithm`` to the
``self.sola_buffer[:] = ...`` statement are cut out of the file by text, dedented and exec'ed unmodified
against a stand-in ``self`` holding the same state the GUI builds at :881-943 (sola_buffer, sin^2 windows),
for several consecutive ticks per stream.

    python oracle/gen_golden_sola.py     (needs /root/reference; writes tests/golden/sola_kat.npz)
"""
import json
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/real-time-gui.py"
# (name, zc, block_time units, crossfade units, ticks, streams): frames follow real-time-gui.py:859-882
CASES = [("rt_default", 441, 9, 2, 3, 3), ("short", 50, 6, 2, 4, 2)]


def reference_snippet():
    lines = open(SRC).read().splitlines()
    a = next(i for i, l in enumerate(lines) if "# SOLA algorithm from" in l)
    b = next(i for i in range(a, len(lines)) if "self.sola_buffer[:] = infer_wav[" in lines[i]) + 3
    code = textwrap.dedent("\n".join(lines[a:b]))
    return compile(code, SRC + ":%d-%d" % (a + 1, b), "exec")


def make_state(zc, block_units, crossfade_units):
    st = types.SimpleNamespace()
    st.zc = zc
    st.block_frame = block_units * zc
    st.crossfade_frame = crossfade_units * zc
    st.sola_buffer_frame = min(st.crossfade_frame, 4 * zc)                    # :881
    st.sola_search_frame = zc                                                 # :882
    st.config = types.SimpleNamespace(device="cpu")
    st.sola_buffer = torch.zeros(st.sola_buffer_frame, dtype=torch.float32)   # :919-921
    st.fade_in_window = torch.sin(0.5 * np.pi * torch.linspace(0.0, 1.0, steps=st.sola_buffer_frame,
                                                                dtype=torch.float32)) ** 2   # :929-942
    st.fade_out_window = 1 - st.fade_in_window                                # :943
    return st


def tick_input(name, stream, tick, n):
    g = torch.Generator().manual_seed(1000 * [c[0] for c in CASES].index(name) + 10 * stream + tick)
    t = torch.arange(n, dtype=torch.float32)
    return 0.4 * torch.sin(0.031 * t + 0.7 * stream + 1.3 * tick) + 0.1 * torch.randn(n, generator=g)


def main():
    code = reference_snippet()
    out, meta = {}, {}
    for name, zc, bu, cu, ticks, streams in CASES:
        sts = [make_state(zc, bu, cu) for _ in range(streams)]
        n = sts[0].sola_buffer_frame + sts[0].sola_search_frame + sts[0].block_frame
        for tick in range(ticks):
            for s, st in enumerate(sts):
                infer_wav = tick_input(name, s, tick, n)
                ns = {"self": st, "infer_wav": infer_wav, "F": F, "torch": torch, "sys": sys, "print": lambda *a: None}
                exec(code, ns)
                out[f"{name}_t{tick}_s{s}_out"] = ns["infer_wav"][:st.block_frame].numpy().copy()
                out[f"{name}_t{tick}_s{s}_buf"] = st.sola_buffer.numpy().copy()
                out[f"{name}_t{tick}_s{s}_off"] = np.int32(ns["sola_offset"])
        meta[name] = dict(zc=zc, block=sts[0].block_frame, sb=sts[0].sola_buffer_frame,
                          search=sts[0].sola_search_frame, ticks=ticks, streams=streams, n=n)
        print(name, meta[name], "offsets", [int(out[f"{name}_t{t}_s0_off"]) for t in range(ticks)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sola_kat.npz"), meta=json.dumps(meta), **out)


if __name__ == "__main__":
    main()
