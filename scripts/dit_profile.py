"""Per-shape timing of the DiT GEMMs inside a real Euler solve at config-2 size (B=32 -> 64 CFG rows, T=2580)."""
import os, sys, torch, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import seedvc_b200
from seedvc_b200 import configs, synth
from seedvc_b200.flow_matching import CFM
DEV = "cuda"
args = configs.v1_model_params("whisper_small")
cfm = CFM(args).to(DEV); cfm.estimator.setup_caches(1, 8192); cfm.set_mode("bf16")
B, T, Tp, NS = 32, 2580, 430, 3
mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, 80, 512)]
lens = torch.full((B,), T, device=DEV); t_span = torch.linspace(0, 1, NS + 1, device=DEV)
run = lambda: cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, t_span, 0.7)
run(); torch.cuda.synchronize()
ops = cfm.estimator.engine().ops
orig = ops.gemm
tags = []
def gemm(segs, N, **kw):
    fl = [k for k in ("bias", "rowbias", "rope", "gate", "res", "out_f32", "out_op") if kw.get(k) is not None]
    if kw.get("accumulate"): fl.append("acc")
    tags.append((N, tuple(s[0].shape[2] for s in segs), kw["B"], kw["T"], kw.get("act", 0), ",".join(fl), bool(kw.get("f32"))))
    return orig(segs, N, **kw)
ops.gemm = gemm
ops.start_profile(); run(); torch.cuda.synchronize()
prof = ops.profile
rows = collections.OrderedDict(); i = 0
for cat, fl, by, e0, e1 in prof:
    if cat in ("gemm_tc", "gemm_f32"):
        key = ("gemm",) + tags[i]; i += 1
    else:
        key = (cat,)
    d = rows.setdefault(key, [0, 0.0, 0.0]); d[0] += 1; d[1] += e0.elapsed_time(e1); d[2] += fl
tot = sum(v[1] for v in rows.values())
print("per Euler step = totals / %d" % NS)
for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    tf = v[2] / v[1] / 1e9 if v[2] else 0
    print(f"{str(k):100s} n={v[0]:4d} {v[1]/NS:8.3f} ms/step {1000*v[1]/v[0]:8.1f} us each {tf:7.1f} TF/s {100*v[1]/tot:5.1f}%")
print("total per step", tot / NS)
ops.gemm = orig
ops.profile = None
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("unprofiled solve: %.3f ms per step (incl. begin/transpose)" % (e0.elapsed_time(e1) / NS))
# first-gemm anomaly probe: gap between the end of cfg_euler and the end of the merge GEMM
for k, (key, v) in enumerate(rows.items()):
    pass
seq = []
i = 0
for cat, fl, by, a0, a1 in prof:
    seq.append((cat, a0, a1))
for j in range(1, len(seq)):
    gap = seq[j - 1][2].elapsed_time(seq[j][1])
    if gap > 0.05:
        print("gap %.3f ms before launch %d (%s) after %s" % (gap, j, seq[j][0], seq[j - 1][0]))
print("--- launches around the start of Euler step 2")
idx = [j for j, (c, a0, a1) in enumerate(seq) if c == "cfg_euler"]
j0 = idx[0]
ti = 0
tagmap = {}
for j, (c, a0, a1) in enumerate(seq):
    if c in ("gemm_tc", "gemm_f32"):
        tagmap[j] = tags[ti]; ti += 1
for j in range(j0 - 4, j0 + 10):
    c, a0, a1 = seq[j]
    print(j, c, "%.1f us" % (a0.elapsed_time(a1) * 1000), tagmap.get(j, ""))
