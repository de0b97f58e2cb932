"""Category breakdown of one BASELINE config-5 tick (512 streams x T=323, tiny DiT, 10 steps, BigVGAN on 65 frames)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import seedvc_b200
from seedvc_b200 import configs, synth
from seedvc_b200.bigvgan import BigVGAN
from seedvc_b200.flow_matching import CFM
DEV = "cuda"
B, T, Tp, steps = 512, 323, 258, int(sys.argv[1]) if len(sys.argv) > 1 else 10
args = configs.v1_model_params("xlsr_tiny")
cfm = CFM(args, mode="bf16").to(DEV); cfm.estimator.setup_caches(B, 8192)
voc = BigVGAN(configs.bigvgan_h(), mode="bf16").to(DEV)
mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, 80, args.DiT.content_dim)]
lens = torch.full((B,), T, device=DEV); ts = torch.linspace(0, 1, steps + 1, device=DEV)
def f():
    mel = cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, ts, 0.7)
    return voc(mel[:, :, Tp:].contiguous())
f(); torch.cuda.synchronize()
od, ov = cfm.estimator.engine().ops, voc._prepare()["ops"]
od.start_profile(); ov.start_profile(); f()
pd, pv = od.stop_profile(), ov.stop_profile()
tot = sum(d["ms"] for p in (pd, pv) for d in p.values())
for name, p in (("dit", pd), ("voc", pv)):
    for cat, d in sorted(p.items(), key=lambda kv: -kv[1]["ms"]):
        tf = d["flops"] / d["ms"] / 1e9 if d["flops"] else 0
        print(f"{name}.{cat:18s} n={d['launches']:5d} {d['ms']:8.2f} ms {100 * d['ms'] / tot:5.1f}%  {tf:7.1f} TF/s")
print("total", round(tot, 1), "ms")
