"""Per-launch timing of one BigVGAN forward at config-2 size (B=32, Tm=2150).  usage: voc_profile.py [bf16|fp16]"""
import os, sys, torch, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import seedvc_b200
from seedvc_b200 import configs, synth
from seedvc_b200.bigvgan import BigVGAN
from seedvc_b200 import ops as ops_mod
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
voc = BigVGAN(configs.bigvgan_h(), mode=mode).to("cuda")
mel = synth.synth_mel(32, 80, 2150).to("cuda")
voc(mel); torch.cuda.synchronize()
ops = voc._prepare()["ops"]
# wrap gemm / snake to tag shapes
orig_gemm, orig_snake = ops.gemm, ops.snake
tags = []
def gemm(segs, N, **kw):
    tags.append(("gemm", N, segs[0][0].shape[2], len(segs), kw["T"]))
    return orig_gemm(segs, N, **kw)
def snake(x, out, a, b):
    tags.append(("snake", x.shape[2], 0, 0, x.shape[1]))
    return orig_snake(x, out, a, b)
ops.gemm, ops.snake = gemm, snake
ops.start_profile()
voc(mel)
prof = ops.profile
torch.cuda.synchronize()
rows = collections.OrderedDict()
i = 0
for cat, fl, by, e0, e1 in prof:
    if cat in ("gemm_tc", "snake_aa"):
        key = tags[i]; i += 1
    else:
        key = (cat,)
    d = rows.setdefault(key, [0, 0.0])
    d[0] += 1; d[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in rows.values())
for k, v in rows.items():
    print(f"{str(k):44s} n={v[0]:3d}  {v[1]:8.3f} ms  {100*v[1]/tot:5.1f}%")
print("total", tot)
