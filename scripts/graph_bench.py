"""Latency of one BASELINE config-1 conversion (tiny DiT, B=1, T=1291, 10 steps + BigVGAN on 861 frames):
eager launch sequence vs one CUDA-graph replay (graphs.GraphedConversion).  CUDA events + wall clock."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import seedvc_b200
from seedvc_b200 import configs, synth
from seedvc_b200.bigvgan import BigVGAN
from seedvc_b200.flow_matching import CFM
from seedvc_b200.graphs import GraphedConversion
DEV = "cuda"
cases = {"config1": ("xlsr_tiny", 1, 1291, 430, 10), "stream1": ("xlsr_tiny", 1, 323, 258, 10)}
for name, (model, B, T, Tp, steps) in cases.items():
    args = configs.v1_model_params(model)
    cfm = CFM(args, mode="bf16").to(DEV); cfm.estimator.setup_caches(B, 8192)
    voc = BigVGAN(configs.bigvgan_h(), mode="bf16").to(DEV)
    mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, 80, args.DiT.content_dim)]
    lens = torch.full((B,), T, device=DEV); ts = torch.linspace(0, 1, steps + 1, device=DEV)
    def eager():
        mel = cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, ts, 0.7)
        return voc(mel[:, :, Tp:].contiguous())
    g = GraphedConversion(cfm, voc, B, T, Tp, steps, 0.7)
    graphed = lambda: g(mu, lens, prompt, style, z)
    for label, fn in (("eager", eager), ("graph", graphed)):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 10
        for _ in range(n): fn()
        torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / n * 1e3
        sec = B * (T - Tp) * 256 / 22050
        print(f"{name} {label}: {ms:7.2f} ms per conversion ({sec / (ms / 1e3):7.1f} audio-s/s)")
    same = torch.equal(eager(), graphed())
    print(f"{name}: graph == eager bit-exact: {same}")
