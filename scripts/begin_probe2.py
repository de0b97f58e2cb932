import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import seedvc_b200
from seedvc_b200 import configs, synth
from seedvc_b200.flow_matching import CFM
from torch.profiler import profile, ProfilerActivity
DEV = "cuda"
args = configs.v1_model_params("whisper_small")
cfm = CFM(args).to(DEV); cfm.estimator.setup_caches(1, 8192); cfm.set_mode("bf16")
B, T, Tp, NS = 32, 2580, 430, 2
mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, 80, 512)]
lens = torch.full((B,), T, device=DEV); t_span = torch.linspace(0, 1, NS + 1, device=DEV)
run = lambda: cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, t_span, 0.7)
run(); run(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = t0
for e in ev[:60]:
    s, en = e.time_range.start, e.time_range.end
    print("%-60s start %9.1f us  dur %8.1f us  idle-before %8.1f" % (e.name[:60], s - t0, en - s, s - prev_end))
    prev_end = max(prev_end, en)
