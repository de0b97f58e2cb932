"""Host-side timeline of DiTEngine.begin(): where does the gap between launches come from?"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import seedvc_b200
from seedvc_b200 import configs, synth
from seedvc_b200.flow_matching import CFM
DEV = "cuda"
args = configs.v1_model_params("whisper_small")
cfm = CFM(args).to(DEV); cfm.estimator.setup_caches(1, 8192); cfm.set_mode("bf16")
B, T, Tp, NS = 32, 2580, 430, 2
mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, 80, 512)]
lens = torch.full((B,), T, device=DEV); t_span = torch.linspace(0, 1, NS + 1, device=DEV)
run = lambda: cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, t_span, 0.7)
run(); run(); torch.cuda.synchronize()
ops = cfm.estimator.engine().ops
log = []
for name in ("gemm", "cast", "bct_to_btc", "norm_mod", "attention", "cfg_euler", "set_rows", "timestep_embedding"):
    orig = getattr(ops, name)
    def mk(orig, name):
        def f(*a, **k):
            t0 = time.perf_counter(); r = orig(*a, **k); log.append((name, t0, time.perf_counter())); return r
        return f
    setattr(ops, name, mk(orig, name))
torch.cuda.synchronize(); T0 = time.perf_counter()
run(); t_host_done = time.perf_counter(); torch.cuda.synchronize(); t_all = time.perf_counter()
print("host enqueue done after %.2f ms, GPU done after %.2f ms" % ((t_host_done - T0) * 1e3, (t_all - T0) * 1e3))
prev = T0
for i, (n, a, b) in enumerate(log[:40]):
    print("%3d %-20s start +%.3f ms (gap %.3f) dur %.3f" % (i, n, (a - T0) * 1e3, (a - prev) * 1e3, (b - a) * 1e3)); prev = b
