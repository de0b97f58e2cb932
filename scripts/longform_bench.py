"""Long-form conversion: one 10-minute source through chunking.convert_chunks
(all windows in one ragged batch, GPU stitching) - audio seconds per second, CUDA events."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import seedvc_b200
from seedvc_b200 import configs
from seedvc_b200.bigvgan import BigVGAN
from seedvc_b200.chunking import chunk_plan, convert_chunks
from seedvc_b200.flow_matching import CFM
DEV = "cuda"
args = configs.v1_model_params("whisper_small")
cfm = CFM(args, mode="bf16").to(DEV); cfm.estimator.setup_caches(1, 8192)
voc = BigVGAN(configs.bigvgan_h(), mode="bf16").to(DEV)
sr, hop = voc.h.sampling_rate, voc.h.hop_size
minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
S, Tp, mcw, steps = int(minutes * 60 * sr / hop), 430, sr // hop * 30, 25
g = torch.Generator().manual_seed(0)
cond = torch.randn(1, S, 512, generator=g).to(DEV)
pc = torch.randn(1, Tp, 512, generator=g).to(DEV)
mel2 = (torch.randn(1, 80, Tp, generator=g) * 2 - 4).to(DEV)
style2 = torch.randn(1, 192, generator=g).to(DEV)
plan = chunk_plan(S, Tp, mcw)
run = lambda: convert_chunks(cfm, voc, cond, pc, mel2, style2, steps, 0.7, mcw, hop=hop)
w = run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); w = run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"long-form: {minutes:.1f} min source = {S} frames -> {len(plan)} windows in one batch, "
      f"{w.shape[1] / sr:.1f} s of audio in {ms:.1f} ms = {w.shape[1] / sr / (ms / 1e3):.1f} audio-s/s "
      f"(finite: {bool(torch.isfinite(w).all())})")
