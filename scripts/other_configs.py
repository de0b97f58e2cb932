"""Full-size runs of the BASELINE configs that are not the bench line (finite outputs + timing):
config 3 (whisper-base-f0-44k + BigVGAN-44k, 50 steps), config 4 (v2 DiT, 3-branch CFG), config 5
(streaming tiny, B=512 x T=323), config 1 (tiny, B=1).  CUDA events, 1 warm-up + 1 timed."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import seedvc_b200
from seedvc_b200 import configs, synth
from seedvc_b200.bigvgan import BigVGAN
from seedvc_b200.flow_matching import CFM
from seedvc_b200.flow_matching_v2 import CFM as CFMv2, DiT as DiTv2
DEV = "cuda"

def timed(fn):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)

def run_v1(name, model, vocn, B, T, Tp, steps, cfg):
    args = configs.v1_model_params(model)
    C, cd = args.DiT.in_channels, args.DiT.content_dim
    cfm = CFM(args, mode="bf16").to(DEV); cfm.estimator.setup_caches(B, 8192)
    voc = BigVGAN(configs.bigvgan_h(vocn), mode="bf16").to(DEV)
    mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, C, cd)]
    lens = torch.full((B,), T, device=DEV); ts = torch.linspace(0, 1, steps + 1, device=DEV)
    def f():
        mel = cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, ts, cfg)
        return voc(mel[:, :, Tp:].contiguous())
    w, ms = timed(f)
    sec = B * (T - Tp) * voc.h.hop_size / voc.h.sampling_rate
    print(f"{name}: B={B} T={T} steps={steps}: {ms:.1f} ms, {sec / (ms / 1e3):.1f} audio-s/s, finite={bool(torch.isfinite(w).all())}")
    del cfm, voc; torch.cuda.empty_cache()

def run_v2(name, B, T, Tp, steps, cfg):
    kw = configs.v2_estimator_kwargs()
    cfm = CFMv2(DiTv2(**kw), mode="bf16").to(DEV) if "mode" in CFMv2.__init__.__code__.co_varnames else CFMv2(DiTv2(**kw)).to(DEV)
    cfm.set_mode("bf16")
    voc = BigVGAN(configs.bigvgan_h("bigvgan_22k"), mode="bf16").to(DEV)
    mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, kw["in_channels"], kw["content_dim"])]
    lens = torch.full((B,), T, device=DEV)
    ts = torch.linspace(0, 1, steps + 1, device=DEV); ts = ts + (-1) * (torch.cos(torch.pi / 2 * ts) - 1 + ts)
    def f():
        mel = cfm.solve_euler(z.clone(), lens, prompt, mu, style, ts, cfg, False)
        return voc(mel[:, :, Tp:].contiguous())
    w, ms = timed(f)
    sec = B * (T - Tp) * 256 / 22050
    print(f"{name}: B={B} T={T} steps={steps}: {ms:.1f} ms, {sec / (ms / 1e3):.1f} audio-s/s, finite={bool(torch.isfinite(w).all())}")
    del cfm, voc; torch.cuda.empty_cache()

which = sys.argv[1:] or ["1", "3", "4", "5"]
if "1" in which: run_v1("config1 tiny", "xlsr_tiny", "bigvgan_22k", 1, 1291, 430, 10, 0.7)
if "3" in which: run_v1("config3 base-f0-44k (per-GPU share of 8-GPU run: B=8)", "whisper_base", "bigvgan_44k", 8, 2580, 430, 50, 0.7)
if "4" in which: run_v2("config4 v2 (per-GPU share: B=16)", 16, 2580, 430, 25, [0.7, 0.7])
if "5" in which: run_v1("config5 streaming tiny (B=512 x T=323, 10 steps)", "xlsr_tiny", "bigvgan_22k", 512, 323, 258, 10, 0.7)
