#!/usr/bin/env python
"""bench.py - Seed-VC conversion hot path (CFM Euler sampler -> DiT -> BigVGAN) on B200.

    python bench.py --gpus N --steps K --warmup W [--workload NAME]          (our arm; torchrun launches N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W [...]     (reference CPU arm, rank 0 only)

A "step" is one pass of the hot path over one batch of synthetic utterances: the whole Euler solve
(n_timesteps estimator calls with batched CFG) followed by the vocoder.  The default workload is BASELINE.json
configs[1] ("config2"); the other BASELINE configs are selectable with --workload:

  config1  xlsr-tiny DiT + BigVGAN-22k, 10 steps, B = 1, T = 1291 (one CUDA-graph replay per conversion)
  config2  whisper-small-wavenet DiT + BigVGAN-22k, 25 steps, 32 utterances per GPU x T = 2580   [weak scaling]
  config3  whisper-base-f0-44k DiT + BigVGAN-44k, 50 steps, 64 utterances split over the ranks    [strong scaling]
  config4  v2 DiT (3-branch CFG, cosine grid) + BigVGAN-22k, 25 steps, 128 utterances split        [strong scaling]
  config5  streaming: 512 concurrent streams, xlsr-tiny, T = 323 (prompt 258), 10 steps, BigVGAN on 65 frames,
           SOLA stitch; one step = one 180 ms tick; adds p50 / p99 tick latency over >= 500 ticks

Metric: audio-seconds converted per second (generated frames * hop / sr, summed over ranks; config5 counts the
audio EMITTED per tick, 0.18 s per stream).  Printed JSON (rank 0, one line) follows the driver's contract; see
DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

_JSON_OUT = None      # real stdout when fd 1 has been redirected (multi-rank runs)

# kind: v1 | v2 | stream.  B: utterances per GPU (weak) or in total (strong).
WORKLOADS = {
    "config2": dict(kind="v1", model="whisper_small", voc="bigvgan_22k", B=32, T=2580, Tp=430, steps=25, cfg=0.7,
                    scaling="weak"),
    "config1": dict(kind="v1", model="xlsr_tiny", voc="bigvgan_22k", B=1, T=1291, Tp=430, steps=10, cfg=0.7,
                    scaling="weak", graph=True),
    "config3": dict(kind="v1", model="whisper_base", voc="bigvgan_44k", B=64, T=2580, Tp=430, steps=50, cfg=0.7,
                    scaling="strong"),
    "config4": dict(kind="v2", model="v2_small", voc="bigvgan_22k", B=128, T=2580, Tp=430, steps=25,
                    cfg=(0.7, 0.7), scaling="strong"),
    "config5": dict(kind="stream", model="xlsr_tiny", voc="bigvgan_22k", B=512, T=323, Tp=258, steps=10, cfg=0.7,
                    scaling="weak", graph=True),
    # config 5 with the vocoder the released xlsr-tiny model is trained against (HiFT)
    "config5_hift": dict(kind="stream", model="xlsr_tiny", voc="hift", B=512, T=323, Tp=258, steps=10, cfg=0.7,
                         scaling="weak", graph=True),
    "smoke": dict(kind="v1", model="whisper_small", voc="bigvgan_22k", B=2, T=323, Tp=65, steps=4, cfg=0.7,
                  scaling="weak"),
    "profile": dict(kind="v1", model="whisper_small", voc="bigvgan_22k", B=8, T=2580, Tp=430, steps=2, cfg=0.7,
                    scaling="weak"),          # config-2 shapes, short (ncu)
}
VOC_FRAMES = 32 * 2150    # mel frames per vocoder call (bounds the stage buffers: ~10 GB at 32 x 2150 frames)
# streaming geometry of config 5 (real-time-gui.py:859-928 with block 0.18 s, crossfade 0.04 s, extra_ce 2.5 s,
# extra_right 0.02 s at 22 050 Hz)
ZC = 441
STREAM = dict(block=9 * ZC, sola_buffer=2 * ZC, sola_search=ZC, tail=1 * ZC, tick_s=0.18)


def voc_spec(name):
    """(hop, sampling rate, mel bins) of a vocoder."""
    return (512, 44100, 128) if name == "bigvgan_44k" else (256, 22050, 80)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the utterance count (per GPU / total)")
    ap.add_argument("--euler-steps", type=int, default=0, help="override the Euler step count (config5: 4 or 10)")
    ap.add_argument("--mode", default="bf16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--latency-ticks", type=int, default=500, help="config5: ticks in the latency distribution")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    return ap.parse_args()


def workload_of(a):
    wl = dict(WORKLOADS[a.workload])
    if a.batch:
        wl["B"] = a.batch
    if a.euler_steps:
        wl["steps"] = a.euler_steps
    return wl


def describe(name, wl, world):
    """``config`` of the JSON line; identical in both arms so the driver can match them."""
    per = "in total, split over the ranks" if wl["scaling"] == "strong" else "per GPU"
    hop, sr, _ = voc_spec(wl["voc"])
    gen = wl["T"] - wl["Tp"]
    s = (f"{name}: {wl['model']} DiT (random init) + {wl['voc']}, {wl['steps']} Euler steps, cfg {wl['cfg']}, "
         f"{wl['B']} utterances {per} x T={wl['T']} frames (prompt {wl['Tp']}, generated {gen} = "
         f"{gen * hop / sr:.2f} s)")
    if wl["kind"] == "stream":
        s += f"; streaming tick {STREAM['tick_s']} s, SOLA stitch, audio counted = emitted block"
    return {"workload": s, "batch": wl["B"], "parallelism": f"utterance-sharded replicas x{world}, no collectives",
            "l2": "per-step working set (GBs of activations; config1: weights + activations re-streamed every "
                  "Euler step) >> 126 MB L2, no flush needed"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# reference / CPU arm: the reference's OWN modules (oracle/_ref staged copy, or /root/reference) on the
# host cores; the oracle port only when neither tree is present.  Bounded sample per step.
# ------------------------------------------------------------------------------------------
_CPU_CACHE = {}
REF_EULER_STEPS = 3       # Euler steps timed per sample (each a full-T CFG estimator call)
REF_VOC_FRAMES = 256      # vocoder frames timed per sample


def _build_reference(wl):
    """(kind, sampler(z, lens, prompt, mu, style, t_span) -> mel, vocoder(mel) -> wav, dims, h)."""
    import seedvc_b200  # noqa: F401
    from seedvc_b200 import configs, synth
    import ref_import

    hop, sr, n_mels = voc_spec(wl["voc"])
    h = configs.to_attr(dict(hop_size=hop, sampling_rate=sr, num_mels=n_mels)) if wl["voc"] == "hift" \
        else configs.bigvgan_h(wl["voc"])
    if ref_import.available():
        ns = ref_import.load()
        from munch import Munch

        def munch(d):
            return Munch({k: munch(v) for k, v in d.items()}) if isinstance(d, dict) else d

        if wl["kind"] == "v2":
            kw = configs.v2_estimator_kwargs()
            est = ns.DiTv2(**kw).eval()
            synth.fill_parameters_(est, seed=0, prefix="estimator.")
            cfm = ns.CFMv2(est).eval()
            dims = (kw["in_channels"], kw["content_dim"])

            def sampler(z, lens, prompt, mu, style, t_span):
                return cfm.solve_euler(z, lens, prompt, mu, style, t_span, list(wl["cfg"]), False)
        else:
            a = configs.v1_model_params(wl["model"])
            cfm = ns.CFM(munch(a)).eval()
            synth.fill_parameters_(cfm, seed=0)
            cfm.estimator.setup_caches(1, 8192)
            dims = (a.DiT.in_channels, a.DiT.content_dim)

            def sampler(z, lens, prompt, mu, style, t_span):
                return cfm.solve_euler(z, lens, prompt, mu, style, None, t_span, wl["cfg"])
        if wl["voc"] == "hift":
            import seedvc_oracle as orc
            hk = {k: v for k, v in orc.HIFT_CFG.items() if k not in ("n_fft", "hop_len")}
            hk["istft_params"] = {"n_fft": orc.HIFT_CFG["n_fft"], "hop_len": orc.HIFT_CFG["hop_len"]}
            voc = ns.HiFTGenerator(**hk, f0_predictor=ns.ConvRNNF0Predictor(num_class=1, in_channels=80,
                                                                            cond_channels=512)).eval()
        else:
            voc = ns.BigVGAN(ns.BigVGANAttrDict(dict(h))).eval()
            voc.remove_weight_norm()
        synth.fill_parameters_(voc, seed=0)
        return ref_import.kind(), sampler, voc, dims, h
    # ---- fallback: the oracle port --------------------------------------------------------
    import seedvc_oracle as orc
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
    sdv = synth.synth_state_dict(man["keys_" + wl["voc"]])
    if wl["voc"] == "hift":
        def port_voc(mel):
            ph, nz = synth.synth_hift_noise(mel.shape[0], orc.HIFT_CFG["nb_harmonics"] + 1, mel.shape[-1] * 256)
            return orc.hift_forward(sdv, mel, ph, nz)
    else:
        def port_voc(mel):
            return orc.bigvgan_forward(sdv, h, mel)
    if wl["kind"] == "v2":
        kw = configs.v2_estimator_kwargs()
        sd = synth.synth_state_dict(man["keys_v2_small"])
        dims = (kw["in_channels"], kw["content_dim"])

        def sampler(z, lens, prompt, mu, style, t_span):
            return orc.solve_euler_v2(sd, kw, z, lens, prompt, mu, style, t_span, wl["cfg"])
    else:
        a = configs.v1_model_params(wl["model"])
        sd = synth.synth_state_dict(man["keys_" + wl["model"]])
        dims = (a.DiT.in_channels, a.DiT.content_dim)

        def sampler(z, lens, prompt, mu, style, t_span):
            return orc.solve_euler_v1(sd, a, z, lens, prompt, mu, style, t_span, wl["cfg"])
    return "port", sampler, port_voc, dims, h


def cpu_sample(name, wl, threads):
    """One bounded sample of the workload on the host cores -> (audio_s_per_s, kind, description, seconds).

    One utterance (the reference cannot batch under CFG, flow_matching.py:90-94): REF_EULER_STEPS whole Euler
    steps at full T through the reference's own ``solve_euler`` and its ``BigVGAN.forward`` on REF_VOC_FRAMES
    frames, scaled linearly to the workload's step / frame counts (the loop is step-homogeneous and the vocoder
    linear in frames).  Units small enough are timed whole (xlsr-tiny: all Euler steps; config5: the 65 frames)."""
    import torch
    from seedvc_b200 import synth

    torch.set_num_threads(threads)
    key = (name, wl["steps"])
    if key not in _CPU_CACHE:
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):      # the reference prints while building ("Removing weight norm...")
            kind, sampler, voc, (C, cd), h = _build_reference(wl)
        T, Tp = wl["T"], wl["Tp"]
        _CPU_CACHE[key] = (kind, sampler, voc, h, synth.synth_batch(1, T, Tp, C, cd),
                           synth.synth_mel(1, h.num_mels, min(REF_VOC_FRAMES, T - Tp)))
    kind, sampler, voc, h, (mu, prompt, style, z), mel = _CPU_CACHE[key]
    T, Tp, N = wl["T"], wl["Tp"], wl["steps"]
    n_s = N if (wl["model"] == "xlsr_tiny") else min(REF_EULER_STEPS, N)
    if wl["kind"] == "v2":
        import seedvc_oracle as orc
        t_span = orc.v2_t_span(N)[:n_s + 1]
    else:
        t_span = torch.linspace(0, 1, N + 1)[:n_s + 1]
    frames = mel.shape[-1]
    gen = T - Tp
    with torch.inference_mode():
        t0 = time.perf_counter()
        sampler(z.clone(), torch.tensor([T]), prompt, mu.clone(), style, t_span)
        t_s = time.perf_counter() - t0
        t1 = time.perf_counter()
        voc(mel)
        t_v = time.perf_counter() - t1
    total = t_s * N / n_s + t_v * gen / frames
    audio_s = STREAM["tick_s"] if wl["kind"] == "stream" else gen * h.hop_size / h.sampling_rate
    what = {"reference": "the reference's own modules (/root/reference)",
            "oracle/_ref": "the reference's own modules (staged copy oracle/_ref)",
            "port": "oracle port (reference tree absent)"}[kind]
    desc = (f"1 utterance of {name}: {n_s} of {N} Euler steps at T={T} (CFG stacked) through solve_euler + the vocoder on "
            f"{frames} of {gen} frames ({wl['voc']}), scaled linearly to the full unit; {what}, fp32, {threads} threads")
    return audio_s / total, kind, desc, t_s + t_v


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload_of(a)
    threads = os.cpu_count() or 1
    vals, secs, desc, kind = [], [], "", "port"
    for i in range(a.warmup + a.steps):
        v, kind, desc, sec = cpu_sample(a.workload, wl, threads)
        if i >= a.warmup:
            vals.append(v)
            secs.append(sec)
    vals.sort()
    v = vals[len(vals) // 2]
    out = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": v, "unit": "audio-s/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(1e3 * sum(secs) / len(secs), 1),
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": describe(a.workload, wl, max(a.gpus, 1)),
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads,
                         "kind": "port" if kind == "port" else "reference", "source": kind, "sample": desc},
        "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    import seedvc_b200  # noqa: F401
    from seedvc_b200 import configs, synth
    from seedvc_b200.bigvgan import BigVGAN
    from seedvc_b200.flow_matching import CFM
    from seedvc_b200.flow_matching_v2 import CFM as CFMv2, DiT as DiTv2
    from seedvc_b200.graphs import GraphedConversion
    from seedvc_b200.sharding import barrier as shard_barrier, max_over_ranks, shard_range
    from seedvc_b200.streaming import SolaStitcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    global _JSON_OUT
    if world > 1:
        # NCCL_DEBUG is left to the caller / driver.  NCCL logs to fd 1, so fd 1 is pointed at stderr for the
        # life of the process and the one JSON line goes to a saved copy of the real stdout.
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    wl = workload_of(a)
    kind, T, Tp, n_steps, cfg = wl["kind"], wl["T"], wl["Tp"], wl["steps"], wl["cfg"]
    # ---- this rank's shard of the utterance list ------------------------------------------------
    B_total = wl["B"] if wl["scaling"] == "strong" else wl["B"] * world
    first_utt, B = shard_range(B_total, world, rank)
    if B < 1:
        raise SystemExit(f"rank {rank}: no utterances (B_total {B_total} < world {world})")
    if kind == "v2":
        kw = configs.v2_estimator_kwargs()
        cfm = CFMv2(DiTv2(**kw)).to(dev)
        cfm.set_mode(a.mode)
        C, cd = kw["in_channels"], kw["content_dim"]
    else:
        args = configs.v1_model_params(wl["model"])
        cfm = CFM(args, mode=a.mode).to(dev)
        cfm.estimator.setup_caches(B, 8192)
        C, cd = args.DiT.in_channels, args.DiT.content_dim
    if wl["voc"] == "hift":
        from seedvc_b200.hifigan import ConvRNNF0Predictor, HiFTGenerator
        hift = HiFTGenerator(f0_predictor=ConvRNNF0Predictor(), mode=a.mode).to(dev)
        voc = lambda mel: hift(mel).unsqueeze(1)            # (B, L) -> (B, 1, L) like BigVGAN
        voc._prepare = hift._prepare
    else:
        voc = BigVGAN(configs.bigvgan_h(wl["voc"]), mode=a.mode).to(dev)
    hop, sr, _ = voc_spec(wl["voc"])
    gen = T - Tp

    mu, prompt, style, z = synth.synth_batch(B, T, Tp, C, cd, first_id=first_utt)
    lens_h = torch.full((B,), T, dtype=torch.int64).pin_memory()
    t_span = torch.linspace(0, 1, n_steps + 1, device=dev)
    if kind == "v2":
        t_span = t_span + (-1) * (torch.cos(torch.pi / 2 * t_span) - 1 + t_span)

    def sample(mu_d, prompt_d, style_d, z_d, lens):
        if kind == "v2":
            return cfm.solve_euler(z_d, lens, prompt_d, mu_d, style_d, t_span, list(cfg), False)
        return cfm.solve_euler(z_d, lens, prompt_d, mu_d, style_d, None, t_span, cfg)

    voc_chunk = max(1, VOC_FRAMES // gen)

    def vocode(mel):
        if mel.shape[0] <= voc_chunk:
            return voc(mel)
        return torch.cat([voc(mel[i:i + voc_chunk]) for i in range(0, mel.shape[0], voc_chunk)])

    def convert_eager(mu_d, prompt_d, style_d, z_d, lens):
        return vocode(sample(mu_d, prompt_d, style_d, z_d, lens)[:, :, Tp:].contiguous())

    use_graph = bool(wl.get("graph")) and not a.no_graph and kind != "v2"
    graphed = GraphedConversion(cfm, voc, B, T, Tp, n_steps, cfg, device=dev) if use_graph else None

    def convert(mu_d, prompt_d, style_d, z_d, lens):
        if graphed is not None:
            return graphed(mu_d, lens, prompt_d, style_d, z_d)
        return convert_eager(mu_d, prompt_d, style_d, z_d.clone(), lens)

    # ---- per-step host inputs (pinned) and what comes back --------------------------------------
    if kind == "stream":
        # per tick the new content window and the noise arrive from the host; prompt / style are per-stream
        # session constants and stay resident.  Returned: the emitted block of every stream.
        host = [t.pin_memory() for t in (mu, z)]
        const = [t.to(dev) for t in (prompt, style)]
        stitch = SolaStitcher(B, STREAM["sola_buffer"], STREAM["sola_search"], STREAM["block"], device=dev)
        need = STREAM["sola_search"] + STREAM["block"] + STREAM["sola_buffer"]
        out_h = torch.empty(B, STREAM["block"], dtype=torch.float32).pin_memory()
        audio_rank, audio_total = B * STREAM["tick_s"], B_total * STREAM["tick_s"]

        def finish(wave):            # real-time-gui.py:150-154 (cut) + :1103-1137 (SOLA)
            w = wave[:, 0]
            return stitch.step(w[:, w.shape[1] - need - STREAM["tail"]: w.shape[1] - STREAM["tail"]])

        def assemble(d):
            return d[0], const[0], const[1], d[1]
    else:
        host = [t.pin_memory() for t in (mu, prompt, style, z)]
        out_h = torch.empty(B, 1, gen * hop, dtype=torch.float32).pin_memory()
        audio_rank, audio_total = B * gen * hop / sr, B_total * gen * hop / sr

        def finish(wave):
            return wave

        def assemble(d):
            return d[0], d[1], d[2], d[3]
    resident = [t.to(dev) for t in host]
    lens_d = lens_h.to(dev)

    def step_resident():
        m, p_, s_, zz = assemble(resident)
        return finish(convert(m, p_, s_, zz, lens_d))

    def step_e2e():
        d = [t.to(dev, non_blocking=True) for t in host]
        lens = lens_h.to(dev, non_blocking=True)
        m, p_, s_, zz = assemble(d)
        out = finish(convert(m, p_, s_, zz, lens))
        out_h.copy_(out, non_blocking=True)
        return out

    def eager_once():
        m, p_, s_, zz = assemble(resident)
        return finish(convert_eager(m, p_, s_, zz.clone(), lens_d))

    def timed(fn, k):
        shard_barrier(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        shard_barrier(dev)
        return max_over_ranks(e0.elapsed_time(e1), dev)

    ops_d, ops_v = cfm.estimator.engine().ops, voc._prepare()["ops"]
    # kernels launched by one conversion (a graph replay launches what one eager pass launches)
    l0 = ops_d.launches + ops_v.launches
    eager_once()
    launches_per_conv = ops_d.launches + ops_v.launches - l0 + (1 if kind == "stream" else 0)
    for _ in range(max(a.warmup, 1)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = audio_total * a.steps / (ms / 1e3)

    step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    e2e_value = audio_total * a.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host) + lens_h.numel() * 8
    d2h = out_h.numel() * 4

    # ---- config 5: per-tick latency distribution (host wall clock around H2D -> ... -> D2H + sync) ----
    latency = None
    if kind == "stream" and a.latency_ticks > 0:
        lat = []
        torch.cuda.synchronize(dev)
        for _ in range(a.latency_ticks):
            t0 = time.perf_counter()
            step_e2e()
            torch.cuda.current_stream(dev).synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat_t = torch.tensor(lat, dtype=torch.float64, device=dev)
        if world > 1:           # a tick is served when the slowest rank has served it
            dist.all_reduce(lat_t, op=dist.ReduceOp.MAX)
        ls = sorted(lat_t.tolist())
        q = lambda f: ls[min(len(ls) - 1, int(f * len(ls)))]
        latency = {"ticks": len(ls), "p50_ms": round(q(0.50), 3), "p90_ms": round(q(0.90), 3),
                   "p99_ms": round(q(0.99), 3), "max_ms": round(ls[-1], 3), "budget_ms": 1e3 * STREAM["tick_s"],
                   "streams_per_gpu": B, "euler_steps": n_steps,
                   "what": "host wall clock per tick: H2D of the tick's inputs, sampler, vocoder, SOLA stitch, "
                           "D2H of the emitted blocks, stream synchronize"}

    # ---- per-kernel attribution (separate profiled pass, CUDA events around every launch) ----
    roof, breakdown, euler_ms = None, None, None
    if rank == 0 and not a.no_profile:
        pk = peaks()
        ops_d.start_profile()
        ops_v.start_profile()
        eager_once()
        prof_d, prof_v = ops_d.stop_profile(), ops_v.stop_profile()
        breakdown = {}
        for name, prof in (("dit", prof_d), ("vocoder", prof_v)):
            for cat, d in prof.items():
                breakdown[f"{name}.{cat}"] = {"launches": d["launches"], "ms": round(d["ms"], 3),
                                              "tflops": round(d["flops"] / d["ms"] / 1e9, 1) if d["flops"] else None,
                                              "gbs": round(d["bytes"] / d["ms"] / 1e6, 1) if d["bytes"] else None}
        tot = sum(d["ms"] for p_ in (prof_d, prof_v) for d in p_.values())
        euler_ms = sum(d["ms"] for d in prof_d.values()) / n_steps
        g = {"flops": 0.0, "ms": 0.0, "launches": 0}
        for p_ in (prof_d, prof_v):
            if "gemm_tc" in p_:
                for k_ in g:
                    g[k_] += p_["gemm_tc"][k_]
        if g["ms"] > 0:
            ach = g["flops"] / g["ms"] / 1e9
            roof = {"kernel": "gemm_tc_kernel (tcgen05 segmented GEMM: DiT linears, WaveNet and BigVGAN convs)",
                    "bound": "tensor", "achieved": round(ach, 1), "peak": pk["tf_sustained"],
                    "unit": "TFLOP/s", "frac": round(ach / pk["tf_sustained"], 4),
                    "peak_source": f"{pk['src']} bf16_tflops_sustained (kernel timed inside a long step)",
                    "share_of_step": round(g["ms"] / tot, 3), "launches_per_step": g["launches"],
                    "avg_launch_ms": round(g["ms"] / g["launches"], 4), "traffic": None}
            # DRAM bytes per launch from the committed ncu pass over one config-2 conversion
            # (profiles/*_gemm_traffic.json, scripts/gpu_ncu.sh); only valid for that workload
            pdir = os.path.join(ROOT, "profiles")
            tfiles = sorted(f for f in os.listdir(pdir) if f.endswith("_gemm_traffic.json")) \
                if os.path.isdir(pdir) else []
            if tfiles and a.workload == "config2":
                with open(os.path.join(pdir, tfiles[-1])) as f:
                    tj = json.load(f)
                if tj.get("launches") == g["launches"]:
                    roof["traffic"] = round(tj["traffic_bytes_per_launch"])
                    roof["traffic_source"] = ("committed ncu pass (dram__bytes_read.sum + dram__bytes_write.sum "
                                              "per launch, profiles/" + tfiles[-1] + "), not measured in this run")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, ckind, desc, _ = cpu_sample(a.workload, wl, threads)
        cpu = {"value": round(v, 4), "unit": "audio-s/s", "cores": threads,
               "kind": "port" if ckind == "port" else "reference", "source": ckind, "sample": desc}
    out = {
        "metric": "audio_seconds_per_second", "value": round(value, 2), "unit": "audio-s/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": a.mode, "data": "synthetic",
        "dtype_detail": {"bf16": "bf16 operands on tcgen05 for the transformer-branch GEMMs and attention; IEEE half for "
                                 "operands that carry a residual stream and for the whole vocoder; fp32 accumulation, "
                                 "residual streams and norms",
                         "fp16": "IEEE half operands on tcgen05; fp32 accumulation, residual streams and norms",
                         "fp32": "fp32 FFMA kernels"}[a.mode],
        "config": describe(a.workload, wl, world),
        "batch_this_rank": B, "cuda_graph": graphed is not None,
        "ms_per_euler_step": round(euler_ms, 3) if euler_ms else None,
        "x_realtime_per_gpu": round(value / world, 1),
        "e2e": {"value": round(e2e_value, 2), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches_per_conv * a.steps,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "latency": latency,
        "kernel_breakdown": breakdown,
    }
    print(json.dumps(out), file=_JSON_OUT or sys.stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
