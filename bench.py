#!/usr/bin/env python
"""bench.py - Seed-VC conversion hot path (CFM Euler sampler -> DiT -> BigVGAN) on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm; torchrun launches N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference CPU arm)

A "step" is one pass of the hot path over one batch of synthetic utterances: the whole Euler
solve (n_timesteps estimator calls with batched CFG) followed by the vocoder.  At N = 1 the
workload is BASELINE.json configs[1]: seed-uvit-whisper-small-wavenet DiT + BigVGAN-22k, 25 Euler
steps with CFG 0.7, batch 32 x 30 s context (2580 mel frames = 430 prompt + 2150 generated).
For N > 1 every rank converts its own batch (utterances shard with no collective; weak scaling).

Metric: audio-seconds converted per second (generated frames * hop / sr, summed over ranks).
Printed JSON (rank 0, one line) follows the driver's contract; see DESIGN.md section "Measurement".
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

WORKLOADS = {
    # name: (v1 model, vocoder, B per GPU, T, Tp, Euler steps, cfg)
    "config2": ("whisper_small", "bigvgan_22k", 32, 2580, 430, 25, 0.7),
    "config1": ("xlsr_tiny", "bigvgan_22k", 1, 1291, 430, 10, 0.7),
    "smoke": ("whisper_small", "bigvgan_22k", 2, 323, 65, 4, 0.7),
    "profile": ("whisper_small", "bigvgan_22k", 8, 2580, 430, 2, 0.7),   # config2 shapes, short (ncu)
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override utterances per GPU")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    return ap.parse_args()


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
# reference / CPU arm: the oracle port timed on the host cores, bounded sample
# ------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def cpu_sample_seconds_per_audio_second(workload, threads):
    """Times a bounded sample of the workload with the CPU oracle and extrapolates linearly
    (the Euler loop is step-homogeneous, the vocoder is linear in frames).
    Returns (audio_s_per_s, description, seconds of timed CPU work)."""
    import torch
    import seedvc_b200  # noqa: F401
    from seedvc_b200 import configs, synth
    import seedvc_oracle as orc

    model, voc, B, T, Tp, n_steps, cfg = WORKLOADS[workload]
    torch.set_num_threads(threads)
    if workload not in _CPU_CACHE:          # weights / inputs are built once, outside the timing
        man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifest.json")))
        args = configs.v1_model_params(model)
        h = configs.bigvgan_h(voc)
        _CPU_CACHE[workload] = (
            args, synth.synth_state_dict(man["keys_" + model]), h,
            synth.synth_state_dict(man["keys_" + voc]),
            synth.synth_batch(1, T, Tp, args.DiT.in_channels, args.DiT.content_dim))
    args, sd, h, sdv, (mu, prompt, style, z) = _CPU_CACHE[workload]
    frames = 48
    mel = synth.synth_mel(1, h.num_mels, frames)
    with torch.inference_mode():
        t0 = time.perf_counter()
        t_span = torch.linspace(0, 1, n_steps + 1)[:2]                       # ONE Euler step
        orc.solve_euler_v1(sd, args, z, torch.tensor([T]), prompt, mu, style, t_span, cfg)
        t_step = time.perf_counter() - t0
        t1 = time.perf_counter()
        orc.bigvgan_forward(sdv, h, mel)
        t_voc = time.perf_counter() - t1
    gen = T - Tp
    audio_s = gen * h.hop_size / h.sampling_rate
    total = t_step * n_steps + t_voc * gen / frames
    desc = (f"1 utterance of {workload}: 1 of {n_steps} Euler steps at T={T} (CFG pair) + BigVGAN on "
            f"{frames} of {gen} frames, extrapolated linearly; oracle port, fp32, {threads} threads")
    return audio_s / total, desc, t_step + t_voc


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals, secs, desc = [], [], ""
    for i in range(a.warmup + a.steps):
        v, desc, sec = cpu_sample_seconds_per_audio_second(a.workload, threads)
        if i >= a.warmup:
            vals.append(v)
            secs.append(sec)
    vals.sort()
    v = vals[len(vals) // 2]
    model, voc, B, T, Tp, n_steps, cfg = WORKLOADS[a.workload]
    out = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": v, "unit": "audio-s/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(1e3 * sum(secs) / len(secs), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{a.workload}: {model} DiT + {voc}, {n_steps} Euler steps, cfg {cfg}, "
                               f"T={T} (prompt {Tp})", "note": "reference CPU path (oracle port) on host cores"},
        "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": desc},
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
    from seedvc_b200.sharding import barrier as shard_barrier, max_over_ranks, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL_DEBUG is left to the caller / driver.  NCCL logs to fd 1, so fd 1 is pointed at stderr for the
        # life of the process and the one JSON line goes to a saved copy of the real stdout.
        global _JSON_OUT
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    model, vocn, B, T, Tp, n_steps, cfg = WORKLOADS[a.workload]
    if a.batch:
        B = a.batch
    args = configs.v1_model_params(model)
    C, cd = args.DiT.in_channels, args.DiT.content_dim
    cfm = CFM(args, mode=a.mode).to(dev)
    cfm.estimator.setup_caches(B, 8192)
    voc = BigVGAN(configs.bigvgan_h(vocn), mode=a.mode).to(dev)
    hop, sr = voc.h.hop_size, voc.h.sampling_rate
    gen = T - Tp

    # synthetic utterances of this rank (ids offset by rank): host pinned + device resident copies
    first_utt, n_utt = shard_range(world * B, world, rank)        # contiguous shard of the utterance list
    assert n_utt == B
    mu, prompt, style, z = synth.synth_batch(B, T, Tp, C, cd, first_id=first_utt)
    host = [t.pin_memory() for t in (mu, prompt, style, z)]
    lens_h = torch.full((B,), T, dtype=torch.int64).pin_memory()
    wav_h = torch.empty(B, 1, gen * hop, dtype=torch.float32).pin_memory()
    resident = [t.to(dev) for t in host]
    lens_d = lens_h.to(dev)
    t_span = torch.linspace(0, 1, n_steps + 1, device=dev)

    def convert(mu_d, prompt_d, style_d, z_d, lens):
        mel = cfm.solve_euler(z_d, lens, prompt_d, mu_d, style_d, None, t_span, cfg)
        return voc(mel[:, :, Tp:].contiguous())

    def step_resident():
        return convert(resident[0], resident[1], resident[2], resident[3].clone(), lens_d)

    def step_e2e():
        d = [t.to(dev, non_blocking=True) for t in host]
        lens = lens_h.to(dev, non_blocking=True)
        wav = convert(d[0], d[1], d[2], d[3], lens)
        wav_h.copy_(wav, non_blocking=True)
        return wav

    def barrier():
        shard_barrier(dev)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev)

    ops_d, ops_v = cfm.estimator.engine().ops, voc._prepare()["ops"]
    for _ in range(max(a.warmup, 1)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops_d.launches + ops_v.launches
    ms = timed(step_resident, a.steps)
    launches = ops_d.launches + ops_v.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    audio_s = B * gen * hop / sr
    value = world * audio_s * a.steps / (ms / 1e3)

    step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    e2e_value = world * audio_s * a.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in host) + lens_h.numel() * 8
    d2h = wav_h.numel() * 4

    # ---- per-kernel attribution (separate profiled pass, CUDA events around every launch) ----
    roof, breakdown, euler_ms = None, None, None
    if rank == 0 and not a.no_profile:
        pk = peaks()
        ops_d.start_profile()
        ops_v.start_profile()
        step_resident()
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
            tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_gemm_traffic.json")) \
                if os.path.isdir(os.path.join(ROOT, "profiles")) else []
            if tfiles and a.workload == "config2":
                with open(os.path.join(ROOT, "profiles", tfiles[-1])) as f:
                    tj = json.load(f)
                if tj.get("launches") == g["launches"]:
                    roof["traffic"] = round(tj["traffic_bytes_per_launch"])
                    roof["traffic_unit"] = "bytes per launch (ncu dram read+write, " + tfiles[-1] + ")"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, desc, _ = cpu_sample_seconds_per_audio_second(a.workload, threads)
        cpu = {"value": round(v, 4), "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": desc}
    out = {
        "metric": "audio_seconds_per_second", "value": round(value, 2), "unit": "audio-s/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": round(ms / a.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.mode, "data": "synthetic",
        "config": {"workload": f"{a.workload}: {model} DiT (random init) + {vocn}, {n_steps} Euler steps, "
                               f"cfg {cfg}, batch {B}/GPU x T={T} frames (prompt {Tp}, generated {gen} = "
                               f"{gen * hop / sr:.2f} s)",
                   "batch_per_gpu": B, "parallelism": f"utterance-sharded replicas x{world}, no collectives",
                   "l2": "working set (GBs of activations per step) >> 126 MB L2, no flush needed"},
        "ms_per_euler_step": round(euler_ms, 3) if euler_ms else None,
        "x_realtime_per_gpu": round(value / world, 1),
        "e2e": {"value": round(e2e_value, 2), "unit": "audio-s/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
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
