"""Parity at the BENCHMARKED sizes: the CUDA path against outputs of the REAL reference
(oracle/gen_golden_full.py) at BASELINE.json shapes - T = 2580 with 25 Euler steps (config 2), T = 323 with 25
steps (config-5 frame counts), tiny T = 1291 x 10 steps (config 1), whisper-base-44k and v2 3-branch at T = 2580,
BigVGAN at 256 and 2150 mel frames - per Euler step (the CFG-combined velocity of EVERY step) and end to end,
in fp32 mode (rel-L2 <= 1e-3) and bf16 mode (rel-L2 <= 1e-2), the tolerances north_star states.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

import seedvc_b200  # noqa: E402,F401
from seedvc_b200 import configs, synth  # noqa: E402
from seedvc_b200.bigvgan import BigVGAN  # noqa: E402
from seedvc_b200.flow_matching import CFM  # noqa: E402
from seedvc_b200.flow_matching_v2 import CFM as CFMv2, DiT as DiTv2  # noqa: E402
from conftest import load_golden, rel_l2  # noqa: E402

DEV = "cuda"
TOL = {"fp32": 1e-3, "bf16": 1e-2, "fp16": 3e-3}
SAMPLER = ["full_small_T2580_n25", "full_small_T323_n25", "full_tiny_T1291_n10", "full_base_T2580_n2",
           "full_v2_T2580_n2"]
VOCODER = ["full_bigvgan22k_256", "full_bigvgan22k_2150", "full_bigvgan44k_256"]
_models = {}


def sampler_model(kind, model, mode):
    key = (kind, model)
    if key not in _models:
        if kind == "v1":
            args = configs.v1_model_params(model)
            cfm = CFM(args).to(DEV)
            cfm.estimator.setup_caches(1, 8192)
            dims = (args.DiT.in_channels, args.DiT.content_dim)
        else:
            kw = configs.v2_estimator_kwargs()
            cfm = CFMv2(DiTv2(**kw)).to(DEV)
            dims = (kw["in_channels"], kw["content_dim"])
        _models[key] = (cfm, dims)
    cfm, dims = _models[key]
    cfm.set_mode(mode)
    return cfm, dims


@pytest.mark.parametrize("mode", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("name", SAMPLER)
def test_full_size_sampler_per_step_and_end_to_end(name, mode):
    g = load_golden(name)
    m = g["meta"]
    cfm, (C, cd) = sampler_model(m["kind"], m["model"], mode)
    T, Tp, N = m["T"], m["Tp"], m["n_steps"]
    mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(1, T, Tp, C, cd, first_id=m["utt_id"])]
    fr = torch.from_numpy(g["frames"]).to(DEV)
    xl = torch.tensor([T], device=DEV)
    vs = []

    def hook(s, v):                       # v (1, T, C) -> (C, frames)
        vs.append(v[0].index_select(0, fr).t().float().cpu())

    if m["kind"] == "v1":
        t_span = torch.linspace(0, 1, N + 1, device=DEV)
        out = cfm.solve_euler(z.clone(), xl, prompt, mu, style, None, t_span, m["cfg"], step_hook=hook)
    else:
        t_span = torch.linspace(0, 1, N + 1, device=DEV)
        t_span = t_span + (-1) * (torch.cos(torch.pi / 2 * t_span) - 1 + t_span)
        out = cfm.solve_euler(z.clone(), xl, prompt, mu, style, t_span, list(m["cfg"]), False, step_hook=hook)
    assert len(vs) == N
    errs = [rel_l2(vs[s], g["v_steps"][s]) for s in range(N)]
    e_out = rel_l2(out[0].index_select(1, fr).cpu(), g["out"])
    print(f"{name} [{mode}] per-step velocity rel-L2: first {errs[0]:.2e} max {max(errs):.2e} "
          f"(step {errs.index(max(errs))}) last {errs[-1]:.2e}; end-to-end {e_out:.2e}")
    assert float(out[:, :, :Tp].abs().max()) == 0.0
    assert abs(float(out.double().abs().mean()) - float(g["out_absmean"])) < TOL[mode] * float(g["out_absmean"])
    assert max(errs) < TOL[mode], errs
    assert e_out < TOL[mode]


@pytest.mark.parametrize("mode", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("name", VOCODER)
def test_full_size_vocoder(name, mode):
    g = load_golden(name)
    m = g["meta"]
    key = ("voc", m["config"])
    if key not in _models:
        _models[key] = BigVGAN(configs.bigvgan_h(m["config"])).to(DEV)
    voc = _models[key]
    voc.set_mode(mode)
    mel = synth.synth_mel(m["B"], voc.h.num_mels, m["Tm"], seed=m["mel_seed"]).to(DEV)
    wav = voc(mel)
    assert wav.shape == (m["B"], 1, m["Tm"] * voc.h.hop_size)
    idx = torch.from_numpy(g["idx"]).to(DEV)
    e = rel_l2(wav.index_select(2, idx).cpu(), g["wav"])
    rms = float(wav.double().pow(2).mean().sqrt())
    print(f"{name} [{mode}] waveform rel-L2 {e:.2e}  rms {rms:.5f} (reference {float(g['rms']):.5f})")
    assert e < TOL[mode]
    assert abs(rms - float(g["rms"])) < TOL[mode] * float(g["rms"])


def test_config2_batch_of_goldens_bf16():
    """Config-2 style batch: the T = 2580 golden utterance inside a batch of 4 gives the same per-utterance
    result (batched CFG is defined as the per-utterance batch-1 result)."""
    g = load_golden("full_small_T2580_n25")
    m = g["meta"]
    cfm, (C, cd) = sampler_model("v1", "whisper_small", "bf16")
    T, Tp = m["T"], m["Tp"]
    B = 4
    parts = [synth.synth_batch(1, T, Tp, C, cd, first_id=i) for i in (11, m["utt_id"], 12, 13)]
    mu, prompt, style, z = [torch.cat([p[k] for p in parts]).to(DEV) for k in range(4)]
    t_span = torch.linspace(0, 1, m["n_steps"] + 1, device=DEV)
    out = cfm.solve_euler(z, torch.full((B,), T, device=DEV), prompt, mu, style, None, t_span, m["cfg"])
    fr = torch.from_numpy(g["frames"]).to(DEV)
    e = rel_l2(out[1].index_select(1, fr).cpu(), g["out"])
    print(f"golden utterance inside a batch of {B} [bf16] end-to-end rel-L2 {e:.2e}")
    assert e < TOL["bf16"]
