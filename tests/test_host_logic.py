"""Host orchestration (weight folding, hoisting, CFG layouts, polyphase upsampling ...) checked on
CPU against the reference's golden outputs, with tests/emu_ops.py standing in for the CUDA
library.  The kernels themselves are checked on the GPU (tests/test_gpu_*.py)."""
import pytest
import torch

import seedvc_b200  # noqa: F401
from seedvc_b200 import configs, synth
from seedvc_b200.bigvgan import BigVGAN
from seedvc_b200.dit_engine import DiTEngine
from seedvc_b200.flow_matching import CFM
from seedvc_b200.flow_matching_v2 import CFM as CFMv2, DiT as DiTv2
from conftest import load_golden, rel_l2
from emu_ops import EmuOps

V1 = ["v1_small_scaled_cfg", "v1_small_scaled_nocfg", "v1_tiny_scaled_cfg", "v1_base_scaled_cfg",
      "v1_tiny_full"]
V2 = ["v2_small_3branch", "v2_small_spk_only", "v2_small_txt_only", "v2_small_nocfg",
      "v2_small_random_voice"]


def emu_engine(estimator, fold=False):
    eng = DiTEngine(estimator.spec, EmuOps(fold_norms=fold))
    eng.load_weights(estimator.state_dict(), "cpu")
    estimator.engine = lambda: eng
    return eng


def v1_case(name, fold=False):
    g = load_golden(name)
    m = g["meta"]
    args = configs.v1_model_params(m["model"])
    if m["scaled"]:
        args = configs.scaled_down(args)
    cfm = CFM(args)
    emu_engine(cfm.estimator, fold)
    cfm.estimator.setup_caches(1, 8192)
    mu, prompt, style, z = synth.synth_batch(1, m["T"], m["Tp"], args.DiT.in_channels,
                                             args.DiT.content_dim)
    return g, m, cfm, (mu, prompt, style, z)


@pytest.mark.parametrize("fold", [False, True], ids=["norm_kernels", "folded_norms"])
@pytest.mark.parametrize("name", V1)
def test_v1_host_logic(name, fold):
    """fold: the engine's folded-RMS-norm orchestration (DiTEngine._fold_begin / _layers_folded) on the emulated ops."""
    g, m, cfm, (mu, prompt, style, z) = v1_case(name, fold)
    assert cfm.estimator.engine().fold == fold
    T, Tp = m["T"], m["Tp"]
    t_span = torch.linspace(0, 1, m["n_steps"] + 1)
    x0 = z.clone()
    px = torch.zeros_like(x0)
    px[..., :Tp] = prompt
    x0[..., :Tp] = 0
    v0 = cfm.estimator(x0, px, torch.tensor([T]), t_span[0:1], style, mu)
    assert rel_l2(v0, g["v0"]) < 1e-4
    out = cfm.solve_euler(z.clone(), torch.tensor([T]), prompt, mu, style, None, t_span, m["cfg"])
    assert rel_l2(out, g["out"]) < 1e-4


@pytest.mark.parametrize("name", V2)
def test_v2_host_logic(name):
    g = load_golden(name)
    m = g["meta"]
    kw = configs.v2_estimator_kwargs()
    cfm = CFMv2(DiTv2(**kw))
    emu_engine(cfm.estimator)
    mu, prompt, style, z = synth.synth_batch(1, m["T"], m["Tp"], kw["in_channels"], kw["content_dim"])
    t_span = torch.linspace(0, 1, m["n_steps"] + 1)
    t_span = t_span + (-1) * (torch.cos(torch.pi / 2 * t_span) - 1 + t_span)
    out = cfm.solve_euler(z.clone(), torch.tensor([m["T"]]), prompt, mu, style, t_span, m["cfg"],
                          m["random_voice"])
    assert rel_l2(out, g["out"]) < 1e-4


def test_batched_equals_per_utterance():
    """Batch of 2 with different lengths == two batch-1 runs (SURVEY App. D-1 semantics)."""
    args = configs.scaled_down(configs.v1_model_params("whisper_small"))
    cfm = CFM(args)
    emu_engine(cfm.estimator)
    T, Tp = 40, 9
    mu, prompt, style, z = synth.synth_batch(2, T, Tp, 80, args.DiT.content_dim)
    lens = torch.tensor([T, 29])
    t_span = torch.linspace(0, 1, 3)
    out = cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, t_span, 0.7)
    for b in range(2):
        n = int(lens[b])
        one = cfm.solve_euler(z[b:b + 1, :, :n].clone(), lens[b:b + 1], prompt[b:b + 1],
                              mu[b:b + 1, :n], style[b:b + 1], None, t_span, 0.7)
        assert rel_l2(out[b:b + 1, :, :n], one) < 1e-5
        assert float(out[b, :, n:].abs().max()) == 0.0 if n < T else True


@pytest.mark.parametrize("name", ["bigvgan_22k_t12", "bigvgan_22k_b2_t7", "bigvgan_44k_t6"])
def test_bigvgan_host_logic(name):
    g = load_golden(name)
    m = g["meta"]
    h = configs.bigvgan_h(m["config"])
    voc = BigVGAN(h)
    w = voc._build_weights(EmuOps())
    voc._prepare = lambda: w
    mel = synth.synth_mel(m["B"], h.num_mels, m["Tm"])
    wav = voc(mel)
    assert wav.shape == g["wav"].shape
    assert rel_l2(wav, g["wav"]) < 1e-4


def test_bigvgan_loads_weight_norm_checkpoint():
    h = configs.bigvgan_h()
    voc = BigVGAN(h)
    sd = voc.state_dict()
    v = sd.pop("conv_pre.weight").clone()
    g = v.flatten(1).norm(dim=1).view(-1, 1, 1) * 1.5
    sd["conv_pre.weight_v"], sd["conv_pre.weight_g"] = v * 3.0, g
    voc.load_state_dict(sd)
    assert torch.allclose(voc.conv_pre.weight, v * 1.5, atol=1e-6)


def test_chunk_plan_matches_reference_loop():
    """Host chunk scheduler (seed-vc_b200/chunking.py) vs the windows the REAL reference loop visited."""
    import json
    import os

    import numpy as np
    from seedvc_b200.chunking import chunk_plan

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "stitch_kat.npz"))
    for name, m in json.loads(str(z["meta"])).items():
        plan = chunk_plan(m["S"], m["Tp"], m["max_context_window"], m["overlap_frame_len"])
        assert [list(p) for p in plan] == [list(p) for p in m["plan"]], name
    import pytest
    with pytest.raises(ValueError):
        chunk_plan(100, 70, 80, 16)


def test_nearest_index_matches_aten():
    """Host nearest-neighbour index table (length_regulator.nearest_index) vs F.interpolate(mode='nearest')."""
    import torch
    import torch.nn.functional as F
    from seedvc_b200.length_regulator import nearest_index

    for n_in, n_out in [(60, 103), (47, 81), (40, 80), (50, 86), (1500, 2580), (2580, 1500), (7, 7), (3, 1000),
                        (999, 1000), (1000, 999), (1293, 2227)]:
        src = torch.arange(n_in, dtype=torch.float32).view(1, 1, n_in)
        want = F.interpolate(src, size=n_out, mode="nearest").view(-1).long().numpy()
        got = nearest_index(n_in, n_out)
        assert (got == want).all(), (n_in, n_out)


def _hift_case(name, m):
    import seedvc_oracle as orc
    mel = synth.synth_mel(m["B"], 80, m["Tm"], seed=m["mel_seed"])
    f0 = synth.synth_f0(m["B"], m["Tm"], seed=m["f0_seed"]) if m["f0_given"] else None
    phase, noise = synth.synth_hift_noise(m["B"], orc.HIFT_CFG["nb_harmonics"] + 1, m["Tm"] * 256, seed=m["noise_seed"])
    return mel, f0, phase, noise


def test_hift_host_logic():
    """HiFTGenerator's orchestration (weight-norm folding, polyphase upsamplers, stride-8 source conv as a
    regrouped 3-tap GEMM, reflection pad, epilogue-fused sums) on the emulated ops vs the REAL reference's outputs."""
    import json
    import os

    import numpy as np
    from seedvc_b200.hifigan import ConvRNNF0Predictor, HiFTGenerator

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hift.npz"))
    gen = HiFTGenerator(f0_predictor=ConvRNNF0Predictor())
    w = gen._build_weights(EmuOps())
    gen._prepare = lambda: w
    for name, m in json.loads(str(z["meta"])).items():
        mel, f0, phase, noise = _hift_case(name, m)
        wav = gen(mel, f0=f0, phase=phase, noise=noise)
        assert wav.shape == z[name + "_wav"].shape
        e = rel_l2(wav, z[name + "_wav"])
        assert e < 1e-4, (name, e)
        f0p = gen._predict_f0(w, w["ops"], mel)
        assert rel_l2(f0p, z[name + "_f0pred"]) < 1e-5


def test_hift_oracle_matches_reference(manifest):
    """oracle.hift_forward / hift_f0_predictor vs the REAL HiFTGenerator (oracle/gen_golden_hift.py)."""
    import json
    import os

    import numpy as np
    import seedvc_oracle as orc

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hift.npz"))
    sd = synth.synth_state_dict(manifest["keys_hift"])
    for name, m in json.loads(str(z["meta"])).items():
        mel, f0, phase, noise = _hift_case(name, m)
        wav = orc.hift_forward(sd, mel, phase, noise, f0=f0)
        assert rel_l2(wav, z[name + "_wav"]) < 2e-5, name
        assert rel_l2(orc.hift_f0_predictor(sd, mel), z[name + "_f0pred"]) < 2e-5


@pytest.mark.parametrize("fold", [False, True], ids=["norm_kernels", "folded_norms"])
@pytest.mark.parametrize("model,cfg", [("whisper_small", 0.7), ("xlsr_tiny", 0.7), ("whisper_base", 0.0), ("v2", (0.7, 0.7))])
def test_launches_per_step_bookkeeping(model, cfg, fold):
    """DiTEngine.launches_per_step (what the one-call C path adds to the launch counter) == the number of ops the
    written-out Python sequence issues for one estimator call."""
    T, Tp = 40, 10
    if model == "v2":
        kw = configs.v2_estimator_kwargs()
        kw.update(hidden_dim=128, num_heads=2, depth=3, content_dim=128)
        cfm = CFMv2(DiTv2(**kw))
        C, cd = kw["in_channels"], kw["content_dim"]
    else:
        args = configs.scaled_down(configs.v1_model_params(model))
        cfm = CFM(args)
        C, cd = args.DiT.in_channels, args.DiT.content_dim
    eng = emu_engine(cfm.estimator, fold)
    if model != "v2":
        cfm.estimator.setup_caches(1, 8192)
    counts = []
    orig = eng.step

    def counting(s, x_op):
        n0 = eng.ops.launches
        v = orig(s, x_op)
        counts.append(eng.ops.launches - n0)
        return v

    eng.step = counting
    mu, prompt, style, z = synth.synth_batch(2, T, Tp, C, cd)
    t_span = torch.linspace(0, 1, 3)
    if model == "v2":
        cfm.solve_euler(z, torch.tensor([T, T]), prompt, mu, style, t_span, list(cfg), False)
    else:
        cfm.solve_euler(z, torch.tensor([T, T]), prompt, mu, style, None, t_span, cfg)
    assert counts and all(c == eng.launches_per_step() for c in counts), (counts, eng.launches_per_step())
