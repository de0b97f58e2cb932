"""End-to-end parity on the B200 through the package's public API (which calls the C ABI):

* every golden fixture (outputs of the REAL reference, tests/golden/) in fp32 mode
  (rel-L2 <= 1e-3, north_star) and bf16 mode (rel-L2 <= 1e-2), per estimator call (v0) and
  after the whole Euler solve;
* the oracle on seeded inputs at sizes it finishes in seconds;
* size-independent properties at BASELINE config-2 frame counts (T = 2580): batch invariance,
  zeroed prompt region, determinism.
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
V1 = ["v1_small_scaled_cfg", "v1_small_scaled_nocfg", "v1_tiny_scaled_cfg", "v1_base_scaled_cfg",
      "v1_tiny_full", "v1_small_full", "v1_base_full"]
V2 = ["v2_small_3branch", "v2_small_spk_only", "v2_small_txt_only", "v2_small_nocfg",
      "v2_small_random_voice"]

_models = {}


def v1_model(model, scaled, mode):
    key = (model, scaled)
    if key not in _models:
        args = configs.v1_model_params(model)
        if scaled:
            args = configs.scaled_down(args)
        cfm = CFM(args).to(DEV)
        cfm.estimator.setup_caches(1, 8192)
        _models[key] = (cfm, args)
    cfm, args = _models[key]
    cfm.set_mode(mode)
    return cfm, args


@pytest.mark.parametrize("mode", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("name", V1)
def test_v1_golden(name, mode):
    g = load_golden(name)
    m = g["meta"]
    cfm, args = v1_model(m["model"], m["scaled"], mode)
    T, Tp = m["T"], m["Tp"]
    mu, prompt, style, z = [t.to(DEV) for t in
                            synth.synth_batch(1, T, Tp, args.DiT.in_channels, args.DiT.content_dim)]
    t_span = torch.linspace(0, 1, m["n_steps"] + 1, device=DEV)
    x0 = z.clone()
    px = torch.zeros_like(x0)
    px[..., :Tp] = prompt
    x0[..., :Tp] = 0
    xl = torch.tensor([T], device=DEV)
    v0 = cfm.estimator(x0, px, xl, t_span[0:1], style, mu)
    e_v = rel_l2(v0.cpu(), g["v0"])
    out = cfm.solve_euler(z.clone(), xl, prompt, mu, style, None, t_span, m["cfg"])
    e_o = rel_l2(out.cpu(), g["out"])
    print(f"{name} [{mode}] velocity rel-L2 {e_v:.2e}  end-to-end rel-L2 {e_o:.2e}")
    assert e_v < TOL[mode] and e_o < TOL[mode]


@pytest.mark.parametrize("mode", ["fp16", "bf16"])
@pytest.mark.parametrize("name", ["v1_small_full", "v1_tiny_full", "v1_base_scaled_cfg"])
def test_v1_golden_folded_norms(name, mode):
    """The optional folded-RMS-norm path (Ops(fold_norms=True): svc_gemm row_ss_out / row_ss_in, DiTEngine._fold_begin)
    against the same reference goldens; off by default because it measures slower (see Ops.fold_norms)."""
    g = load_golden(name)
    m = g["meta"]
    cfm, args = v1_model(m["model"], m["scaled"], mode)
    est = cfm.estimator
    est.fold_norms = True
    try:
        assert est.engine().fold
        T, Tp = m["T"], m["Tp"]
        mu, prompt, style, z = [t.to(DEV) for t in
                                synth.synth_batch(1, T, Tp, args.DiT.in_channels, args.DiT.content_dim)]
        t_span = torch.linspace(0, 1, m["n_steps"] + 1, device=DEV)
        xl = torch.tensor([T], device=DEV)
        out = cfm.solve_euler(z.clone(), xl, prompt, mu, style, None, t_span, m["cfg"])
        e_o = rel_l2(out.cpu(), g["out"])
        print(f"{name} [{mode}, folded norms] end-to-end rel-L2 {e_o:.2e}")
        assert e_o < TOL[mode]
    finally:
        est.fold_norms = False
    assert not est.engine().fold


@pytest.mark.parametrize("mode", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("name", V2)
def test_v2_golden(name, mode):
    g = load_golden(name)
    m = g["meta"]
    if "v2" not in _models:
        _models["v2"] = CFMv2(DiTv2(**configs.v2_estimator_kwargs())).to(DEV)
    cfm = _models["v2"]
    cfm.set_mode(mode)
    kw = configs.v2_estimator_kwargs()
    mu, prompt, style, z = [t.to(DEV) for t in
                            synth.synth_batch(1, m["T"], m["Tp"], kw["in_channels"], kw["content_dim"])]
    t_span = torch.linspace(0, 1, m["n_steps"] + 1, device=DEV)
    t_span = t_span + (-1) * (torch.cos(torch.pi / 2 * t_span) - 1 + t_span)
    out = cfm.solve_euler(z.clone(), torch.tensor([m["T"]], device=DEV), prompt, mu, style, t_span,
                          m["cfg"], m["random_voice"])
    e = rel_l2(out.cpu(), g["out"])
    print(f"{name} [{mode}] end-to-end rel-L2 {e:.2e}")
    assert e < TOL[mode]


@pytest.mark.parametrize("mode", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("name", ["bigvgan_22k_t12", "bigvgan_22k_b2_t7", "bigvgan_44k_t6"])
def test_bigvgan_golden(name, mode):
    g = load_golden(name)
    m = g["meta"]
    key = "voc" if m["config"] == "bigvgan_22k" else "voc44"
    if key not in _models:
        _models[key] = BigVGAN(configs.bigvgan_h(m["config"])).to(DEV)
    voc = _models[key]
    voc.set_mode(mode)
    mel = synth.synth_mel(m["B"], voc.h.num_mels, m["Tm"]).to(DEV)
    wav = voc(mel)
    assert wav.shape == g["wav"].shape
    e = rel_l2(wav.cpu(), g["wav"])
    print(f"{name} [{mode}] waveform rel-L2 {e:.2e}")
    assert e < TOL[mode]


@pytest.mark.parametrize("mode", ["fp32", "fp16", "bf16"])
def test_v1_against_oracle_ragged_batch(mode, manifest):
    """Seeded batch of 3 utterances of different lengths vs the oracle's per-utterance runs."""
    import seedvc_oracle as orc

    cfm, args = v1_model("whisper_small", True, mode)
    sd = synth.synth_state_dict(manifest["keys_whisper_small_scaled"])
    T, Tp = 300, 41
    mu, prompt, style, z = synth.synth_batch(3, T, Tp, 80, args.DiT.content_dim, first_id=50)
    lens = torch.tensor([300, 257, 129])
    t_span = torch.linspace(0, 1, 4)
    want = orc.solve_euler_v1(sd, args, z, lens, prompt, mu, style, t_span, 0.7)
    out = cfm.solve_euler(z.to(DEV), lens.to(DEV), prompt.to(DEV), mu.to(DEV), style.to(DEV), None,
                          t_span.to(DEV), 0.7).cpu()
    for b in range(3):
        n = int(lens[b])
        e = rel_l2(out[b, :, :n], want[b, :, :n])
        print(f"ragged[{b}] len {n} [{mode}] rel-L2 {e:.2e}")
        assert e < TOL[mode]
        assert float(out[b, :, n:].abs().max() if n < T else 0.0) == 0.0
        assert float(out[b, :, :Tp].abs().max()) == 0.0


def test_bigvgan_against_oracle_longer(manifest):
    import seedvc_oracle as orc

    if "voc" not in _models:
        _models["voc"] = BigVGAN(configs.bigvgan_h()).to(DEV)
    voc = _models["voc"]
    sd = synth.synth_state_dict(manifest["keys_bigvgan_22k"])
    mel = synth.synth_mel(1, 80, 65, seed=11)       # one streaming block (config 5: 65 frames)
    want = orc.bigvgan_forward(sd, voc.h, mel)
    for mode in ("fp32", "bf16"):
        voc.set_mode(mode)
        wav = voc(mel.to(DEV)).cpu()
        e = rel_l2(wav, want)
        print(f"bigvgan 65 frames [{mode}] rel-L2 {e:.2e}")
        assert e < TOL[mode]


def test_full_length_properties():
    """T = 2580 (30 s context, BASELINE config 2): utterance results do not depend on what else is
    in the batch, the prompt region stays zero, and two runs are bit-identical."""
    cfm, args = v1_model("whisper_small", False, "bf16")
    T, Tp, B = 2580, 430, 3
    mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, 80, 512, first_id=7)]
    lens = torch.full((B,), T, device=DEV)
    t_span = torch.linspace(0, 1, 3, device=DEV)
    out = cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, t_span, 0.7)
    out2 = cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, t_span, 0.7)
    assert torch.equal(out, out2)
    assert torch.isfinite(out).all()
    assert float(out[:, :, :Tp].abs().max()) == 0.0
    one = cfm.solve_euler(z[1:2].clone(), lens[1:2], prompt[1:2], mu[1:2], style[1:2], None, t_span, 0.7)
    assert rel_l2(out[1:2].cpu(), one.cpu()) < 1e-6


def test_inference_api_and_cpu_refusal():
    cfm, args = v1_model("xlsr_tiny", True, "bf16")
    mu, prompt, style, _ = [t.to(DEV) for t in synth.synth_batch(2, 60, 20, 80, args.DiT.content_dim)]
    torch.manual_seed(3)
    out = cfm.inference(mu, torch.tensor([60, 60], device=DEV), prompt, style, None, 4,
                        inference_cfg_rate=0.7)
    assert out.shape == (2, 80, 60) and torch.isfinite(out).all()
    cpu = CFM(configs.scaled_down(configs.v1_model_params("xlsr_tiny")))
    with pytest.raises(RuntimeError):
        cpu.inference(mu.cpu(), torch.tensor([60, 60]), prompt.cpu(), style.cpu(), None, 2)


def test_long_form_chunks_match_sequential_loop():
    """one batched ragged conversion + GPU stitching == the reference's sequential
    window loop (run here with the same CUDA modules per window, stitched by the oracle)."""
    import seedvc_oracle as orc
    from seedvc_b200.chunking import chunk_plan, convert_chunks

    cfm, args = v1_model("whisper_small", True, "fp32")
    if "voc" not in _models:
        _models["voc"] = BigVGAN(configs.bigvgan_h()).to(DEV)
    voc = _models["voc"]
    voc.set_mode("fp32")
    S, Tp, mcw, hop, steps, cfg = 150, 24, 84, 256, 2, 0.7
    D = args.DiT.content_dim
    g = torch.Generator().manual_seed(4)
    cond = torch.randn(1, S, D, generator=g).to(DEV)
    prompt_cond = torch.randn(1, Tp, D, generator=g).to(DEV)
    mel2 = (torch.randn(1, 80, Tp, generator=g) * 2 - 4).to(DEV)
    style2 = torch.randn(1, 192, generator=g).to(DEV)
    plan = chunk_plan(S, Tp, mcw)
    assert len(plan) == 4 and plan[-1][1] < plan[0][1]
    Tmax = Tp + plan[0][1]
    z = torch.randn(len(plan), 80, Tmax, generator=g).to(DEV)
    wave = convert_chunks(cfm, voc, cond, prompt_cond, mel2, style2, steps, cfg, mcw, hop=hop, z=z)
    # the reference's loop: one window at a time
    t_span = torch.linspace(0, 1, steps + 1, device=DEV)
    waves = []
    for k, (start, length, _) in enumerate(plan):
        cat = torch.cat([prompt_cond, cond[:, start:start + length]], dim=1)
        T = Tp + length
        mel = cfm.solve_euler(z[k:k + 1, :, :T].clone(), torch.tensor([T], device=DEV), mel2, cat, style2,
                              None, t_span, cfg)
        waves.append(voc(mel[:, :, Tp:].contiguous())[0, 0].cpu().numpy())
    want = orc.stitch_chunks(waves, 16 * hop)
    assert wave.shape == (1, want.shape[0])
    e = rel_l2(wave[0].cpu(), torch.from_numpy(want))
    print(f"long-form {len(plan)} windows vs sequential loop [fp32] rel-L2 {e:.2e}")
    assert e < 1e-3


@pytest.mark.parametrize("mode", ["fp32", "fp16", "bf16"])
def test_length_regulator_golden(mode):
    """InterpolateRegulator on the CUDA kernels vs the REAL reference module's outputs."""
    import json
    import os

    import numpy as np
    import gen_golden_lr as gl
    from seedvc_b200.length_regulator import InterpolateRegulator

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "length_regulator.npz"))
    for name, m in json.loads(str(z["meta"])).items():
        v2 = bool(m.get("v2"))
        pre = "cfm_length_regulator." if v2 else "length_regulator."
        lr = InterpolateRegulator(**m["kw"], mode=mode, v2=v2)
        sd = synth.synth_state_dict({pre + k: v for k, v in m["keys"].items()})
        lr.load_state_dict({k[len(pre):]: v for k, v in sd.items()}, strict=True)    # the reference's own keys
        lr = lr.to(DEV)
        if v2:
            tok = gl.tokens(name, m["B"], m["Tin"], m["kw"]["codebook_size"]).to(DEV)
            y, olens = lr(tok, ylens=torch.tensor(m["ylens"], device=DEV), f0=None)
        else:
            x, f0 = gl.inputs(name, m["B"], m["Tin"], m["kw"]["in_channels"], Tf0=m["Tin"] + 3 if m["f0"] else None)
            y, olens, *_ = lr(x.to(DEV), ylens=torch.tensor(m["ylens"], device=DEV), n_quantizers=3,
                              f0=None if f0 is None else f0.to(DEV))
        e = rel_l2(y.cpu(), z[name])
        print(f"length regulator {name} [{mode}] rel-L2 {e:.2e}")
        assert tuple(y.shape) == z[name].shape and e < TOL[mode]
        assert [int(v) for v in olens] == m["ylens"]


def test_mel_frontend_golden():
    """mel_spectrogram as reflect pad + segmented-GEMM STFT + mel GEMM vs the REAL reference
    function's outputs (torch.stft path); fp32 arithmetic, tolerance 1e-3 like the fp32 mode of the path."""
    import json
    import os

    import numpy as np
    import gen_golden_mel as gm
    from seedvc_b200.audio import mel_spectrogram

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mel_kat.npz"))
    for name, m in json.loads(str(z["meta"])).items():
        y = gm.audio(name, m["B"], m["L"]).to(DEV)
        got = mel_spectrogram(y, **m["kw"])
        e = rel_l2(got.cpu(), z[name])
        print(f"mel front-end {name} rel-L2 {e:.2e} max|d| {float((got.cpu() - torch.from_numpy(z[name])).abs().max()):.2e}")
        assert tuple(got.shape) == z[name].shape and e < 1e-3
    with pytest.raises(RuntimeError):
        mel_spectrogram(torch.zeros(1, 4096), 1024, 80, 22050, 256, 1024, 0, None)


@pytest.mark.parametrize("mode", ["fp32", "fp16", "bf16"])
def test_hift_golden(mode):
    """HiFTGenerator on the CUDA kernels vs the REAL reference's outputs (oracle/gen_golden_hift.py),
    SineGen's phase / noise draws injected; the F0 predictor separately."""
    import json
    import os

    import numpy as np
    from seedvc_b200.hifigan import ConvRNNF0Predictor, HiFTGenerator

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "hift.npz"))
    if "hift" not in _models:
        _models["hift"] = HiFTGenerator(f0_predictor=ConvRNNF0Predictor()).to(DEV)
    gen = _models["hift"]
    gen.set_mode(mode)
    for name, m in json.loads(str(z["meta"])).items():
        mel = synth.synth_mel(m["B"], 80, m["Tm"], seed=m["mel_seed"]).to(DEV)
        f0 = synth.synth_f0(m["B"], m["Tm"], seed=m["f0_seed"]).to(DEV) if m["f0_given"] else None
        phase, noise = [t.to(DEV) for t in synth.synth_hift_noise(m["B"], 9, m["Tm"] * 256, seed=m["noise_seed"])]
        wav = gen(mel, f0=f0, phase=phase, noise=noise)
        e = rel_l2(wav.cpu(), z[name + "_wav"])
        w = gen._prepare()
        e_f0 = rel_l2(gen._predict_f0(w, w["ops"], mel).cpu(), z[name + "_f0pred"])
        print(f"hift {name} [{mode}] waveform rel-L2 {e:.2e}  f0 predictor rel-L2 {e_f0:.2e}")
        # 16-bit modes: the north_star tolerance (HiFT runs on IEEE-half operands in both, see hifigan.py)
        assert tuple(wav.shape) == z[name + "_wav"].shape and e < (1e-3 if mode == "fp32" else 1e-2)
        assert e_f0 < {"fp32": 1e-5, "fp16": 3e-3, "bf16": 3e-3}[mode]
    out = gen.inference(synth.synth_mel(1, 80, 20).to(DEV))          # own phase / noise draws, predictor F0
    assert out.shape == (1, 20 * 256) and torch.isfinite(out).all()


@pytest.mark.parametrize("shape", [(1, 323, 258, 4), (3, 200, 60, 3)])
def test_graphed_conversion_bit_exact(shape):
    """graphs.GraphedConversion (one CUDA-graph replay per conversion: the launch-bound config 1 / config 5 shapes)
    returns bit-exactly what the eager launch sequence returns, also after the inputs change between replays."""
    from seedvc_b200.graphs import GraphedConversion

    B, T, Tp, steps = shape
    cfm, args = v1_model("xlsr_tiny", False, "bf16")
    if "voc" not in _models:
        _models["voc"] = BigVGAN(configs.bigvgan_h()).to(DEV)
    voc = _models["voc"]
    voc.set_mode("bf16")
    g = GraphedConversion(cfm, voc, B, T, Tp, steps, 0.7)
    t_span = torch.linspace(0, 1, steps + 1, device=DEV)
    lens = torch.full((B,), T, device=DEV)
    for first_id in (0, 40, 41):
        mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, 80, args.DiT.content_dim,
                                                                      first_id=first_id)]
        want = voc(cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, t_span, 0.7)[:, :, Tp:].contiguous())
        got = g(mu, lens, prompt, style, z)
        assert torch.equal(got, want), first_id


@pytest.mark.parametrize("model,mode", [("whisper_small", "bf16"), ("whisper_small", "fp32"), ("xlsr_tiny", "fp16"),
                                        ("whisper_base", "bf16")])
def test_dit_step_c_entry_point_equals_python_sequence(model, mode):
    """svc_dit_step (one C call per estimator call, csrc/graph.cu) issues the launch sequence DiTEngine.step writes
    out in Python: the two give bit-identical results (the Python sequence runs when per-launch profiling is on)."""
    cfm, args = v1_model(model, True, mode)
    T, Tp, B = 140, 33, 2
    mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(B, T, Tp, args.DiT.in_channels,
                                                                  args.DiT.content_dim, first_id=21)]
    lens = torch.tensor([T, 101], device=DEV)
    t_span = torch.linspace(0, 1, 4, device=DEV)
    ops = cfm.estimator.engine().ops
    n0 = ops.launches
    a = cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, t_span, 0.7)
    n_c = ops.launches - n0
    ops.start_profile()
    n0 = ops.launches
    b = cfm.solve_euler(z.clone(), lens, prompt, mu, style, None, t_span, 0.7)
    n_py = ops.launches - n0
    prof = ops.stop_profile()
    assert torch.equal(a, b)
    assert n_c == n_py and sum(d["launches"] for d in prof.values()) == n_py


def test_dit_step_c_entry_point_v2():
    if "v2" not in _models:
        _models["v2"] = CFMv2(DiTv2(**configs.v2_estimator_kwargs())).to(DEV)
    cfm = _models["v2"]
    cfm.set_mode("bf16")
    kw = configs.v2_estimator_kwargs()
    mu, prompt, style, z = [t.to(DEV) for t in synth.synth_batch(2, 90, 25, kw["in_channels"], kw["content_dim"])]
    t_span = torch.linspace(0, 1, 3, device=DEV)
    lens = torch.tensor([90, 77], device=DEV)
    ops = cfm.estimator.engine().ops
    a = cfm.solve_euler(z.clone(), lens, prompt, mu, style, t_span, [0.7, 0.7], False)
    ops.start_profile()
    b = cfm.solve_euler(z.clone(), lens, prompt, mu, style, t_span, [0.7, 0.7], False)
    ops.stop_profile()
    assert torch.equal(a, b)


@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_bigvgan_c_entry_point_equals_python_sequence(mode):
    """svc_bigvgan_forward (one C call, csrc/graph.cu) == the launch sequence BigVGAN.forward writes out in Python
    (which runs when per-launch profiling is on), bit for bit, for both vocoder configs."""
    for key, cfgname, Tm in (("voc", "bigvgan_22k", 23), ("voc44", "bigvgan_44k", 9)):
        if key not in _models:
            _models[key] = BigVGAN(configs.bigvgan_h(cfgname)).to(DEV)
        voc = _models[key]
        voc.set_mode(mode)
        mel = synth.synth_mel(2, voc.h.num_mels, Tm, seed=5).to(DEV)
        ops = voc._prepare()["ops"]
        n0 = ops.launches
        a = voc(mel)
        n_c = ops.launches - n0
        ops.start_profile()
        n0 = ops.launches
        b = voc(mel)
        n_py = ops.launches - n0
        ops.stop_profile()
        assert voc._prepare()["c"] is not None
        assert torch.equal(a, b) and n_c == n_py
