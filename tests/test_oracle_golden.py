"""Pin oracle/seedvc_oracle.py to the committed outputs of the real reference
(tests/golden/*.npz, made by oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import seedvc_b200  # noqa: F401
from seedvc_b200 import configs, synth
import seedvc_oracle as orc
from conftest import load_golden, rel_l2

V1 = ["v1_small_scaled_cfg", "v1_small_scaled_nocfg", "v1_tiny_scaled_cfg", "v1_base_scaled_cfg",
      "v1_tiny_full", "v1_small_full", "v1_base_full"]
V2 = ["v2_small_3branch", "v2_small_spk_only", "v2_small_txt_only", "v2_small_nocfg",
      "v2_small_random_voice"]
TOL = 2e-5


def test_fir_taps_and_snake_kat():
    g = load_golden("snake_kat")
    h = orc.kaiser_sinc_filter12()
    assert np.allclose(h.numpy(), g["filter"], atol=1e-8)
    # taps quoted in SURVEY App. A.7
    assert abs(float(h[5]) - 0.4432097971) < 1e-7 and abs(float(h[0]) - 0.0020289647) < 1e-8
    y0 = orc.snake_aa(torch.tensor(g["x0"]), torch.zeros(4), torch.zeros(4))
    assert rel_l2(y0, g["y0"]) < 1e-6
    assert abs(float(y0[0, 0, 0]) - 0.0032857) < 1e-6        # SURVEY section 8c KAT row 0
    y1 = orc.snake_aa(torch.tensor(g["x1"]), torch.tensor(g["alpha"]), torch.tensor(g["beta"]))
    assert rel_l2(y1, g["y1"]) < 1e-6


@pytest.mark.parametrize("name", V1)
def test_v1_sampler_matches_reference(name, manifest):
    g = load_golden(name)
    m = g["meta"]
    args = configs.v1_model_params(m["model"])
    if m["scaled"]:
        args = configs.scaled_down(args)
    sd = synth.synth_state_dict(manifest[f"keys_{m['model']}{'_scaled' if m['scaled'] else ''}"])
    T, Tp = m["T"], m["Tp"]
    mu, prompt, style, z = synth.synth_batch(1, T, Tp, args.DiT.in_channels, args.DiT.content_dim)
    t_span = torch.linspace(0, 1, m["n_steps"] + 1)
    x0 = z.clone()
    px = torch.zeros_like(x0)
    px[..., :Tp] = prompt
    x0[..., :Tp] = 0
    v0 = orc.dit_v1_forward(sd, args, x0, px, torch.tensor([T]), t_span[0:1], style, mu)
    assert rel_l2(v0, g["v0"]) < TOL
    out = orc.solve_euler_v1(sd, args, z, torch.tensor([T]), prompt, mu, style, t_span, m["cfg"])
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("name", V2)
def test_v2_sampler_matches_reference(name, manifest):
    g = load_golden(name)
    m = g["meta"]
    kw = configs.v2_estimator_kwargs()
    sd = synth.synth_state_dict(manifest["keys_v2_small"])
    mu, prompt, style, z = synth.synth_batch(1, m["T"], m["Tp"], kw["in_channels"], kw["content_dim"])
    t_span = orc.v2_t_span(m["n_steps"])
    out = orc.solve_euler_v2(sd, kw, z, torch.tensor([m["T"]]), prompt, mu, style, t_span,
                             m["cfg"], m["random_voice"])
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("name", ["bigvgan_22k_t12", "bigvgan_22k_b2_t7", "bigvgan_44k_t6"])
def test_bigvgan_matches_reference(name, manifest):
    g = load_golden(name)
    m = g["meta"]
    h = configs.bigvgan_h(m["config"])
    sd = synth.synth_state_dict(manifest["keys_" + m["config"]])
    mel = synth.synth_mel(m["B"], h.num_mels, m["Tm"])
    wav = orc.bigvgan_forward(sd, h, mel)
    assert rel_l2(wav, g["wav"]) < TOL


def test_chunk_loop_against_reference_golden():
    """oracle.stitch_chunks / chunk_plan vs the REAL reference methods (oracle/gen_golden_chunks.py)."""
    import json
    import os

    import numpy as np
    import seedvc_oracle as orc

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "stitch_kat.npz"))
    meta = json.loads(str(z["meta"]))
    for name, m in meta.items():
        plan = orc.chunk_plan(m["S"], m["Tp"], m["max_context_window"], m["overlap_frame_len"])
        assert [list(p) for p in plan] == [list(p) for p in m["plan"]], name
        waves = [z[f"{name}_w{k}"] for k in range(len(plan))]
        got = orc.stitch_chunks(waves, m["overlap_frame_len"] * m["hop"])
        assert got.dtype == np.float32 and got.shape == z[name + "_out"].shape
        assert np.array_equal(got, z[name + "_out"]), name      # bit-exact
    short = orc.crossfade(z["xf_c1"].copy(), z["xf_c2"].copy(), 64)
    assert np.array_equal(short, z["xf_out"])


def test_length_regulator_oracle_against_reference_golden():
    """oracle.interpolate_regulator vs outputs of the REAL InterpolateRegulator (oracle/gen_golden_lr.py)."""
    import json
    import os

    import numpy as np
    import torch
    import gen_golden_lr as gl
    import seedvc_oracle as orc
    from seedvc_b200 import synth

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "length_regulator.npz"))
    meta = json.loads(str(z["meta"]))
    for name, m in meta.items():
        pre = "cfm_length_regulator." if m.get("v2") else "length_regulator."
        sd = synth.synth_state_dict({pre + k: v for k, v in m["keys"].items()})
        sd = {k[len(pre):]: v for k, v in sd.items()}
        if m.get("v2"):
            tok = gl.tokens(name, m["B"], m["Tin"], m["kw"]["codebook_size"])
            y = orc.interpolate_regulator_v2(sd, tok, torch.tensor(m["ylens"]), n_blocks=len(m["kw"]["sampling_ratios"]))
        else:
            x, f0 = gl.inputs(name, m["B"], m["Tin"], m["kw"]["in_channels"], Tf0=m["Tin"] + 3 if m["f0"] else None)
            y = orc.interpolate_regulator(sd, x, torch.tensor(m["ylens"]), f0=f0,
                                          f0_condition=m["kw"].get("f0_condition", False),
                                          n_f0_bins=m["kw"].get("n_f0_bins", 512))
        want = torch.from_numpy(z[name])
        assert y.shape == want.shape
        e = float((y - want).norm() / want.norm())
        assert e < 1e-5, (name, e)


def test_mel_frontend_oracle_against_reference_golden():
    """oracle.mel_spectrogram vs the REAL reference function (oracle/gen_golden_mel.py); and the package's
    Slaney filterbank vs the oracle's independent restatement (librosa itself is not available)."""
    import json
    import os

    import numpy as np
    import torch
    import gen_golden_mel as gm
    import seedvc_oracle as orc
    from seedvc_b200.audio import mel_filterbank

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "mel_kat.npz"))
    for name, m in json.loads(str(z["meta"])).items():
        y = gm.audio(name, m["B"], m["L"])
        got = orc.mel_spectrogram(y, **m["kw"])
        want = torch.from_numpy(z[name])
        assert got.shape == want.shape
        assert float((got - want).norm() / want.norm()) < 1e-6, name
    for sr, n_fft, n_mels, fmax in [(22050, 1024, 80, None), (44100, 2048, 128, None), (22050, 1024, 80, 8000)]:
        a = mel_filterbank(sr, n_fft, n_mels, 0, fmax)
        b = orc.slaney_mel_filterbank(sr, n_fft, n_mels, 0, fmax)
        assert a.shape == b.shape == (n_mels, n_fft // 2 + 1)
        assert np.abs(a - b).max() < 1e-7 * max(1.0, np.abs(b).max())
        assert (a.sum(1) > 0).all()


def test_sola_oracle_against_reference_golden():
    """oracle.sola_step vs the REAL reference lines of real-time-gui.py (oracle/gen_golden_sola.py)."""
    import json
    import os

    import numpy as np
    import torch
    import gen_golden_sola as gs
    import seedvc_oracle as orc

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "sola_kat.npz"))
    for name, m in json.loads(str(z["meta"])).items():
        for s in range(m["streams"]):
            st = gs.make_state(m["zc"], m["block"] // m["zc"], m["sb"] // m["zc"] if m["sb"] % m["zc"] == 0 else 2)
            buf = torch.zeros(m["sb"])
            for t in range(m["ticks"]):
                x = gs.tick_input(name, s, t, m["n"])
                out, buf, off = orc.sola_step(x, buf, st.fade_in_window, st.fade_out_window, m["sb"], m["search"],
                                              m["block"])
                assert off == int(z[f"{name}_t{t}_s{s}_off"])
                assert np.array_equal(out.numpy(), z[f"{name}_t{t}_s{s}_out"])
                assert np.array_equal(buf.numpy(), z[f"{name}_t{t}_s{s}_buf"])


# ---------------------------------------------------------------------------------------------
# full-size fixtures (oracle/gen_golden_full.py): the oracle at BASELINE frame counts.  Cases are sized so
# the CPU suite stays within minutes: whole solves where they take seconds, the first Euler step otherwise.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,steps", [("full_small_T323_n25", 25), ("full_tiny_T1291_n10", 10),
                                        ("full_small_T2580_n25", 1), ("full_base_T2580_n2", 1)])
def test_oracle_v1_full_size(name, steps, manifest):
    g = load_golden(name)
    m = g["meta"]
    args = configs.v1_model_params(m["model"])
    sd = synth.synth_state_dict(manifest["keys_" + m["model"]])
    T, Tp, N = m["T"], m["Tp"], m["n_steps"]
    mu, prompt, style, z = synth.synth_batch(1, T, Tp, args.DiT.in_channels, args.DiT.content_dim,
                                             first_id=m["utt_id"])
    t_span = torch.linspace(0, 1, N + 1)[:steps + 1]
    out, vs = orc.solve_euler_v1(sd, args, z, torch.tensor([T]), prompt, mu, style, t_span, m["cfg"],
                                 return_steps=True)
    fr = torch.from_numpy(g["frames"])
    for s in range(steps):
        assert rel_l2(vs[0][s][0][:, fr], g["v_steps"][s]) < 2e-5, (name, s)
    if steps == N:
        assert rel_l2(out[0][:, fr], g["out"]) < 2e-5


def test_oracle_v2_full_size_first_step(manifest):
    g = load_golden("full_v2_T2580_n2")
    m = g["meta"]
    kw = configs.v2_estimator_kwargs()
    sd = synth.synth_state_dict(manifest["keys_v2_small"])
    T, Tp = m["T"], m["Tp"]
    mu, prompt, style, z = synth.synth_batch(1, T, Tp, kw["in_channels"], kw["content_dim"], first_id=m["utt_id"])
    t_span = orc.v2_t_span(m["n_steps"])[:2]
    out = orc.solve_euler_v2(sd, kw, z, torch.tensor([T]), prompt, mu, style, t_span, m["cfg"])
    fr = torch.from_numpy(g["frames"])
    x0 = z.clone()
    x0[..., :Tp] = 0
    v0 = (out - x0) / float(t_span[1] - t_span[0])
    assert rel_l2(v0[0][:, fr], g["v_steps"][0]) < 5e-5


def test_oracle_bigvgan_full_size_256(manifest):
    g = load_golden("full_bigvgan22k_256")
    m = g["meta"]
    h = configs.bigvgan_h(m["config"])
    sd = synth.synth_state_dict(manifest["keys_" + m["config"]])
    wav = orc.bigvgan_forward(sd, h, synth.synth_mel(m["B"], h.num_mels, m["Tm"], seed=m["mel_seed"]))
    assert rel_l2(wav[:, :, torch.from_numpy(g["idx"])], g["wav"]) < 2e-5


@pytest.mark.parametrize("sr,n_fft,n_mels,fmin,fmax", [(22050, 1024, 80, 0, None), (44100, 2048, 128, 0, None),
                                                        (22050, 1024, 80, 0, 8000), (16000, 400, 80, 0, 8000)])
def test_mel_filterbank_pinned_to_torchaudio(sr, n_fft, n_mels, fmin, fmax):
    """SURVEY 8f N3 pin: librosa is not installed, so the Slaney filterbank that stands in for
    ``librosa.filters.mel`` (modules/audio.py:4,66) - the oracle's restatement AND the package's own - is pinned to an
    independent published implementation, torchaudio's ``melscale_fbanks(norm='slaney', mel_scale='slaney')``."""
    ta = pytest.importorskip("torchaudio")
    from seedvc_b200 import audio
    want = ta.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax if fmax else sr / 2), n_mels, sr,
                                         norm="slaney", mel_scale="slaney").t()
    got_o = torch.as_tensor(orc.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)).float()
    got_p = torch.as_tensor(audio.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)).float().cpu()
    assert got_o.shape == want.shape == got_p.shape
    assert float((got_o - want).abs().max()) < 2e-7 and rel_l2(got_o, want) < 1e-5
    assert float((got_p - want).abs().max()) < 2e-7 and rel_l2(got_p, want) < 1e-5
