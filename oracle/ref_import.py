"""TEST INFRASTRUCTURE ONLY - imports the *real* reference modules from /root/reference.

Only usable in the authoring container (``/root/reference`` does not exist on the
GPU box).  Used by ``oracle/gen_golden.py`` to produce the fixtures under
``tests/golden/`` and by CPU tests that pin ``oracle/seedvc_oracle.py`` to the
reference directly when the tree is present.

The reference imports three packages that are not installed here and that the
hot path never calls: ``munch`` (modules/commons.py:6),
``matplotlib`` (modules/bigvgan/utils.py:7-11) and ``librosa``
(modules/bigvgan/meldataset.py:13-15).  Tiny stand-ins are put in
``sys.modules`` before importing.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("SEEDVC_REFERENCE", "/root/reference")
# staged copy of the hot-path files (oracle/build_ref.py): what the GPU box has instead of /root/reference
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
if not os.path.isdir(os.path.join(REF_ROOT, "modules")) and os.path.isdir(os.path.join(STAGED_ROOT, "modules")):
    REF_ROOT = STAGED_ROOT


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "modules"))


def kind() -> str:
    """'reference' (full tree), 'oracle/_ref' (staged copy) or 'none'."""
    if not available():
        return "none"
    return "oracle/_ref" if os.path.abspath(REF_ROOT) == os.path.abspath(STAGED_ROOT) else "reference"


def _install_stubs():
    if "munch" not in sys.modules:
        m = types.ModuleType("munch")

        class Munch(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

            def __setattr__(self, k, v):
                self[k] = v

        m.Munch = Munch
        sys.modules["munch"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        pylab = types.ModuleType("matplotlib.pylab")
        pyplot = types.ModuleType("matplotlib.pyplot")
        mpl.pylab, mpl.pyplot = pylab, pyplot
        sys.modules.update({"matplotlib": mpl, "matplotlib.pylab": pylab,
                            "matplotlib.pyplot": pyplot})
    if "dac" not in sys.modules:      # modules/length_regulator.py:7 imports the (unused here) VQ layer
        dac = types.ModuleType("dac")
        dnn = types.ModuleType("dac.nn")
        dq = types.ModuleType("dac.nn.quantize")

        class VectorQuantize:           # never instantiated: the v1 presets set vector_quantize: false
            def __init__(self, *a, **k):
                raise RuntimeError("dac is not installed")

        dq.VectorQuantize = VectorQuantize
        dac.nn, dnn.quantize = dnn, dq
        sys.modules.update({"dac": dac, "dac.nn": dnn, "dac.nn.quantize": dq})
    if "librosa" not in sys.modules:
        lb = types.ModuleType("librosa")
        util = types.ModuleType("librosa.util")
        util.normalize = lambda x, *a, **k: x
        filters = types.ModuleType("librosa.filters")
        def _mel(*a, **k):      # modules/audio.py:4 - librosa itself is absent: the oracle's restatement
            import seedvc_oracle
            return seedvc_oracle.slaney_mel_filterbank(*a, **k)

        filters.mel = _mel
        lb.util, lb.filters = util, filters
        sys.modules.update({"librosa": lb, "librosa.util": util,
                            "librosa.filters": filters})


def load():
    """Return a namespace with the reference classes of the hot path."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings

    warnings.filterwarnings("ignore", category=FutureWarning)
    warnings.filterwarnings("ignore", category=UserWarning)
    from modules.flow_matching import CFM  # noqa: E402
    from modules.diffusion_transformer import DiT  # noqa: E402
    from modules.bigvgan.bigvgan import BigVGAN  # noqa: E402
    from modules.bigvgan.env import AttrDict as BigVGANAttrDict  # noqa: E402
    from modules.bigvgan.alias_free_activation.torch.act import Activation1d  # noqa: E402
    from modules.bigvgan.activations import SnakeBeta, Snake  # noqa: E402
    from modules.v2.cfm import CFM as CFMv2  # noqa: E402
    from modules.v2.dit_wrapper import DiT as DiTv2  # noqa: E402
    from modules.wavenet import WN  # noqa: E402
    from modules.length_regulator import InterpolateRegulator  # noqa: E402
    from modules.audio import mel_spectrogram  # noqa: E402
    from modules.v2.length_regulator import InterpolateRegulator as InterpolateRegulatorV2  # noqa: E402
    from modules.hifigan.generator import HiFTGenerator  # noqa: E402
    from modules.hifigan.f0_predictor import ConvRNNF0Predictor  # noqa: E402
    import modules.hifigan.generator as hift_module  # noqa: E402

    ns = types.SimpleNamespace(
        CFM=CFM, DiT=DiT, BigVGAN=BigVGAN, BigVGANAttrDict=BigVGANAttrDict,
        Activation1d=Activation1d, SnakeBeta=SnakeBeta, Snake=Snake,
        CFMv2=CFMv2, DiTv2=DiTv2, WN=WN, InterpolateRegulator=InterpolateRegulator, mel_spectrogram=mel_spectrogram, InterpolateRegulatorV2=InterpolateRegulatorV2,
        HiFTGenerator=HiFTGenerator, ConvRNNF0Predictor=ConvRNNF0Predictor, hift_module=hift_module,
    )
    return ns
