"""Model hyper-parameters of the hot path, as attribute dictionaries.

The reference reads these from YAML presets through ``recursive_munch``
(reference: modules/commons.py:482-488, inference.py:71-73) and ``hasattr``
defaults (modules/diffusion_transformer.py:413-415).  The values below restate
the presets so the path can be built where ``/root/reference`` is absent:

* ``xlsr_tiny``      configs/presets/config_dit_mel_seed_uvit_xlsr_tiny.yml:57-79
* ``whisper_small``  configs/presets/config_dit_mel_seed_uvit_whisper_small_wavenet.yml:56-86
* ``whisper_base``   configs/presets/config_dit_mel_seed_uvit_whisper_base_f0_44k.yml:63-93
* ``v2_small``       configs/v2/vc_wrapper.yaml:15-31
* ``bigvgan_22k``    modules/bigvgan/config.json:11-18,44-52
* ``bigvgan_44k``    nvidia/bigvgan_v2_44khz_128band_512x (not in the tree)

``load_yaml_model_params`` accepts the reference's own preset files, so a
caller holding the reference tree can keep using them.
"""
from __future__ import annotations

import copy


class AttrDict(dict):
    """dict with attribute access (what ``Munch`` / bigvgan ``AttrDict`` give)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # hasattr() must see AttributeError
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(d):
    if isinstance(d, dict):
        return AttrDict({k: to_attr(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return type(d)(to_attr(v) for v in d)
    return d


def _dit(hidden, heads, depth, in_ch, content_dim, final, long_skip, tat, sat):
    return dict(
        hidden_dim=hidden, num_heads=heads, depth=depth, class_dropout_prob=0.1,
        block_size=8192, in_channels=in_ch, style_condition=True,
        final_layer_type=final, target="mel", content_dim=content_dim,
        content_codebook_size=1024, content_type="discrete", f0_condition=False,
        n_f0_bins=512, content_codebooks=1, is_causal=False,
        long_skip_connection=long_skip, zero_prompt_speech_token=False,
        time_as_token=tat, style_as_token=sat, uvit_skip_connection=True,
        add_resblock_in_transformer=False,
    )


def _wavenet(hidden):
    return dict(hidden_dim=hidden, num_layers=8, kernel_size=5, dilation_rate=1,
                p_dropout=0.2, style_condition=True)


_V1 = {
    "xlsr_tiny": dict(
        dit_type="DiT", reg_loss_type="l1", style_encoder=dict(dim=192),
        DiT=_dit(384, 6, 9, 80, 384, "mlp", False, True, True),
    ),
    "whisper_small": dict(
        dit_type="DiT", reg_loss_type="l1", style_encoder=dict(dim=192),
        DiT=_dit(512, 8, 13, 80, 512, "wavenet", True, False, False),
        wavenet=_wavenet(512),
    ),
    "whisper_base": dict(
        dit_type="DiT", reg_loss_type="l1", style_encoder=dict(dim=192),
        DiT=dict(_dit(768, 12, 17, 128, 768, "mlp", False, False, False),
                 f0_condition=True, n_f0_bins=256),
        wavenet=_wavenet(768),
    ),
}

# v2 estimator keyword arguments (configs/v2/vc_wrapper.yaml:15-31)
_V2 = {
    "v2_small": dict(
        time_as_token=True, style_as_token=True, uvit_skip_connection=False,
        block_size=8192, depth=13, num_heads=8, hidden_dim=512, in_channels=80,
        content_dim=512, style_encoder_dim=192, class_dropout_prob=0.1,
        dropout_rate=0.0, attn_dropout_rate=0.0,
    ),
}

_BIGVGAN = {
    "bigvgan_22k": dict(
        resblock="1", upsample_rates=[4, 4, 2, 2, 2, 2],
        upsample_kernel_sizes=[8, 8, 4, 4, 4, 4], upsample_initial_channel=1536,
        resblock_kernel_sizes=[3, 7, 11],
        resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        use_tanh_at_final=False, use_bias_at_final=False, activation="snakebeta",
        snake_logscale=True, num_mels=80, n_fft=1024, hop_size=256, win_size=1024,
        sampling_rate=22050, fmin=0, fmax=None,
    ),
    "bigvgan_44k": dict(
        resblock="1", upsample_rates=[8, 4, 2, 2, 2, 2],
        upsample_kernel_sizes=[16, 8, 4, 4, 4, 4], upsample_initial_channel=1536,
        resblock_kernel_sizes=[3, 7, 11],
        resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        use_tanh_at_final=False, use_bias_at_final=False, activation="snakebeta",
        snake_logscale=True, num_mels=128, n_fft=2048, hop_size=512, win_size=2048,
        sampling_rate=44100, fmin=0, fmax=None,
    ),
}


def v1_model_params(name: str) -> AttrDict:
    """``args`` for ``CFM(args)`` (reference: modules/flow_matching.py:160-167)."""
    return to_attr(copy.deepcopy(_V1[name]))


def v2_estimator_kwargs(name: str = "v2_small") -> dict:
    """kwargs for the v2 ``DiT(**kw)`` (reference: modules/v2/dit_wrapper.py:58-73)."""
    return copy.deepcopy(_V2[name])


def bigvgan_h(name: str = "bigvgan_22k") -> AttrDict:
    """``h`` for ``BigVGAN(h)`` (reference: modules/bigvgan/bigvgan.py:266-269)."""
    return to_attr(copy.deepcopy(_BIGVGAN[name]))


def scaled_down(params: AttrDict, hidden=128, heads=2, depth=5, wn_layers=2) -> AttrDict:
    """A structurally identical but small v1 model for fast parity cases."""
    p = to_attr(copy.deepcopy(dict(params)))
    p.DiT.hidden_dim = hidden
    p.DiT.num_heads = heads
    p.DiT.depth = depth
    p.DiT.content_dim = hidden
    if "wavenet" in p:
        p.wavenet.hidden_dim = hidden
        p.wavenet.num_layers = wn_layers
    return p


def load_yaml_model_params(path: str) -> AttrDict:
    """Read ``model_params`` from one of the reference's preset YAML files."""
    import yaml

    with open(path) as f:
        return to_attr(yaml.safe_load(f)["model_params"])


V1_NAMES = tuple(_V1)
V2_NAMES = tuple(_V2)
BIGVGAN_NAMES = tuple(_BIGVGAN)
