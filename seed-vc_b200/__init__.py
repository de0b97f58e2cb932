"""seedvc_b200 - B200-native implementation of Seed-VC's conversion hot path.

Public surface mirrors the reference:

* ``CFM(args)`` with ``inference`` / ``solve_euler`` / ``estimator.setup_caches``
  (reference: modules/flow_matching.py:30-112,159-167)
* ``CFMv2(estimator)`` / ``DiTv2(**kw)`` (reference: modules/v2/cfm.py, dit_wrapper.py)
* ``BigVGAN(h)`` (reference: modules/bigvgan/bigvgan.py:266-386)

All compute runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/seedvc_b200.h`` (``csrc/`` -> ``libseedvc_b200.so``).  There is no CPU
fallback: using a module without the library or without a CUDA device raises.
"""
from . import configs, synth  # noqa: F401

__all__ = ["configs", "synth", "CFM", "CFMv2", "DiTv2", "BigVGAN", "load_library"]


def __getattr__(name):
    if name in ("CFM", "DiT"):
        from . import flow_matching as _m
        return getattr(_m, name)
    if name in ("CFMv2", "DiTv2"):
        from . import flow_matching_v2 as _m
        return getattr(_m, name)
    if name == "BigVGAN":
        from . import bigvgan as _m
        return _m.BigVGAN
    if name == "load_library":
        from . import _lib
        return _lib.load_library
    raise AttributeError(name)
