"""Streaming block stitching (SOLA) for many concurrent streams (streaming half).

Mirror of the reference's per-stream state and arithmetic in ``real-time-gui.py``: the kept tail
``sola_buffer`` (:919-921), the ``sin^2`` fade windows (:929-943) and the per-tick search / cut / crossfade
(:1103-1137).  The reference runs it for one stream with two ``F.conv1d`` calls, an ``argmax().item()`` host
round trip and in-place slicing; here B streams are one kernel launch (``svc_sola_stitch``) and the offsets
stay on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from .ops import Ops


class SolaStitcher:
    def __init__(self, n_streams, sola_buffer_frame, sola_search_frame, block_frame, device="cuda"):
        self.B, self.sb, self.search, self.block = n_streams, sola_buffer_frame, sola_search_frame, block_frame
        dev = torch.device(device)
        self.sola_buffer = torch.zeros(n_streams, sola_buffer_frame, dtype=torch.float32, device=dev)   # :919-921
        # windows evaluated on the host (the reference evaluates them once at start-up on its device, :929-943);
        # host evaluation makes them identical on every platform
        fi = torch.sin(0.5 * np.pi * torch.linspace(0.0, 1.0, steps=sola_buffer_frame, dtype=torch.float32)) ** 2
        self.fade_in_window = fi.to(dev)
        self.fade_out_window = (1 - fi).to(dev)
        self.offsets = torch.zeros(n_streams, dtype=torch.int32, device=dev)
        self.ops = Ops("fp32")

    @torch.no_grad()
    def step(self, infer_wav):
        """infer_wav (B, >= search + block + sola_buffer_frame) fp32 on the device -> (B, block_frame);
        ``self.offsets`` holds the chosen SOLA offsets, ``self.sola_buffer`` the tails for the next tick."""
        if infer_wav.device.type != "cuda":
            raise RuntimeError("seedvc_b200 SolaStitcher runs on a CUDA (sm_100a) device only")
        x = infer_wav.to(torch.float32)
        if x.stride(1) != 1:
            x = x.contiguous()
        assert x.shape[0] == self.B
        out = torch.empty(self.B, self.block, dtype=torch.float32, device=x.device)
        ops = self.ops
        ops._t0("misc")
        from ._lib import check
        check(ops.lib.svc_sola_stitch(x.data_ptr(), x.stride(0), x.shape[1], self.sola_buffer.data_ptr(),
                                      self.sola_buffer.stride(0), self.fade_in_window.data_ptr(),
                                      self.fade_out_window.data_ptr(), out.data_ptr(), out.stride(0),
                                      self.offsets.data_ptr(), self.B, self.sb, self.search, self.block,
                                      ops._stream()), "svc_sola_stitch")
        ops._t1()
        return out
