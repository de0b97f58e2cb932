"""CUDA-graph capture of one whole conversion (Euler sampler + vocoder) for fixed shapes.

Small batches are launch-bound: BASELINE config 1 (B = 1, tiny DiT, 10 steps) is ~1 000 launches of a
few microseconds each, and every launch costs ~20 us of host time through ctypes.  All launches of the
path go to torch's current stream through the C ABI, take no host round trip and allocate through
torch's caching allocator, so the whole conversion can be captured once per shape and replayed as one
graph launch.  Tensor maps are encoded on the host at capture time and travel inside the kernel
parameters; the graph's private memory pool keeps every buffer they point to alive.
"""
from __future__ import annotations

import torch


class GraphedConversion:
    """``wave = g(mu, x_lens, prompt, style, z)`` == ``vocoder(cfm.solve_euler(z, ...)[:, :, Tp:])`` for the
    shapes / step count / cfg rate given at construction (v1 ``CFM``).  The returned tensor is a static
    buffer that the next call overwrites."""

    def __init__(self, cfm, vocoder, B, T, Tp, n_timesteps, inference_cfg_rate, device="cuda"):
        dev = torch.device(device)
        C = cfm.in_channels
        self.cfm, self.vocoder, self.Tp = cfm, vocoder, Tp
        self.cfg, self.n = float(inference_cfg_rate), int(n_timesteps)
        self._shape = (B, T, Tp, C)
        self._static = None
        self._graph = None
        self._dev = dev

    def _build(self, mu, x_lens, prompt, style, z):
        dev = self._dev
        self._static = [t.detach().clone().to(dev) for t in (mu, x_lens, prompt, style, z)]
        t_span = torch.linspace(0, 1, self.n + 1)                         # host: no D2H in the capture
        ts = t_span.float()
        t_vals, t = [], ts[0].clone()
        for step in range(1, len(ts)):                                    # same accumulation as solve_euler
            t_vals.append(t.clone())
            t = t + (ts[step] - ts[step - 1])
        self._t_dev = torch.stack(t_vals).to(dev)
        self._t_span = t_span

        def run():
            m, l, p, s, zz = self._static
            mel = self.cfm.solve_euler(zz.clone(), l, p, m, s, None, self._t_span, self.cfg,
                                       t_values_dev=self._t_dev)
            return self.vocoder(mel[:, :, self.Tp:].contiguous())

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                                            # warm-up: caches, lazy tables
                run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._out = run()

    @torch.no_grad()
    def __call__(self, mu, x_lens, prompt, style, z):
        B, T, Tp, C = self._shape
        assert tuple(z.shape) == (B, C, T) and prompt.shape[-1] == Tp and mu.shape[:2] == (B, T)
        if self._graph is None:
            self._build(mu, x_lens, prompt, style, z)
        for dst, src in zip(self._static, (mu, x_lens, prompt, style, z)):
            dst.copy_(src, non_blocking=True)
        self._graph.replay()
        return self._out
