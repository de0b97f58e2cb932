#!/usr/bin/env python
"""Kernel micro-benchmarks at BASELINE config-2 shapes (CUDA events, 3 warm-up + N timed).
usage: python scripts/kbench.py [attn] [gemm] [snake] [norm]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import seedvc_b200  # noqa: E402
from seedvc_b200 import _lib  # noqa: E402
from seedvc_b200.dit_engine import rope_table  # noqa: E402
from seedvc_b200.ops import Ops  # noqa: E402

DEV = "cuda"
ops = Ops("bf16")
what = set(sys.argv[1:]) or {"attn", "gemm", "snake", "norm"}


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def bf(*s):
    return torch.randn(*s, device=DEV).to(torch.bfloat16)


if "attn" in what:
    for B, T, H in ((64, 2580, 8), (16, 2580, 8), (64, 325, 6), (8, 1293, 6)):
        D = H * 64
        qkv = bf(B, T, 3 * D)
        qkv[..., :D] *= 0.125
        out = torch.empty(B, T, D, dtype=torch.bfloat16, device=DEV)
        kv = torch.full((B,), T, dtype=torch.int32, device=DEV)
        ms = timeit(lambda: ops.attention(qkv, out, H, kv))
        fl = 4.0 * B * T * T * D
        print(f"attention B={B} T={T} H={H}: {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s")

if "gemm" in what:
    M_B, M_T = 64, 2580
    D, I, Dw, C = 512, 1536, 512, 80
    tab = rope_table(M_T + 4).to(DEV)
    cases = [
        ("qkv  K512 N1536 rope->bf16", [D], 3 * D, dict(act=_lib.ACT_ROPE, rope=(tab, 2 * D, 0, D, 0.125)), "op"),
        ("qkvN K512 N1536 plain->bf16", [D], 3 * D, dict(), "op"),
        ("qkvB K512 N1536 bias->bf16", [D], 3 * D, dict(bias=True), "op"),
        ("wo   K512 N512 +res->f32", [D], D, dict(res=True), "f32"),
        ("w13  K512 N3072 swiglu->bf16", [D], 2 * I, dict(act=_lib.ACT_SWIGLU_PAIR), "op"),
        ("w2   K1536 N512 +res->f32", [I], D, dict(res=True), "f32"),
        ("skip K512+512 N512 ->f32", [D, D], D, dict(), "f32"),
        ("merge K80 N512 +res->f32", [C], D, dict(res=True), "f32"),
        ("wn_in 5xK512 N1024 gate->bf16", [Dw] * 5, 2 * Dw, dict(act=_lib.ACT_TANH_SIG_PAIR), "op"),
        ("wn_rs K512 N512 +res->f32+bf16", [Dw], Dw, dict(res=True), "both"),
        ("wn_rsB K512 N512 bias+res->f32+bf16", [Dw], Dw, dict(res=True, bias=True), "both"),
        ("w2d  K1536 N512 +res->f32+bf16", [I], D, dict(res=True), "both"),
        ("wod  K512 N512 +res->f32+bf16", [D], D, dict(res=True), "both"),
        ("conv2 K512 N80 ->f32", [Dw], C, dict(), "f32"),
    ]
    flt = os.environ.get("KB_FILTER", "")
    for name, Ks, N, kw, outk in cases:
        if flt and flt not in name:
            continue
        A = bf(M_B, M_T + 4, Ks[0])
        segs = []
        for i, K in enumerate(Ks):
            W = bf(N, K) * 0.05
            segs.append((A if len(Ks) == 5 else (A[:, :M_T] if i == 0 else bf(M_B, M_T, K)), i if len(Ks) == 5 else 0, W))
        pair = kw.get("act") in (_lib.ACT_SWIGLU_PAIR, _lib.ACT_TANH_SIG_PAIR)
        n_out = N // 2 if pair else N
        kws = dict(kw)
        if kws.pop("res", False):
            kws["res"] = torch.randn(M_B, M_T, n_out, device=DEV)
        if kws.pop("bias", False):
            kws["bias"] = torch.randn(N, device=DEV)
        of = torch.empty(M_B, M_T, n_out, device=DEV) if outk in ("f32", "both") else None
        oo = torch.empty(M_B, M_T, n_out, dtype=torch.bfloat16, device=DEV) if outk in ("op", "both") else None
        if "res" in kws and of is not None:
            of = kws["res"]          # in-place residual update like the engine does
        ms = timeit(lambda: ops.gemm(segs, N, B=M_B, T=M_T, out_f32=of, out_op=oo, **kws))
        fl = 2.0 * M_B * M_T * N * sum(Ks)
        print(f"gemm {name:34s}: {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s")
    # BigVGAN conv shapes (B=32, Tm=2150)
    for ch, Lm, k, d in () if flt else ((768, 4, 3, 1), (768, 4, 11, 5), (384, 16, 7, 3), (192, 32, 7, 1), (96, 64, 11, 1),
                         (48, 128, 7, 3), (24, 256, 3, 1), (24, 256, 11, 5)):
        Bv, L = 32, 2150 * Lm
        A = bf(Bv, L, ch)
        W = bf(k, ch, ch) * 0.05
        of = torch.empty(Bv, L, ch, device=DEV)
        res = torch.randn(Bv, L, ch, device=DEV)
        half = (k - 1) // 2
        segs = [(A, (t - half) * d, W[t]) for t in range(k)]
        ms = timeit(lambda: ops.gemm(segs, ch, B=Bv, T=L, res=res, out_f32=of), n=5)
        fl = 2.0 * Bv * L * ch * ch * k
        by = Bv * L * ch * (2 + 4 + 4)
        print(f"conv ch={ch:4d} L={L:7d} k={k:2d} d={d}: {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s  {by / ms / 1e6:7.0f} GB/s")

if "snake" in what:
    for ch, Lm in ((768, 4), (384, 16), (192, 32), (96, 64), (48, 128), (24, 256)):
        Bv, L = 32, 2150 * Lm
        x = torch.randn(Bv, L, ch, device=DEV)
        out = torch.empty(Bv, L, ch, dtype=torch.bfloat16, device=DEV)
        a = torch.rand(ch, device=DEV) + 0.5
        ib = torch.rand(ch, device=DEV) + 0.5
        ms = timeit(lambda: ops.snake(x, out, a, ib), n=5)
        print(f"snake ch={ch:4d} L={L:7d}: {ms:8.3f} ms  {Bv * L * ch * 6 / ms / 1e6:7.0f} GB/s")
        xh = x.half()
        ms = timeit(lambda: ops.snake(xh, out, a, ib), n=5)
        print(f"snake ch={ch:4d} L={L:7d} half in: {ms:8.3f} ms  {Bv * L * ch * 4 / ms / 1e6:7.0f} GB/s")
        del xh

if "norm" in what:
    x = torch.randn(64, 2580, 512, device=DEV)
    out = torch.empty(64, 2580, 512, dtype=torch.bfloat16, device=DEV)
    g = torch.randn(512, device=DEV)
    ms = timeit(lambda: ops.norm_mod(x, out, gamma=g, mul=g, add=g))
    print(f"norm_mod 165k x 512: {ms:8.3f} ms  {x.numel() * 6 / ms / 1e6:7.0f} GB/s")
