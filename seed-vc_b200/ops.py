"""Tensor-level wrappers over the C ABI (``include/seedvc_b200.h``).

PyTorch is used only for device memory and the current stream; every method
hands raw pointers, sizes and strides to ``libseedvc_b200.so``.  ``Ops`` refuses
non-CUDA tensors: there is no CPU path.

``mode``: ``"fp16"`` (IEEE-half operands on tcgen05 tensor cores, fp32 accumulate, fp32
residual streams - the reference's own default GPU precision, fp16 autocast, inference.py:499),
``"bf16"`` (same kernels with bf16 operands: 8 instead of 11 mantissa bits, wider range) or
``"fp32"`` (fp32 operands, FFMA kernels).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_ROPE, BACKEND_AUTO, BACKEND_SIMT, SVC_BF16, SVC_F16, SVC_F32, GemmDesc,
                   check)

MODES = {"bf16": (torch.bfloat16, SVC_BF16), "fp16": (torch.float16, SVC_F16), "fp32": (torch.float32, SVC_F32)}


def _ptr(t):
    return None if t is None else t.data_ptr()


class Ops:
    def __init__(self, mode: str = "bf16", force_simt: bool = False, fold_norms: bool = False):
        if mode not in MODES:
            raise ValueError("mode must be 'fp16', 'bf16' or 'fp32'")
        self.lib = _lib.load_library()
        self.mode = mode
        self.op_dtype, self.op_code = MODES[mode]
        # dtype of operands that carry a residual STREAM itself (U-ViT skip tensors, long skip, head chain, the
        # x / prompt / content inputs) rather than a normalised branch input.  In bf16 mode these are IEEE half:
        # rounding the stream to 8 mantissa bits is the largest single error source of the bf16 path (measured,
        # scripts/analysis/bf16_ablation.py), the GEMMs that read them are a few % of the FLOPs, and both formats
        # run on the same tcgen05 kind::f16 path.
        self.stream_dtype = torch.float16 if mode == "bf16" else self.op_dtype
        self.backend = BACKEND_SIMT if force_simt else BACKEND_AUTO
        self.precise = 1 if mode == "fp32" else 0
        # RMS norms folded into the GEMMs around them (svc_gemm_desc.row_ss_*; tensor-core epilogues only).  OFF by
        # default: measured on config 2 it saves the two norm passes of a layer (2 x 98 us) but the K = 512 GEMMs
        # are epilogue-bound, and the row scale + bias in the wqkv / w13 epilogues (+96 / +121 us) and the second
        # output of wo / w2 (+38 / +59 us) cost more than that (44.2 against 41.7 ms per Euler step).
        self.fold_norms = bool(fold_norms) and mode != "fp32" and not force_simt
        self.launches = 0
        self._rope_t = {}        # rope table data_ptr -> (table, pair-major copy)
        self.use_rope_t = True   # False: pass only the position-major table (exercises the fallback epilogue)
        self.profile = None      # list of [category, flops, bytes, start_event, end_event] when on

    # ------------------------------------------------------------------ optional per-launch timing
    def start_profile(self):
        self.profile = []

    def stop_profile(self):
        """-> {category: dict(launches, ms, flops, bytes)} (synchronises)."""
        torch.cuda.synchronize()
        out = {}
        for cat, fl, by, e0, e1 in self.profile:
            d = out.setdefault(cat, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["bytes"] += by
        self.profile = None
        return out

    def _t0(self, cat, flops=0.0, nbytes=0.0):
        self.launches += 1
        if self.profile is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            self.profile.append([cat, flops, nbytes, e0, e1])

    def _t1(self):
        if self.profile is not None:
            self.profile[-1][4].record()

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    @staticmethod
    def _chk(*ts):
        for t in ts:
            if t is not None and not t.is_cuda:
                raise _lib.SvcError("seedvc_b200 kernels need CUDA tensors (no CPU fallback)")

    def _code(self, dt):
        if dt == torch.bfloat16:
            return SVC_BF16
        if dt == torch.float16:
            return SVC_F16
        if dt == torch.float32:
            return SVC_F32
        raise TypeError(f"unsupported dtype {dt}")

    def empty(self, *shape, dtype=None, device="cuda"):
        return torch.empty(*shape, dtype=dtype or self.op_dtype, device=device)

    def zeros(self, *shape, dtype=None, device="cuda"):
        return torch.zeros(*shape, dtype=dtype or self.op_dtype, device=device)

    # ------------------------------------------------------------------ GEMM
    def gemm(self, segs, N, *, B, T, bias=None, rowbias=None, act=ACT_NONE, rope=None, gate=None,
             res=None, alpha=1.0, accumulate=False, out_f32=None, out_op=None, f32=False,
             algo_flops=None, row_ss_out=None, row_scale=None):
        """out[b,t,:] = epilogue(sum_s A_s[b, t+shift_s, :] @ W_s.T); see svc_gemm.

        segs: [(A (B, rows, K), shift, W (N, K))]; rope: (table, rope_cols, pos0, q_cols, q_scale).
        ``f32=True`` forces fp32 operands whatever the mode (small conditioning GEMMs).
        Folded RMS norm (svc_gemm_desc.row_ss_*): ``row_ss_out`` (B*T, 4) fp32 receives the row sums of squares of
        the fp32 output; ``row_scale=(ss, dim, eps)`` scales accumulator row r by rsqrt(sum(ss[r]) / dim + eps).
        """
        d = GemmDesc()
        dt = segs[0][0].dtype              # operand type of this call: fp32, or one of the 16-bit formats
        if f32:
            assert dt == torch.float32
        d.dtype = self._code(dt)
        d.B, d.T, d.N, d.n_seg = B, T, N, len(segs)
        for i, (A, shift, W) in enumerate(segs):
            self._chk(A, W)
            assert A.dtype == dt and W.dtype == dt, (A.dtype, W.dtype, dt)
            assert A.dim() == 3 and A.stride(2) == 1 and W.dim() == 2 and W.stride(1) == 1
            assert A.shape[0] == B and W.shape[0] == N and W.shape[1] == A.shape[2]
            d.a_ptr[i] = A.data_ptr()
            d.a_bstride[i], d.a_rstride[i] = A.stride(0), A.stride(1)
            d.a_rows[i], d.a_shift[i] = A.shape[1], shift
            d.w_ptr[i], d.w_rstride[i], d.K[i] = W.data_ptr(), W.stride(0), W.shape[1]
        pair = act in (_lib.ACT_SWIGLU_PAIR, _lib.ACT_TANH_SIG_PAIR)
        n_out = N // 2 if pair else N
        self._chk(bias, rowbias, gate, res, out_f32, out_op)
        if bias is not None:
            assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
            d.bias = bias.data_ptr()
        if rowbias is not None:
            assert rowbias.dtype == torch.float32 and rowbias.shape == (B, N) and rowbias.stride(1) == 1
            d.rowbias, d.rowbias_bstride = rowbias.data_ptr(), rowbias.stride(0)
        d.act = act
        if rope is not None:
            tab, rope_cols, pos0, q_cols, q_scale = rope
            assert act == ACT_ROPE and tab.dtype == torch.float32 and tab.is_contiguous()
            assert tab.shape[0] >= pos0 + T and tab.shape[1:] == (32, 2)
            d.rope_tab, d.rope_cols, d.rope_pos0 = tab.data_ptr(), rope_cols, pos0
            tt = self._rope_t.get(tab.data_ptr())
            if tt is None or tt[0] is not tab:      # pair-major copy for the row-layout epilogue
                tt = (tab, tab.permute(1, 0, 2).contiguous())
                self._rope_t[tab.data_ptr()] = tt
            if self.use_rope_t:
                d.rope_tab_t, d.rope_ld = tt[1].data_ptr(), tab.shape[0]
            d.q_cols, d.q_scale = q_cols, q_scale
        if gate is not None:
            assert gate.dtype == torch.float32 and gate.shape == (B, n_out) and gate.stride(1) == 1
            d.gate, d.gate_bstride = gate.data_ptr(), gate.stride(0)
        if res is not None:
            assert res.dtype == torch.float32 and res.shape == (B, T, n_out) and res.stride(2) == 1
            d.res, d.res_bstride, d.res_rstride = res.data_ptr(), res.stride(0), res.stride(1)
        d.alpha, d.accumulate = float(alpha), int(bool(accumulate))
        if out_f32 is not None:
            assert out_f32.dtype == torch.float32 and out_f32.shape == (B, T, n_out)
            assert out_f32.stride(2) == 1
            d.out_f32 = out_f32.data_ptr()
            d.of_bstride, d.of_rstride = out_f32.stride(0), out_f32.stride(1)
        if out_op is not None:
            assert out_op.shape == (B, T, n_out) and out_op.stride(2) == 1
            if out_op.dtype != dt:         # 16-bit operand copy in the other 16-bit format
                assert dt != torch.float32 and out_op.dtype in (torch.bfloat16, torch.float16)
                d.out_op_dtype_p1 = 1 + self._code(out_op.dtype)
            d.out_op = out_op.data_ptr()
            d.oo_bstride, d.oo_rstride = out_op.stride(0), out_op.stride(1)
        if row_ss_out is not None:
            self._chk(row_ss_out)
            assert row_ss_out.dtype == torch.float32 and row_ss_out.is_contiguous()
            assert row_ss_out.shape == (B * T, _lib.SS_SLOTS)
            d.row_ss_out = row_ss_out.data_ptr()
        if row_scale is not None:
            ss, dim, eps = row_scale
            self._chk(ss)
            assert ss.dtype == torch.float32 and ss.is_contiguous() and ss.shape == (B * T, _lib.SS_SLOTS)
            d.row_ss_in, d.rs_inv_dim, d.rs_eps = ss.data_ptr(), 1.0 / float(dim), float(eps)
        ktot = sum(W.shape[1] for _, _, W in segs)
        cat = "gemm_f32" if dt == torch.float32 else ("gemm_tc" if self.backend == BACKEND_AUTO
                                                               else "gemm_simt")
        # algo_flops: algorithmic work when the launch computes a zero-padded regrouping
        self._t0(cat, 2.0 * B * T * N * ktot if algo_flops is None else float(algo_flops))
        check(self.lib.svc_gemm(C.byref(d), self.backend, self._stream()), "svc_gemm")
        self._t1()

    # ------------------------------------------------------------------ graph-level estimator call
    def dit_step(self, c_weights, c_state, s, x_op, n_launches):
        """svc_dit_step: every kernel of one estimator call, sequenced on the C side (csrc/graph.cu)."""
        self.launches += n_launches
        check(self.lib.svc_dit_step(C.byref(c_weights), C.byref(c_state), int(s), x_op.data_ptr(), self._stream()),
              "svc_dit_step")

    def bigvgan_forward(self, c_weights, mel, out, B, Tm, n_launches):
        """svc_bigvgan_forward: the whole vocoder, sequenced on the C side; the workspace comes from torch's
        caching allocator."""
        self._chk(mel, out)
        n = int(self.lib.svc_bigvgan_workspace_bytes(C.byref(c_weights), B, Tm))
        if n <= 0:
            raise _lib.SvcError("svc_bigvgan_workspace_bytes failed")
        ws = torch.empty(n, dtype=torch.uint8, device=mel.device)
        self.launches += n_launches
        check(self.lib.svc_bigvgan_forward(C.byref(c_weights), mel.data_ptr(), ws.data_ptr(), out.data_ptr(), B, Tm,
                                           self._stream()), "svc_bigvgan_forward")

    # ------------------------------------------------------------------ attention
    def attention(self, qkv, out, H, kv_len):
        """qkv: (B, T, 3*H*64) packed [q|k|v] (RoPE + scale applied); out: (B, T, H*64)."""
        self._chk(qkv, out, kv_len)
        B, T, W = qkv.shape
        D = H * 64
        assert W == 3 * D and qkv.stride(2) == 1 and out.shape == (B, T, D) and out.stride(2) == 1
        assert qkv.dtype == out.dtype and kv_len.dtype == torch.int32 and kv_len.numel() == B
        es = qkv.element_size()
        base = qkv.data_ptr()
        self._t0("attention", 4.0 * B * T * T * D)
        check(self.lib.svc_attention(base, base + D * es, base + 2 * D * es, qkv.stride(0),
                                     qkv.stride(1), out.data_ptr(), out.stride(0), out.stride(1),
                                     B, T, H, kv_len.data_ptr(), self._code(qkv.dtype), self.backend,
                                     self._stream()), "svc_attention")
        self._t1()

    # ------------------------------------------------------------------ norms
    def norm_mod(self, x, out, *, gamma=None, mul=None, add=None, eps=1e-5, mode=0, raw_out=None):
        """out = norm(x) * gamma * mul + add; mode 0 RMSNorm, 1 LayerNorm(no affine).
        ``raw_out`` (same shape / dtype / strides as out): also store the un-normalised rows (operand copy)."""
        self._chk(x, out, gamma, mul, add, raw_out)
        B, T, D = x.shape
        assert x.dtype == torch.float32 and x.stride(2) == 1 and out.shape == x.shape
        assert out.stride(2) == 1
        for v in (gamma, mul, add):
            assert v is None or (v.dtype == torch.float32 and v.numel() == D and v.is_contiguous())
        nbytes = float(B) * T * D * (4 + out.element_size() * (2 if raw_out is not None else 1))
        self._t0("norm_mod", 0.0, nbytes)
        if raw_out is None:
            check(self.lib.svc_norm_mod(x.data_ptr(), x.stride(0), x.stride(1), _ptr(gamma), _ptr(mul),
                                        _ptr(add), float(eps), mode, out.data_ptr(), out.stride(0),
                                        out.stride(1), B, T, D, self._code(out.dtype), self._stream()),
                  "svc_norm_mod")
        else:
            assert raw_out.shape == out.shape and raw_out.stride() == out.stride()
            assert raw_out.element_size() == out.element_size()      # same layout; the 16-bit format may differ
            check(self.lib.svc_norm_mod_copy(x.data_ptr(), x.stride(0), x.stride(1), _ptr(gamma), _ptr(mul),
                                             _ptr(add), float(eps), mode, out.data_ptr(), raw_out.data_ptr(),
                                             out.stride(0), out.stride(1), B, T, D, self._code(out.dtype),
                                             self._code(raw_out.dtype), self._stream()), "svc_norm_mod_copy")
        self._t1()

    # ------------------------------------------------------------------ BigVGAN activations
    def snake(self, x, out, a, inv_b):
        """Anti-aliased SnakeBeta on contiguous (B, L, C)."""
        self._chk(x, out, a, inv_b)
        B, L, Cc = x.shape
        assert x.is_contiguous() and out.is_contiguous() and out.shape == x.shape
        assert a.dtype == torch.float32 and a.numel() == Cc and inv_b.numel() == Cc
        self._t0("snake_aa", 0.0, float(B) * L * Cc * (x.element_size() + out.element_size()))
        check(self.lib.svc_snake_aa(x.data_ptr(), self._code(x.dtype), out.data_ptr(),
                                    self._code(out.dtype), a.data_ptr(), inv_b.data_ptr(), B, L, Cc,
                                    self.precise, self._stream()), "svc_snake_aa")
        self._t1()

    def conv_post(self, act, w, bias, out, use_tanh):
        """act (B, L, C) 16-bit (already activated) -> out (B, L) fp32; w (k, C) fp32."""
        self._chk(act, w, bias, out)
        B, L, Cc = act.shape
        assert act.is_contiguous() and out.shape == (B, L) and out.is_contiguous() and w.shape[1] == Cc
        self._t0("snake_conv_post", 0.0, float(B) * L * (Cc * 2 + 4))
        check(self.lib.svc_conv_post(act.data_ptr(), self._code(act.dtype), w.data_ptr(), _ptr(bias), out.data_ptr(),
                                     B, L, Cc, w.shape[0], int(bool(use_tanh)), self._stream()), "svc_conv_post")
        self._t1()

    def snake_conv_post(self, x, a, inv_b, w, bias, out, use_tanh):
        """x (B, L, C) fp32 -> out (B, L) fp32; w (k, C) fp32."""
        self._chk(x, a, inv_b, w, bias, out)
        B, L, Cc = x.shape
        assert x.dtype == torch.float32 and x.is_contiguous() and out.shape == (B, L)
        assert out.is_contiguous() and w.is_contiguous() and w.shape[1] == Cc
        self._t0("snake_conv_post", 0.0, float(B) * L * (Cc + 1) * 4)
        check(self.lib.svc_snake_conv_post(x.data_ptr(), a.data_ptr(), inv_b.data_ptr(), w.data_ptr(),
                                           _ptr(bias), out.data_ptr(), B, L, Cc, w.shape[0],
                                           int(bool(use_tanh)), self.precise, self._stream()),
              "svc_snake_conv_post")
        self._t1()

    # ------------------------------------------------------------------ sampler
    def cfg_euler(self, x, v, coefs, dt, prompt_len, x_lens=None, x_op=None):
        """x += dt * sum_i coefs[i] * v[i*B:(i+1)*B]; zero prompt / padded rows."""
        self._chk(x, v, x_lens, x_op)
        B, T, Cc = x.shape
        nb = len(coefs)
        assert x.dtype == torch.float32 and x.is_contiguous() and v.is_contiguous()
        assert v.shape == (nb * B, T, Cc) and v.dtype == torch.float32
        c = list(coefs) + [0.0] * (3 - nb)
        if x_op is not None:
            assert x_op.is_contiguous() and x_op.shape == x.shape
        self._t0("cfg_euler", 0.0, float(B) * T * Cc * 4 * (2 + nb))
        check(self.lib.svc_cfg_euler(x.data_ptr(), v.data_ptr(), nb, c[0], c[1], c[2], float(dt), B, T,
                                     Cc, int(prompt_len), _ptr(x_lens), _ptr(x_op),
                                     self._code(x_op.dtype) if x_op is not None else self.op_code,
                                     self._stream()), "svc_cfg_euler")
        self._t1()

    def bct_to_btc(self, inp, out, zero_from=0, zero_to=0):
        """(B, C, T) fp32 contiguous -> (B, T, C) view `out` (any float dtype)."""
        self._chk(inp, out)
        B, Cc, T = inp.shape
        assert inp.dtype == torch.float32 and inp.is_contiguous()
        assert out.shape == (B, T, Cc) and out.stride(2) == 1
        self._t0("misc")
        check(self.lib.svc_bct_to_btc(inp.data_ptr(), out.data_ptr(), out.stride(0), out.stride(1), B,
                                      Cc, T, zero_from, zero_to, self._code(out.dtype),
                                      self._stream()), "svc_bct_to_btc")
        self._t1()

    def btc_to_bct(self, inp, out):
        self._chk(inp, out)
        B, T, Cc = inp.shape
        assert inp.dtype == torch.float32 and inp.is_contiguous() and out.is_contiguous()
        assert out.shape == (B, Cc, T) and out.dtype == torch.float32
        self._t0("misc")
        check(self.lib.svc_btc_to_bct(inp.data_ptr(), out.data_ptr(), B, T, Cc, self._stream()),
              "svc_btc_to_bct")
        self._t1()

    def cast(self, inp, out):
        self._chk(inp, out)
        assert inp.dtype == torch.float32 and inp.is_contiguous() and out.is_contiguous()
        assert inp.numel() == out.numel()
        self._t0("misc")
        check(self.lib.svc_cast(inp.data_ptr(), out.data_ptr(), inp.numel(), self._code(out.dtype),
                                self._stream()), "svc_cast")
        self._t1()

    def scale_cols(self, W, g, mul, out):
        """out[s] = W * g[None, :] * mul[s][None, :]  (folded RMS norm: weight side).  W (N, K) fp32, g (K,) or
        None, mul (S, K) fp32 view or None, out (S, N, K) in any operand dtype."""
        self._chk(W, g, mul, out)
        S, N, K = out.shape
        assert W.dtype == torch.float32 and W.shape == (N, K) and W.stride(1) == 1 and out.is_contiguous()
        assert mul is None or (mul.shape == (S, K) and mul.stride(1) == 1 and mul.dtype == torch.float32)
        assert mul is not None or S == 1
        self._t0("misc")
        check(self.lib.svc_scale_cols(W.data_ptr(), W.stride(0), _ptr(g), _ptr(mul),
                                      mul.stride(0) if mul is not None else 0, out.data_ptr(),
                                      self._code(out.dtype), S, N, K, self._stream()), "svc_scale_cols")
        self._t1()

    def reflect_halo(self, buf, T, pad, lens=None):
        """buf: (B, T + 2*pad, C) with the body at rows [pad, pad+T)."""
        self._chk(buf, lens)
        B, R, Cc = buf.shape
        assert R == T + 2 * pad and buf.stride(2) == 1
        self._t0("misc")
        check(self.lib.svc_reflect_halo(buf.data_ptr(), buf.stride(0), buf.stride(1), B, T, Cc, pad,
                                        _ptr(lens), self._code(buf.dtype), self._stream()),
              "svc_reflect_halo")
        self._t1()

    def timestep_embedding(self, t, freqs, out):
        self._chk(t, freqs, out)
        n, half = t.numel(), freqs.numel()
        assert out.shape == (n, 2 * half) and out.is_contiguous() and out.dtype == torch.float32
        assert t.dtype == torch.float32 and freqs.dtype == torch.float32
        self._t0("misc")
        check(self.lib.svc_timestep_embedding(t.data_ptr(), freqs.data_ptr(), out.data_ptr(), n, half,
                                              self._stream()), "svc_timestep_embedding")
        self._t1()

    def set_rows(self, src, dst):
        """dst[b, :] = src[b or 0, :]; dst is a (B, D) strided view, src (B, D) or (1, D)."""
        self._chk(src, dst)
        B, D = dst.shape
        assert src.dtype == torch.float32 and dst.dtype == torch.float32
        assert src.stride(1) == 1 and dst.stride(1) == 1 and src.shape[1] == D
        sb = 0 if src.shape[0] == 1 else src.stride(0)
        self._t0("misc")
        check(self.lib.svc_set_rows(src.data_ptr(), sb, dst.data_ptr(), dst.stride(0), B, D,
                                    self._stream()), "svc_set_rows")
        self._t1()

    def crossfade_stitch(self, waves, lens, overlap):
        """Stitch vocoded chunks (n, Lmax) fp32 with per-chunk sample counts ``lens`` into one
        waveform; cos^2 crossfade over ``overlap`` samples.  Reference: inference.py:343-350,505-527."""
        import numpy as np
        self._chk(waves)
        n = waves.shape[0]
        assert waves.dtype == torch.float32 and waves.stride(1) == 1 and len(lens) == n
        lens = [int(v) for v in lens]
        assert all(0 < v <= waves.shape[1] for v in lens)
        if n > 1:
            assert all(v >= overlap for v in lens[:-1]), "only the last chunk may be shorter than the overlap"
        offs = [0]
        for k in range(n - 1):
            offs.append(offs[-1] + lens[k] - overlap)
        total = offs[-1] + lens[-1]
        dev = waves.device
        # the reference's ramps, fp64 (inference.py:344-345)
        fo = np.cos(np.linspace(0, np.pi / 2, overlap)) ** 2
        fi = np.cos(np.linspace(np.pi / 2, 0, overlap)) ** 2
        fo_d = torch.from_numpy(fo).to(dev)
        fi_d = torch.from_numpy(fi).to(dev)
        lens_d = torch.tensor(lens, dtype=torch.int32, device=dev)
        offs_d = torch.tensor(offs, dtype=torch.int64, device=dev)
        out = torch.empty(total, dtype=torch.float32, device=dev)
        self._t0("misc")
        check(self.lib.svc_crossfade_stitch(waves.data_ptr(), waves.stride(0), lens_d.data_ptr(),
                                            offs_d.data_ptr(), n, overlap, fi_d.data_ptr(), fo_d.data_ptr(),
                                            out.data_ptr(), total, self._stream()), "svc_crossfade_stitch")
        self._t1()
        return out

    # ------------------------------------------------------------------ length regulator pieces
    def interp_rows(self, src, idx, out, add_vec=None, emb=None, emb_q=None, emb_idx=None):
        """out[b,t,:] = src[b,idx[t],:] (+ add_vec) (+ emb[emb_q[b, emb_idx[t]]]); length_regulator.py:115-129."""
        self._chk(src, idx, out, add_vec, emb, emb_q, emb_idx)
        B, Tout, D = out.shape
        assert src.dtype == torch.float32 and src.stride(2) == 1 and out.stride(2) == 1 and src.shape[2] == D
        assert idx.dtype == torch.int32 and idx.numel() == Tout
        self._t0("misc")
        check(self.lib.svc_interp_rows(
            src.data_ptr(), src.stride(0), src.stride(1), idx.data_ptr(),
            add_vec.data_ptr() if add_vec is not None else None,
            emb.data_ptr() if emb is not None else None,
            emb_q.data_ptr() if emb_q is not None else None,
            emb_q.stride(0) if emb_q is not None else 0,
            emb_idx.data_ptr() if emb_idx is not None else None,
            out.data_ptr(), out.stride(0), out.stride(1), B, Tout, D, self._code(out.dtype),
            self._stream()), "svc_interp_rows")
        self._t1()

    def groupnorm1_mish(self, x, gamma, beta, out, eps=1e-5):
        """out = Mish(GroupNorm(1, C)(x)), x (B, T, C) fp32; length_regulator.py:50-53."""
        self._chk(x, gamma, beta, out)
        B, T, Cc = x.shape
        assert x.dtype == torch.float32 and x.stride(2) == 1 and out.stride(2) == 1 and out.shape == x.shape
        ws = torch.empty(2 * B, dtype=torch.float64, device=x.device)
        self._t0("misc")
        check(self.lib.svc_groupnorm1_mish(x.data_ptr(), x.stride(0), x.stride(1), gamma.data_ptr(),
                                           beta.data_ptr(), eps, ws.data_ptr(), out.data_ptr(), out.stride(0),
                                           out.stride(1), B, T, Cc, self._code(out.dtype), self.precise,
                                           self._stream()), "svc_groupnorm1_mish")
        self._t1()

    def mask_rows(self, x, lens):
        self._chk(x, lens)
        B, T, D = x.shape
        assert x.dtype == torch.float32 and x.stride(2) == 1 and lens.dtype == torch.int32
        self._t0("misc")
        check(self.lib.svc_mask_rows(x.data_ptr(), x.stride(0), x.stride(1), lens.data_ptr(), B, T, D,
                                     self._stream()), "svc_mask_rows")
        self._t1()

    # ------------------------------------------------------------------ mel front-end pieces
    def reflect_pad1d(self, y, pad, out):
        """out[b, :L+2*pad] = reflect-padded y[b]; zeros after (audio.py:58-61)."""
        self._chk(y, out)
        B, L = y.shape
        assert y.dtype == torch.float32 and out.dtype == torch.float32 and y.stride(1) == 1 and out.stride(1) == 1
        self._t0("misc")
        check(self.lib.svc_reflect_pad1d(y.data_ptr(), y.stride(0), B, L, pad, out.data_ptr(), out.stride(0),
                                         out.shape[1], self._stream()), "svc_reflect_pad1d")
        self._t1()

    def stft_mag(self, spec, n_bins, mag, eps=1e-9):
        self._chk(spec, mag)
        rows = spec.shape[0] * spec.shape[1]
        assert spec.is_contiguous() and mag.is_contiguous() and mag.shape[-1] == n_bins
        self._t0("misc")
        check(self.lib.svc_stft_mag(spec.data_ptr(), spec.shape[-1], rows, n_bins, eps, mag.data_ptr(), n_bins,
                                    self._stream()), "svc_stft_mag")
        self._t1()

    def log_clamp(self, x, clip=1e-5):
        self._chk(x)
        assert x.dtype == torch.float32 and x.is_contiguous()
        self._t0("misc")
        check(self.lib.svc_log_clamp(x.data_ptr(), x.numel(), clip, self._stream()), "svc_log_clamp")
        self._t1()


    # ------------------------------------------------------------------ HiFT vocoder pieces
    UNARY_LRELU, UNARY_ELU, UNARY_SNAKE, UNARY_ABS = 0, 1, 2, 3

    def unary(self, x, out, kind, slope=0.0, alpha=None):
        """out = f(x) on (B, T, C) fp32 -> out's dtype; see svc_unary."""
        self._chk(x, out, alpha)
        B, T, Cc = x.shape
        assert x.dtype == torch.float32 and x.stride(2) == 1 and out.shape == x.shape and out.stride(2) == 1
        self._t0("hift_unary", 0.0, float(B) * T * Cc * (4 + out.element_size()))
        check(self.lib.svc_unary(x.data_ptr(), x.stride(0), x.stride(1), out.data_ptr(), out.stride(0), out.stride(1),
                                 B, T, Cc, kind, float(slope), _ptr(alpha), self._code(out.dtype), self.precise,
                                 self._stream()), "svc_unary")
        self._t1()

    def hift_source(self, f0, phase, noise, lin_w, lin_b, out, scale, sr, sine_amp, noise_std, voiced_thr):
        """f0 (B, Tm) fp32, phase (B, H), noise (B, H, Tm*scale) or None -> out (B, Tm*scale) fp32."""
        self._chk(f0, phase, noise, lin_w, out)
        B, Tm = f0.shape
        H = phase.shape[1]
        assert f0.dtype == torch.float32 and f0.stride(1) == 1 and phase.is_contiguous() and phase.shape == (B, H)
        assert out.shape == (B, Tm * scale) and out.stride(1) == 1 and lin_w.numel() == H and lin_w.is_contiguous()
        if noise is not None:
            assert noise.shape == (B, H, Tm * scale) and noise.is_contiguous() and noise.dtype == torch.float32
        ws = torch.empty(B * H * Tm, dtype=torch.float64, device=f0.device)
        self._t0("hift_source")
        check(self.lib.svc_hift_source(f0.data_ptr(), f0.stride(0), phase.data_ptr(), _ptr(noise), lin_w.data_ptr(),
                                       float(lin_b), ws.data_ptr(), out.data_ptr(), out.stride(0), B, Tm, H, scale,
                                       float(sr), float(sine_amp), float(noise_std), float(voiced_thr),
                                       self._stream()), "svc_hift_source")
        self._t1()

    def hift_stft(self, s, out):
        """s (B, L) fp32 -> out (B, rows, Cpad): [re | im | 0] per frame, rows beyond L/4+1 zeroed."""
        self._chk(s, out)
        B, L = s.shape
        assert s.dtype == torch.float32 and s.stride(1) == 1 and out.shape[0] == B and out.stride(2) == 1
        self._t0("hift_stft")
        check(self.lib.svc_hift_stft(s.data_ptr(), s.stride(0), out.data_ptr(), out.stride(0), out.stride(1), B, L,
                                     out.shape[1], out.shape[2], self._code(out.dtype), self._stream()),
              "svc_hift_stft")
        self._t1()

    def hift_istft(self, x, wav, clip_mag=1e2, audio_limit=0.99):
        """x (B, TT, >=18) fp32 -> wav (B, 4*(TT-1)) fp32."""
        self._chk(x, wav)
        B, TT, Cx = x.shape
        assert x.dtype == torch.float32 and x.stride(2) == 1 and Cx >= 18
        assert wav.shape == (B, 4 * (TT - 1)) and wav.stride(1) == 1 and wav.dtype == torch.float32
        self._t0("hift_istft")
        check(self.lib.svc_hift_istft(x.data_ptr(), x.stride(0), x.stride(1), wav.data_ptr(), wav.stride(0), B, TT,
                                      float(clip_mag), float(audio_limit), self.precise, self._stream()),
              "svc_hift_istft")
        self._t1()
