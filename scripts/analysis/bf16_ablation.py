"""CPU study (no GPU needed): where does the bf16-mode velocity error come from?

Runs the package's own host orchestration (DiTEngine) on tests/emu_ops.py with bf16 ROUNDING emulated at
exactly the points where the CUDA path rounds (every operand-dtype store, every weight), keeps an unrounded
fp32 shadow of every operand buffer, and lets individual GEMM sites read the exact A operand / exact weights.
Compares one estimator call (CFG pair, step 0) against the fp32 emulation.

    python scripts/analysis/bf16_ablation.py [T] [Tp]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import seedvc_b200  # noqa
from seedvc_b200 import configs, synth
from seedvc_b200.dit_engine import DiTEngine
from seedvc_b200.flow_matching import CFM
from emu_ops import EmuOps, ACT_ROPE, ACT_SWIGLU_PAIR, ACT_TANH_SIG_PAIR, ACT_SILU
import torch.nn.functional as F

BF = torch.float16 if os.environ.get("EMU_F16") else torch.bfloat16


def r(x):
    return x.to(BF).float()


class EmuBf16(EmuOps):
    def __init__(self):
        super().__init__()
        self.op_dtype = torch.float32       # weights load as fp32; switched to bf16 before begin()
        self.shadow = {}
        self.exact_A = set()                # gemm site names reading the unrounded A operand
        self.exact_W = set()
        self.exact_attn_in = False
        self.exact_P = True                 # P rounding emulated when False
        self.sites = []
        self.record = True
        self.namer = None

    # ---- shadows -----------------------------------------------------------------------
    def _sh(self, t, create=True):
        if t.dtype != BF:
            return None
        key = t.untyped_storage().data_ptr()
        if key not in self.shadow:
            if not create:
                return None
            self.shadow[key] = torch.zeros(t.untyped_storage().nbytes() // 2, dtype=torch.float32)
        return torch.as_strided(self.shadow[key], t.shape, t.stride(), t.storage_offset())

    def _store(self, dst, v):
        if dst.dtype == BF:
            self._sh(dst).copy_(v)
        dst.copy_(v)

    def _read(self, t, exact):
        if t.dtype == BF and exact:
            s = self._sh(t, create=False)
            if s is not None:
                return s
        return t.float()

    def empty(self, *shape, dtype=None, device="cpu"):
        return torch.zeros(*shape, dtype=dtype or self.op_dtype)

    zeros = empty

    # ---- ops ---------------------------------------------------------------------------
    def gemm(self, segs, N, *, B, T, bias=None, rowbias=None, act=0, rope=None, gate=None, res=None, alpha=1.0,
             accumulate=False, out_f32=None, out_op=None, f32=False, algo_flops=None):
        ktot = sum(W.shape[1] for _, _, W in segs)
        name = self.namer(act, N, ktot, len(segs), f32, res is not None, out_f32 is not None, out_op is not None)
        if self.record:
            self.sites.append(name)
        exA, exW = name in self.exact_A or f32, name in self.exact_W or f32
        segs2 = []
        for A, sh, W in segs:
            a = self._read(A, exA)
            w = W.float() if exW else r(W.float())
            segs2.append((a, sh, w))
        tmp_f = torch.zeros(B, T, N // 2 if act in (ACT_SWIGLU_PAIR, ACT_TANH_SIG_PAIR) else N)
        super().gemm(segs2, N, B=B, T=T, bias=bias, rowbias=rowbias, act=act, rope=rope, gate=gate, res=res,
                     alpha=alpha, accumulate=False, out_f32=tmp_f)
        if accumulate:
            tmp_f = tmp_f + out_f32
        if out_f32 is not None:
            out_f32.copy_(tmp_f)
        if out_op is not None:
            self._store(out_op, tmp_f)

    def attention(self, qkv, out, H, kv_len):
        q = self._read(qkv, self.exact_attn_in)
        B, T, W = q.shape
        D = W // 3
        qq, k, v = q.split([D, D, D], dim=-1)
        qq = qq.view(B, T, H, 64).transpose(1, 2)
        k = k.view(B, T, H, 64).transpose(1, 2)
        v = v.view(B, T, H, 64).transpose(1, 2)
        s = qq @ k.transpose(-1, -2)
        ok = torch.arange(T)[None, :] < kv_len[:, None]
        s = s.masked_fill(~ok[:, None, None, :], float("-inf"))
        m = s.max(-1, keepdim=True).values
        p = torch.exp(s - m)
        l = p.sum(-1, keepdim=True)
        if not self.exact_P:
            p = r(p)
        y = (p @ v) / l
        self._store(out, y.transpose(1, 2).reshape(B, T, D))

    def norm_mod(self, x, out, *, gamma=None, mul=None, add=None, eps=1e-5, mode=0, raw_out=None):
        if raw_out is not None:
            self._store(raw_out, x)
        if mode == 0:
            y = x * torch.rsqrt(torch.mean(x * x, -1, keepdim=True) + eps)
        else:
            y = F.layer_norm(x, (x.shape[-1],), eps=eps)
        for g in (gamma, mul):
            if g is not None:
                y = y * g
        if add is not None:
            y = y + add
        self._store(out, y)

    def cfg_euler(self, x, v, coefs, dt, prompt_len, x_lens=None, x_op=None):
        super().cfg_euler(x, v, coefs, dt, prompt_len, x_lens, None)
        if x_op is not None:
            self._store(x_op, x)

    def bct_to_btc(self, inp, out, zero_from=0, zero_to=0):
        v = inp.transpose(1, 2).clone()
        v[:, zero_from:zero_to] = 0
        self._store(out, v)

    def cast(self, inp, out):
        self._store(out, inp.view(out.shape))

    def reflect_halo(self, buf, T, pad, lens=None):
        super().reflect_halo(buf, T, pad, lens)
        s = self._sh(buf, create=False)
        if s is not None:
            super().reflect_halo(s, T, pad, lens)


def small_namer():
    """Names for the whisper-small GEMM sites from their descriptors."""
    def f(act, N, K, nseg, f32, res, of, oo):
        if f32:
            return "cond_f32"
        if act == ACT_ROPE:
            return "wqkv"
        if act == ACT_SWIGLU_PAIR:
            return "w13"
        if act == ACT_TANH_SIG_PAIR:
            return "wn_in"
        if K == 80:
            return "merge_x"
        if K == 1536:
            return "w2"
        if nseg == 2 and K == 1024:
            return "skip_in"
        if nseg == 2 and K == 592:
            return "long_skip"
        if nseg == 2 and K == 4608:
            return "wn_skip"
        if N == 80:
            return "conv2"
        if nseg == 2 and not res:
            return "merge_const"
        if K == 512 and res and of and not oo:
            return "wo"
        if K == 512 and res and of and oo:
            return "wn_rs"
        if K == 512 and of and oo:
            return "conv1"
        if K == 512 and oo and not of:
            return "fl_or_cond"
        return f"other_{act}_{N}_{K}_{nseg}"
    return f


def run(ops_cfg, T, Tp):
    args = configs.v1_model_params("whisper_small")
    cfm = CFM(args)
    ops = EmuBf16()
    ops.namer = small_namer()
    for k, v in ops_cfg.items():
        setattr(ops, k, v)
    eng = DiTEngine(cfm.estimator.spec, ops)
    eng.load_weights(cfm.estimator.state_dict(), "cpu")
    bf = ops_cfg.get("bf16", True)
    ops.op_dtype = BF if bf else torch.float32
    if not bf:
        ops.exact_A = ops.exact_W = type("All", (), {"__contains__": lambda s, x: True})()
    cfm.estimator.engine = lambda: eng
    cfm.estimator.setup_caches(1, 8192)
    mu, prompt, style, z = synth.synth_batch(1, T, Tp, 80, 512, first_id=301)
    vs = []
    t_span = torch.linspace(0, 1, 26)[:2]
    cfm.solve_euler(z.clone(), torch.tensor([T]), prompt, mu, style, None, t_span, 0.7,
                    step_hook=lambda s, v: vs.append(v.clone()))
    return vs[0][0, Tp:], ops


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


if __name__ == "__main__":
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 323
    Tp = int(sys.argv[2]) if len(sys.argv) > 2 else 258
    torch.set_num_threads(8)
    ref, _ = run({"bf16": False}, T, Tp)
    base, ops = run({}, T, Tp)
    names = sorted(set(ops.sites))
    print("sites:", {n: ops.sites.count(n) for n in names})
    print(f"all bf16 (exact P): velocity rel-L2 {rel(base, ref):.3e}")
    v, _ = run({"exact_P": False}, T, Tp)
    print(f"all bf16 + P rounded to bf16: {rel(v, ref):.3e}")
    for n in names:
        if n == "cond_f32":
            continue
        va, _ = run({"exact_A": {n}}, T, Tp)
        vw, _ = run({"exact_W": {n}}, T, Tp)
        vb, _ = run({"exact_A": {n}, "exact_W": {n}}, T, Tp)
        print(f"  {n:12s} exact A: {rel(va, ref):.3e}   exact W: {rel(vw, ref):.3e}   both: {rel(vb, ref):.3e}")
    v, _ = run({"exact_attn_in": True}, T, Tp)
    print(f"  attention reads exact qkv: {rel(v, ref):.3e}")
