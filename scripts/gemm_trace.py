import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["SEEDVC_B200_LIB"] = os.path.join(ROOT, "seed-vc_b200", "libseedvc_b200_trace.so")
import seedvc_b200
from seedvc_b200 import _lib
from seedvc_b200.ops import Ops
from seedvc_b200.dit_engine import rope_table
import numpy as np
ops = Ops("bf16")
B, T, D = 64, 2580, 512
which = sys.argv[1] if len(sys.argv) > 1 else "qkv"
A = (torch.randn(B, T, D, device="cuda")).to(torch.bfloat16)
if which == "qkv":
    N = 3 * D; W = (torch.randn(N, D, device="cuda") * 0.05).to(torch.bfloat16)
    tab = rope_table(T + 4).to("cuda")
    out = torch.empty(B, T, N, dtype=torch.bfloat16, device="cuda")
    fn = lambda: ops.gemm([(A, 0, W)], N, B=B, T=T, act=_lib.ACT_ROPE, rope=(tab, 2 * D, 0, D, 0.125), out_op=out)
else:
    N = 6 * D; W = (torch.randn(N, D, device="cuda") * 0.05).to(torch.bfloat16)
    out = torch.empty(B, T, N // 2, dtype=torch.bfloat16, device="cuda")
    fn = lambda: ops.gemm([(A, 0, W)], N, B=B, T=T, act=_lib.ACT_SWIGLU_PAIR, out_op=out)
fn(); torch.cuda.synchronize()
n = 2 * 128 * 8
buf = (C.c_longlong * n)()
ops.lib.svc_debug_gemm_trace.argtypes = [C.c_void_p, C.c_int]
print("rc", ops.lib.svc_debug_gemm_trace(buf, n))
a = np.array(buf[:]).reshape(2, 128, 8)
t0 = a[1, 0, 0]
print("MMA thread per tile: 0 start, 1 tmem_empty ok, 2 all issued")
for i in range(12): print(i, " ".join(f"{int(x - t0):8d}" for x in a[1, i, :3]))
print("epilogue warp 2 per item: 0 start, 1 ld+prefetch issued, 2 ld done, 3 item done")
for i in range(40): print(i, " ".join(f"{int(x - t0):8d}" for x in a[0, i, :4]))
