import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["SEEDVC_B200_LIB"] = os.path.join(ROOT, "seed-vc_b200", "libseedvc_b200_trace.so")
import seedvc_b200
from seedvc_b200.ops import Ops
ops = Ops("bf16")
B, T, H = 64, 2580, 8
D = H * 64
qkv = torch.randn(B, T, 3 * D, device="cuda").to(torch.bfloat16); qkv[..., :D] *= 0.125
out = torch.empty(B, T, D, dtype=torch.bfloat16, device="cuda")
kv = torch.full((B,), T, dtype=torch.int32, device="cuda")
import ctypes
ops.attention(qkv, out, H, kv); torch.cuda.synchronize()
z = (C.c_longlong * (4 * 64 * 8))()
# zero the trace between runs is not possible from the host without a symbol write; one launch only
torch.cuda.synchronize()
buf = (C.c_longlong * (4 * 64 * 8))()
ops.lib.svc_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
print("rc", ops.lib.svc_debug_attn_trace(buf, 4 * 64 * 8))
import numpy as np
a = np.array(buf[:]).reshape(4, 64, 8)
t0 = a[1, 0, 0]
print("cols: 0 s_empty[0](j) seen, 1 k_full(j) seen, 2 S0(j) issued, 3 K_j load issued, 4 V_j load issued, 5 PV0(j) issued, 6 PV1(j) issued")
for j in range(0, 21):
    print(j, " ".join(f"{int(x - t0):8d}" if x else "       -" for x in a[0, j, :7]))
print("softmax g0 lane0: 0 enter 1 s_full 2 ld done 3 exp done 4 p_empty 5 P stored 6 arrived")
for j in range(0, 21):
    print(j, " ".join(f"{int(x - t0):8d}" for x in a[1, j, :7]))

print("S-load done per softmax warp 4..11 (g0: 4-7, g1: 8-11)")
for j in range(0, 21):
    print(j, " ".join(f"{int(x - t0):8d}" if x else "       -" for x in a[2, j, :8]))

print("MMA thread g=1: 0 PV1 enter, 1 after fence, 2 after 8 MMAs, 3 S1 enter, 4 S1 issued+commit")
for j in range(0, 14):
    print(j, " ".join(f"{int(x - t0):8d}" if x else "       -" for x in a[3, j, :5]))
