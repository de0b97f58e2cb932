import ctypes as C, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
os.environ["SEEDVC_B200_LIB"] = os.path.join(ROOT, "seed-vc_b200", "libseedvc_b200_v_ws.so")
import seedvc_b200
from seedvc_b200.ops import Ops
ops = Ops("bf16")
B, T, H = 64, 2580, 8
D = H * 64
qkv = torch.randn(B, T, 3 * D, device="cuda").to(torch.bfloat16); qkv[..., :D] *= 0.125
out = torch.empty(B, T, D, dtype=torch.bfloat16, device="cuda")
kv = torch.full((B,), T, dtype=torch.int32, device="cuda")
for _ in range(3): ops.attention(qkv, out, H, kv)
torch.cuda.synchronize()
buf = (C.c_longlong * 72)()
ops.lib.svc_debug_attn_waitstat.argtypes = [C.c_void_p]
print("rc", ops.lib.svc_debug_attn_waitstat(buf))
for w in range(12):
    ws, wp, tot, nb = buf[4*w:4*w+4]
    print(f"softmax warp {w} (group {w//4}): s_full wait {ws/nb:7.1f} cyc/block, p_empty wait {wp/nb:7.1f}, total {tot/nb:7.1f} cyc/block over {nb} blocks")

for g in range(3):
    a = buf[48 + 8*g: 48 + 8*g + 7]; nb = a[6]
    print(f"issuer g{g}: per block: wait k_full {a[0]/nb:6.1f}  wait s_empty {a[1]/nb:6.1f}  issue S {a[2]/nb:6.1f}  wait v_full {a[3]/nb:6.1f}  wait p_full {a[4]/nb:6.1f}  issue PV {a[5]/nb:6.1f}")
