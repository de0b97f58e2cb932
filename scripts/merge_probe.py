import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import seedvc_b200
from seedvc_b200.ops import Ops
ops = Ops("bf16"); DEV = "cuda"
def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
B, T, D = 32, 2580, 512
for K in (80, 64, 128, 512):
    x = torch.randn(B, T, K, device=DEV).bfloat16()
    Wfull = (torch.randn(D, 864, device=DEV) * 0.05).bfloat16()
    W = Wfull[:, :K]
    Wc = W.contiguous()
    res = torch.randn(B, T, D, device=DEV)
    h = torch.empty(2 * B, T + 1, D, device=DEV)
    out_s = h[:B, 1:, :]
    out_c = torch.empty(B, T, D, device=DEV)
    bias = torch.randn(D, device=DEV)
    print(f"K={K}")
    print("  bias, out contiguous      %8.1f us" % timeit(lambda: ops.gemm([(x, 0, W)], D, B=B, T=T, bias=bias, out_f32=out_c)))
    print("  bias, out strided         %8.1f us" % timeit(lambda: ops.gemm([(x, 0, W)], D, B=B, T=T, bias=bias, out_f32=out_s)))
    print("  res!=out, out contiguous  %8.1f us" % timeit(lambda: ops.gemm([(x, 0, W)], D, B=B, T=T, res=res, out_f32=out_c)))
    print("  res!=out, out strided     %8.1f us" % timeit(lambda: ops.gemm([(x, 0, W)], D, B=B, T=T, res=res, out_f32=out_s)))
    print("  res!=out, W contiguous    %8.1f us" % timeit(lambda: ops.gemm([(x, 0, Wc)], D, B=B, T=T, res=res, out_f32=out_c)))
    print("  res==out (in place)       %8.1f us" % timeit(lambda: ops.gemm([(x, 0, W)], D, B=B, T=T, res=out_c, out_f32=out_c)))
