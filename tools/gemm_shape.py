"""Time / profile one GEMM shape through the C ABI: gemm_shape.py M N K [residual] [geglu]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from univst_b200 import ops
M, N, K = (int(a) for a in sys.argv[1:4])
res = "residual" in sys.argv
geglu = "geglu" in sys.argv
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda").half()
w = (torch.randn(N, K, device="cuda") * K ** -0.5).half()
b = torch.randn(N, device="cuda").half()
r = torch.randn(M, N // 2 if geglu else N, device="cuda").half() if res else None
out = torch.empty(M, N // 2 if geglu else N, device="cuda", dtype=torch.float16)
fn = lambda: ops.gemm(a, w, bias=b, residual=r, geglu=geglu, out=out)
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
byt = (M * K + N * K + M * out.shape[1] * (2 if res else 1)) * 2
print(f"gemm {M}x{N}x{K} residual={res} geglu={geglu}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.0f} TF/s  {byt / ms / 1e6:.0f} GB/s (algorithmic)")
