"""Time one 3x3 convolution shape through the C ABI: conv_shape.py NB H W Cin Cout"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from univst_b200 import ops
from univst_b200.pack import pack_conv3x3
NB, H, W, Cin, Cout = (int(a) for a in sys.argv[1:6])
torch.manual_seed(0)
x = torch.randn(NB, H, W, Cin, device="cuda").half()
w = pack_conv3x3((torch.randn(Cout, Cin, 3, 3) * (9 * Cin) ** -0.5), 0).cuda().half()
b = torch.randn(Cout, device="cuda").half()
r = torch.randn(NB * H * W, Cout, device="cuda").half()
out = torch.empty(NB * H * W, Cout, device="cuda", dtype=torch.float16)
fn = lambda: ops.conv3x3(x, w, bias=b, residual=r, out=out)
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
print(f"conv3x3 {NB}x{H}x{W} {Cin}->{Cout} BN_OVERRIDE={os.environ.get('UNIVST_BN_OVERRIDE', '-')}: {ms * 1e3:.1f} us  {2.0 * NB * H * W * Cout * 9 * Cin / ms / 1e9:.0f} TF/s")
