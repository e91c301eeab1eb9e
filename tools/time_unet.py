"""Quick device timing of one full-size UNet forward (SD-1.5 shapes, seeded random weights)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_oracle as uo
from univst_b200.unet import UNetPseudo3DConditionModel
from univst_b200 import pnp_utils, ops
from types import SimpleNamespace

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
AD = "--animatediff" in sys.argv
if AD:
    from oracle import animatediff_oracle as ao
    from univst_b200.animatediff import UNet3DConditionModel as UNetPseudo3DConditionModel  # noqa: F811
    cfg = ao.AD_SD15_CONFIG
    shapes = ao.unet_param_shapes(cfg)
else:
    cfg = uo.SD21_CONFIG if "--sd21" in sys.argv else uo.SD15_CONFIG
    shapes = uo.unet_param_shapes(cfg)
t0 = time.time()
g = torch.Generator(device="cuda").manual_seed(33)
sd = {}
for k, s in shapes.items():
    if "attn_temporal.to_out.0.weight" in k:
        sd[k] = torch.zeros(s, device="cuda", dtype=torch.float16)
    elif k.endswith("weight") and len(s) == 1:
        sd[k] = torch.ones(s, device="cuda", dtype=torch.float16)
    elif k.endswith("bias"):
        sd[k] = (0.02 * torch.randn(s, device="cuda", generator=g)).half()
    else:
        fan = 1
        for d in s[1:]:
            fan *= d
        sd[k] = (torch.randn(s, device="cuda", generator=g) * fan ** -0.5).half()
unet = UNetPseudo3DConditionModel(sd, cfg)
print("built in", time.time() - t0, "s")
pipe = SimpleNamespace(unet=unet)
pnp_utils.register_spatial_attention_pnp(pipe)
pnp_utils.register_time(pipe, 5)
x = torch.randn(3, 4, F, 64, 64, device="cuda").half()
ctx = torch.randn(3, 77, cfg["cross_attention_dim"], device="cuda").half()
for _ in range(2):
    y = unet(x, 981, encoder_hidden_states=ctx).sample
torch.cuda.synchronize()
print("finite:", torch.isfinite(y).all().item(), "absmean", y.float().abs().mean().item())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = ops.launch_count
e0.record()
for _ in range(iters):
    y = unet(x, 981, encoder_hidden_states=ctx).sample
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"UNet 3x{F}x64x64 forward: {ms:.2f} ms  ({(ops.launch_count - n0) // iters} launches)  "
      f"{3 * F * 0.9751 / ms:.1f} TFLOP/s live  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")

if "--shapes" in sys.argv:
    import collections
    ops.profile_start({"gemm", "conv3x3", "sc_attention", "temporal_attention", "cross_attention"})
    unet(x, 981, encoder_hidden_states=ctx)
    prof = ops.profile_stop()
    for name, recs in prof.items():
        agg = collections.defaultdict(lambda: [0, 0.0])
        for ms_, meta in recs:
            agg[meta][0] += 1
            agg[meta][1] += ms_
        print(f"== {name}: {sum(v[1] for v in agg.values()):.2f} ms")
        for meta, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            if name in ("sc_attention", "cross_attention"):
                fl = 4.0 * meta[3] * meta[4] * meta[1] * meta[2] * meta[0]
            elif name == "temporal_attention":   # (B, F, N, H, d): report GB/s of Q/K/V + output traffic instead
                fl = 4.0 * meta[0] * meta[1] * meta[2] * meta[3] * meta[4] * 2 * 1e3
            else:
                fl = 2.0 * meta[0] * meta[1] * meta[2]
            print(f"  {t:8.3f} ms n={n:3d} avg={t / n * 1e3:8.1f} us {fl / (t / n) / 1e9:8.1f} TF/s  {meta}")
