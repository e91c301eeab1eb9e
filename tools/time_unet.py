"""Device timing of one UNet forward (SD-1.5 shapes, seeded random weights) on one GPU.

    python tools/time_unet.py [F] [iters] [--branches B] [--idx I] [--truncate] [--graph] [--splitk] [--shapes] [--kernels]
                              [--animatediff] [--sd21]

F frames per branch (16 = the whole clip; 2 = what one rank of an 8-GPU frame-sharded run evaluates, without its
synchronisations), B = 3 (shift window open unless --idx > 25) or 1 (edit branch alone).  --graph replays the forward
as a CUDA graph; --shapes prints the GEMM / conv / attention launches by shape with their achieved TFLOP/s;
--kernels prints device time by entry point (CUDA events around every launch)."""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from univst_b200 import ops, pnp_utils
from univst_b200.weights import random_state_dict


def arg(name, default):
    return type(default)(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


pos = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith("--") and not sys.argv[i - 1] in ("--branches", "--idx", "--hw")]
F = int(pos[0]) if len(pos) > 0 else 16
iters = int(pos[1]) if len(pos) > 1 else 3
B, idx, hw = arg("--branches", 3), arg("--idx", 5), arg("--hw", 64)
AD = "--animatediff" in sys.argv
if AD:
    from univst_b200.animatediff import AD_SD15_CONFIG as cfg, UNet3DConditionModel as Model
else:
    from univst_b200.unet import SD15_CONFIG, SD21_CONFIG, UNetPseudo3DConditionModel as Model
    cfg = SD21_CONFIG if "--sd21" in sys.argv else SD15_CONFIG
t0 = time.time()
unet = Model(random_state_dict(cfg, seed=33, device="cuda", animatediff=AD), cfg)
print("built in", round(time.time() - t0, 2), "s")
pipe = SimpleNamespace(unet=unet)
pnp_utils.register_spatial_attention_pnp(pipe)
pnp_utils.register_time(pipe, idx)
if "--splitk" in sys.argv:
    ops.gemm_splitk(74)
unet.truncate_dead_branches = "--truncate" in sys.argv
unet.use_cuda_graphs = "--graph" in sys.argv
x = torch.randn(B, 4, F, hw, hw, device="cuda").half()
ctx = torch.randn(B, 77, cfg["cross_attention_dim"], device="cuda").half()
for _ in range(3):
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
print(f"UNet {B}x{F}x{hw}x{hw} idx {idx} truncate={unet.truncate_dead_branches} graph={unet.use_cuda_graphs}: {ms:.2f} ms  "
      f"({(ops.launch_count - n0) // iters} launches)  {B * F * 0.9751 * (hw / 64) ** 2 / ms:.1f} TFLOP/s live  "
      f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
unet.use_cuda_graphs = False

if "--shapes" in sys.argv:
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

if "--kernels" in sys.argv:
    # device time by entry point: wrap every ops function that launches in CUDA events
    recs = collections.defaultdict(list)
    originals = {}
    for name in list(ops._LAUNCHES):
        fn = getattr(ops, name, None) or getattr(ops, name + "_", None)
        if fn is None:
            continue
        real = name if hasattr(ops, name) else name + "_"

        def wrap(fn=fn, name=name):
            def inner(*a, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                out = fn(*a, **k)
                e.record()
                recs[name].append((s, e))
                return out
            return inner
        originals[real] = fn
        setattr(ops, real, wrap())
    unet(x, 981, encoder_hidden_states=ctx)
    torch.cuda.synchronize()
    for real, fn in originals.items():
        setattr(ops, real, fn)
    tot = 0.0
    rows = []
    for name, ev in recs.items():
        t = sum(s.elapsed_time(e) for s, e in ev)
        tot += t
        rows.append((t, name, len(ev)))
    print(f"== device time by entry point (events around each call; sum {tot:.2f} ms)")
    for t, name, n in sorted(rows, reverse=True):
        print(f"  {t:8.3f} ms  {100 * t / tot:5.1f}%  n={n:4d} avg={t / n * 1e3:7.1f} us  {name}")
