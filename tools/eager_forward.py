"""SURVEY.md 8(d) "library kernel to beat": the same three-branch UNet call (3 x 16 x 64 x 64, shift window open / closed)
through torch fp16 eager on the same GPU -- cuDNN convolutions, cuBLAS linears, SDPA (flash) attention, i.e. what the
reference's own modules run -- next to this library's forward.  The eager graph is the oracle's restatement of the
reference forward (oracle/unet_oracle.py), so this is a measurement tool, not a product path.
Prints one JSON line; device events, median of the timed iterations."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace
from oracle import unet_oracle as uo
from univst_b200.unet import UNetPseudo3DConditionModel
from univst_b200 import pnp_utils


def timed(fn, iters):
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return out, ts[len(ts) // 2]


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    cfg = uo.SD15_CONFIG
    sd32 = uo.seeded_state_dict(cfg, seed=33)
    sd16 = {k: v.cuda().half() for k, v in sd32.items()}
    del sd32
    unet = UNetPseudo3DConditionModel(dict(sd16), cfg)
    pipe = SimpleNamespace(unet=unet)
    pnp_utils.register_spatial_attention_pnp(pipe)
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(3, 4, F, 64, 64, device="cuda", generator=g).half()
    ctx = torch.randn(3, 77, cfg["cross_attention_dim"], device="cuda", generator=g).half()
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "shape": [3, 4, F, 64, 64], "dtype": "fp16",
           "cudnn_benchmark": True}
    torch.backends.cudnn.benchmark = True
    for idx, t in ((5, 881), (30, 381)):   # shift window open / closed
        pnp_utils.register_time(pipe, idx)
        with torch.no_grad():
            ours, ms_ours = timed(lambda: unet(x, t, encoder_hidden_states=ctx).sample, iters)
            eager, ms_eager = timed(lambda: uo.unet_forward(sd16, cfg, x, t, ctx, patched=True, idx=idx), iters)
        diff = (ours.float() - eager.float()).norm() / eager.float().norm()
        res[f"idx{idx}"] = {"ms_univst_b200": round(ms_ours, 2), "ms_torch_eager_fp16": round(ms_eager, 2),
                            "speedup": round(ms_eager / ms_ours, 2), "rel_l2_between": float(diff)}
    res["peak_mem_gib"] = round(torch.cuda.max_memory_allocated() / 2**30, 1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
