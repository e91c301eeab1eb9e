"""Frame-sharded UNet forward (one clip over P GPUs) against the single-GPU forward of the same clip.
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node=P --master-addr 127.0.0.1 tools/check_frame_sharding.py [F]
Prints one JSON line on rank 0: max abs / rel-L2 difference, per-forward time of both modes (device events, max over ranks)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from types import SimpleNamespace
from oracle import unet_oracle as uo
from univst_b200.unet import UNetPseudo3DConditionModel
from univst_b200 import pnp_utils


AD = "--animatediff" in sys.argv


def build_unet(cfg):
    g = torch.Generator(device="cuda").manual_seed(33)
    sd = {}
    if AD:
        from oracle import animatediff_oracle as ao
        from univst_b200.animatediff import UNet3DConditionModel
        shapes, cls = ao.unet_param_shapes(cfg), UNet3DConditionModel
    else:
        shapes, cls = uo.unet_param_shapes(cfg), UNetPseudo3DConditionModel
    for k, s in shapes.items():
        if "attn_temporal.to_out.0.weight" in k or "conv_temporal.bias" in k:
            sd[k] = torch.zeros(s, device="cuda", dtype=torch.float16)
        elif "conv_temporal.weight" in k:   # the Dirac identity of the pseudo-3D inflation (resnet.py:54)
            sd[k] = torch.zeros(s, device="cuda", dtype=torch.float16)
            i = torch.arange(min(s[0], s[1]))
            sd[k][i, i, s[2] // 2] = 1
        elif k.endswith("weight") and len(s) == 1:
            sd[k] = torch.ones(s, device="cuda", dtype=torch.float16)
        elif k.endswith("bias"):
            sd[k] = (0.02 * torch.randn(s, device="cuda", generator=g)).half()
        else:
            fan = 1
            for d in s[1:]:
                fan *= d
            sd[k] = (torch.randn(s, device="cuda", generator=g) * fan ** -0.5).half()
    return cls(sd, cfg)


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t)


def op_profile(fn):
    """Device time by entry point of ``univst_b200.ops`` for one call of ``fn`` (CUDA events around every op call)."""
    import collections
    from univst_b200 import ops
    recs, originals = collections.defaultdict(list), {}
    for name in list(ops._LAUNCHES):
        real = name if hasattr(ops, name) else name + "_"
        f = getattr(ops, real, None)
        if f is None:
            continue

        def wrap(f=f, name=name):
            def inner(*a, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                out = f(*a, **k)
                e.record()
                recs[name].append((s, e))
                return out
            return inner
        originals[real] = f
        setattr(ops, real, wrap())
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        for real, f in originals.items():
            setattr(ops, real, f)
    return {name: [round(sum(s.elapsed_time(e) for s, e in ev), 3), len(ev)] for name, ev in recs.items()}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    F = int(args[0]) if len(args) > 0 else 16
    hw = int(args[1]) if len(args) > 1 else 64
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    if AD:
        from oracle import animatediff_oracle as ao
        unet = build_unet(ao.AD_SD15_CONFIG)
    else:
        unet = build_unet(uo.SD15_CONFIG)
    pipe = SimpleNamespace(unet=unet)
    pnp_utils.register_spatial_attention_pnp(pipe)
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(3, 4, F, hw, hw, device="cuda", generator=g).half()
    ctx = torch.randn(3, 77, 768, device="cuda", generator=g).half()
    res = {"world": world, "frames": F, "latent": hw, "backbone": "animatediff" if AD else "sd"}
    for idx, t in ((5, 881), (30, 381)):   # shift window open / closed
        pnp_utils.register_time(pipe, idx)
        unet.set_frame_sharding_off()
        ref, t_single = timed(lambda: unet(x, t, encoder_hidden_states=ctx).sample, 2)
        unet.set_frame_sharding(**({"push_exchange": False} if AD else {"push_halo": False}))   # the NCCL exchanges
        out, t_shard = timed(lambda: unet(x, t, encoder_hidden_states=ctx).sample, 2)
        diff = (out.float() - ref.float())
        res[f"idx{idx}"] = {"max_abs": float(diff.abs().max()), "rel_l2": float(diff.norm() / ref.float().norm()),
                            "ms_single": t_single, "ms_sharded": t_shard}
        if "--push" in sys.argv and AD:   # frames <-> pixels exchange stored straight into the peers' symmetric memory
            unet.set_frame_sharding(push_exchange=True)
            out_p, t_push = timed(lambda: unet(x, t, encoder_hidden_states=ctx).sample, 2)
            res[f"idx{idx}"].update({"ms_sharded_push_exchange": t_push,
                                     "push_vs_single_max_abs": float((out_p.float() - ref.float()).abs().max())})
        if "--push" in sys.argv and not AD:   # K/V halo stored straight into the peers' symmetric-memory banks
            unet.set_frame_sharding(push_halo=True)
            out_p, t_push = timed(lambda: unet(x, t, encoder_hidden_states=ctx).sample, 2)
            res[f"idx{idx}"].update({"ms_sharded_push_halo": t_push,
                                     "push_vs_nccl_max_abs": float((out_p.float() - out.float()).abs().max())})
        if "--xrank" in sys.argv:   # everything through peer memory + device-side flags (csrc/xrank.cu), eager and as a CUDA graph
            unet.set_frame_sharding(transport="xrank")
            out_x, t_x = timed(lambda: unet(x, t, encoder_hidden_states=ctx).sample, 3)
            unet.use_cuda_graphs = True
            unet(x, t, encoder_hidden_states=ctx)                     # capture
            n0, w0 = unet._xr.wait_stats()
            out_g, t_g = timed(lambda: unet(x, t, encoder_hidden_states=ctx).sample, 5)
            n1, w1 = unet._xr.wait_stats()
            unet.use_cuda_graphs = False
            unet._xr.check()
            wt = torch.tensor([(w1 - w0) / 6.0], device="cuda")       # timed() runs the call 1 + 5 times
            wl = [torch.zeros_like(wt) for _ in range(world)]
            dist.all_gather(wl, wt)
            res[f"idx{idx}"].update({"xrank_syncs_per_forward": (n1 - n0) // 6,
                                     "xrank_wait_ms_per_forward_by_rank": [round(float(w), 3) for w in wl]})
            dx = out_x.float() - ref.float()
            res[f"idx{idx}"].update({"ms_sharded_xrank": t_x, "ms_sharded_xrank_graph": t_g,
                                     "xrank_vs_single_max_abs": float(dx.abs().max()),
                                     "xrank_vs_single_rel_l2": float(dx.norm() / ref.float().norm()),
                                     "xrank_graph_vs_eager_max_abs": float((out_g.float() - out_x.float()).abs().max())})
            if "--profile" in sys.argv:   # where a rank's time goes: per entry point, this rank's shard vs the same shard alone
                fn = lambda: unet(x, t, encoder_hidden_states=ctx)
                fn()
                prof = op_profile(fn)
                unet.set_frame_sharding_off()
                Fl = F // world
                xs = x[:, :, rank * Fl:(rank + 1) * Fl].contiguous()
                from univst_b200 import ops as _ops
                _ops.gemm_splitk(74 if world >= 4 else 0)
                fn1 = lambda: unet(xs, t, encoder_hidden_states=ctx)
                fn1()
                prof1 = op_profile(fn1)
                _ops.gemm_splitk(0)
                unet.set_frame_sharding(transport="xrank")
                pl = [None] * world
                dist.all_gather_object(pl, {"sharded": prof, "alone": prof1})
                res[f"idx{idx}"]["op_ms_by_rank"] = pl
            # all ranks must hold the same result (the statistics are added in the same order everywhere)
            gathered = [torch.empty_like(out_x) for _ in range(world)]
            dist.all_gather(gathered, out_x.contiguous())
            res[f"idx{idx}"]["xrank_ranks_agree"] = all(bool(torch.equal(gathered[0], g_)) for g_ in gathered)
        if "--fused" in sys.argv and not AD:   # K/V halo read from peer memory by the attention kernel itself
            unet.set_frame_sharding(fused_halo=True)
            out_f, t_fused = timed(lambda: unet(x, t, encoder_hidden_states=ctx).sample, 2)
            res[f"idx{idx}"].update({"ms_sharded_fused_halo": t_fused,
                                     "fused_vs_nccl_max_abs": float((out_f.float() - out.float()).abs().max())})
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
