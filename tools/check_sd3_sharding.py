"""SD3 MMDiT forward with its 16 frames per branch sharded over P GPUs against the single-GPU forward of the same batch.
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node=P --master-addr 127.0.0.1 tools/check_sd3_sharding.py
Prints one JSON line on rank 0.  Everything in the MMDiT is per image except the cross-frame attention, whose halo is exchanged
verbatim: the sharded result must be bit-identical on every rank, inside and outside the shift window."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from types import SimpleNamespace
from oracle import sd3_transformer_oracle as to
from univst_b200 import sd3
from univst_b200.sd3_transformer import SD3Transformer2DModel


def main():
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = to.TINY_CONFIG
    model = SD3Transformer2DModel(to.seeded_state_dict(cfg, seed=71), cfg)
    sd3.register_spatial_attention_pnp(SimpleNamespace(transformer=model))
    g = torch.Generator(device="cuda").manual_seed(3)
    BF = 48
    x = torch.randn(BF, 16, 8, 12, device="cuda", generator=g).half()
    enc = torch.randn(BF, 10, cfg["joint_attention_dim"], device="cuda", generator=g).half()
    pooled = torch.randn(BF, cfg["pooled_projection_dim"], device="cuda", generator=g).half()
    t = torch.full((BF,), 640.0, device="cuda")
    res = {"world": world}
    for idx in (5, 40):
        kw = dict(encoder_hidden_states=enc, pooled_projections=pooled, timestep=t, joint_attention_kwargs={"idx": idx})
        model.set_frame_sharding_off()
        ref = model(x, **kw).sample.clone()
        model.set_frame_sharding()
        out = model(x, **kw).sample.clone()
        out2 = model(x, **kw).sample.clone()      # second call: the other halves of the double-buffered halo banks
        model._xr.check()
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        res[f"idx{idx}"] = {"max_abs_vs_single": float((out.float() - ref.float()).abs().max()),
                            "second_call_equal": bool(torch.equal(out, out2)),
                            "ranks_agree": all(bool(torch.equal(gathered[0], g_)) for g_ in gathered)}
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
