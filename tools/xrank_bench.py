"""Micro-benchmark of the cross-rank primitives (csrc/xrank.cu) on P GPUs of one box:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node=P --master-addr 127.0.0.1 tools/xrank_bench.py
Per-call device time (CUDA events around a back-to-back loop, max over ranks): the bare synchronisation, the GroupNorm with
cross-rank statistics next to the local one, the K/V halo push to the next rank, and the frame-0 "broadcast" from rank 0 to
every rank with one store per peer vs one switch-replicated (multicast) store."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from univst_b200 import ops
from univst_b200.xrank import XRank


def timed(fn, iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t), 2)


def main():
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    rank, world = dist.get_rank(), dist.get_world_size()
    xr = XRank()
    res = {"world": world, "multicast_ok": xr.multicast_ok, "unit": "us per call"}
    res["barrier"] = timed(lambda: ops.xrank_barrier(xr), 500)
    g = torch.Generator(device="cuda").manual_seed(rank)
    for name, rows, C in (("gn_8x8_c1280", 2 * 64, 1280), ("gn_64x64_c320", 2 * 4096, 320)):
        x = torch.randn(3 * rows, C, device="cuda", generator=g).half()
        gamma, beta = torch.ones(C, device="cuda").half(), torch.zeros(C, device="cuda").half()
        res[name + "_local"] = timed(lambda: ops.groupnorm(x, gamma, beta, NB=3, rows=rows, silu=True), 200)
        res[name + "_xrank"] = timed(lambda: ops.groupnorm_xrank(x, gamma, beta, NB=3, rows=rows, xr=xr, silu=True), 200)
        t = torch.zeros(3, 32, 2, device="cuda")
        res[name + "_nccl_allreduce_only"] = timed(lambda: dist.all_reduce(t), 200)
    # K/V halo of one 64 x 64 layer: 3 branches x 4096 tokens x 640 halves = 15.7 MB
    N, C2, B = 4096, 640, 3
    src = torch.randn(B * N, C2, device="cuda", generator=g).half()
    buf, ptrs = xr.buffer("bench", (B * N, C2))
    mc = xr.multicast("bench")
    nxt = [0] * world
    if rank + 1 < world:
        nxt[rank + 1] = ptrs[rank + 1]
    res["halo_push_next_15.7MB"] = timed(lambda: ops.xrank_push(xr, [dict(src=src, src_blk_rows=N, dst=nxt, ld_dst=C2, dst_blk_rows=N, nblk=B, rows=N)] if rank + 1 < world else []), 50)
    allp = [0] + ptrs[1:]
    res["first_push_unicast_15.7MB_x%d" % (world - 1)] = timed(lambda: ops.xrank_push(xr, [dict(src=src, src_blk_rows=N, dst=allp, ld_dst=C2, dst_blk_rows=N, nblk=B, rows=N)] if rank == 0 else []), 50)
    if mc:
        res["first_push_multicast_15.7MB"] = timed(lambda: ops.xrank_push(xr, [dict(src=src, src_blk_rows=N, dst=[0] * world, ld_dst=C2, dst_blk_rows=N, nblk=B, rows=N, mc=mc)] if rank == 0 else []), 50)
        torch.cuda.synchronize()
        dist.barrier()
        want = [torch.empty_like(src) for _ in range(world)]
        dist.all_gather(want, src)
        res["multicast_data_ok"] = bool(torch.equal(buf, want[0]))
    xr.check()
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
