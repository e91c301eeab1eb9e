"""world_size-2 gloo tests (CPU) of the N > 1 host logic of bench.py: rank-seeded clips are independent (clip-parallel,
no data-path collective) and the max-over-ranks timing reduction works."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gloo_clip_sharding(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, json, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r})
        import bench
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        clip = bench.synthetic_clip(seed_offset=1000 * r)
        sig = clip["traj_c"][50].float().sum().reshape(1)
        sigs = [torch.zeros(1) for _ in range(w)]
        dist.all_gather(sigs, sig)
        t = torch.tensor([10.0 + r, 20.0 - r], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if r == 0:
            print(json.dumps({{"sigs": [float(s) for s in sigs], "max": t.tolist(), "world": w,
                              "frames": len(clip["traj_c"]), "h2d": bench.h2d_bytes(clip)}}))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["world"] == 2 and res["frames"] == 51
    assert res["sigs"][0] != res["sigs"][1], "each rank must get its own clip"
    assert res["max"] == [11.0, 20.0]
    assert res["h2d"] == 2 * 50 * 4 * 16 * 64 * 64 * 2 + 16 * 512 * 512 + 77 * 768 * 2


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
