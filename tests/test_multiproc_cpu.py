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


def test_sharded_source_tables_match_global_table():
    """A frame shard's K/V source table (local images + two halo banks) names the same frames as the global one."""
    from univst_b200.unet import kv_source_table, kv_source_table_sharded
    B, F = 3, 8
    for P in (2, 4, 8):
        Fl = F // P
        NI = B * Fl
        for mode in ("prev_first", "prev_self_first", "self"):
            g = kv_source_table(B, F, mode)
            for r in range(P):
                t = kv_source_table_sharded(B, Fl, mode, r)

                def glob(i):
                    if i < NI:
                        b, fl = divmod(i, Fl)
                        return b * F + r * Fl + fl
                    i -= NI
                    if i < B:
                        return i * F + r * Fl - 1      # halo bank 1: last frame of the previous rank
                    return (i - B) * F                 # halo bank 2: frame 0 of the clip
                for b in range(B):
                    for fl in range(Fl):
                        assert [glob(i) for i in t[b * Fl + fl].tolist()] == g[b * F + r * Fl + fl].tolist()


def test_two_rank_gloo_kv_halo_exchange(tmp_path):
    """The per-layer K/V halo exchange of the frame-sharded UNet (boundary frame to the next rank, frame 0 broadcast)."""
    script = tmp_path / "h.py"
    script.write_text(textwrap.dedent(f"""
        import sys, json, torch, torch.distributed as dist
        from types import SimpleNamespace
        sys.path.insert(0, {ROOT!r})
        from univst_b200.unet import UNetPseudo3DConditionModel as U
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        B, Fl, N, C3 = 3, 2, 5, 6
        NI = B * Fl
        qkv = torch.zeros((NI + 2 * B) * N, C3)
        # value of local image (b, fl) on rank r: 100 r + 10 b + fl
        for b in range(B):
            for fl in range(Fl):
                qkv[(b * Fl + fl) * N:(b * Fl + fl + 1) * N] = 100 * r + 10 * b + fl
        U._exchange_kv_halo(SimpleNamespace(_shard=(None, r, w)), qkv, B, Fl, N)
        prev = [float(qkv[(NI + b) * N, 0]) for b in range(B)]
        first = [float(qkv[(NI + B + b) * N, 0]) for b in range(B)]
        out = [None] * w
        dist.all_gather_object(out, {{"prev": prev, "first": first}})
        if r == 0:
            print(json.dumps(out))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("[")][-1])
    assert res[0]["first"] == [0.0, 10.0, 20.0] and res[1]["first"] == [0.0, 10.0, 20.0]   # frame 0 of rank 0
    assert res[1]["prev"] == [1.0, 11.0, 21.0]                                             # last frame of rank 0
    assert res[0]["prev"] == [0.0, 0.0, 0.0]                                               # rank 0 has no predecessor


def test_two_rank_gloo_frames_pixels_all_to_all(tmp_path):
    """The AnimateDiff motion-module exchange: frames <-> pixels all-to-all and its inverse, world size 2 on gloo."""
    script = tmp_path / "a.py"
    script.write_text(textwrap.dedent(f"""
        import sys, json, torch, torch.distributed as dist
        sys.path.insert(0, {ROOT!r})
        from univst_b200.animatediff import frames_to_pixels, pixels_to_frames
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        B, Fl, N, C = 3, 2, 6, 4
        F = Fl * w
        # global tensor value at (b, f, pix, c) = 1000 b + 100 f + 10 pix + c; this rank owns frames [r Fl, (r+1) Fl)
        b, f, p, c = torch.meshgrid(torch.arange(B), torch.arange(F), torch.arange(N), torch.arange(C), indexing="ij")
        glob = (1000 * b + 100 * f + 10 * p + c).float()
        mine = glob[:, r * Fl:(r + 1) * Fl].reshape(B * Fl * N, C).contiguous()
        got = frames_to_pixels(mine, B, Fl, N, w)
        n = N // w
        want = glob[:, :, r * n:(r + 1) * n].reshape(B * F * n, C)
        ok1 = bool(torch.equal(got, want))
        back = pixels_to_frames(got, B, Fl, N, w)
        ok2 = bool(torch.equal(back, mine))
        out = [None] * w
        dist.all_gather_object(out, [ok1, ok2])
        if r == 0:
            print(json.dumps(out))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29619", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("[")][-1])
    assert res == [[True, True], [True, True]]


def test_pushed_kv_halo_offsets_for_2_to_4_ranks(monkeypatch):
    """Host side of the pushed K/V halo (unet._push_kv_halo): with ``ops.halo_push`` replaced by a byte-exact emulation
    that treats the destination pointers as addresses inside the ranks' (CPU) buffers, every rank's two halo banks must end
    up holding the K|V columns of the previous rank's last frame and of the clip's first frame -- for 2, 3 and 4 ranks."""
    import torch
    from types import SimpleNamespace
    from univst_b200 import ops
    from univst_b200.unet import UNetPseudo3DConditionModel

    B, Fl, N, C = 3, 2, 5, 8
    ld = 3 * C
    for P in (2, 3, 4):
        NI = B * Fl
        g = torch.Generator().manual_seed(P)
        clip = torch.randn(B, P * Fl, N, ld, generator=g).half()           # the projection of the whole clip
        bufs = [torch.zeros((NI + 2 * B) * N, ld, dtype=torch.float16) for _ in range(P)]
        for r in range(P):
            bufs[r][: NI * N] = clip[:, r * Fl:(r + 1) * Fl].reshape(NI * N, ld)
        bases = [b.data_ptr() for b in bufs]

        def fake_halo_push(src, src_blk_rows, dst_ptrs, ld_dst, dst_blk_rows, nblk, rows):
            cols = src.shape[1]
            for r, p in enumerate(dst_ptrs):
                if not p:
                    continue
                owner = next(i for i in range(P) if bases[i] <= p < bases[i] + bufs[i].numel() * 2)   # whose buffer
                off = (p - bases[owner]) // 2
                flat = bufs[owner].view(-1)
                for blk in range(nblk):
                    for row in range(rows):
                        d0 = off + (blk * dst_blk_rows + row) * ld_dst
                        flat[d0:d0 + cols] = src[blk * src_blk_rows + row]

        monkeypatch.setattr(ops, "halo_push", fake_halo_push)
        hdl = SimpleNamespace(barrier=lambda channel=0: None)
        for r in range(P):
            fake_self = SimpleNamespace(_shard=(None, r, P))
            UNetPseudo3DConditionModel._push_kv_halo(fake_self, bufs[r], bases, hdl, B, Fl, N, C)
        for r in range(P):
            prev = bufs[r][NI * N:(NI + B) * N].view(B, N, ld)
            first = bufs[r][(NI + B) * N:].view(B, N, ld)
            assert torch.equal(first[:, :, C:], clip[:, 0, :, C:]), (P, r)
            assert torch.all(first[:, :, :C] == 0)                        # the Q columns are not sent
            if r > 0:
                assert torch.equal(prev[:, :, C:], clip[:, r * Fl - 1, :, C:]), (P, r)
            else:
                assert torch.all(prev == 0)


def test_pushed_frames_pixels_exchange_for_2_to_4_ranks(monkeypatch):
    """Host side of the AnimateDiff push exchange (animatediff._exchange: arena allocation, per-direction buffer offsets,
    peer pointer list) with ``ops.exchange_push`` replaced by a Python restatement of the kernel's index maps and torch
    symmetric memory by per-rank CPU tensors: after every rank has pushed, each rank must hold "all frames, my pixels" in
    (b, frame, pixel) order, and the inverse exchange must give back "my frames, all pixels" -- for 2, 3 and 4 ranks."""
    import torch
    import torch.distributed._symmetric_memory as symm_mem
    from types import SimpleNamespace
    from univst_b200 import ops
    from univst_b200.animatediff import UNet3DConditionModel

    B, Fl, n, C = 3, 2, 4, 8
    for P, xrank in ((2, False), (3, False), (4, False), (2, True), (4, True)):
        N = n * P
        g = torch.Generator().manual_seed(10 + P)
        clip = torch.randn(B, P * Fl, N, C, generator=g).half()
        arenas, cur = {}, {"rank": 0}

        def fake_empty(*shape, dtype=None, device=None):
            if cur["rank"] not in arenas:   # a peer that pushed earlier in this sequential emulation has already mapped it
                arenas[cur["rank"]] = torch.zeros(*shape, dtype=dtype)
            return arenas[cur["rank"]]

        def fake_rendezvous(t, group):
            def get_buffer(r, shape, dtype):
                if r not in arenas:
                    arenas[r] = torch.zeros(*shape, dtype=dtype)
                return arenas[r]
            return SimpleNamespace(get_buffer=get_buffer, barrier=lambda channel=0: None)

        def fake_exchange_push(direction, src, dst_ptrs, rank, P_, B_, Fl_, N_, xr=None):
            n_ = N_ // P_
            flats = {r: arenas[r].view(-1) for r in range(P_)}
            base = {r: arenas[r].data_ptr() for r in range(P_)}
            for row in range(src.shape[0]):
                if direction == 0:      # rows (b, fl, pix) -> rank pix / n, row (b, rank Fl + fl, pix % n)
                    pix, bf = row % N_, row // N_
                    fl, b = bf % Fl_, bf // Fl_
                    owner = pix // n_
                    drow = (b * (P_ * Fl_) + rank * Fl_ + fl) * n_ + pix % n_
                else:                   # rows (b, f_global, pix_local) -> rank f_global / Fl, row (b, f % Fl, rank n + pix_local)
                    pl, bf = row % n_, row // n_
                    fg, b = bf % (P_ * Fl_), bf // (P_ * Fl_)
                    owner = fg // Fl_
                    drow = (b * Fl_ + fg % Fl_) * N_ + rank * n_ + pl
                off = (dst_ptrs[owner] - base[owner]) // 2 + drow * src.shape[1]
                flats[owner][off:off + src.shape[1]] = src[row]

        monkeypatch.setattr(symm_mem, "empty", fake_empty)
        monkeypatch.setattr(symm_mem, "rendezvous", fake_rendezvous)
        monkeypatch.setattr(ops, "exchange_push", fake_exchange_push)
        ranks = [SimpleNamespace(_shard=(object(), r, P), _push=True, _arena=None, _xr=None, device="cpu") for r in range(P)]
        if xrank:   # the xrank transport: buffers come from XRank.buffer (same arenas), the kernel's tail synchronises
            def fake_buffer(key, shape, dtype=torch.float16):
                for r in range(P):
                    if r not in arenas:
                        arenas[r] = torch.zeros(*shape, dtype=dtype)
                return arenas[cur["rank"]], [arenas[r].data_ptr() for r in range(P)]
            for ns in ranks:
                ns._xr, ns._push = SimpleNamespace(buffer=fake_buffer), False

        def exchange(direction, ys):
            outs = []
            for r in range(P):
                cur["rank"] = r
                outs.append(UNet3DConditionModel._exchange(ranks[r], direction, ys[r], B, Fl, N))
            return outs

        local = [clip[:, r * Fl:(r + 1) * Fl].reshape(B * Fl * N, C).contiguous() for r in range(P)]
        pix = exchange(0, local)                      # views of the arenas: complete once every rank has pushed
        for r in range(P):
            assert torch.equal(pix[r], clip[:, :, r * n:(r + 1) * n].reshape(-1, C)), (P, r)
        back = exchange(1, [p.clone() for p in pix])
        for r in range(P):
            assert torch.equal(back[r], local[r]), (P, r)
            assert back[r].data_ptr() != pix[r].data_ptr()   # one buffer per direction


def _run_sharded_forward(tmp_path, world, backbone, port):
    script = tmp_path / f"fs_{backbone}_{world}.py"
    script.write_text(textwrap.dedent(f"""
        import sys, json, torch, torch.distributed as dist
        from types import SimpleNamespace
        sys.path.insert(0, {ROOT!r})
        sys.path.insert(0, {os.path.join(ROOT, "tests")!r})
        import _torch_ops
        _torch_ops.install()
        from univst_b200 import pnp_utils
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        if {backbone!r} == "animatediff":
            from oracle import animatediff_oracle as ao
            from univst_b200.animatediff import UNet3DConditionModel as U
            cfg, sd, kw = ao.AD_TINY_CONFIG, ao.seeded_state_dict(ao.AD_TINY_CONFIG, seed=33), dict(push_exchange=False)
        else:
            from oracle import unet_oracle as uo
            from univst_b200.unet import UNetPseudo3DConditionModel as U
            cfg, sd, kw = uo.TINY_CONFIG, uo.seeded_state_dict(uo.TINY_CONFIG, seed=33), dict(push_halo=False)
        unet = U(sd, cfg, device="cpu")
        pipe = SimpleNamespace(unet=unet)
        pnp_utils.register_spatial_attention_pnp(pipe)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(3, 4, 4, 16, 16, generator=g).half()
        ctx = torch.randn(3, 77, cfg["cross_attention_dim"], generator=g).half()
        res = {{}}
        for idx, t in ((5, 881), (30, 381)):   # shift window open / closed
            pnp_utils.register_time(pipe, idx)
            unet.set_frame_sharding_off()
            ref = unet(x, t, encoder_hidden_states=ctx).sample.float()
            unet.set_frame_sharding(**kw)
            out = unet(x, t, encoder_hidden_states=ctx).sample.float()
            res[str(idx)] = float((out - ref).norm() / ref.norm())
        allres = [None] * w
        dist.all_gather_object(allres, res)
        if r == 0:
            print(json.dumps(allres))
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    import json
    return json.loads([l for l in out.stdout.splitlines() if l.startswith("[")][-1])


def test_frame_sharded_unet_forward_on_gloo(tmp_path):
    """The WHOLE frame-sharded forward of the SD UNet mirror -- frame slicing, sharded source tables, K/V halo exchange, cross-rank
    GroupNorm, the final all-gather -- on 2 and on 4 CPU ranks (gloo), kernels replaced by torch definitions
    (tests/_torch_ops.py): every rank must return the single-process result for the whole clip (up to the order of the GroupNorm
    partial sums and fp16 storage)."""
    for world, port in ((2, 29671), (4, 29673)):
        res = _run_sharded_forward(tmp_path, world, "sd", port)
        assert len(res) == world
        for r in res:
            assert r["5"] < 5e-3 and r["30"] < 5e-3, (world, res)


def test_frame_sharded_animatediff_forward_on_gloo(tmp_path):
    """The same for the AnimateDiff mirror (per-frame GroupNorm and attention; the temporal transformer block runs
    pixel-sharded between two all-to-alls per motion module).  With the real kernels the sharded result is bit-identical
    (checked on 2 GPUs, tests/test_frame_sharding_gpu.py); the CPU matmuls behind the torch definitions block differently for
    different row counts, so here the bar is the fp16 noise floor."""
    for world, port in ((2, 29675), (4, 29677)):
        res = _run_sharded_forward(tmp_path, world, "animatediff", port)
        assert len(res) == world
        for r in res:
            assert r["5"] < 5e-3 and r["30"] < 5e-3, (world, res)
