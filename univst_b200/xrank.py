"""Cross-rank plumbing of the frame-sharded UNet: one peer-mapped control block per rank plus symmetric data buffers.

PyTorch supplies the memory (``torch.distributed._symmetric_memory``: allocation + rendezvous = every rank's buffer
mapped into every process over NVLink) and nothing else: signalling, waiting and the exchanges themselves are the
library's own kernels (``csrc/xrank.cu``), so a frame-sharded forward contains no collective-library call and can be
captured in a CUDA graph.  The reference is single-GPU (SURVEY.md 8e lists what frame sharding has to exchange).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class XRank:
    """Control block + symmetric buffers of one process group (ranks of one NVLink domain)."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 16:
            raise ValueError("at most 16 ranks (one NVLink domain)")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._symm_mem = symm_mem
        nbytes = int(_lib.lib().univst_xrank_ctl_bytes())
        self._ctl = symm_mem.empty(nbytes, dtype=torch.uint8, device=self.device)
        self._ctl.zero_()
        self._ctl_hdl = symm_mem.rendezvous(self._ctl, self.group)
        ptrs = [self._ctl_hdl.get_buffer(r, (nbytes,), torch.uint8).data_ptr() for r in range(self.world)]
        self.ctl = (C.c_void_p * self.world)(*ptrs)
        self._buffers = {}
        # nobody may signal into a block that its owner has not zeroed yet
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)
        torch.cuda.synchronize(self.device)
        self.multicast_ok = self._multicast_selftest()

    def _multicast_selftest(self) -> bool:
        """One multicast push from rank 0, verified on every rank: only then do later pushes use switch-replicated stores."""
        import torch.distributed as dist
        from . import ops
        self.multicast_ok = True
        t, ptrs = self.buffer("_mc_selftest", (64, 64))
        t.zero_()
        mc = self.multicast("_mc_selftest")
        ok = torch.ones(1, device=self.device)
        try:
            torch.cuda.synchronize(self.device)
            dist.barrier(self.group)
            pushes = []
            if self.rank == 0 and mc:
                src = (torch.arange(64 * 64, device=self.device) % 251).to(torch.float16).view(64, 64)
                pushes = [dict(src=src, src_blk_rows=64, dst=[0] * self.world, ld_dst=64, dst_blk_rows=64, nblk=1, rows=64, mc=mc)]
            ops.xrank_push(self, pushes)
            torch.cuda.synchronize(self.device)
            want = (torch.arange(64 * 64, device=self.device) % 251).to(torch.float16).view(64, 64)
            if not mc or not torch.equal(t, want):
                ok.zero_()
        except Exception:  # noqa: BLE001
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        return bool(ok.item() > 0)

    def buffer(self, key, shape, dtype=torch.float16):
        """Symmetric buffer ``key`` of ``shape`` (same call sequence on every rank): (local tensor, [device pointer of every
        rank's copy])."""
        ent = self._entry(key, shape, dtype)
        return ent[0], ent[1]

    def multicast(self, key) -> int:
        """Multicast address of an existing symmetric buffer (a store to it lands at the same offset on every rank, replicated
        by the NVSwitch) or 0 where the fabric has no multicast support (or UNIVST_XRANK_MULTICAST=0)."""
        return self._buffers[key][3]

    def _entry(self, key, shape, dtype):
        ent = self._buffers.get(key)
        if ent is None:
            import os
            t = self._symm_mem.empty(*shape, dtype=dtype, device=self.device)
            hdl = self._symm_mem.rendezvous(t, self.group)
            ptrs = [hdl.get_buffer(r, tuple(shape), dtype).data_ptr() for r in range(self.world)]
            mc = 0
            if os.environ.get("UNIVST_XRANK_MULTICAST", "1") != "0":
                try:   # multicast_ptr follows the convention of buffer_ptrs (allocation base): add the tensor's offset
                    base = int(hdl.multicast_ptr or 0)
                    mc = base + (t.data_ptr() - int(hdl.buffer_ptrs[self.rank])) if base else 0
                except Exception:  # noqa: BLE001
                    mc = 0
            if not getattr(self, "multicast_ok", True):
                mc = 0
            ent = self._buffers[key] = (t, ptrs, hdl, mc)
        return ent

    def error(self) -> int:
        """0, or 1 + the rank a wait timed out on (sticky).  Synchronises the device."""
        torch.cuda.synchronize(self.device)
        return int(self._ctl[72:76].view(torch.int32).item())

    def wait_stats(self):
        """(synchronisations completed, milliseconds this rank has spent waiting for its peers inside them) since start-up
        -- the device-side counters of the control block.  Synchronises the device."""
        torch.cuda.synchronize(self.device)
        ns = int(self._ctl[80:88].view(torch.int64).item())
        return int(self._ctl[88:92].view(torch.int32).item()), ns / 1e6

    def check(self):
        e = self.error()
        if e:
            raise RuntimeError(f"cross-rank wait timed out on rank {self.rank} waiting for rank {e - 1}: the results of "
                               "this process group are invalid")


def push_kv_halo(xr: XRank, qkv: torch.Tensor, ptrs, mc: int, B: int, F: int, N: int, C: int):
    """K/V halo of a frame-sharded sparse-causal / cross-frame attention.  ``qkv``: this rank's fused projection buffer
    [(B F + 2 B) N, 3 C] in symmetric memory (local images, then bank 1 = "previous frame of my first frame", bank 2 = "first
    frame of the clip", B images each); ``ptrs`` / ``mc``: every rank's copy / the multicast address of the same buffer.
    One launch: K|V columns (C .. 3C) of my last frame -> bank 1 of rank + 1; rank 0: its first frame -> bank 2 of every other
    rank (one switch-replicated store where the fabric multicasts); the kernel's tail is the cross-rank synchronisation."""
    from . import ops
    rank, world = xr.rank, xr.world
    NI, ld = B * F, qkv.stride(0)
    kv = qkv[:, C:]
    pushes = []
    if rank + 1 < world:
        dst = [0] * world
        dst[rank + 1] = ptrs[rank + 1] + (NI * N * ld + C) * 2
        pushes.append(dict(src=kv[(F - 1) * N:], src_blk_rows=F * N, dst=dst, ld_dst=ld, dst_blk_rows=N, nblk=B, rows=N))
    if rank == 0:
        off = ((NI + B) * N * ld + C) * 2
        pushes.append(dict(src=kv, src_blk_rows=F * N, dst=[0] + [p + off for p in ptrs[1:]], ld_dst=ld, dst_blk_rows=N, nblk=B,
                           rows=N, mc=mc + off if mc else 0))
    ops.xrank_push(xr, pushes)


class HaloBuffers:
    """Double-buffered symmetric projection buffers with halo banks, one pair per (rows, cols) shape: a buffer is rewritten
    only two synchronisations after it was last read (see unet.py::_xr_halo)."""

    def __init__(self, xr: XRank, tag: str):
        self.xr, self.tag, self._par = xr, tag, {}

    def next(self, rows: int, cols: int):
        key = (self.tag, rows, cols)
        par = self._par.get(key, 0)
        self._par[key] = par ^ 1
        t, ptrs = self.xr.buffer(key, (2, rows, cols))
        mc = self.xr.multicast(key)
        off = par * rows * cols * 2
        return t[par], [p + off for p in ptrs], (mc + off if mc else 0)
