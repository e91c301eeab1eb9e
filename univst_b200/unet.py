"""B200 host-side mirror of the reference's ``UNetPseudo3DConditionModel``
(backbones/video_diffusion_sd/models/unet_3d_condition.py:45, forward :306-443) for the SD-1.5 / SD-2.1 backbones.

Same call signature, same ``state_dict`` key names, same monkey-patch protocol on
``unet.up_blocks[r].attentions[b].transformer_blocks[0].attn1`` (``idx`` / ``eta1`` / ``eta2`` attributes and an
instance-level ``forward`` override installed by ``register_spatial_attention_pnp``, pnp_utils.py:7-15,104-111),
but every arithmetic operation runs in the sm_100a kernels behind ``include/univst_b200.h``:

* activations live channels-last ``[(b f), h, w, C]`` fp16 for the whole forward; the reference's
  ``(b f) c h w <-> b c f h w`` / ``b (h w) c`` rearranges do not exist;
* 3x3 convolutions are implicit GEMMs on tcgen05 (TMA taps, zero-fill padding), 1x1 convolutions and all Linear
  layers are the same GEMM; bias / time-embedding / residual / GEGLU / the dead temporal attention's bias are
  epilogues of the producing GEMM;
* the skip-connection concat exists only as the output of the fused GroupNorm+SiLU and as a second K-range of the
  shortcut GEMM;
* attn1 is one fused kernel over K/V tiles fetched from the neighbour frames (no concatenated K/V), preceded --
  while the AdaIN-guided shift is active -- by an in-place shift of the edit branch's Q/K/V;
* exact simplifications: the temporal Conv1d is the identity and is skipped; the temporal
  attention has a zero output projection, so it is its output bias (verified at pack time, else refused).

There is no fallback path: without the CUDA library or off sm_100 every call raises.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from . import ops
from .pack import pack_conv3x3, pack_geglu

SD15_CONFIG = dict(block_out_channels=(320, 640, 1280, 1280), attention_head_dim=8, cross_attention_dim=768,
                   layers_per_block=2, norm_num_groups=32, norm_eps=1e-5, in_channels=4, out_channels=4,
                   use_linear_projection=False)
SD21_CONFIG = dict(SD15_CONFIG, attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024, use_linear_projection=True)

CIN_PAD = 64  # conv_in reads a latent zero-padded to one 128-byte swizzle row of channels


class _Config(dict):
    """The UNet configuration with diffusers' FrozenDict-style access: ``config["x"]`` and ``config.x``."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None


def kv_source_table(B: int, F: int, mode: str) -> torch.Tensor:
    """int32 [B*F, nsrc]: K/V source images of image (b, f) for the sparse-causal modes of the reference:
    ``prev_first`` = SparseCausalAttention_index [-1, 'first'] (patched attn1, pnp_utils.py:25),
    ``prev_self_first`` = [-1, 0, 'first'] (stock, models/attention.py:356), ``self`` = plain self-attention,
    ``branch`` = the per-branch 77-token context of the cross-attention."""
    rows = []
    for b in range(B):
        for f in range(F):
            prev, first, me = b * F + max(f - 1, 0), b * F, b * F + f
            rows.append({"prev_first": [prev, first], "prev_self_first": [prev, me, first], "self": [me],
                         "branch": [b]}[mode])
    return torch.tensor(rows, dtype=torch.int32)


def kv_source_table_sharded(B: int, Fl: int, mode: str, rank: int) -> torch.Tensor:
    """Source table of a frame shard (local frames [rank * Fl, (rank + 1) * Fl) of every branch).  Local images are
    0 .. B*Fl-1; two halo banks follow in the K/V buffer: B*Fl + b = last frame of the previous rank (branch b),
    B*Fl + B + b = frame 0 of the clip (owned by rank 0).  Rank 0 needs neither (frame -1 clips to frame 0)."""
    NI = B * Fl
    rows = []
    for b in range(B):
        for fl in range(Fl):
            me = b * Fl + fl
            prev = me - 1 if fl > 0 else (NI + b if rank > 0 else me)
            first = b * Fl if rank == 0 else NI + B + b
            rows.append({"prev_first": [prev, first], "prev_self_first": [prev, me, first], "self": [me],
                         "branch": [b]}[mode])
    return torch.tensor(rows, dtype=torch.int32)


class UNetPseudo3DConditionOutput(dict):
    """Attribute *and* item access, like diffusers' BaseOutput (ddim_inversion.py:209-211 uses ``["sample"]``)."""

    def __init__(self, sample):
        super().__init__(sample=sample)
        self.sample = sample


class _AttnHandle:
    """What the reference's ``register_time`` / ``register_spatial_attention_pnp`` touch on attn1 / attn2."""

    def __init__(self, heads: int):
        self.heads = heads
        self.group_norm = None
        self.idx = None
        self.eta1 = 0.0
        self.eta2 = 0.5

    def forward(self, *a, **k):  # placeholder: an instance-level override marks the layer as patched
        raise RuntimeError("attention runs inside the fused UNet forward; this handle only carries patch state")

    @property
    def patched(self) -> bool:
        return "forward" in self.__dict__ or self.__dict__.get("_patched", False)


class _Block:
    def __init__(self, heads):
        self.attn1 = _AttnHandle(heads)
        self.attn2 = _AttnHandle(heads)


class _Transformer:
    def __init__(self, prefix, heads):
        self.prefix = prefix
        self.heads = heads
        self.transformer_blocks = [_Block(heads)]


class _UNetBlock:
    def __init__(self):
        self.attentions: List[_Transformer] = []
        self.resnets: List[str] = []
        self.has_cross_attention = False


class UNetPseudo3DConditionModel:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: Optional[dict] = None, device="cuda"):
        cfg = dict(SD15_CONFIG)
        cfg.update(config or {})
        self.config = _Config(cfg)   # item and attribute access (stable_diffusion.py:572 reads unet.config.in_channels)
        self.device = torch.device(device)
        self.dtype = torch.float16
        boc = cfg["block_out_channels"]
        self.nlev = len(boc)
        self._tables = {}
        self._ctx_cache = None
        self._shard = None  # (process group, rank, world) when frames are sharded over GPUs
        self._fused_halo = False
        self._push_halo = False
        # SURVEY 2.3 D3, second half: while the shift is live the content / style branches only feed Q_content and
        # K/V_style to the patched attn1 layers, so after the LAST live patched layer's projection (+ shift) they are dead.
        # When set (the pipeline sets it together with skip_dead_branches) a three-branch call evaluates the rest of the
        # network on the edit branch alone and ``.sample`` holds that branch only: (1, C, F, h, w).
        self.truncate_dead_branches = False
        self.last_edit_branch = None   # index of the edit branch inside ``last_eps_rows`` after a call
        self._xr = None                # cross-rank control block + symmetric buffers (xrank.XRank) of the sharded forward
        # CUDA graphs: one captured forward per (shapes, attention plan, sharding) key, replayed with the step's inputs
        # copied into static buffers and its scalars (timestep, shift parameters) written to device memory
        self.use_cuda_graphs = False
        self._graphs, self._graph_pool, self._graph_stream = {}, None, None
        self._step_params = None       # device float32 [64] while a graph-safe forward is being issued
        self._build_tree()
        for i, tr in enumerate(self._all_transformers()):
            tr.index = i
        self._pack(state_dict)

    # ------------------------------------------------------------------------------------------ construction
    @classmethod
    def from_reference(cls, module, device="cuda"):
        """Build from a reference ``UNetPseudo3DConditionModel`` nn.Module (weights are copied and packed once)."""
        c = module.config
        cfg = {k: c[k] for k in SD15_CONFIG if k in c}
        if "sample_size" in c:
            cfg["sample_size"] = c["sample_size"]
        return cls(module.state_dict(), cfg, device=device)

    def set_frame_sharding(self, group=None, fused_halo: bool = False, push_halo=None, transport=None, split_k=None):
        """Shard the frames of every clip over the ranks of ``group`` (default: the world group): rank r evaluates
        frames [r F/P, (r+1) F/P) of all branches.  What crosses ranks: B x 32 x 2 partial sums per cross-frame
        GroupNorm, the K/V of the neighbouring frames that attn1 needs (last frame of the previous rank, frame 0 of the
        clip), the predicted noise at the end.

        ``transport="xrank"`` (the default on NCCL groups when neither ``push_halo`` nor ``fused_halo`` is given): all of
        it goes through peer-mapped memory with the library's own kernels (csrc/xrank.cu) -- the halo is stored into
        the peers' banks and the GroupNorm sums into the peers' control blocks by the kernel that produced them, whose
        tail is the cross-rank synchronisation; the noise prediction is stored into every rank's full-clip buffer.  No
        collective-library call is left in the forward (any world size up to 16), so it can be captured in a CUDA graph.

        ``split_k`` (default: on from 4 ranks up): shards of a few images leave the deep UNet levels with a handful of
        output tiles per GEMM / conv; their K loops are then split over the idle SMs (``ops.gemm_splitk``) --
        deterministic, but the fp32 summation order, hence the last bits of some fp16 outputs, differ from the unsplit
        kernel's.

        ``transport="nccl"`` keeps the collectives of the first implementation (GroupNorm all-reduce, noise all-gather;
        works on any backend, e.g. gloo), with the halo exchanged in one of these ways:
        * ``push_halo=True`` (the default on NCCL groups): the fused projection lives in torch symmetric memory with two
          halo banks behind the local images; one kernel stores the K|V columns of the boundary frame into the next
          rank's bank, and rank 0's first frame into every rank's bank, over NVLink (``univst_halo_push_f16``), followed
          by one cross-rank barrier (double buffering makes a second one unnecessary);
        * ``push_halo=False``: the same banks filled by NCCL send/recv + broadcast;
        * ``fused_halo=True``: no exchange step at all -- the fused projection is written into torch symmetric memory
          and the attention kernel's TMA producer reads the peers' K/V tiles straight over NVLink while the tensor pipe
          works on the previous tile (``univst_sc_attention_sharded_f16``); one cross-rank barrier per layer orders
          the reads after the peers' projections, double buffering makes a second barrier unnecessary."""
        import torch.distributed as dist
        if group is None and not dist.is_initialized():
            self._shard = None
            return
        world = dist.get_world_size(group)
        self._shard = (group, dist.get_rank(group), world) if world > 1 else None
        self._tables = {}
        self.drop_cuda_graphs()
        if split_k is None:
            split_k = world >= 4
        self._split_k = bool(split_k) and self._shard is not None
        if self._split_k or getattr(self, "_split_k_was_on", False):
            ops.gemm_splitk(74 if self._split_k else 0)   # at most half of the SMs' worth of tiles
        self._split_k_was_on = self._split_k
        nccl = self._shard is not None and dist.get_backend(group) == "nccl"
        if transport is None:
            transport = "xrank" if (nccl and push_halo is None and not fused_halo) else "nccl"
        if transport not in ("xrank", "nccl"):
            raise ValueError(transport)
        self._xr = None
        if transport == "xrank" and self._shard is not None:
            from .xrank import XRank
            key = id(group) if group is not None else 0
            if not hasattr(self, "_xr_cache"):
                self._xr_cache = {}
            if key not in self._xr_cache:
                self._xr_cache[key] = XRank(group, self.device)
            self._xr = self._xr_cache[key]
            fused_halo, push_halo = False, False
        self._fused_halo = bool(fused_halo) and self._shard is not None
        if push_halo is None:
            push_halo = nccl
        self._push_halo = bool(push_halo) and self._shard is not None and not self._fused_halo
        if not hasattr(self, "_symm"):
            self._symm = {}

    def _symm_qkv(self, rows, cols):
        """Double-buffered symmetric-memory projection buffer of one UNet level: (local [rows, cols] view to write,
        previous rank's view, rank 0's view, handle).  Allocation + rendezvous happen on first use (same order on
        every rank); the parity flips per use so that a buffer is rewritten only two barriers after it was read."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group, rank, world = self._shard
        key = (rows, cols)
        if key not in self._symm:
            t = symm_mem.empty(2, rows, cols, dtype=torch.float16, device=self.device)
            hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
            peers = {r: hdl.get_buffer(r, (2, rows, cols), torch.float16) for r in {max(rank - 1, 0), 0} if r != rank}
            self._symm[key] = [t, hdl, peers, 0]
        ent = self._symm[key]
        t, hdl, peers, par = ent
        ent[3] = par ^ 1
        prev = peers[rank - 1][par] if rank > 0 else None
        first = peers[0][par] if rank > 0 else None
        return t[par], prev, first, hdl

    def _symm_halo(self, rows, cols):
        """Double-buffered symmetric-memory projection buffer WITH halo banks: (local [rows, cols] tensor to fill, device
        pointers of every rank's copy of the same buffer, handle)."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group, rank, world = self._shard
        key = ("halo", rows, cols)
        if key not in self._symm:
            t = symm_mem.empty(2, rows, cols, dtype=torch.float16, device=self.device)
            hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
            ptrs = [hdl.get_buffer(r, (2, rows, cols), torch.float16).data_ptr() for r in range(world)]
            self._symm[key] = [t, hdl, ptrs, 0]
        ent = self._symm[key]
        t, hdl, ptrs, par = ent
        ent[3] = par ^ 1
        return t[par], [p + par * rows * cols * 2 for p in ptrs], hdl

    def _push_kv_halo(self, qkv, ptrs, hdl, B, F, N, C):
        """Frame-sharded attn1, push flavour of :meth:`_exchange_kv_halo`: K|V columns (C .. 3C) of my last frame -> bank 1
        of rank + 1, of the clip's first frame (rank 0) -> bank 2 of every rank; then the barrier."""
        group, rank, world = self._shard
        NI, ld = B * F, qkv.stride(0)
        kv = qkv[:, C:]                                   # [rows, 2C] strided view of the K|V columns
        if rank + 1 < world:
            dst = [0] * world
            dst[rank + 1] = ptrs[rank + 1] + (NI * N * ld + C) * 2
            ops.halo_push(kv[(F - 1) * N:], F * N, dst, ld, N, B, N)
        if rank == 0:
            dst = [p + ((NI + B) * N * ld + C) * 2 for p in ptrs]
            ops.halo_push(kv, F * N, dst, ld, N, B, N)
        hdl.barrier(channel=0)

    def _xr_halo(self, rows, cols):
        """Double-buffered symmetric projection buffer with halo banks (xrank transport): (local [rows, cols] tensor,
        device pointers of every rank's copy of the same half)."""
        key = ("qkv", rows, cols)
        if not hasattr(self, "_xr_par"):
            self._xr_par = {}
        par = self._xr_par.get(key, 0)
        self._xr_par[key] = par ^ 1
        t, ptrs = self._xr.buffer(key, (2, rows, cols))
        mc = self._xr.multicast(key)
        return t[par], [p + par * rows * cols * 2 for p in ptrs], (mc + par * rows * cols * 2 if mc else 0)

    def _xr_push_kv_halo(self, qkv, ptrs, mc, B, F, N, C):
        """xrank transport of the K/V halo (xrank.push_kv_halo): one launch, whose tail is the synchronisation."""
        from .xrank import push_kv_halo
        push_kv_halo(self._xr, qkv, ptrs, mc, B, F, N, C)

    def set_frame_sharding_off(self):
        self._shard = None
        self._xr = None
        self._fused_halo = False
        self._push_halo = False
        self._tables = {}
        self.drop_cuda_graphs()
        if getattr(self, "_split_k_was_on", False):
            ops.gemm_splitk(0)
            self._split_k_was_on = self._split_k = False

    def _heads(self, level):
        h = self.config["attention_head_dim"]
        return h if isinstance(h, int) else h[level]

    def _build_tree(self):
        n, lpb = self.nlev, self.config["layers_per_block"]
        self.down_blocks, self.up_blocks = [], []
        for i in range(n):
            b = _UNetBlock()
            b.has_cross_attention = i < n - 1
            for j in range(lpb):
                b.resnets.append(f"down_blocks.{i}.resnets.{j}.")
                if i < n - 1:
                    b.attentions.append(_Transformer(f"down_blocks.{i}.attentions.{j}.", self._heads(i)))
            self.down_blocks.append(b)
        self.mid_block = _UNetBlock()
        self.mid_block.attentions.append(_Transformer("mid_block.attentions.0.", self._heads(n - 1)))
        for i in range(n):
            b = _UNetBlock()
            b.has_cross_attention = i > 0
            for j in range(lpb + 1):
                b.resnets.append(f"up_blocks.{i}.resnets.{j}.")
                if i > 0:
                    b.attentions.append(_Transformer(f"up_blocks.{i}.attentions.{j}.", self._heads(n - 1 - i)))
            self.up_blocks.append(b)

    def _pack(self, sd):
        dev = self.device
        h = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()
        W: Dict[str, torch.Tensor] = {}
        temb_w, temb_b, self._temb_slices, off = [], [], {}, 0
        for key, t in sd.items():
            if "conv_temporal" in key:
                # Dirac / zero-bias identity (resnet.py:54-55); never loaded (unet_3d_condition.py:503) -- verified, since a
                # checkpoint with trained temporal convolutions would silently give wrong results otherwise
                if key.endswith("weight"):
                    ident = torch.zeros_like(t)
                    k = t.shape[-1]
                    idx = torch.arange(min(t.shape[0], t.shape[1]))
                    ident.reshape(t.shape[0], t.shape[1], k)[idx, idx, k // 2] = 1
                    ok = t.shape[0] == t.shape[1] and bool(torch.equal(t, ident))
                else:
                    ok = not bool((t != 0).any())
                if not ok:
                    raise NotImplementedError(f"{key} is not the Dirac / zero-bias identity of the pseudo-3D inflation: "
                                              "trained temporal convolutions are not on the UniVST path")
                continue
            if "attn_temporal" in key or "norm_temporal" in key:
                if key.endswith("attn_temporal.to_out.0.weight") and bool((t != 0).any()):
                    raise NotImplementedError(
                        f"{key} is non-zero: this backbone has a live temporal attention (not the SD-1.5/2.1 "
                        "pseudo-3D inflation, whose temporal output projection is zero-initialised and never loaded)")
                if key.endswith("attn_temporal.to_out.0.bias"):
                    W[key] = h(t)
                continue
            if key.endswith("time_emb_proj.weight"):
                pre = key[: -len("time_emb_proj.weight")]
                self._temb_slices[pre] = (off, t.shape[0])
                off += t.shape[0]
                temb_w.append(t)
                temb_b.append(sd[pre + "time_emb_proj.bias"])
                continue
            if key.endswith("time_emb_proj.bias"):
                continue
            if t.dim() == 4 and t.shape[-1] == 3:
                W[key] = h(pack_conv3x3(t, CIN_PAD if key == "conv_in.weight" else 0))
            elif t.dim() == 4:
                W[key] = h(t.reshape(t.shape[0], t.shape[1]))
            else:
                W[key] = h(t)
        W["temb_all.weight"] = h(torch.cat(temb_w, 0))
        W["temb_all.bias"] = h(torch.cat(temb_b, 0))
        for tr in self._all_transformers():
            b = tr.prefix + "transformer_blocks.0."
            W[b + "attn1.to_qkv.weight"] = torch.cat([W.pop(b + "attn1.to_q.weight"), W.pop(b + "attn1.to_k.weight"),
                                                      W.pop(b + "attn1.to_v.weight")], 0).contiguous()
            W[b + "attn2.to_kv.weight"] = torch.cat([W.pop(b + "attn2.to_k.weight"), W.pop(b + "attn2.to_v.weight")],
                                                    0).contiguous()
            W[b + "ff.net.0.proj.weight"], W[b + "ff.net.0.proj.bias"] = pack_geglu(W[b + "ff.net.0.proj.weight"],
                                                                                  W[b + "ff.net.0.proj.bias"])
        self.W = W
        self._pack_extra()

    # ------------------------------------------------------------------------------------------ backbone flavour
    # The SD "pseudo-3D" inflation is the default; backbones/animatediff overrides these (univst_b200/animatediff.py).
    def _pack_extra(self):
        pass

    def _gn_span(self, B, F, HW):
        """(groups of rows, rows per group) the ResNet GroupNorm statistics span: all frames of a branch (resnet.py:338)."""
        return B, F * HW

    def _attn1_plan(self, a1):
        """K/V source mode of attn1 and, while the AdaIN-guided shift is active, its (alpha, beta, gamma)."""
        if not a1.patched:
            return "prev_self_first", None   # stock SparseCausalAttention, models/attention.py:356
        if a1.idx is None:
            raise RuntimeError("patched attn1 called before register_time() set .idx")
        shift = None
        if a1.idx >= a1.eta1 and a1.idx <= a1.eta2 * 50:  # pnp_utils.py:47
            beta = (0.9 - 0.1) / (a1.eta1 * 50 - a1.eta2 * 50) * (a1.idx - a1.eta2 * 50) + 0.1
            shift = (0.65, beta, 3.0)
        return "prev_first", shift

    def shift_live(self, a1) -> bool:
        """False once the content / style branches can no longer influence the edit branch (patched, window closed)."""
        return (not a1.patched) or self._attn1_plan(a1)[1] is not None

    def _ff_out_bias2(self, b):
        return self.W[b + "attn_temporal.to_out.0.bias"]   # the dead temporal attention (models/attention.py:331-346)

    def _motion(self, prefix, x, B, F, H, Wd):
        return x

    def _all_transformers(self):
        for blk in self.down_blocks + [self.mid_block] + self.up_blocks:
            yield from blk.attentions

    def _table(self, B, F, mode):
        key = (B, F, mode)
        if key not in self._tables:
            if self._shard is None or mode == "branch":
                t = kv_source_table(B, F, mode)
            else:
                t = kv_source_table_sharded(B, F, mode, self._shard[1])
            self._tables[key] = t.to(self.device)
        return self._tables[key]

    def _gn(self, x, gamma, beta, *, NB, rows, eps, silu, x2=None):
        """GroupNorm whose statistics span all frames of a branch (all ranks when the frames are sharded)."""
        g = self.config["norm_num_groups"]
        if self._shard is None:
            return ops.groupnorm(x, gamma, beta, NB=NB, rows=rows, groups=g, eps=eps, silu=silu, x2=x2)
        if self._xr is not None:
            return ops.groupnorm_xrank(x, gamma, beta, NB=NB, rows=rows, xr=self._xr, groups=g, eps=eps, silu=silu, x2=x2)
        return ops.groupnorm_sharded(x, gamma, beta, NB=NB, rows=rows, group=self._shard[0], world=self._shard[2],
                                     groups=g, eps=eps, silu=silu, x2=x2)

    def _exchange_kv_halo(self, qkv, B, F, N):
        """Frame-sharded attn1: qkv is [(B*F + 2B) * N, 3C]; fill the two halo banks (see kv_source_table_sharded)."""
        import torch.distributed as dist
        group, rank, world = self._shard
        NI = B * F
        v = qkv[: NI * N].view(B, F, N, -1)
        halo_prev = qkv[NI * N: (NI + B) * N].view(B, N, -1)
        halo_first = qkv[(NI + B) * N: (NI + 2 * B) * N].view(B, N, -1)
        ops_ = []
        if rank + 1 < world:
            ops_.append(dist.P2POp(dist.isend, v[:, F - 1].contiguous(), dist.get_global_rank(group, rank + 1) if group else rank + 1, group))
        if rank > 0:
            ops_.append(dist.P2POp(dist.irecv, halo_prev, dist.get_global_rank(group, rank - 1) if group else rank - 1, group))
        if rank == 0:
            halo_first.copy_(v[:, 0])
        if ops_:
            for w in dist.batch_isend_irecv(ops_):
                w.wait()
        dist.broadcast(halo_first, src=dist.get_global_rank(group, 0) if group else 0, group=group)

    # ------------------------------------------------------------------------------------------ building blocks
    def _resnet(self, pre, x, skip, temb_all, B, F, H, Wd):
        """ResnetBlockPseudo3D.forward (resnet.py:335-394).  x: [M, C1], skip: [M, C2] or None."""
        W, cfg = self.W, self.config
        NI = B * F
        NBg, rows = self._gn_span(B, F, H * Wd)
        eps = cfg["norm_eps"]
        h = self._gn(x, W[pre + "norm1.weight"], W[pre + "norm1.bias"], NB=NBg, rows=rows, eps=eps, silu=True, x2=skip)
        off, cout = self._temb_slices[pre]
        h = ops.conv3x3(h.view(NI, H, Wd, -1), W[pre + "conv1.weight"], bias=W[pre + "conv1.bias"],
                        rowvec=temb_all[:, off:off + cout], rows_per_group=F * H * Wd)
        h = self._gn(h, W[pre + "norm2.weight"], W[pre + "norm2.bias"], NB=NBg, rows=rows, eps=eps, silu=True)
        if pre + "conv_shortcut.weight" in W:
            sc = ops.gemm(x, W[pre + "conv_shortcut.weight"], a2=skip, bias=W[pre + "conv_shortcut.bias"])
        else:
            assert skip is None
            sc = x
        return ops.conv3x3(h.view(NI, H, Wd, -1), W[pre + "conv2.weight"], bias=W[pre + "conv2.bias"], residual=sc)

    def _transformer(self, tr: _Transformer, x, ctx_kv_of, B, F, H, Wd, cut: bool = False):
        """SpatioTemporalTransformerModel.forward (attention.py:104-153) + its block (:280-334).  x: [M, C].
        ``cut``: this is the last layer in which the content / style branches matter -- after the fused projection and
        the shift only the edit branch (the last F images) goes on: returns [F * N, C]."""
        W, cfg = self.W, self.config
        pre, heads = tr.prefix, tr.heads
        NI, N, C = B * F, H * Wd, x.shape[1]
        d = C // heads
        b = pre + "transformer_blocks.0."
        y = ops.groupnorm(x, W[pre + "norm.weight"], W[pre + "norm.bias"], NB=NI, rows=N, groups=cfg["norm_num_groups"],
                          eps=1e-6, silu=False)
        y = ops.gemm(y, W[pre + "proj_in.weight"], bias=W[pre + "proj_in.bias"])
        # 1. sparse-causal self-attention (stock or patched)
        n1 = ops.layernorm(y, W[b + "norm1.weight"], W[b + "norm1.bias"])
        NIkv = NI
        a1 = tr.transformer_blocks[0].attn1
        mode, shift = self._attn1_plan(a1)
        halo = self._shard is not None and mode != "self"   # per-frame self-attention needs no neighbour K/V
        if not halo:
            qkv = ops.gemm(n1, W[b + "attn1.to_qkv.weight"])
        elif self._fused_halo:  # projection into symmetric memory: the peers' attention kernels read it over NVLink
            qkv_buf, qkv_prev, qkv_first, symm_hdl = self._symm_qkv(NI * N, 3 * C)
            qkv = ops.gemm(n1, W[b + "attn1.to_qkv.weight"], out=qkv_buf)
        else:  # two halo banks of B images each behind the local images
            NIkv = NI + 2 * B
            if self._xr is not None:
                qkv_all, halo_ptrs, halo_mc = self._xr_halo(NIkv * N, 3 * C)
            elif self._push_halo:
                qkv_all, halo_ptrs, symm_hdl = self._symm_halo(NIkv * N, 3 * C)
            else:
                qkv_all = torch.empty((NIkv * N, 3 * C), dtype=torch.float16, device=x.device)
            qkv = ops.gemm(n1, W[b + "attn1.to_qkv.weight"], out=qkv_all[: NI * N])
        if shift is not None:
            if B != 3:
                raise ValueError("the AdaIN-guided shift needs the three-branch batch [content, style, edit]")
            if self._step_params is not None:   # graph-safe: (alpha, beta, gamma) of this layer in device memory
                ops.attn_shift_dev_(qkv, F, N, C, self._step_params[8 + 3 * tr.index: 11 + 3 * tr.index])
            else:
                ops.attn_shift_(qkv, F, N, C, *shift)
        if halo and self._fused_halo:
            symm_hdl.barrier(channel=0)   # every rank's projection (and shift) of this layer is complete and visible
            o = ops.sc_attention_sharded(qkv, qkv_prev, qkv_first, self._table(B, F, mode), B=B, Fl=F, H=heads, d=d, N=N)
        else:
            if halo:
                if self._xr is not None:
                    self._xr_push_kv_halo(qkv_all, halo_ptrs, halo_mc, B, F, N, C)
                elif self._push_halo:
                    self._push_kv_halo(qkv_all, halo_ptrs, symm_hdl, B, F, N, C)
                else:
                    self._exchange_kv_halo(qkv_all, B, F, N)
                kv = qkv_all
            else:
                kv = qkv
            table, qrows = self._table(B, F, mode), qkv
            if cut:  # queries of the edit branch only; its K/V sources keep their image numbers in the full buffer
                e0 = (B - 1) * F
                table, qrows, x, y, NI = table[e0:], qkv[e0 * N:], x[e0 * N:], y[e0 * N:], F
            o = ops.sc_attention(qrows[:, :C], kv[:, C:2 * C], kv[:, 2 * C:], table, NI=NI, NIkv=NIkv,
                                 H=heads, d=d, N=N, Nkv=N)
        y = ops.gemm(o, W[b + "attn1.to_out.0.weight"], bias=W[b + "attn1.to_out.0.bias"], residual=y)
        # 2. cross-attention over the (per-branch) context
        n2 = ops.layernorm(y, W[b + "norm2.weight"], W[b + "norm2.bias"])
        q2 = ops.gemm(n2, W[b + "attn2.to_q.weight"])
        kv2, L = ctx_kv_of(b)
        Bx = B
        if cut:
            kv2, Bx = kv2[(B - 1) * L:], 1
        o2 = ops.cross_attention(q2, kv2[:, :C], kv2[:, C:], self._table(Bx, F, "branch"), NI=NI, NIkv=Bx, H=heads, d=d, N=N,
                                 Nkv=L)
        y = ops.gemm(o2, W[b + "attn2.to_out.0.weight"], bias=W[b + "attn2.to_out.0.bias"], residual=y)
        # 3. GEGLU feed-forward; the dead temporal attention (attention.py:331-346) is its bias, added in the epilogue
        n3 = ops.layernorm(y, W[b + "norm3.weight"], W[b + "norm3.bias"])
        g = ops.gemm(n3, W[b + "ff.net.0.proj.weight"], bias=W[b + "ff.net.0.proj.bias"], geglu=True)
        y = ops.gemm(g, W[b + "ff.net.2.weight"], bias=W[b + "ff.net.2.bias"], residual=y, bias2=self._ff_out_bias2(b))
        return ops.gemm(y, W[pre + "proj_out.weight"], bias=W[pre + "proj_out.bias"], residual=x)

    # ------------------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, attention_mask=None,
                ft_indices=None, ft_timesteps=None, ft_path=None, **kwargs):
        if class_labels is not None or attention_mask is not None:
            raise NotImplementedError("class_labels / attention_mask are unused on the UniVST path")
        if self.use_cuda_graphs and ft_path is None:
            return self._forward_graphed(sample, timestep, encoder_hidden_states)
        return self._forward_impl(sample, timestep, encoder_hidden_states, ft_indices, ft_timesteps, ft_path)

    # ------------------------------------------------------------------------------------------ CUDA graphs
    def drop_cuda_graphs(self):
        """Forget the captured forwards (their memory pool goes with them: a pool whose graphs are gone cannot be captured
        into again, so the next capture starts a new one)."""
        self._graphs = {}
        self._graph_pool = None

    def _plan_signature(self):
        sig = []
        for tr in self._all_transformers():
            mode, shift = self._attn1_plan(tr.transformer_blocks[0].attn1)
            sig.append((mode, shift is not None))
        return tuple(sig)

    def _step_values(self, timestep, B):
        """The 64 floats a captured forward reads from device memory: [0, B) the timestep of every sample, then from 8 on
        (alpha, beta, gamma) of the AdaIN-guided shift of every transformer (zeros where it is not active)."""
        if B > 8:
            raise ValueError("a graphed forward takes at most 8 samples (branches)")
        if torch.is_tensor(timestep):
            tl = [float(v) for v in timestep.reshape(-1).tolist()]   # (synchronises: pass Python numbers on the hot path)
            tl = tl * B if len(tl) == 1 else tl
        else:
            tl = [float(timestep)] * B
        vals = tl + [0.0] * (8 - B)
        for tr in self._all_transformers():
            shift = self._attn1_plan(tr.transformer_blocks[0].attn1)[1]
            vals += list(shift) if shift is not None else [0.0] * 3
        if len(vals) > 64:
            raise ValueError("too many transformer blocks for the step-parameter block")
        return vals

    def _forward_graphed(self, sample, timestep, encoder_hidden_states):
        """The forward as a replayed CUDA graph.  One graph per (input shapes, attention plan, truncation, sharding) key is
        captured on first use -- after an eager warm-up call that sizes workspaces, source tables and symmetric buffers
        -- and replayed afterwards: the step's latents / context are copied into the graph's static inputs, its scalars
        (timestep; alpha, beta, gamma of the shift) go to device memory through one tiny launch, and the ~500 kernel
        launches of the forward cost one ``cudaGraphLaunch``.  Under frame sharding the cross-rank synchronisations are
        kernels of the graph (device-side epochs), so every rank replays independently."""
        dev = self.device
        sample = sample.to(device=dev, dtype=torch.float16)
        ctx = encoder_hidden_states.to(device=dev, dtype=torch.float16)
        B = sample.shape[0]
        key = (tuple(sample.shape), tuple(ctx.shape), self._plan_signature(), bool(self.truncate_dead_branches),
               None if self._shard is None else (self._shard[1], self._shard[2], self._xr is not None))
        if self._shard is not None and self._xr is None:
            raise NotImplementedError("CUDA graphs under frame sharding need the xrank transport (no collective-library calls)")
        vals = self._step_values(timestep, B)
        ent = self._graphs.get(key)
        if ent is None:
            st = {"sample": torch.empty_like(sample, memory_format=torch.contiguous_format),
                  "ctx": torch.empty_like(ctx, memory_format=torch.contiguous_format),
                  "params": torch.zeros(64, dtype=torch.float32, device=dev)}
            st["sample"].copy_(sample)
            st["ctx"].copy_(ctx)
            ops.set_floats(st["params"], vals)
            if self._graph_stream is None:
                self._graph_stream = torch.cuda.Stream(device=dev)
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()
            side, cur = self._graph_stream, torch.cuda.current_stream(dev)

            def run():
                self._step_params = st["params"]
                try:
                    out = self._forward_impl(st["sample"], st["params"][:B], st["ctx"], None, None, None)
                finally:
                    self._step_params = None
                return out, self.last_eps_rows, self.last_edit_branch

            side.wait_stream(cur)
            with torch.cuda.stream(side):
                run()                      # eager warm-up on the capture stream (allocations, tables, symmetric buffers)
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.launch_count
            with torch.cuda.graph(graph, pool=self._graph_pool, stream=side):
                res = run()
            ent = self._graphs[key] = (graph, st, res, ops.launch_count - n0)
            ops.launch_count = n0          # capturing launches nothing
        graph, st, (out, eps_rows, edit_branch), nkernels = ent
        st["sample"].copy_(sample)
        st["ctx"].copy_(ctx)
        ops.set_floats(st["params"], vals)
        graph.replay()
        ops.launch_count += nkernels       # the kernels of this library inside the replayed graph
        self.last_eps_rows, self.last_edit_branch = eps_rows, edit_branch
        return out

    def _forward_impl(self, sample, timestep, encoder_hidden_states, ft_indices=None, ft_timesteps=None, ft_path=None):
        W, cfg = self.W, self.config
        dev = self.device
        B, Cin, F, H, Wd = sample.shape
        q = 1 << (self.nlev - 1)
        if H % q or Wd % q:
            raise ValueError(f"latent height / width must be multiples of {q} (the UNet halves them {self.nlev - 1} times)")
        boc = cfg["block_out_channels"]
        sample = sample.to(device=dev, dtype=torch.float16).contiguous()
        F_total = F
        if self._shard is not None:
            _, rank, world = self._shard
            if F % world:
                raise ValueError(f"{F} frames do not shard evenly over {world} ranks")
            if ft_path is not None:
                raise NotImplementedError("feature dumps are not available under frame sharding")
            F = F // world
            sample = sample[:, :, rank * F:(rank + 1) * F].contiguous()
        x = ops.pack_latents([sample[b] for b in range(B)], Cpad=CIN_PAD)

        # time embedding: sinusoid -> Linear -> SiLU -> Linear; every ResNet consumes silu(emb) (resnet.py:355), so
        # the second Linear stores silu(emb) and one GEMM evaluates all time_emb_proj layers at once
        if torch.is_tensor(timestep):
            t = timestep.to(device=dev, dtype=torch.float32).reshape(-1)
            t = t.expand(B).contiguous() if t.numel() == 1 else t.contiguous()
        else:
            t = torch.full((B,), float(timestep), dtype=torch.float32, device=dev)
        e = ops.timestep_embedding(t, boc[0])
        e = ops.gemm(e, W["time_embedding.linear_1.weight"], bias=W["time_embedding.linear_1.bias"], act=True)
        e = ops.gemm(e, W["time_embedding.linear_2.weight"], bias=W["time_embedding.linear_2.bias"], act=True)
        temb_all = ops.gemm(e, W["temb_all.weight"], bias=W["temb_all.bias"])

        # cross-attention K/V of the context: one tiny GEMM per layer (the context does not depend on the frame)
        ctx = encoder_hidden_states.to(device=dev, dtype=torch.float16)
        if ctx.shape[0] != B:
            ctx = ctx.expand(B, -1, -1)
        L = ctx.shape[1]
        ctx2d = ctx.reshape(B * L, -1).contiguous()

        # SURVEY 2.3 D3 (second half): the transformer after whose projection the content / style branches are dead
        cut_tr = None
        if self.truncate_dead_branches and B == 3 and not self._fused_halo:
            for tr in self._all_transformers():
                a1 = tr.transformer_blocks[0].attn1
                if a1.patched and self._attn1_plan(a1)[1] is not None:
                    cut_tr = tr   # the last one in execution order
        live = {"B": B, "off": 0}   # branches still evaluated / index of the first of them in the call's batch

        def ctx_kv_of(bprefix):   # context rows of the branches still evaluated
            return ops.gemm(ctx2d[live["off"] * L:], W[bprefix + "attn2.to_kv.weight"]), L

        def transformer(tr, x, h, w):
            if tr is not cut_tr:
                return self._transformer(tr, x, ctx_kv_of, live["B"], F, h, w)
            x = self._transformer(tr, x, ctx_kv_of, B, F, h, w, cut=True)
            live["B"], live["off"] = 1, B - 1
            return x

        def cur(t, rows_per_branch):
            """Rows of the branches that are still evaluated of a tensor produced before the cut."""
            if live["off"] and t.shape[0] == B * rows_per_branch:
                return t[live["off"] * rows_per_branch:]
            return t

        x = ops.conv3x3(x, W["conv_in.weight"], bias=W["conv_in.bias"])
        h, w = H, Wd
        skips = [x]
        n, lpb = self.nlev, cfg["layers_per_block"]
        for i, blk in enumerate(self.down_blocks):
            for j in range(lpb):
                x = self._resnet(blk.resnets[j], x, None, temb_all[live["off"]:], live["B"], F, h, w)
                if blk.attentions:
                    x = transformer(blk.attentions[j], x, h, w)
                x = self._motion(f"down_blocks.{i}.motion_modules.{j}.", x, live["B"], F, h, w)
                skips.append(x)
            if i < n - 1:
                planes = ops.space_to_depth2(x.view(live["B"] * F, h, w, -1))
                h, w = h // 2, w // 2
                pre = f"down_blocks.{i}.downsamplers.0.conv."
                x = ops.conv3x3(planes, W[pre + "weight"], stride=2, bias=W[pre + "bias"])
                skips.append(x)
        x = self._resnet("mid_block.resnets.0.", x, None, temb_all[live["off"]:], live["B"], F, h, w)
        x = transformer(self.mid_block.attentions[0], x, h, w)
        x = self._motion("mid_block.motion_modules.0.", x, live["B"], F, h, w)
        x = self._resnet("mid_block.resnets.1.", x, None, temb_all[live["off"]:], live["B"], F, h, w)
        for i, blk in enumerate(self.up_blocks):
            for j in range(lpb + 1):
                x = self._resnet(blk.resnets[j], x, cur(skips.pop(), F * h * w), temb_all[live["off"]:], live["B"], F, h, w)
                if blk.attentions:
                    x = transformer(blk.attentions[j], x, h, w)
                x = self._motion(f"up_blocks.{i}.motion_modules.{j}.", x, live["B"], F, h, w)
            if i < n - 1:
                up = ops.upsample2x(x.view(live["B"] * F, h, w, -1))
                h, w = h * 2, w * 2
                pre = f"up_blocks.{i}.upsamplers.0.conv."
                x = ops.conv3x3(up, W[pre + "weight"], bias=W[pre + "bias"])
            if ft_indices is not None and ft_timesteps is not None and ft_path is not None:
                # unet_3d_condition.py:430-436: sample[0].permute(1, 2, 3, 0) == our channels-last rows of branch 0
                if i in ft_indices and timestep in ft_timesteps:
                    if live["off"]:
                        raise NotImplementedError("feature dumps need branch 0: do not combine with truncate_dead_branches")
                    path = os.path.join(ft_path, f"inversion_feature_map_{i}_block_{timestep}_step.pt")
                    torch.save(x[: F * h * w].view(F, h, w, -1).clone(), path)
                    print(f"save feature map at: {path}")
        Bo = live["B"]   # branches in the output (1 after a cut: the edit branch)
        NBg, rows = self._gn_span(Bo, F, h * w)
        y = self._gn(x, W["conv_norm_out.weight"], W["conv_norm_out.bias"], NB=NBg, rows=rows, eps=cfg["norm_eps"],
                     silu=True)
        eps_rows = ops.conv3x3(y.view(Bo * F, h, w, -1), W["conv_out.weight"], bias=W["conv_out.bias"],
                               out=torch.empty((Bo * F * h * w, 8), dtype=torch.float16, device=dev))
        if self._xr is not None:   # store my frames of the (tiny) noise prediction into every rank's full-clip buffer
            _, rank, world = self._shard
            key = ("eps", Bo, F_total, h * w)
            full, ptrs = self._xr.buffer(key, (Bo * F_total * h * w, 8))
            mc = self._xr.multicast(key)
            ops.xrank_push(self._xr, [dict(src=eps_rows, src_blk_rows=F * h * w, dst=[p + rank * F * h * w * 16 for p in ptrs],
                                           ld_dst=8, dst_blk_rows=F_total * h * w, nblk=Bo, rows=F * h * w,
                                           mc=mc + rank * F * h * w * 16 if mc else 0)])
            eps_rows, F = full, F_total
        elif self._shard is not None:  # all-gather the (tiny) noise prediction: [P][B][Fl] -> [B][P Fl]
            import torch.distributed as dist
            group, rank, world = self._shard
            gathered = torch.empty((world, Bo, F, h * w, 8), dtype=torch.float16, device=dev)
            # output given in the concatenated form (rank-major dim 0), which every backend accepts
            dist.all_gather_into_tensor(gathered.view(world * Bo, F, h * w, 8), eps_rows.view(Bo, F, h * w, 8), group=group)
            eps_rows = gathered.permute(1, 0, 2, 3, 4).reshape(Bo * F_total * h * w, 8).contiguous()
            F = F_total
        self.last_eps_rows = eps_rows
        self.last_edit_branch = Bo - 1
        out = ops.unpack_latents(self.last_eps_rows, Bo, cfg["out_channels"], F, h, w)
        return UNetPseudo3DConditionOutput(sample=out)

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)
