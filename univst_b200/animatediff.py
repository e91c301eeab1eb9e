"""B200 host-side mirror of the reference's AnimateDiff-v2 backbone (``backbones/animatediff``):
``UNet3DConditionModel`` (models/unet.py:41, forward :322-466) with its ``VanillaTemporalModule`` motion modules
(models/motion_module.py:52-337) and ``AnimationPipeline.video_style_transfer`` (pipelines/pipeline_animation.py:
449-603).  Same ``state_dict`` key names, same patch protocol (``backbones/animatediff/pnp_utils.py`` sets ``idx`` /
``eta1`` / ``eta2`` / an instance-level ``forward`` on ``up_blocks[r].attentions[b].transformer_blocks[0].attn1``).

What differs from the SD "pseudo-3D" backbone (unet.py), all of it read off the reference:
* GroupNorm statistics are per frame (``InflatedGroupNorm``, models/resnet.py:21-29);
* attn1 is per-frame self-attention: the patched forward is called without ``clip_length`` (models/attention.py:333,
  pnp_utils.py:57), so K/V = the frame itself; shift window ``eta1*50 <= idx < eta2*50``, alpha 0.8, gamma 2.0
  (pnp_utils.py:45-49);
* the transformer block has no (dead) temporal attention;
* a motion module follows every (resnet, attention) pair: GroupNorm -> Linear -> 2 x [LayerNorm -> +positional
  encoding -> attention over the F frames of each pixel -> +residual] -> GEGLU feed-forward -> Linear -> +residual.

On the GPU the motion module is five tcgen05 GEMMs with fused epilogues plus the temporal-attention kernel
(csrc/temporal_attn.cu).  The sinusoidal positional encoding is folded through the (bias-free) Q/K/V projection:
``W (n + pe_f) = W n + W pe_f``, so ``W pe_f`` is computed once at pack time per attention block and added as a
per-frame row vector in the epilogue of the fused QKV GEMM -- the "(b f) d c -> (b d) f c" rearrange and the
elementwise add of the reference do not exist.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops
from .pack import pack_geglu
from .pipeline import SpatioTemporalStableDiffusionPipeline
from .unet import SD15_CONFIG, UNetPseudo3DConditionModel

AD_SD15_CONFIG = dict(SD15_CONFIG, motion_heads=8, motion_max_len=24)  # animatediff-v2.yaml:8-14 + motion_module.py:60


def positional_encoding(d_model: int, max_len: int) -> torch.Tensor:
    """PositionalEncoding buffer of the reference (models/motion_module.py:232-243), fp32 (max_len, d_model)."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def frames_to_pixels(x: torch.Tensor, B: int, Fl: int, N: int, P: int, group=None) -> torch.Tensor:
    """Rows (b, local frame, pixel) of every rank -> rows (b, global frame, local pixel): rank r ends up with pixel
    slice [r N/P, (r+1) N/P) of ALL P*Fl frames (frame order = rank order).  One all_to_all_single."""
    import torch.distributed as dist
    if N % P:
        raise ValueError(f"{N} pixels do not split evenly over {P} ranks")
    n, C = N // P, x.shape[1]
    send = x.view(B, Fl, P, n, C).permute(2, 0, 1, 3, 4).contiguous()          # [dest rank, b, fl, pix, C]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)                           # [source rank, b, fl, pix, C]
    return recv.permute(1, 0, 2, 3, 4).reshape(B * P * Fl * n, C)             # (b, (rank, fl) = global frame, pix)


def pixels_to_frames(x: torch.Tensor, B: int, Fl: int, N: int, P: int, group=None) -> torch.Tensor:
    """Inverse of :func:`frames_to_pixels`: rows (b, global frame, local pixel) -> rows (b, local frame, pixel)."""
    import torch.distributed as dist
    n, C = N // P, x.shape[1]
    send = x.view(B, P, Fl, n, C).permute(1, 0, 2, 3, 4).contiguous()          # [frame owner, b, fl, pix, C]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)                           # [pixel-slice owner, b, fl, pix, C]
    return recv.permute(1, 2, 0, 3, 4).reshape(B * Fl * N, C)                 # (b, fl, (slice, pix) = pixel)


class UNet3DConditionModel(UNetPseudo3DConditionModel):
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: Optional[dict] = None, device="cuda"):
        cfg = dict(AD_SD15_CONFIG)
        cfg.update(config or {})
        if cfg.get("use_linear_projection"):
            raise NotImplementedError("AnimateDiff-v2 sits on SD-1.5 (1x1-conv proj_in / proj_out)")
        self._pe_rows = {}
        self._push, self._arena = False, None
        super().__init__(state_dict, cfg, device=device)

    def set_frame_sharding(self, group=None, push_exchange=None, transport=None, split_k=None):
        """Shard the frames of every clip over the ranks of ``group``.  Everything spatial is frame-local in this backbone
        (per-frame GroupNorm, per-frame attn1); the only exchange is inside the motion modules, whose attention runs over
        the frames of one pixel: an all-to-all turns "my frames, all pixels" into "all frames, my pixels" before the
        temporal transformer block and back after it (two all-to-alls of C-wide rows per motion module over NVLink).
        ``push_exchange``: instead of permute + NCCL all-to-all + permute, one kernel stores every row at its place in the
        owner's symmetric-memory buffer over NVLink (``univst_exchange_push_f16``), followed by a cross-rank barrier --
        the default on NCCL process groups (2 GPUs, 16 frames at 64 x 64: 78.0 ms on one GPU, 47.4 ms with the all-to-all,
        43.5 ms pushed; all three bit-identical, profiles/r01_animatediff_sharding_exchange_2gpu.json)."""
        import torch.distributed as dist
        # default on NCCL groups: the xrank transport (unet.py) -- the exchange kernel's tail is the cross-rank
        # synchronisation and the noise prediction is stored into every rank's buffer: no collective-library call at all
        if transport is None and push_exchange is not None:
            transport = "nccl"
        super().set_frame_sharding(group, transport=transport, split_k=split_k)
        self._pe_rows = {}
        if push_exchange is None:
            push_exchange = self._shard is not None and dist.get_backend(group) == "nccl"
        self._push = bool(push_exchange) and self._shard is not None
        self._arena = None

    def _exchange(self, direction, y, B, F, N):
        """frames -> pixels (0) / pixels -> frames (1) of the [rows, C] activations ``y``."""
        group, rank, P = self._shard
        if self._xr is not None:
            rows, C = y.shape
            need = rows * C
            # two buffers (one per direction): a buffer is rewritten only after the synchronisation that follows the OTHER
            # direction's push, which every rank passes after its last read of this one
            t, ptrs = self._xr.buffer(("arena", need), (2, need))
            ops.exchange_push(direction, y, [p + direction * need * 2 for p in ptrs], rank, P, B, F, N, xr=self._xr)
            return t[direction].view(rows, C)
        if not self._push:
            return (frames_to_pixels if direction == 0 else pixels_to_frames)(y, B, F, N, P, group)
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        rows, C = y.shape
        need = rows * C
        if self._arena is None or self._arena[0].shape[1] < need:
            # two buffers (one per direction): a buffer is rewritten only after the barrier that follows the OTHER
            # direction's push, which every rank passes after its last read of this one
            t = symm_mem.empty(2, need, dtype=torch.float16, device=self.device)
            hdl = symm_mem.rendezvous(t, group if group is not None else dist.group.WORLD)
            ptrs = [hdl.get_buffer(r, (2, need), torch.float16).data_ptr() for r in range(P)]
            self._arena = (t, hdl, ptrs)
        t, hdl, ptrs = self._arena
        off = direction * t.shape[1] * 2   # bytes
        ops.exchange_push(direction, y, [p + off for p in ptrs], rank, P, B, F, N)
        hdl.barrier(channel=0)             # every rank's rows have landed (and are visible) before anybody reads
        return t[direction, :need].view(rows, C)

    def set_frame_sharding_off(self):
        super().set_frame_sharding_off()
        self._pe_rows = {}
        self._push = False

    # ------------------------------------------------------------------------------------------ flavour hooks
    def _pack_extra(self):
        W, cfg = self.W, self.config
        self._motion_prefixes = sorted({k[: k.index("temporal_transformer.")] for k in W if "motion_modules" in k})
        for pre in self._motion_prefixes:
            b = pre + "temporal_transformer.transformer_blocks.0."
            C = W[b + "norms.0.weight"].shape[0]
            pe = positional_encoding(C, cfg["motion_max_len"]).to(self.device)
            for i in range(2):
                a = b + f"attention_blocks.{i}."
                wqkv = torch.cat([W.pop(a + "to_q.weight"), W.pop(a + "to_k.weight"), W.pop(a + "to_v.weight")], 0)
                W[a + "to_qkv.weight"] = wqkv.contiguous()
                # W (n + pe_f) = W n + W pe_f: the positional term of every frame as an epilogue row vector
                W[a + "pe_qkv"] = (pe @ wqkv.float().T).to(torch.float16).contiguous()
            W[b + "ff.net.0.proj.weight"], W[b + "ff.net.0.proj.bias"] = pack_geglu(W[b + "ff.net.0.proj.weight"],
                                                                                  W[b + "ff.net.0.proj.bias"])

    def _gn_span(self, B, F, HW):
        return B * F, HW   # InflatedGroupNorm: per-frame statistics (models/resnet.py:21-29)

    def _gn(self, x, gamma, beta, *, NB, rows, eps, silu, x2=None):
        return ops.groupnorm(x, gamma, beta, NB=NB, rows=rows, groups=self.config["norm_num_groups"], eps=eps, silu=silu,
                             x2=x2)

    def _attn1_plan(self, a1):
        if not a1.patched:
            return "self", None
        if a1.idx is None:
            raise RuntimeError("patched attn1 called before register_time() set .idx")
        shift = None
        if a1.idx >= a1.eta1 * 50 and a1.idx < a1.eta2 * 50:  # backbones/animatediff/pnp_utils.py:45
            beta = (0.9 - 0.1) / (a1.eta1 * 50 - a1.eta2 * 50) * (a1.idx - a1.eta2 * 50) + 0.1
            shift = (0.8, beta, 2.0)
        return "self", shift

    def _ff_out_bias2(self, b):
        return None

    def _pe_table(self, key, B, F):
        """[B*F, 3C] row vectors for the QKV epilogue: image (b, f) takes row f of W pe (``F`` = all frames of the clip)."""
        ck = (key, B, F)
        if ck not in self._pe_rows:
            t = self.W[key]
            if F > t.shape[0]:
                raise ValueError(f"{F} frames exceed the motion modules' positional-encoding length {t.shape[0]}")
            self._pe_rows[ck] = t[:F].repeat(B, 1).contiguous()
        return self._pe_rows[ck]

    def _motion(self, prefix, x, B, F, H, Wd):
        """VanillaTemporalModule.forward (models/motion_module.py:83-89 -> :138-163, :218-229).  x: [B*F*H*W, C].

        Frame-sharded: GroupNorm (per frame) and the two projections at the ends run on "my frames, all pixels"; the
        transformer block in between -- LayerNorms, both attentions over the frames of a pixel, feed-forward: all of it
        row-wise or per pixel -- runs on "all frames, my pixels".  One all-to-all of the C-wide activations each way per
        module (instead of one of the 3C-wide projections and one of the output around each of the two attentions)."""
        W, cfg = self.W, self.config
        t = prefix + "temporal_transformer."
        if t + "norm.weight" not in W:
            return x
        NI, N, C = B * F, H * Wd, x.shape[1]
        heads = cfg["motion_heads"]
        y = ops.groupnorm(x, W[t + "norm.weight"], W[t + "norm.bias"], NB=NI, rows=N, groups=cfg["norm_num_groups"],
                          eps=1e-6, silu=False)
        y = ops.gemm(y, W[t + "proj_in.weight"], bias=W[t + "proj_in.bias"])
        Fa, Na = F, N                       # frames / pixels per clip of the rows the block works on
        if self._shard is not None:
            group, _, P = self._shard
            y = self._exchange(0, y, B, F, N)
            Fa, Na = P * F, N // P
        b = t + "transformer_blocks.0."
        for i in range(2):
            a = b + f"attention_blocks.{i}."
            n = ops.layernorm(y, W[b + f"norms.{i}.weight"], W[b + f"norms.{i}.bias"])
            qkv = ops.gemm(n, W[a + "to_qkv.weight"], rowvec=self._pe_table(a + "pe_qkv", B, Fa), rows_per_group=Na)
            o = ops.temporal_attention(qkv, B=B, F=Fa, N=Na, H=heads, d=C // heads)
            y = ops.gemm(o, W[a + "to_out.0.weight"], bias=W[a + "to_out.0.bias"], residual=y)
        n = ops.layernorm(y, W[b + "ff_norm.weight"], W[b + "ff_norm.bias"])
        g = ops.gemm(n, W[b + "ff.net.0.proj.weight"], bias=W[b + "ff.net.0.proj.bias"], geglu=True)
        y = ops.gemm(g, W[b + "ff.net.2.weight"], bias=W[b + "ff.net.2.bias"], residual=y)
        if self._shard is not None:
            y = self._exchange(1, y, B, F, N)
        return ops.gemm(y, W[t + "proj_out.weight"], bias=W[t + "proj_out.bias"], residual=x)


class AnimationPipeline(SpatioTemporalStableDiffusionPipeline):
    """``AnimationPipeline.video_style_transfer`` (pipelines/pipeline_animation.py:449-603): the SD loop with the
    late latent AdaIN starting one step earlier (``i >= 0.8 n``, :515) and the trajectory index hard-coded to
    ``50 - i`` (:505-506)."""

    @staticmethod
    def _traj_index(i, n):
        return 50 - i

    def _trajectory(self, path_or_list, n):
        # the reference reads ddim_latents_{50 - i}.pt whatever num_inference_steps is: load exactly the files the loop
        # will touch (indices 50 - n + 1 .. 50), leaving the others None
        if isinstance(path_or_list, (list, tuple)):
            return super()._trajectory(path_or_list, n)
        from .util import load_ddim_latents_at_t
        lat = [None] * 51
        for i in range(n):
            k = 50 - i
            if k < 1:
                raise ValueError(f"{n} steps: the reference's hard-coded index 50 - i (pipeline_animation.py:505) leaves the trajectory")
            lat[k] = load_ddim_latents_at_t(k, path_or_list).to(self.device, torch.float16).contiguous()
        return lat

    @staticmethod
    def _late_adain(i, n):
        return i >= 0.8 * n and i <= 0.9 * n
