"""B200 host-side mirror of the SD3 / SD3.5 MMDiT that the reference's SD3 path drives (BASELINE.json
configs[4]): diffusers' ``SD3Transformer2DModel`` with the forward of the reference's subclass
``CustomSD3Transformer2DModel`` (backbones/video_diffusion_sd3/models/transformer_3D_model.py:12-113 -- a copy of the
library forward plus the ``idx`` / ``ft_indices`` / ``ft_timesteps`` / ``ft_path`` feature dump, :77-84).  Same constructor
data (a state dict with diffusers' key names), same call signature, same ``attn_processors`` / ``set_attn_processor``
protocol, so the reference's processors' mirrors (``univst_b200.sd3``) are installed exactly like the reference installs its
own (run_video_style_transfer_sd3.py:55-63, pnp_utils.py:276-284) and its pipeline loops' mirrors (``univst_b200.
sd3_pipeline``) call it like custom_pipeline.py:316 does.

PARITY UNPINNED outside the attention processors: the MMDiT is third-party (diffusers 0.35.1) and absent from the reference
checkout and from this image, and no weights exist offline; it is restated from the published state-dict layout and
checked against an independent fp32 evaluation of the same definition (oracle/sd3_transformer_oracle.py).  The processors
inside the blocks are pinned to golden vectors of the reference's own classes (tests/test_sd3_gpu.py).

On the GPU a joint block is: one small GEMM for the adaLN modulation of both streams, ``layernorm_modulate`` (LayerNorm
without affine x (1 + scale) + shift per sample), the processor (fused QKV / added-QKV GEMMs, per-head RMS norm, shift,
fused joint attention over [first, previous, self] + text tokens), ``gated_add`` residuals, and the feed-forward as two
tcgen05 GEMMs with the tanh-GELU in the first one's epilogue.  Patch embedding = one GEMM over 2 x 2 patches with the
cropped sin-cos table as its residual; the conditioning MLPs end in an epilogue SiLU because every consumer applies one.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops
from .sd3 import JointAttnProcessor

SD35_MEDIUM_CONFIG = dict(sample_size=128, patch_size=2, in_channels=16, num_layers=24, attention_head_dim=64,
                          num_attention_heads=24, joint_attention_dim=4096, caption_projection_dim=1536,
                          pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=384,
                          dual_attention_layers=tuple(range(13)), qk_norm="rms_norm")


class _Config(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None


class _Attention:
    """What the processors read off a diffusers ``Attention`` module: projections, RMS-norm weights, ``heads``,
    ``context_pre_only`` -- plus the ``processor`` slot ``set_attn_processor`` fills."""

    def __init__(self, W, pre, heads, context_pre_only, joint):
        lin = lambda n: SimpleNamespace(weight=W[pre + n + ".weight"], bias=W[pre + n + ".bias"])
        norm = lambda n: SimpleNamespace(weight=W[pre + n + ".weight"], eps=1e-6) if pre + n + ".weight" in W else None
        self.heads, self.context_pre_only = heads, context_pre_only
        self.to_q, self.to_k, self.to_v = lin("to_q"), lin("to_k"), lin("to_v")
        self.to_out = [lin("to_out.0")]
        self.norm_q, self.norm_k = norm("norm_q"), norm("norm_k")
        self.add_q_proj = self.add_k_proj = self.add_v_proj = self.to_add_out = None
        self.norm_added_q = self.norm_added_k = None
        if joint:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = lin("add_q_proj"), lin("add_k_proj"), lin("add_v_proj")
            self.norm_added_q, self.norm_added_k = norm("norm_added_q"), norm("norm_added_k")
            if not context_pre_only:
                self.to_add_out = lin("to_add_out")
        self.processor = JointAttnProcessor()


class SD3Transformer2DModel:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: Optional[dict] = None, device="cuda"):
        cfg = dict(SD35_MEDIUM_CONFIG)
        cfg.update(config or {})
        self.config = _Config(cfg)
        self.device = torch.device(device)
        self.dtype = torch.float16
        self.out_channels = cfg["out_channels"]
        self.inner_dim = cfg["attention_head_dim"] * cfg["num_attention_heads"]
        h = lambda t: t.detach().to(device=self.device, dtype=torch.float16).contiguous()
        W = {k: h(v) for k, v in state_dict.items()}
        W["pos_embed.proj.weight"] = W["pos_embed.proj.weight"].reshape(self.inner_dim, -1).contiguous()   # [D, C p p]
        self.W = W
        self._pos = {}
        self.transformer_blocks = []
        for i in range(cfg["num_layers"]):
            b = f"transformer_blocks.{i}."
            last = i == cfg["num_layers"] - 1
            blk = SimpleNamespace(context_pre_only=last, use_dual_attention=i in cfg["dual_attention_layers"], prefix=b,
                                  attn=_Attention(W, b + "attn.", cfg["num_attention_heads"], last, joint=True), attn2=None)
            if blk.use_dual_attention:
                blk.attn2 = _Attention(W, b + "attn2.", cfg["num_attention_heads"], False, joint=False)
            self.transformer_blocks.append(blk)

    # ------------------------------------------------------------------------------------------ processor protocol
    def _attentions(self):
        for i, blk in enumerate(self.transformer_blocks):
            yield f"transformer_blocks.{i}.attn.processor", blk.attn
            if blk.attn2 is not None:
                yield f"transformer_blocks.{i}.attn2.processor", blk.attn2

    @property
    def attn_processors(self):
        return {name: a.processor for name, a in self._attentions()}

    def set_attn_processor(self, processor):
        """diffusers semantics: one processor for every attention, or a dict keyed like :attr:`attn_processors`."""
        for name, a in self._attentions():
            a.processor = processor[name] if isinstance(processor, dict) else processor

    # ------------------------------------------------------------------------------------------ frame sharding
    def set_frame_sharding(self, group=None, frames_per_clip: int = 16):
        """Shard the ``frames_per_clip`` frames of every branch over the ranks of ``group`` (BASELINE.json
        configs[4]).  Everything in the MMDiT is per image except the cross-frame attention of the reference's processors,
        whose [first, previous] K/V cross ranks: per attention the boundary frame's K|V go to the next rank's halo bank and
        the clip's first frame's to every rank's, stored over NVLink by one kernel whose tail is the synchronisation
        (xrank.push_kv_halo); the velocity prediction of the local frames is stored into every rank's full-batch buffer
        at the end.  No collective-library call.  The callers keep passing (and receiving) the whole batch."""
        import torch.distributed as dist
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world == 1:
            return self.set_frame_sharding_off()
        if frames_per_clip % world:
            raise ValueError(f"{frames_per_clip} frames do not shard evenly over {world} ranks")
        from .xrank import HaloBuffers, XRank
        if getattr(self, "_xr", None) is None:
            self._xr = XRank(group, self.device)
        self._shard = SimpleNamespace(xr=self._xr, Fl=frames_per_clip // world, F=frames_per_clip,
                                      buffers=HaloBuffers(self._xr, "sd3qkv"))
        for _, a in self._attentions():
            a._shard = self._shard

    def set_frame_sharding_off(self):
        self._shard = None
        for _, a in self._attentions():
            a._shard = None

    def _local(self, t, B, F, Fl, rank):
        """Rows (branch, frame) of a batch-major tensor -> this rank's frames of every branch."""
        return t.view(B, F, *t.shape[1:])[:, rank * Fl:(rank + 1) * Fl].reshape(B * Fl, *t.shape[1:]).contiguous()

    # ------------------------------------------------------------------------------------------ pieces
    def _pos_rows(self, BF, h, w):
        """The cropped sin-cos table (diffusers PatchEmbed.cropped_pos_embed: centre crop of the max x max grid) repeated for
        every image, as the residual of the patch-embedding GEMM."""
        key = (BF, h, w)
        if key not in self._pos:
            mx, D = self.config["pos_embed_max_size"], self.inner_dim
            if h > mx or w > mx:
                raise ValueError(f"{h} x {w} patches exceed pos_embed_max_size {mx}")
            top, left = (mx - h) // 2, (mx - w) // 2
            pos = self.W["pos_embed.pos_embed"].reshape(mx, mx, D)[top:top + h, left:left + w].reshape(h * w, D)
            self._pos[key] = pos.repeat(BF, 1).contiguous()
        return self._pos[key]

    def _ff(self, pre, x):
        W = self.W
        g = ops.gemm(x, W[pre + "net.0.proj.weight"], bias=W[pre + "net.0.proj.bias"], act="gelu_tanh")
        return ops.gemm(g, W[pre + "net.2.weight"], bias=W[pre + "net.2.bias"])

    def _block(self, blk, hs, ctx, emb, BF, N, L, kwargs):
        """diffusers JointTransformerBlock.forward (adaLN-Zero on both streams; SD3.5: a second, image-only attention)."""
        W, D, b = self.W, self.inner_dim, blk.prefix
        m = ops.gemm(emb, W[b + "norm1.linear.weight"], bias=W[b + "norm1.linear.bias"])          # [BF, 6D | 9D]
        col = lambda t, k: t[:, k * D:(k + 1) * D]
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = (col(m, k) for k in range(6))
        nh = ops.layernorm_modulate(hs, scale_msa, shift_msa, N)
        # SD35AdaLayerNormZeroX: the block's INPUT is modulated a second time for the image-only attention
        nh2 = ops.layernorm_modulate(hs, col(m, 7), col(m, 6), N) if blk.use_dual_attention else None
        c = ops.gemm(emb, W[b + "norm1_context.linear.weight"], bias=W[b + "norm1_context.linear.bias"])
        if blk.context_pre_only:      # AdaLayerNormContinuous: (scale, shift)
            nc = ops.layernorm_modulate(ctx, col(c, 0), col(c, 1), L)
        else:                         # AdaLayerNormZero: (shift, scale, gate) x (msa, mlp)
            nc = ops.layernorm_modulate(ctx, col(c, 1), col(c, 0), L)
        ao, co = blk.attn.processor(blk.attn, nh.view(BF, N, D), encoder_hidden_states=nc.view(BF, L, D), **kwargs)
        hs = ops.gated_add(hs, ao.reshape(BF * N, D).contiguous(), gate_msa, N)
        if blk.use_dual_attention:
            a2 = blk.attn2.processor(blk.attn2, nh2.view(BF, N, D), **kwargs)
            hs = ops.gated_add(hs, a2.reshape(BF * N, D).contiguous(), col(m, 8), N)
        nh = ops.layernorm_modulate(hs, scale_mlp, shift_mlp, N)
        hs = ops.gated_add(hs, self._ff(b + "ff.", nh), gate_mlp, N)
        if blk.context_pre_only:
            return None, hs
        ctx = ops.gated_add(ctx, co.reshape(BF * L, D).contiguous(), col(c, 2), L)
        nc = ops.layernorm_modulate(ctx, col(c, 4), col(c, 3), L)
        ctx = ops.gated_add(ctx, self._ff(b + "ff_context.", nc), col(c, 5), L)
        return ctx, hs

    # ------------------------------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, hidden_states, encoder_hidden_states=None, pooled_projections=None, timestep=None,
                block_controlnet_hidden_states=None, joint_attention_kwargs=None, return_dict=True, skip_layers=None,
                idx=0, ft_indices=None, ft_timesteps=None, ft_path=None):
        if block_controlnet_hidden_states is not None:
            raise NotImplementedError("ControlNet residuals are unused on the UniVST path")
        kwargs = dict(joint_attention_kwargs or {})
        kwargs.pop("scale", None)
        if "ip_adapter_image_embeds" in kwargs:
            raise NotImplementedError("IP-Adapter inputs are unused on the UniVST path")
        W, cfg, D, dev = self.W, self.config, self.inner_dim, self.device
        shard = getattr(self, "_shard", None)
        BF_total = hidden_states.shape[0]
        if shard is not None:
            if BF_total % shard.F:
                raise ValueError(f"the batch must hold whole clips of {shard.F} frames under frame sharding")
            if ft_path is not None:
                raise NotImplementedError("feature dumps are not available under frame sharding")
            Bc, rk = BF_total // shard.F, shard.xr.rank
            loc = lambda t: self._local(t.to(dev), Bc, shard.F, shard.Fl, rk)
            hidden_states, encoder_hidden_states, pooled_projections = loc(hidden_states), loc(encoder_hidden_states), loc(pooled_projections)
            if torch.is_tensor(timestep) and timestep.numel() == BF_total:
                timestep = loc(timestep.reshape(-1))
        BF, Cin, H, Wd = hidden_states.shape
        p = cfg["patch_size"]
        if H % p or Wd % p:
            raise ValueError(f"latent height / width must be multiples of the patch size {p}")
        h, w = H // p, Wd // p
        N = h * w
        # patch embedding: Conv2d(k = stride = p) == GEMM over (c, py, px) patches; + cropped positional table
        x16 = hidden_states.to(device=dev, dtype=torch.float16)
        patches = x16.view(BF, Cin, h, p, w, p).permute(0, 2, 4, 1, 3, 5).reshape(BF * N, Cin * p * p).contiguous()
        hs = ops.gemm(patches, W["pos_embed.proj.weight"], bias=W["pos_embed.proj.bias"], residual=self._pos_rows(BF, h, w))
        # conditioning: silu(timestep_embedder(t) + text_embedder(pooled)) -- every adaLN applies the SiLU first
        t = timestep if torch.is_tensor(timestep) else torch.tensor([float(timestep)])
        t = t.to(device=dev, dtype=torch.float32).reshape(-1)
        t = (t.expand(BF) if t.numel() == 1 else t).contiguous()
        te = ops.gemm(ops.timestep_embedding(t, 256), W["time_text_embed.timestep_embedder.linear_1.weight"],
                      bias=W["time_text_embed.timestep_embedder.linear_1.bias"], act="silu")
        te = ops.gemm(te, W["time_text_embed.timestep_embedder.linear_2.weight"],
                      bias=W["time_text_embed.timestep_embedder.linear_2.bias"])
        pooled = pooled_projections.to(device=dev, dtype=torch.float16).contiguous()
        pe = ops.gemm(pooled, W["time_text_embed.text_embedder.linear_1.weight"],
                      bias=W["time_text_embed.text_embedder.linear_1.bias"], act="silu")
        emb = ops.gemm(pe, W["time_text_embed.text_embedder.linear_2.weight"],
                       bias=W["time_text_embed.text_embedder.linear_2.bias"], residual=te, act="silu")
        enc = encoder_hidden_states.to(device=dev, dtype=torch.float16)
        L = enc.shape[1]
        ctx = ops.gemm(enc.reshape(BF * L, -1).contiguous(), W["context_embedder.weight"], bias=W["context_embedder.bias"])
        for i, blk in enumerate(self.transformer_blocks):
            if skip_layers is not None and i in skip_layers:
                continue
            ctx, hs = self._block(blk, hs, ctx, emb, BF, N, L, kwargs)
            if ft_indices is not None and ft_timesteps and ft_path is not None and i in ft_indices and idx in ft_timesteps:
                path = os.path.join(ft_path, f"inversion_feature_map_{i}_block_{idx}_step.pt")   # :77-84
                print(f"save feature map at: {path}")
                torch.save(hs.view(BF, h, w, -1).detach().clone(), path)
        m = ops.gemm(emb, W["norm_out.linear.weight"], bias=W["norm_out.linear.bias"])     # AdaLayerNormContinuous: (scale, shift)
        hs = ops.layernorm_modulate(hs, m[:, :D], m[:, D:], N)
        out = ops.gemm(hs, W["proj_out.weight"], bias=W["proj_out.bias"])                   # [BF N, p p C_out]
        co = self.out_channels
        out = torch.einsum("nhwpqc->nchpwq", out.view(BF, h, w, p, p, co)).reshape(BF, co, h * p, w * p)
        if shard is not None:   # my frames -> their place in every rank's full-batch buffer (one multicast store per 16 B)
            per = co * H * Wd
            key = ("sd3out", BF_total, per)
            full, ptrs = shard.xr.buffer(key, (BF_total, per))
            mc = shard.xr.multicast(key)
            off = shard.xr.rank * shard.Fl * per * 2
            ops.xrank_push(shard.xr, [dict(src=out.contiguous().view(BF, per), src_blk_rows=shard.Fl, dst=[q + off for q in ptrs],
                                           ld_dst=per, dst_blk_rows=shard.F, nblk=BF_total // shard.F, rows=shard.Fl,
                                           mc=mc + off if mc else 0)])
            out = full.view(BF_total, co, H, Wd)
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)


CustomSD3Transformer2DModel = SD3Transformer2DModel   # the reference's subclass name (transformer_3D_model.py:12)


def random_state_dict(cfg=None, seed: int = 71, device="cuda"):
    """Seeded random weights with diffusers' key names and shapes (no checkpoint exists offline); the positional table is
    the sin-cos one diffusers builds."""
    import math
    cfg = dict(SD35_MEDIUM_CONFIG, **(cfg or {}))
    D, p, hd = cfg["attention_head_dim"] * cfg["num_attention_heads"], cfg["patch_size"], cfg["attention_head_dim"]
    g = torch.Generator(device=device).manual_seed(seed)
    shapes: Dict[str, tuple] = {}
    lin = lambda k, o, i: shapes.update({k + ".weight": (o, i), k + ".bias": (o,)})
    shapes["pos_embed.proj.weight"], shapes["pos_embed.proj.bias"] = (D, cfg["in_channels"], p, p), (D,)
    lin("time_text_embed.timestep_embedder.linear_1", D, 256)
    lin("time_text_embed.timestep_embedder.linear_2", D, D)
    lin("time_text_embed.text_embedder.linear_1", D, cfg["pooled_projection_dim"])
    lin("time_text_embed.text_embedder.linear_2", D, D)
    lin("context_embedder", D, cfg["joint_attention_dim"])
    for i in range(cfg["num_layers"]):
        b = f"transformer_blocks.{i}."
        last, dual = i == cfg["num_layers"] - 1, i in cfg["dual_attention_layers"]
        lin(b + "norm1.linear", (9 if dual else 6) * D, D)
        lin(b + "norm1_context.linear", (2 if last else 6) * D, D)
        for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0"):
            lin(b + "attn." + n, D, D)
        if not last:
            lin(b + "attn.to_add_out", D, D)
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            shapes[b + "attn." + n + ".weight"] = (hd,)
        if dual:
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(b + "attn2." + n, D, D)
            for n in ("norm_q", "norm_k"):
                shapes[b + "attn2." + n + ".weight"] = (hd,)
        lin(b + "ff.net.0.proj", 4 * D, D)
        lin(b + "ff.net.2", D, 4 * D)
        if not last:
            lin(b + "ff_context.net.0.proj", 4 * D, D)
            lin(b + "ff_context.net.2", D, 4 * D)
    lin("norm_out.linear", 2 * D, D)
    lin("proj_out", p * p * cfg["out_channels"], D)
    out = {}
    for k, shp in shapes.items():
        if ".norm_" in k:
            t = torch.ones(shp, device=device)
        elif k.endswith("bias"):
            t = 0.02 * torch.randn(shp, device=device, generator=g)
        else:
            fan = 1
            for d in shp[1:]:
                fan *= d
            t = torch.randn(shp, device=device, generator=g) * fan ** -0.5
        out[k] = t.half()
    # diffusers get_2d_sincos_pos_embed(D, max, base_size = sample_size / patch): (max^2, D), width coordinate first
    mx, base = cfg["pos_embed_max_size"], cfg["sample_size"] // p
    coord = torch.arange(mx, dtype=torch.float64, device=device) / (mx / base)
    gw, gh = torch.meshgrid(coord, coord, indexing="xy")

    def one(dim, pos):
        omega = 1.0 / 10000 ** (torch.arange(dim // 2, dtype=torch.float64, device=device) / (dim / 2.0))
        o = pos.reshape(-1)[:, None] * omega[None]
        return torch.cat([o.sin(), o.cos()], 1)
    out["pos_embed.pos_embed"] = torch.cat([one(D // 2, gw), one(D // 2, gh)], 1)[None].half()
    return out
