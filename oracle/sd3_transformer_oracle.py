"""ORACLE (test infrastructure, not product code): plain-PyTorch restatement of diffusers' ``SD3Transformer2DModel`` forward
as the reference drives it (backbones/video_diffusion_sd3/models/transformer_3D_model.py:12-113 is a verbatim copy of the
library forward plus the feature dump; custom_pipeline.py:316 calls it with ``joint_attention_kwargs={'idx': i}``).

PARITY UNPINNED for everything outside the attention processors: the MMDiT (patch embedding, time / text conditioning,
adaLN-Zero joint blocks, SD3.5 dual-attention blocks, final adaLN + projection) is third-party (diffusers 0.35.1,
``models/transformers/transformer_sd3.py``, ``models/attention.py::JointTransformerBlock``, ``models/normalization.py``,
``models/embeddings.py``), absent from /root/reference and from this image, and no weights exist offline: it is restated
here from the published state-dict layout.  The attention inside the blocks IS pinned: it is oracle/sd3_oracle.py, checked
against golden vectors of the reference's own processor classes.  Only tests/ may import this.

Functional over a state dict with diffusers' key names."""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from . import sd3_oracle as so

SD35_MEDIUM_CONFIG = dict(sample_size=128, patch_size=2, in_channels=16, num_layers=24, attention_head_dim=64,
                          num_attention_heads=24, joint_attention_dim=4096, caption_projection_dim=1536,
                          pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=384,
                          dual_attention_layers=tuple(range(13)), qk_norm="rms_norm")
TINY_CONFIG = dict(SD35_MEDIUM_CONFIG, num_layers=3, attention_head_dim=32, num_attention_heads=4, joint_attention_dim=96,
                   caption_projection_dim=128, pooled_projection_dim=64, pos_embed_max_size=24, dual_attention_layers=(0,))


def inner_dim(cfg):
    return cfg["attention_head_dim"] * cfg["num_attention_heads"]


def sincos_pos_embed(embed_dim: int, grid: int, base_size: int, patch_size_scale: float = 1.0) -> torch.Tensor:
    """diffusers get_2d_sincos_pos_embed(embed_dim, grid, base_size=base_size, interpolation_scale=1): (grid^2, embed_dim)."""
    gh = torch.arange(grid, dtype=torch.float64) / (grid / base_size) / patch_size_scale
    gw = torch.arange(grid, dtype=torch.float64) / (grid / base_size) / patch_size_scale
    g = torch.stack(torch.meshgrid(gw, gh, indexing="xy"), 0).reshape(2, 1, grid, grid)     # w first

    def one(dim, pos):
        omega = 1.0 / 10000 ** (torch.arange(dim // 2, dtype=torch.float64) / (dim / 2.0))
        out = pos.reshape(-1)[:, None] * omega[None]
        return torch.cat([out.sin(), out.cos()], 1)
    return torch.cat([one(embed_dim // 2, g[0]), one(embed_dim // 2, g[1])], 1).float()


def param_shapes(cfg) -> Dict[str, tuple]:
    D, p, hd = inner_dim(cfg), cfg["patch_size"], cfg["attention_head_dim"]
    s: Dict[str, tuple] = {}
    lin = lambda k, o, i: s.update({k + ".weight": (o, i), k + ".bias": (o,)})
    s["pos_embed.proj.weight"], s["pos_embed.proj.bias"] = (D, cfg["in_channels"], p, p), (D,)
    s["pos_embed.pos_embed"] = (1, cfg["pos_embed_max_size"] ** 2, D)
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
            s[b + "attn." + n + ".weight"] = (hd,)
        if dual:
            for n in ("to_q", "to_k", "to_v", "to_out.0"):
                lin(b + "attn2." + n, D, D)
            for n in ("norm_q", "norm_k"):
                s[b + "attn2." + n + ".weight"] = (hd,)
        lin(b + "ff.net.0.proj", 4 * D, D)
        lin(b + "ff.net.2", D, 4 * D)
        if not last:
            lin(b + "ff_context.net.0.proj", 4 * D, D)
            lin(b + "ff_context.net.2", D, 4 * D)
    lin("norm_out.linear", 2 * D, D)
    lin("proj_out", p * p * cfg["out_channels"], D)
    return s


def seeded_state_dict(cfg, seed: int = 71) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if k == "pos_embed.pos_embed":
            sd[k] = sincos_pos_embed(shp[2], cfg["pos_embed_max_size"], cfg["sample_size"] // cfg["patch_size"])[None]
        elif ".norm_" in k:
            sd[k] = 1.0 + 0.2 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.05 * torch.randn(shp, generator=g)
        else:
            fan = 1
            for d in shp[1:]:
                fan *= d
            sd[k] = torch.randn(shp, generator=g) * fan ** -0.5
    return sd


def _ln(x):
    return F.layer_norm(x, x.shape[-1:], eps=1e-6)


def _lin(sd, k, x):
    return F.linear(x, sd[k + ".weight"], sd[k + ".bias"])


def _attn_weights(sd, pre):
    return {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}


def timestep_embedding(t, dim=256):
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    e = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(e), torch.sin(e)], -1)


def forward(sd, cfg, x, enc, pooled, timestep, idx=None, eta1=0.0, eta2=0.6, cross_frame=True, feature_blocks=()):
    """x (BF, C, H, W), enc (BF, L, joint_dim), pooled (BF, pooled_dim), timestep (BF,) -> (BF, C_out, H, W) [, features].
    ``cross_frame`` / ``idx``: which attention processor sits in the blocks (sd3_oracle.joint_attention)."""
    D, p, heads = inner_dim(cfg), cfg["patch_size"], cfg["num_attention_heads"]
    BF, C, H, W = x.shape
    h, w = H // p, W // p
    hs = F.conv2d(x, sd["pos_embed.proj.weight"], sd["pos_embed.proj.bias"], stride=p).flatten(2).transpose(1, 2)
    mx = cfg["pos_embed_max_size"]
    top, left = (mx - h) // 2, (mx - w) // 2
    pos = sd["pos_embed.pos_embed"].reshape(1, mx, mx, D)[:, top:top + h, left:left + w].reshape(1, h * w, D)
    hs = hs + pos
    te = _lin(sd, "time_text_embed.timestep_embedder.linear_2", F.silu(_lin(sd, "time_text_embed.timestep_embedder.linear_1",
                                                                          timestep_embedding(timestep))))
    pe = _lin(sd, "time_text_embed.text_embedder.linear_2", F.silu(_lin(sd, "time_text_embed.text_embedder.linear_1", pooled)))
    emb = F.silu(te + pe)                                   # every consumer applies SiLU first
    ctx = _lin(sd, "context_embedder", enc)
    feats = {}
    for i in range(cfg["num_layers"]):
        b = f"transformer_blocks.{i}."
        last, dual = i == cfg["num_layers"] - 1, i in cfg["dual_attention_layers"]
        m = _lin(sd, b + "norm1.linear", emb)[:, None].chunk(9 if dual else 6, dim=-1)
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = m[:6]
        nh0 = _ln(hs)
        nh = nh0 * (1 + scale_msa) + shift_msa
        if last:
            scale, shift = _lin(sd, b + "norm1_context.linear", emb)[:, None].chunk(2, dim=-1)
            nc = _ln(ctx) * (1 + scale) + shift
        else:
            c_shift_msa, c_scale_msa, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = \
                _lin(sd, b + "norm1_context.linear", emb)[:, None].chunk(6, dim=-1)
            nc = _ln(ctx) * (1 + c_scale_msa) + c_shift_msa
        ao, co = so.joint_attention(_attn_weights(sd, b + "attn."), nh, nc, heads, idx=idx, eta1=eta1, eta2=eta2,
                                    cross_frame=cross_frame)
        hs = hs + gate_msa * ao
        if dual:
            nh2 = nh0 * (1 + m[7]) + m[6]
            hs = hs + m[8] * so.joint_attention(_attn_weights(sd, b + "attn2."), nh2, None, heads, idx=idx, eta1=eta1,
                                                eta2=eta2, cross_frame=cross_frame)
        nh = _ln(hs) * (1 + scale_mlp) + shift_mlp
        hs = hs + gate_mlp * _lin(sd, b + "ff.net.2", F.gelu(_lin(sd, b + "ff.net.0.proj", nh), approximate="tanh"))
        if not last:
            ctx = ctx + c_gate_msa * co
            nc = _ln(ctx) * (1 + c_scale_mlp) + c_shift_mlp
            ctx = ctx + c_gate_mlp * _lin(sd, b + "ff_context.net.2", F.gelu(_lin(sd, b + "ff_context.net.0.proj", nc),
                                                                          approximate="tanh"))
        if i in feature_blocks:
            feats[i] = hs.view(BF, h, w, -1).clone()
    scale, shift = _lin(sd, "norm_out.linear", emb)[:, None].chunk(2, dim=-1)
    hs = _lin(sd, "proj_out", _ln(hs) * (1 + scale) + shift)
    co_ = cfg["out_channels"]
    out = torch.einsum("nhwpqc->nchpwq", hs.reshape(BF, h, w, p, p, co_)).reshape(BF, co_, h * p, w * p)
    return (out, feats) if feature_blocks else out
