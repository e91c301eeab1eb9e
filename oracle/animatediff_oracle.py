"""ORACLE (test infrastructure, not product code): CPU / plain-PyTorch restatement of the reference's AnimateDiff-v2
backbone -- the inflated SD-1.5 UNet with ``VanillaTemporalModule`` motion modules -- including the AnimateDiff
flavour of the AdaIN-guided attention patch.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU arm may import this.
Parity status: PINNED against the reference's own module code -- ``oracle/gen_golden_animatediff.py`` runs
``/root/reference/backbones/animatediff/models/unet.py`` (``UNet3DConditionModel`` with the ``animatediff-v2.yaml``
kwargs, on the test-only diffusers shim in ``oracle/_shim``) and ``backbones/animatediff/pnp_utils.py`` on seeded
weights and commits input/output vectors under ``tests/golden/animatediff_tiny.pt``; ``tests/test_oracle_cpu.py``
checks this file against them.

File:line citations are relative to ``/root/reference/backbones/animatediff``.  Differences to the SD "pseudo-3D"
backbone (oracle/unet_oracle.py): GroupNorm statistics are per frame (``InflatedGroupNorm``, models/resnet.py:21-29),
attn1 is plain per-frame self-attention (``unet_use_cross_frame_attention`` False, models/attention.py:330-333, so
the patched forward is called with ``clip_length=None``, pnp_utils.py:57), there is no (dead) temporal attention in
the transformer block, and a motion module follows every (resnet, attention) pair (models/unet_blocks.py:407-411).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import unet_oracle as uo

AD_SD15_CONFIG = dict(uo.SD15_CONFIG, motion_heads=8, motion_max_len=24)
AD_TINY_CONFIG = dict(uo.TINY_CONFIG, motion_heads=8, motion_max_len=24)
PATCHED = uo.PATCHED  # the same eight decoder layers (pnp_utils.py:8, :101)


def shift_params(idx: int, eta1: float = 0.0, eta2: float = 0.5):
    """pnp_utils.py:45-49: window ``eta1*50 <= idx < eta2*50``, alpha 0.8, gamma 2.0, beta 0.9 -> 0.1."""
    active = idx >= eta1 * 50 and idx < eta2 * 50
    beta = (0.9 - 0.1) / (eta1 * 50 - eta2 * 50) * (idx - eta2 * 50) + 0.1
    return active, 0.8, beta, 2.0


def self_attention(sd, pre, x, heads, patched: bool, idx: Optional[int], eta1=0.0, eta2=0.5):
    """attn1: diffusers Attention (models/attention.py:230-237) or the patched forward (pnp_utils.py:20-98) called
    without ``clip_length`` -> K/V of the frame itself.  x: (3*F, N, C) with the branches content | style | edit."""
    q = F.linear(x, sd[pre + "to_q.weight"])
    k = F.linear(x, sd[pre + "to_k.weight"])
    v = F.linear(x, sd[pre + "to_v.weight"])
    BF, N, C = q.shape
    if patched:
        chunk = BF // 3
        active, alpha, beta, gamma = shift_params(idx, eta1, eta2)
        if active:
            q, k, v = q.clone(), k.clone(), v.clone()
            q[2 * chunk:] = alpha * q[:chunk] + (1 - alpha) * q[2 * chunk:]
            k[2 * chunk:] = beta * uo.attention_adain(k[2 * chunk:], k[chunk:2 * chunk]) + (1 - beta) * k[chunk:2 * chunk]
            v[2 * chunk:] = beta * uo.attention_adain(v[2 * chunk:], v[chunk:2 * chunk]) + (1 - beta) * v[chunk:2 * chunk]
            q[2 * chunk:] = gamma * q[2 * chunk:]
    d = C // heads
    q = q.view(BF, -1, heads, d).transpose(1, 2)
    k = k.view(BF, -1, heads, d).transpose(1, 2)
    v = v.view(BF, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(BF, -1, C)
    return F.linear(o, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])


def transformer(sd, pre, x, ctx, heads, F_, patched, idx, cfg, eta1=0.0, eta2=0.5):
    """Transformer3DModel.forward (models/attention.py:163-212) with its BasicTransformerBlock (:319-361; no
    temporal attention: ``unet_use_temporal_attention`` False).  x: (B*F, C, h, w); ctx: (B, L, D)."""
    BF, C, h, w = x.shape
    res = x
    ctx_rep = ctx.repeat_interleave(F_, 0)
    y = F.group_norm(x, cfg["norm_num_groups"], sd[pre + "norm.weight"], sd[pre + "norm.bias"], eps=1e-6)
    y = F.conv2d(y, sd[pre + "proj_in.weight"], sd[pre + "proj_in.bias"])
    y = y.permute(0, 2, 3, 1).reshape(BF, h * w, C)
    b = pre + "transformer_blocks.0."
    ln = lambda t, name: F.layer_norm(t, (C,), sd[b + name + ".weight"], sd[b + name + ".bias"])
    y = self_attention(sd, b + "attn1.", ln(y, "norm1"), heads, patched, idx, eta1, eta2) + y
    y = uo.cross_attention(sd, b + "attn2.", ln(y, "norm2"), ctx_rep, heads) + y
    y = uo.feed_forward(sd, b + "ff.", ln(y, "norm3")) + y
    y = y.reshape(BF, h, w, C).permute(0, 3, 1, 2)
    y = F.conv2d(y, sd[pre + "proj_out.weight"], sd[pre + "proj_out.bias"])
    return y + res


def positional_encoding(d_model: int, max_len: int) -> torch.Tensor:
    """PositionalEncoding buffer (models/motion_module.py:232-243): (max_len, d_model), sin on even, cos on odd."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def temporal_self_attention(sd, pre, x, heads, F_, pe):
    """VersatileAttention.forward, mode "Temporal" (models/motion_module.py:276-336): tokens of one pixel across the F
    frames attend to each other; the positional encoding is added to the (already layer-normed) input of q, k and v;
    explicit softmax path (get_attention_scores + bmm).  x: (B*F, N, C)."""
    BF, N, C = x.shape
    B = BF // F_
    h = x.view(B, F_, N, C).permute(0, 2, 1, 3).reshape(B * N, F_, C)
    h = h + pe[:F_].to(h.dtype)[None]
    q, k, v = F.linear(h, sd[pre + "to_q.weight"]), F.linear(h, sd[pre + "to_k.weight"]), F.linear(h, sd[pre + "to_v.weight"])
    d = C // heads
    q = q.view(B * N, F_, heads, d).transpose(1, 2)
    k = k.view(B * N, F_, heads, d).transpose(1, 2)
    v = v.view(B * N, F_, heads, d).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B * N, F_, C)
    o = F.linear(o, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])
    return o.view(B, N, F_, C).permute(0, 2, 1, 3).reshape(BF, N, C)


def motion_module(sd, pre, x, F_, cfg):
    """VanillaTemporalModule -> TemporalTransformer3DModel.forward (models/motion_module.py:138-163) with one
    TemporalTransformerBlock (:218-229; two "Temporal_Self" attention blocks, then GEGLU FF).  x: (B*F, C, h, w)."""
    BF, C, h, w = x.shape
    t = pre + "temporal_transformer."
    res = x
    y = F.group_norm(x, cfg["norm_num_groups"], sd[t + "norm.weight"], sd[t + "norm.bias"], eps=1e-6)
    y = y.permute(0, 2, 3, 1).reshape(BF, h * w, C)
    y = F.linear(y, sd[t + "proj_in.weight"], sd[t + "proj_in.bias"])
    b = t + "transformer_blocks.0."
    pe = positional_encoding(C, cfg["motion_max_len"]).to(x.device)
    for i in range(2):
        n = F.layer_norm(y, (C,), sd[b + f"norms.{i}.weight"], sd[b + f"norms.{i}.bias"])
        y = temporal_self_attention(sd, b + f"attention_blocks.{i}.", n, cfg["motion_heads"], F_, pe) + y
    n = F.layer_norm(y, (C,), sd[b + "ff_norm.weight"], sd[b + "ff_norm.bias"])
    y = uo.feed_forward(sd, b + "ff.", n) + y
    y = F.linear(y, sd[t + "proj_out.weight"], sd[t + "proj_out.bias"])
    y = y.reshape(BF, h, w, C).permute(0, 3, 1, 2)
    return y + res


def conv(sd, pre, x, stride=1, padding=1):
    """InflatedConv3d.forward (models/resnet.py:12-19): the 2-D convolution per frame."""
    return F.conv2d(x, sd[pre + "weight"], sd[pre + "bias"], stride=stride, padding=padding)


def resnet(sd, pre, x, emb, F_, cfg):
    """ResnetBlock3D.forward (models/resnet.py:180-209) with InflatedGroupNorm (per-frame statistics, :21-29)."""
    g, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    h = F.silu(F.group_norm(x, g, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps))
    h = conv(sd, pre + "conv1.", h)
    temb = F.linear(F.silu(emb), sd[pre + "time_emb_proj.weight"], sd[pre + "time_emb_proj.bias"])
    h = h + temb.repeat_interleave(F_, 0)[:, :, None, None]
    h = F.silu(F.group_norm(h, g, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps))
    h = conv(sd, pre + "conv2.", h)
    if pre + "conv_shortcut.weight" in sd:
        x = conv(sd, pre + "conv_shortcut.", x, padding=0)
    return x + h


def unet_forward(sd: Dict[str, torch.Tensor], cfg, sample, timestep, ctx, *, patched: bool = False,
                 idx: Optional[int] = None, eta1: float = 0.0, eta2: float = 0.5, features: Optional[dict] = None):
    """UNet3DConditionModel.forward (models/unet.py:322-466) with the block forwards of models/unet_blocks.py
    (:271-278 mid, :382-421 / :493-521 down, :621-667 / :735-760 up).  sample (B, 4, F, h, w); ctx (B, L, D)."""
    B, Cin, F_, h, w = sample.shape
    dt = sample.dtype
    boc = cfg["block_out_channels"]
    x = sample.permute(0, 2, 1, 3, 4).reshape(B * F_, Cin, h, w)
    t = torch.full((B,), float(timestep), device=sample.device)
    temb = uo.timestep_embedding(t, boc[0]).to(dt)
    emb = F.linear(F.silu(F.linear(temb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])),
                   sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
    x = conv(sd, "conv_in.", x)
    skips = [x]
    nlev = len(boc)
    for i in range(nlev):
        for j in range(cfg["layers_per_block"]):
            x = resnet(sd, f"down_blocks.{i}.resnets.{j}.", x, emb, F_, cfg)
            if i < nlev - 1:
                x = transformer(sd, f"down_blocks.{i}.attentions.{j}.", x, ctx, uo.heads_of(cfg, i), F_, False, idx, cfg)
            x = motion_module(sd, f"down_blocks.{i}.motion_modules.{j}.", x, F_, cfg)
            skips.append(x)
        if i < nlev - 1:
            x = conv(sd, f"down_blocks.{i}.downsamplers.0.conv.", x, stride=2)
            skips.append(x)
    x = resnet(sd, "mid_block.resnets.0.", x, emb, F_, cfg)
    x = transformer(sd, "mid_block.attentions.0.", x, ctx, uo.heads_of(cfg, nlev - 1), F_, False, idx, cfg)
    x = motion_module(sd, "mid_block.motion_modules.0.", x, F_, cfg)
    x = resnet(sd, "mid_block.resnets.1.", x, emb, F_, cfg)
    for i in range(nlev):
        for j in range(cfg["layers_per_block"] + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet(sd, f"up_blocks.{i}.resnets.{j}.", x, emb, F_, cfg)
            if i > 0:
                x = transformer(sd, f"up_blocks.{i}.attentions.{j}.", x, ctx, uo.heads_of(cfg, nlev - 1 - i), F_,
                                patched and (i, j) in PATCHED, idx, cfg, eta1, eta2)
            x = motion_module(sd, f"up_blocks.{i}.motion_modules.{j}.", x, F_, cfg)
        if i < nlev - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv(sd, f"up_blocks.{i}.upsamplers.0.conv.", x)
        if features is not None:
            features[i] = x.view(B, F_, -1, x.shape[-2], x.shape[-1]).permute(0, 2, 1, 3, 4)
    x = F.silu(F.group_norm(x, cfg["norm_num_groups"], sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], cfg["norm_eps"]))
    x = conv(sd, "conv_out.", x)
    return x.view(B, F_, -1, h, w).permute(0, 2, 1, 3, 4)


# ----------------------------------------------------------------------------------------------- seeded weights
def unet_param_shapes(cfg) -> Dict[str, tuple]:
    """Key -> shape of the reference ``UNet3DConditionModel`` tree with the animatediff-v2.yaml kwargs (verified
    against the reference constructor in gen_golden_animatediff.py): the SD tree without the ``*_temporal*`` keys,
    plus one motion module per (resnet, attention) pair."""
    s = {k: v for k, v in uo.unet_param_shapes(cfg).items() if "_temporal" not in k}
    boc, lpb, nlev = cfg["block_out_channels"], cfg["layers_per_block"], len(cfg["block_out_channels"])

    def mm(pre, c):
        t = pre + "temporal_transformer."
        s[t + "norm.weight"], s[t + "norm.bias"] = (c,), (c,)
        s[t + "proj_in.weight"], s[t + "proj_in.bias"] = (c, c), (c,)
        s[t + "proj_out.weight"], s[t + "proj_out.bias"] = (c, c), (c,)
        b = t + "transformer_blocks.0."
        for i in range(2):
            a = b + f"attention_blocks.{i}."
            s[a + "to_q.weight"] = s[a + "to_k.weight"] = s[a + "to_v.weight"] = s[a + "to_out.0.weight"] = (c, c)
            s[a + "to_out.0.bias"] = (c,)
            s[b + f"norms.{i}.weight"], s[b + f"norms.{i}.bias"] = (c,), (c,)
        s[b + "ff.net.0.proj.weight"], s[b + "ff.net.0.proj.bias"] = (8 * c, c), (8 * c,)
        s[b + "ff.net.2.weight"], s[b + "ff.net.2.bias"] = (c, 4 * c), (c,)
        s[b + "ff_norm.weight"], s[b + "ff_norm.bias"] = (c,), (c,)

    for i in range(nlev):
        for j in range(lpb):
            mm(f"down_blocks.{i}.motion_modules.{j}.", boc[i])
    mm("mid_block.motion_modules.0.", boc[-1])
    rev = list(reversed(boc))
    for i in range(nlev):
        for j in range(lpb + 1):
            mm(f"up_blocks.{i}.motion_modules.{j}.", rev[i])
    return s


def seeded_state_dict(cfg, seed: int = 33, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic weights as in ``unet_oracle.seeded_state_dict``.  The motion modules' ``proj_out`` is drawn like
    every other Linear (non-zero): the constructor zeroes it (motion_module.py:80-81) but ``mm_sd_v15_v2.ckpt``
    overwrites it (utils/util.py:105-121), so the modules are live in production and must be live in the tests."""
    out = {}
    for key, shape in unet_param_shapes(cfg).items():
        g = torch.Generator().manual_seed(uo._seed_of(key, seed))
        if key.endswith("weight") and len(shape) == 1:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("bias"):
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) * fan_in ** -0.5
        out[key] = t.to(dtype)
    return out
