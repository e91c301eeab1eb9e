"""ORACLE (test infrastructure, not product code): CPU / plain-PyTorch restatement of the reference's three-branch
UNet forward for the SD-1.5 / SD-2.1 "pseudo-3D" backbone, including the AdaIN-guided attention patch.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm may import this.
Parity status: PINNED against the reference's own module code -- ``oracle/gen_golden.py`` runs
``/root/reference/backbones/video_diffusion_sd/models/unet_3d_condition.py`` (on the test-only diffusers shim in
``oracle/_shim``) and ``pnp_utils.py`` on seeded weights and commits input/output vectors under ``tests/golden/``;
``tests/test_oracle_cpu.py`` checks this file against them.  The reference itself ships no tests or golden vectors.

All functions are functional over a ``state_dict`` with the reference's key names (= SD ``UNet2DConditionModel``
names plus the never-loaded ``*_temporal*`` keys).  File:line citations are relative to
``/root/reference/backbones/video_diffusion_sd``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD15_CONFIG = dict(block_out_channels=(320, 640, 1280, 1280), attention_head_dim=8, cross_attention_dim=768,
                   layers_per_block=2, norm_num_groups=32, norm_eps=1e-5, in_channels=4, out_channels=4,
                   use_linear_projection=False)
SD21_CONFIG = dict(SD15_CONFIG, attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024, use_linear_projection=True)
TINY_CONFIG = dict(SD15_CONFIG, block_out_channels=(64, 128, 128, 128), attention_head_dim=4, cross_attention_dim=64)
# SD-2.1 layout in miniature: head dim 64 on every level (heads 1/2/2/2), Linear proj_in / proj_out
TINY_SD21_CONFIG = dict(TINY_CONFIG, attention_head_dim=(1, 2, 2, 2), use_linear_projection=True, cross_attention_dim=128)

# layers whose attn1.forward is replaced by register_spatial_attention_pnp (pnp_utils.py:104-111)
PATCHED = {(1, 1), (1, 2), (2, 0), (2, 1), (2, 2), (3, 0), (3, 1), (3, 2)}


def heads_of(cfg, level: int) -> int:
    h = cfg["attention_head_dim"]  # passed as *number of heads* (models/unet_3d_blocks.py:269-271)
    return h if isinstance(h, int) else h[level]


# ----------------------------------------------------------------------------------------------- pnp_utils.py
def attention_adain(cnt, sty):
    """pnp_utils.py:114-125.  cnt, sty: (frames, tokens, channels)."""
    sty_mean = sty.mean(dim=[1], keepdim=True)
    sty_std = sty.std(dim=[1], keepdim=True)
    return (F.instance_norm(cnt) * sty_std + sty_mean).to(cnt.dtype)


def latent_adain(cnt, sty):
    """pnp_utils.py:128-139.  cnt, sty: (1, C, F, h, w)."""
    sty_mean = sty.mean(dim=[0, 3, 4], keepdim=True)
    sty_std = sty.std(dim=[0, 3, 4], keepdim=True)
    return (F.instance_norm(cnt) * sty_std + sty_mean).to(cnt.dtype)


def shift_params(idx: int, eta1: float = 0.0, eta2: float = 0.5):
    """pnp_utils.py:47-51: active window and the beta schedule (0.9 at eta1*50 -> 0.1 at eta2*50)."""
    active = idx >= eta1 and idx <= eta2 * 50
    beta = (0.9 - 0.1) / (eta1 * 50 - eta2 * 50) * (idx - eta2 * 50) + 0.1
    return active, 0.65, beta, 3.0


def frame_sources(F_: int, index_list):
    """Frame-index gather of SparseCausalAttention (models/attention.py:386-406; pnp_utils.py:63-79)."""
    out = []
    for index in index_list:
        if index == "first":
            out.append(torch.zeros(F_, dtype=torch.long))
        elif index == "last":
            out.append(torch.full((F_,), F_ - 1, dtype=torch.long))
        elif index in ("mid", "middle"):
            out.append(torch.full((F_,), (F_ - 1) // 2, dtype=torch.long))
        else:
            out.append((torch.arange(F_) + index).clip(0, F_ - 1))
    return out


def sc_attention(sd, pre, x, heads, F_, patched: bool, idx: Optional[int], eta1=0.0, eta2=0.5):
    """attn1: stock SparseCausalAttention.forward (models/attention.py:350-430, KV = [prev, self, first]) or the
    patched forward of pnp_utils.py:20-100 (AdaIN shift while active, KV = [prev, first]).  x: (B*F, N, C)."""
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
            k[2 * chunk:] = beta * attention_adain(k[2 * chunk:], k[chunk:2 * chunk]) + (1 - beta) * k[chunk:2 * chunk]
            v[2 * chunk:] = beta * attention_adain(v[2 * chunk:], v[chunk:2 * chunk]) + (1 - beta) * v[chunk:2 * chunk]
            q[2 * chunk:] = gamma * q[2 * chunk:]
        index_list = [-1, "first"]
    else:
        index_list = [-1, 0, "first"]
    B = BF // F_
    k5, v5 = k.view(B, F_, N, C), v.view(B, F_, N, C)
    srcs = frame_sources(F_, index_list)
    k = torch.cat([k5[:, s] for s in srcs], dim=2).reshape(BF, -1, C)
    v = torch.cat([v5[:, s] for s in srcs], dim=2).reshape(BF, -1, C)
    d = C // heads
    q = q.view(BF, -1, heads, d).transpose(1, 2)
    k = k.view(BF, -1, heads, d).transpose(1, 2)
    v = v.view(BF, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(BF, -1, C)
    return F.linear(o, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])


def cross_attention(sd, pre, x, ctx, heads):
    """attn2: diffusers Attention + AttnProcessor2_0 (called at models/attention.py:316-323).  ctx: (B*F, L, D)."""
    q = F.linear(x, sd[pre + "to_q.weight"])
    k = F.linear(ctx, sd[pre + "to_k.weight"])
    v = F.linear(ctx, sd[pre + "to_v.weight"])
    BF, N, C = q.shape
    d = C // heads
    q = q.view(BF, -1, heads, d).transpose(1, 2)
    k = k.view(BF, -1, heads, d).transpose(1, 2)
    v = v.view(BF, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(BF, -1, C)
    return F.linear(o, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])


def feed_forward(sd, pre, x):
    """diffusers FeedForward(geglu): Linear(C, 8C) -> h * gelu(gate) -> Linear(4C, C) (models/attention.py:329)."""
    h, gate = F.linear(x, sd[pre + "net.0.proj.weight"], sd[pre + "net.0.proj.bias"]).chunk(2, dim=-1)
    return F.linear(h * F.gelu(gate), sd[pre + "net.2.weight"], sd[pre + "net.2.bias"])


def temporal_attention(sd, pre, x, heads, F_):
    """apply_temporal_attention (models/attention.py:336-346): full attention over the F frames of every token.
    With the reference's zero ``to_out.0.weight`` this is ``x + to_out.0.bias`` -- evaluated in full here anyway."""
    BF, N, C = x.shape
    B = BF // F_
    h = x.view(B, F_, N, C).permute(0, 2, 1, 3).reshape(B * N, F_, C)
    n = F.layer_norm(h, (C,), sd[pre + "norm_temporal.weight"], sd[pre + "norm_temporal.bias"])
    a = pre + "attn_temporal."
    q, k, v = F.linear(n, sd[a + "to_q.weight"]), F.linear(n, sd[a + "to_k.weight"]), F.linear(n, sd[a + "to_v.weight"])
    d = C // heads
    q = q.view(B * N, F_, heads, d).transpose(1, 2)
    k = k.view(B * N, F_, heads, d).transpose(1, 2)
    v = v.view(B * N, F_, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, F_, C)
    h = F.linear(o, sd[a + "to_out.0.weight"], sd[a + "to_out.0.bias"]) + h
    return h.view(B, N, F_, C).permute(0, 2, 1, 3).reshape(BF, N, C)


def transformer(sd, pre, x, ctx, heads, F_, patched, idx, cfg, eta1=0.0, eta2=0.5):
    """SpatioTemporalTransformerModel.forward (models/attention.py:104-153) with its single
    SpatioTemporalTransformerBlock (:280-334).  x: (B*F, C, h, w); ctx: (B, L, D)."""
    BF, C, h, w = x.shape
    res = x
    ctx_rep = ctx.repeat_interleave(F_, 0)
    y = F.group_norm(x, cfg["norm_num_groups"], sd[pre + "norm.weight"], sd[pre + "norm.bias"], eps=1e-6)
    if not cfg["use_linear_projection"]:
        y = F.conv2d(y, sd[pre + "proj_in.weight"], sd[pre + "proj_in.bias"])
        y = y.permute(0, 2, 3, 1).reshape(BF, h * w, C)
    else:
        y = y.permute(0, 2, 3, 1).reshape(BF, h * w, C)
        y = F.linear(y, sd[pre + "proj_in.weight"], sd[pre + "proj_in.bias"])
    b = pre + "transformer_blocks.0."
    ln = lambda t, name: F.layer_norm(t, (C,), sd[b + name + ".weight"], sd[b + name + ".bias"])
    y = y + sc_attention(sd, b + "attn1.", ln(y, "norm1"), heads, F_, patched, idx, eta1, eta2)
    y = cross_attention(sd, b + "attn2.", ln(y, "norm2"), ctx_rep, heads) + y
    y = feed_forward(sd, b + "ff.", ln(y, "norm3")) + y
    y = temporal_attention(sd, b, y, heads, F_)
    if not cfg["use_linear_projection"]:
        y = y.reshape(BF, h, w, C).permute(0, 3, 1, 2)
        y = F.conv2d(y, sd[pre + "proj_out.weight"], sd[pre + "proj_out.bias"])
    else:
        y = F.linear(y, sd[pre + "proj_out.weight"], sd[pre + "proj_out.bias"])
        y = y.reshape(BF, h, w, C).permute(0, 3, 1, 2)
    return y + res


def group_norm_5d(x, F_, groups, w, b, eps):
    """GroupNorm applied by the reference to a (B, C, F, h, w) tensor: statistics span the frames (resnet.py:338)."""
    BF, C, h, wd = x.shape
    B = BF // F_
    x5 = x.view(B, F_, C, h, wd).permute(0, 2, 1, 3, 4)
    y5 = F.group_norm(x5, groups, w, b, eps)
    return y5.permute(0, 2, 1, 3, 4).reshape(BF, C, h, wd)


def conv_pseudo3d(sd, pre, x, stride=1, padding=1):
    """PseudoConv3d.forward (resnet.py:57-80): 2-D conv per frame; the temporal Conv1d is Dirac / zero-bias
    (resnet.py:54-55) and never loaded (unet_3d_condition.py:503), i.e. the identity."""
    return F.conv2d(x, sd[pre + "weight"], sd[pre + "bias"], stride=stride, padding=padding)


def resnet(sd, pre, x, emb, F_, cfg):
    """ResnetBlockPseudo3D.forward (resnet.py:335-394), time_embedding_norm='default', output_scale_factor 1."""
    g, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    h = F.silu(group_norm_5d(x, F_, g, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps))
    h = conv_pseudo3d(sd, pre + "conv1.", h)
    temb = F.linear(F.silu(emb), sd[pre + "time_emb_proj.weight"], sd[pre + "time_emb_proj.bias"])
    h = h + temb.repeat_interleave(F_, 0)[:, :, None, None]
    h = F.silu(group_norm_5d(h, F_, g, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps))
    h = conv_pseudo3d(sd, pre + "conv2.", h)
    if pre + "conv_shortcut.weight" in sd:
        x = conv_pseudo3d(sd, pre + "conv_shortcut.", x, padding=0)
    return x + h


def timestep_embedding(t, dim):
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    e = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(e), torch.sin(e)], dim=-1)


def unet_forward(sd: Dict[str, torch.Tensor], cfg, sample, timestep, ctx, *, patched: bool = False,
                 idx: Optional[int] = None, eta1: float = 0.0, eta2: float = 0.5, features: Optional[dict] = None):
    """UNetPseudo3DConditionModel.forward (models/unet_3d_condition.py:306-443).
    sample (B, 4, F, h, w); timestep: int; ctx (B, L, D).  ``features[i]`` receives the output of up block i as
    (B, C, F, h, w) (the reference dumps ``sample[0].permute(1, 2, 3, 0)`` of it, :430-436)."""
    B, Cin, F_, h, w = sample.shape
    dt = sample.dtype
    boc = cfg["block_out_channels"]
    x = sample.permute(0, 2, 1, 3, 4).reshape(B * F_, Cin, h, w)
    t = torch.full((B,), float(timestep), device=sample.device)
    temb = timestep_embedding(t, boc[0]).to(dt)
    emb = F.linear(F.silu(F.linear(temb, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])),
                   sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])
    x = conv_pseudo3d(sd, "conv_in.", x)
    skips = [x]
    nlev = len(boc)
    for i in range(nlev):
        for j in range(cfg["layers_per_block"]):
            x = resnet(sd, f"down_blocks.{i}.resnets.{j}.", x, emb, F_, cfg)
            if i < nlev - 1:
                x = transformer(sd, f"down_blocks.{i}.attentions.{j}.", x, ctx, heads_of(cfg, i), F_, False, idx, cfg)
            skips.append(x)
        if i < nlev - 1:
            x = conv_pseudo3d(sd, f"down_blocks.{i}.downsamplers.0.conv.", x, stride=2)
            skips.append(x)
    x = resnet(sd, "mid_block.resnets.0.", x, emb, F_, cfg)
    x = transformer(sd, "mid_block.attentions.0.", x, ctx, heads_of(cfg, nlev - 1), F_, False, idx, cfg)
    x = resnet(sd, "mid_block.resnets.1.", x, emb, F_, cfg)
    for i in range(nlev):
        for j in range(cfg["layers_per_block"] + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet(sd, f"up_blocks.{i}.resnets.{j}.", x, emb, F_, cfg)
            if i > 0:
                x = transformer(sd, f"up_blocks.{i}.attentions.{j}.", x, ctx, heads_of(cfg, nlev - 1 - i), F_,
                                patched and (i, j) in PATCHED, idx, cfg, eta1, eta2)
        if i < nlev - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = conv_pseudo3d(sd, f"up_blocks.{i}.upsamplers.0.conv.", x)
        if features is not None:
            features[i] = x.view(B, F_, -1, x.shape[-2], x.shape[-1]).permute(0, 2, 1, 3, 4)
    x = F.silu(group_norm_5d(x, F_, cfg["norm_num_groups"], sd["conv_norm_out.weight"], sd["conv_norm_out.bias"],
                             cfg["norm_eps"]))
    x = conv_pseudo3d(sd, "conv_out.", x)
    return x.view(B, F_, -1, h, w).permute(0, 2, 1, 3, 4)


# ----------------------------------------------------------------------------------------------- seeded weights
def _seed_of(key: str, seed: int) -> int:
    v = seed
    for ch in key.encode():
        v = (v * 1000003 + ch) % (2 ** 31 - 1)
    return v


def unet_param_shapes(cfg) -> Dict[str, tuple]:
    """Key -> shape of the reference module tree (verified against the reference constructor in gen_golden.py)."""
    boc, lpb = cfg["block_out_channels"], cfg["layers_per_block"]
    D, temb_dim, nlev = cfg["cross_attention_dim"], boc[0] * 4, len(boc)
    s: Dict[str, tuple] = {}

    def conv(pre, cin, cout, k):
        s[pre + "weight"], s[pre + "bias"] = (cout, cin, k, k), (cout,)
        if k > 1:
            s[pre + "conv_temporal.weight"], s[pre + "conv_temporal.bias"] = (cout, cout, k), (cout,)

    def lin(pre, cin, cout, bias=True):
        s[pre + "weight"] = (cout, cin)
        if bias:
            s[pre + "bias"] = (cout,)

    def norm(pre, c):
        s[pre + "weight"], s[pre + "bias"] = (c,), (c,)

    def res(pre, cin, cout):
        norm(pre + "norm1.", cin), conv(pre + "conv1.", cin, cout, 3), lin(pre + "time_emb_proj.", temb_dim, cout)
        norm(pre + "norm2.", cout), conv(pre + "conv2.", cout, cout, 3)
        if cin != cout:
            conv(pre + "conv_shortcut.", cin, cout, 1)

    def attn(pre, c, kv):
        lin(pre + "to_q.", c, c, False), lin(pre + "to_k.", kv, c, False), lin(pre + "to_v.", kv, c, False)
        lin(pre + "to_out.0.", c, c)

    def tr(pre, c):
        norm(pre + "norm.", c)
        if cfg["use_linear_projection"]:
            lin(pre + "proj_in.", c, c), lin(pre + "proj_out.", c, c)
        else:
            conv(pre + "proj_in.", c, c, 1), conv(pre + "proj_out.", c, c, 1)
        b = pre + "transformer_blocks.0."
        attn(b + "attn1.", c, c), norm(b + "norm1.", c), attn(b + "attn2.", c, D), norm(b + "norm2.", c)
        attn(b + "attn_temporal.", c, c), norm(b + "norm_temporal.", c)
        lin(b + "ff.net.0.proj.", c, 8 * c), lin(b + "ff.net.2.", 4 * c, c), norm(b + "norm3.", c)

    conv("conv_in.", cfg["in_channels"], boc[0], 3)
    lin("time_embedding.linear_1.", boc[0], temb_dim), lin("time_embedding.linear_2.", temb_dim, temb_dim)
    skip_ch = [boc[0]]
    cout = boc[0]
    for i in range(nlev):
        cin, cout = cout, boc[i]
        for j in range(lpb):
            res(f"down_blocks.{i}.resnets.{j}.", cin if j == 0 else cout, cout)
            if i < nlev - 1:
                tr(f"down_blocks.{i}.attentions.{j}.", cout)
            skip_ch.append(cout)
        if i < nlev - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv.", cout, cout, 3)
            skip_ch.append(cout)
    res("mid_block.resnets.0.", boc[-1], boc[-1]), tr("mid_block.attentions.0.", boc[-1])
    res("mid_block.resnets.1.", boc[-1], boc[-1])
    rev = list(reversed(boc))
    x_ch = boc[-1]
    for i in range(nlev):
        cout = rev[i]
        for j in range(lpb + 1):
            res(f"up_blocks.{i}.resnets.{j}.", x_ch + skip_ch.pop(), cout)
            x_ch = cout
            if i > 0:
                tr(f"up_blocks.{i}.attentions.{j}.", cout)
        if i < nlev - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv.", cout, cout, 3)
    norm("conv_norm_out.", boc[0]), conv("conv_out.", boc[0], cfg["out_channels"], 3)
    return s


def seeded_state_dict(cfg, seed: int = 33, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic, construction-order-independent weights (no checkpoint exists in this environment): every tensor
    is drawn from its own generator seeded by (seed, key).  Weights ~ N(0, 1/fan_in), biases ~ N(0, 0.05^2), norm
    scales 1 + N(0, 0.1^2); the never-loaded temporal parts keep the reference's constructor values:
    ``conv_temporal`` Dirac / zero (resnet.py:54-55), ``attn_temporal.to_out.0.weight`` zero (attention.py:233)."""
    out = {}
    for key, shape in unet_param_shapes(cfg).items():
        g = torch.Generator().manual_seed(_seed_of(key, seed))
        if "conv_temporal.weight" in key:
            t = torch.zeros(shape)
            torch.nn.init.dirac_(t)
        elif "conv_temporal.bias" in key:
            t = torch.zeros(shape)
        elif "attn_temporal.to_out.0.weight" in key:
            t = torch.zeros(shape)
        elif key.endswith("weight") and len(shape) == 1:
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
