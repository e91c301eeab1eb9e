"""ORACLE (test infrastructure, not product code): plain-PyTorch restatement of diffusers' ``AutoencoderKLTemporalDecoder``
(the SVD VAE the reference decodes with: run_video_style_transfer_sd.py:33-36; called at stable_diffusion.py:385, :810,
:830 and ddim_inversion.py:29) -- KL encoder, temporal decoder, posterior sample.

PARITY UNPINNED: the network is third-party (diffusers 0.35.1, ``models/autoencoders/autoencoder_kl_temporal_decoder.py``,
``models/autoencoders/vae.py``, ``models/unets/unet_3d_blocks.py``, ``models/resnet.py``) and absent from /root/reference,
diffusers is not installed here and no weights exist offline, so this file restates the published architecture from its
state-dict layout and cannot be checked against the library or against golden vectors of the reference.  What it pins is
the PRODUCT (univst_b200/vae.py) against an independent fp32 evaluation of the same definition.  Only tests/ may import it.

Functional over a state dict with diffusers' key names (``encoder.*``, ``decoder.*``, ``quant_conv.*``)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                  scaling_factor=0.18215)
TINY_VAE_CONFIG = dict(VAE_CONFIG, block_out_channels=(64, 64, 128, 128), layers_per_block=1)


def param_shapes(cfg) -> Dict[str, tuple]:
    boc, lpb, lc = cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"]
    s: Dict[str, tuple] = {}

    def resnet(pre, cin, cout):
        s[pre + "norm1.weight"], s[pre + "norm1.bias"] = (cin,), (cin,)
        s[pre + "conv1.weight"], s[pre + "conv1.bias"] = (cout, cin, 3, 3), (cout,)
        s[pre + "norm2.weight"], s[pre + "norm2.bias"] = (cout,), (cout,)
        s[pre + "conv2.weight"], s[pre + "conv2.bias"] = (cout, cout, 3, 3), (cout,)
        if cin != cout:
            s[pre + "conv_shortcut.weight"], s[pre + "conv_shortcut.bias"] = (cout, cin, 1, 1), (cout,)

    def attn(pre, c):
        s[pre + "group_norm.weight"], s[pre + "group_norm.bias"] = (c,), (c,)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            s[pre + n + ".weight"], s[pre + n + ".bias"] = (c, c), (c,)

    def st_block(pre, cin, cout):
        resnet(pre + "spatial_res_block.", cin, cout)
        t = pre + "temporal_res_block."
        for n in ("norm1", "norm2"):
            s[t + n + ".weight"], s[t + n + ".bias"] = (cout,), (cout,)
        for n in ("conv1", "conv2"):
            s[t + n + ".weight"], s[t + n + ".bias"] = (cout, cout, 3, 1, 1), (cout,)
        s[pre + "time_mixer.mix_factor"] = (1,)

    # encoder (diffusers Encoder: DownEncoderBlock2D x len(boc), UNetMidBlock2D with one single-head attention)
    s["encoder.conv_in.weight"], s["encoder.conv_in.bias"] = (boc[0], cfg["in_channels"], 3, 3), (boc[0],)
    cin = boc[0]
    for i, c in enumerate(boc):
        for j in range(lpb):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}.", cin, c)
            cin = c
        if i < len(boc) - 1:
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (c, c, 3, 3)
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (c,)
    resnet("encoder.mid_block.resnets.0.", boc[-1], boc[-1])
    attn("encoder.mid_block.attentions.0.", boc[-1])
    resnet("encoder.mid_block.resnets.1.", boc[-1], boc[-1])
    s["encoder.conv_norm_out.weight"], s["encoder.conv_norm_out.bias"] = (boc[-1],), (boc[-1],)
    s["encoder.conv_out.weight"], s["encoder.conv_out.bias"] = (2 * lc, boc[-1], 3, 3), (2 * lc,)
    s["quant_conv.weight"], s["quant_conv.bias"] = (2 * lc, 2 * lc, 1, 1), (2 * lc,)
    # temporal decoder (diffusers TemporalDecoder: MidBlockTemporalDecoder, UpBlockTemporalDecoder x len(boc))
    s["decoder.conv_in.weight"], s["decoder.conv_in.bias"] = (boc[-1], lc, 3, 3), (boc[-1],)
    for j in range(lpb):
        st_block(f"decoder.mid_block.resnets.{j}.", boc[-1], boc[-1])
    attn("decoder.mid_block.attentions.0.", boc[-1])
    rev = list(reversed(boc))
    cin = rev[0]
    for i, c in enumerate(rev):
        for j in range(lpb + 1):
            st_block(f"decoder.up_blocks.{i}.resnets.{j}.", cin, c)
            cin = c
        if i < len(boc) - 1:
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (c, c, 3, 3)
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (c,)
    s["decoder.conv_norm_out.weight"], s["decoder.conv_norm_out.bias"] = (boc[0],), (boc[0],)
    s["decoder.conv_out.weight"], s["decoder.conv_out.bias"] = (cfg["out_channels"], boc[0], 3, 3), (cfg["out_channels"],)
    s["decoder.time_conv_out.weight"] = (cfg["out_channels"], cfg["out_channels"], 3, 1, 1)
    s["decoder.time_conv_out.bias"] = (cfg["out_channels"],)
    return s


def seeded_state_dict(cfg, seed: int = 55) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in param_shapes(cfg).items():
        if k.endswith("mix_factor"):
            sd[k] = torch.randn(shp, generator=g)                    # learned blend (sigmoid -> 0.2 .. 0.8)
        elif k.endswith("weight") and len(shp) == 1:
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.02 * torch.randn(shp, generator=g)
        else:
            fan = 1
            for d in shp[1:]:
                fan *= d
            sd[k] = torch.randn(shp, generator=g) * fan ** -0.5
    return sd


def _resnet(sd, pre, x, eps=1e-6):
    h = F.silu(F.group_norm(x, 32, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps))
    h = F.conv2d(h, sd[pre + "conv1.weight"], sd[pre + "conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, 32, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps))
    h = F.conv2d(h, sd[pre + "conv2.weight"], sd[pre + "conv2.bias"], padding=1)
    if pre + "conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[pre + "conv_shortcut.weight"], sd[pre + "conv_shortcut.bias"])
    return x + h


def _attn(sd, pre, x):
    """diffusers Attention(heads=1, dim_head=C, norm_num_groups=32, eps=1e-6, bias=True, residual_connection=True)."""
    n, c, hh, ww = x.shape
    y = F.group_norm(x, 32, sd[pre + "group_norm.weight"], sd[pre + "group_norm.bias"], 1e-6).view(n, c, hh * ww).transpose(1, 2)
    q, k, v = (F.linear(y, sd[pre + f"to_{t}.weight"], sd[pre + f"to_{t}.bias"]) for t in "qkv")
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])
    return x + o.transpose(1, 2).reshape(n, c, hh, ww)


def _st_block(sd, pre, x, num_frames):
    """SpatioTemporalResBlock(eps=1e-6, temporal_eps=1e-5, merge_strategy="learned", switch_spatial_to_temporal_mix=True)."""
    x = _resnet(sd, pre + "spatial_res_block.", x)
    bf, c, hh, ww = x.shape
    xs = x.view(bf // num_frames, num_frames, c, hh, ww).permute(0, 2, 1, 3, 4)      # (b, c, t, h, w)
    t = pre + "temporal_res_block."
    h = F.silu(F.group_norm(xs, 32, sd[t + "norm1.weight"], sd[t + "norm1.bias"], 1e-5))
    h = F.conv3d(h, sd[t + "conv1.weight"], sd[t + "conv1.bias"], padding=(1, 0, 0))
    h = F.silu(F.group_norm(h, 32, sd[t + "norm2.weight"], sd[t + "norm2.bias"], 1e-5))
    h = F.conv3d(h, sd[t + "conv2.weight"], sd[t + "conv2.bias"], padding=(1, 0, 0))
    xt = xs + h
    alpha = 1.0 - torch.sigmoid(sd[pre + "time_mixer.mix_factor"])                   # switch_spatial_to_temporal_mix
    out = alpha * xs + (1.0 - alpha) * xt
    return out.permute(0, 2, 1, 3, 4).reshape(bf, c, hh, ww)


def encode_moments(sd, cfg, x):
    """x: (N, 3, H, W) in [-1, 1] -> (N, 2 C_lat, H/8, W/8) = [mean | logvar]."""
    boc, lpb = cfg["block_out_channels"], cfg["layers_per_block"]
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(len(boc)):
        for j in range(lpb):
            h = _resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}.", h)
        if i < len(boc) - 1:
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv."
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[p + "weight"], sd[p + "bias"], stride=2)
    h = _resnet(sd, "encoder.mid_block.resnets.0.", h)
    h = _attn(sd, "encoder.mid_block.attentions.0.", h)
    h = _resnet(sd, "encoder.mid_block.resnets.1.", h)
    h = F.silu(F.group_norm(h, 32, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"], 1e-6))
    h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def sample_latents(moments, noise, scaling):
    """DiagonalGaussianDistribution.sample() x scaling, "(b f) c h w -> b c f h w" (ddim_inversion.py:29-31)."""
    mean, logvar = moments.chunk(2, dim=1)
    z = mean if noise is None else mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise
    return (scaling * z).permute(1, 0, 2, 3).unsqueeze(0)


def decode(sd, cfg, z, num_frames):
    """z: (N, C_lat, h, w) (already divided by the scaling factor) -> (N, 3, 8h, 8w)."""
    boc, lpb = cfg["block_out_channels"], cfg["layers_per_block"]
    h = F.conv2d(z, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = _st_block(sd, "decoder.mid_block.resnets.0.", h, num_frames)
    for j in range(1, lpb):
        h = _attn(sd, "decoder.mid_block.attentions.0.", h)
        h = _st_block(sd, f"decoder.mid_block.resnets.{j}.", h, num_frames)
    for i in range(len(boc)):
        for j in range(lpb + 1):
            h = _st_block(sd, f"decoder.up_blocks.{i}.resnets.{j}.", h, num_frames)
        if i < len(boc) - 1:
            p = f"decoder.up_blocks.{i}.upsamplers.0.conv."
            h = F.conv2d(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd[p + "weight"], sd[p + "bias"], padding=1)
    h = F.silu(F.group_norm(h, 32, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], 1e-6))
    h = F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
    bf, c, hh, ww = h.shape
    h = h.view(bf // num_frames, num_frames, c, hh, ww).permute(0, 2, 1, 3, 4)
    h = F.conv3d(h, sd["decoder.time_conv_out.weight"], sd["decoder.time_conv_out.bias"], padding=(1, 0, 0))
    return h.permute(0, 2, 1, 3, 4).reshape(bf, c, hh, ww)


def frames_to_u8(frames):
    """stable_diffusion.py:812-814: (x / 2 + 0.5).clamp(0, 1) -> round(255 x) uint8, (N, 3, H, W) -> (N, H, W, 3)."""
    f = (frames / 2 + 0.5).clamp(0, 1).float()
    return (f * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
