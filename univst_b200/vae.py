"""B200 host-side mirror of the VAE around the loop: diffusers' ``AutoencoderKLTemporalDecoder`` -- the
SVD VAE the reference builds in run_video_style_transfer_sd.py:33-36 and calls at stable_diffusion.py:385 (``decode_latents``),
:810 / :830 (the smoother's ``get_images_from_latents`` / ``get_latent_image``) and inversion_tools/ddim_inversion.py:29
(``vae.encode(video).latent_dist.sample()``).

PARITY UNPINNED: the network is third-party (diffusers 0.35.1) and absent from the reference checkout, diffusers is not in
this image and its weights cannot be downloaded, so the architecture is restated from the published state-dict layout
(key names and shapes below are diffusers'; the real checkpoint has the same 97.7 M parameters) and checked against an
independent fp32 evaluation of the same definition (oracle/vae_oracle.py), not against the library.

Everything runs on the kernels behind ``include/univst_b200.h`` in the channels-last ``[(frame) h w, C]`` layout of the UNet:
3x3 convolutions (any frame size), GroupNorm(+SiLU) with per-frame or per-clip statistics, the (3, 1, 1) temporal
convolutions as a three-tap implicit GEMM over the frame axis (zero frames beyond the clip come from TMA out-of-bounds
fill), the stride-2 encoder convolutions with diffusers' pad-after convention, the learned spatial / temporal blend as one
axpby.  The single-head mid-block attention (head dim = 512) exceeds the fused attention kernel's tensor-memory budget and
runs as QK^T GEMM -> row softmax -> PV GEMM per frame (one layer at 1/64 of the pixels: 0.2 % of the decoder's FLOPs).
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops
from .pack import pack_conv3x3

VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                  scaling_factor=0.18215)
CPAD = 64   # 3- / 4- / 8-channel tensors are zero-padded to one 128-byte swizzle row of channels where a convolution reads them


class _Posterior:
    """What the callers use of diffusers' DiagonalGaussianDistribution: ``sample(generator)`` and ``mode()`` -> (N, C, h, w)."""

    def __init__(self, vae, moments, N, h, w):
        self._vae, self._m, self._shape = vae, moments, (N, h, w)

    def _latents(self, noise):
        N, h, w = self._shape
        C = self._vae.config.latent_channels
        z = ops.vae_sample(self._m, noise, C, N, h * w, 1.0)                      # (1, C, N, hw)
        return z.view(C, N, h, w).permute(1, 0, 2, 3).contiguous()

    def sample(self, generator=None):
        N, h, w = self._shape
        C = self._vae.config.latent_channels
        noise = torch.randn((N, C, h, w), generator=generator, device=self._m.device, dtype=torch.float16)   # randn_tensor(mean.shape)
        return self._latents(noise.contiguous())

    def mode(self):
        return self._latents(None)


class AutoencoderKLTemporalDecoder:
    def __init__(self, state_dict: Dict[str, torch.Tensor], config: Optional[dict] = None, device="cuda"):
        cfg = dict(VAE_CONFIG)
        cfg.update(config or {})
        self.config = SimpleNamespace(**cfg)
        self.device = torch.device(device)
        self.dtype = torch.float16
        self._pack(state_dict)

    # ------------------------------------------------------------------------------------------ weights
    def _pack(self, sd):
        dev = self.device
        h = lambda t: t.detach().to(device=dev, dtype=torch.float16).contiguous()
        W: Dict[str, torch.Tensor] = {}
        self._alpha = {}
        for key, t in sd.items():
            if key.endswith("time_mixer.mix_factor"):
                # AlphaBlender("learned", switch_spatial_to_temporal_mix=True): x = a x_spatial + (1 - a) x_temporal with
                # a = 1 - sigmoid(mix_factor); the weights are rounded to fp16 as the reference's fp16 module would hold them
                a = 1.0 - torch.sigmoid(t.detach().float().reshape(-1)[0].half().float()).half().float().item()
                self._alpha[key[: -len("time_mixer.mix_factor")]] = a
            elif t.dim() == 5:   # Conv3d (3, 1, 1): [Cout, Cin, 3, 1, 1] -> [Cout, 3, Cin] tap-major, channels padded where tiny
                cout, cin = t.shape[:2]
                w = t.detach().float().reshape(cout, cin, 3).permute(0, 2, 1)
                if cin < 8:
                    w = torch.nn.functional.pad(w, (0, CPAD - cin))
                W[key] = h(w.reshape(cout, -1))
            elif t.dim() == 4 and t.shape[-1] == 3:
                W[key] = h(pack_conv3x3(t, CPAD if t.shape[1] < 8 else 0))
            elif t.dim() == 4:   # 1x1 convolutions
                w = t.detach().float().reshape(t.shape[0], t.shape[1])
                if t.shape[1] < 64:   # quant_conv reads the 8 moment channels out of a CPAD-wide buffer
                    w = torch.nn.functional.pad(w, (0, CPAD - t.shape[1]))
                W[key] = h(w)
            else:
                W[key] = h(t)
        self.W = W

    # ------------------------------------------------------------------------------------------ building blocks
    def _resnet(self, pre, x, NI, H, Wd):
        """diffusers ResnetBlock2D(temb_channels=None, groups=32, eps=1e-6): per-frame statistics."""
        W = self.W
        hd = ops.groupnorm(x, W[pre + "norm1.weight"], W[pre + "norm1.bias"], NB=NI, rows=H * Wd, eps=1e-6, silu=True)
        hd = ops.conv3x3(hd.view(NI, H, Wd, -1), W[pre + "conv1.weight"], bias=W[pre + "conv1.bias"])
        hd = ops.groupnorm(hd, W[pre + "norm2.weight"], W[pre + "norm2.bias"], NB=NI, rows=H * Wd, eps=1e-6, silu=True)
        sc = x
        if pre + "conv_shortcut.weight" in W:
            sc = ops.gemm(x, W[pre + "conv_shortcut.weight"], bias=W[pre + "conv_shortcut.bias"])
        return ops.conv3x3(hd.view(NI, H, Wd, -1), W[pre + "conv2.weight"], bias=W[pre + "conv2.bias"], residual=sc)

    def _st_block(self, pre, x, NB, F, H, Wd):
        """SpatioTemporalResBlock: spatial ResnetBlock2D -> TemporalResnetBlock (GroupNorm over all frames of a clip,
        eps 1e-5, two (3, 1, 1) convolutions, residual) -> learned blend of the two."""
        W = self.W
        xs = self._resnet(pre + "spatial_res_block.", x, NB * F, H, Wd)
        t = pre + "temporal_res_block."
        rows = F * H * Wd
        hd = ops.groupnorm(xs, W[t + "norm1.weight"], W[t + "norm1.bias"], NB=NB, rows=rows, eps=1e-5, silu=True)
        hd = ops.conv_temporal3(hd, W[t + "conv1.weight"], NB=NB, F=F, HW=H * Wd, bias=W[t + "conv1.bias"])
        hd = ops.groupnorm(hd, W[t + "norm2.weight"], W[t + "norm2.bias"], NB=NB, rows=rows, eps=1e-5, silu=True)
        xt = ops.conv_temporal3(hd, W[t + "conv2.weight"], NB=NB, F=F, HW=H * Wd, bias=W[t + "conv2.bias"], residual=xs)
        a = self._alpha[pre]
        return ops.axpby(xs, xt, a, 1.0 - a)

    def _attn(self, pre, x, NI, N):
        """diffusers Attention(heads=1, dim_head=C, norm_num_groups=32, eps=1e-6, bias=True, residual_connection=True)."""
        W = self.W
        C = x.shape[1]
        xn = ops.groupnorm(x, W[pre + "group_norm.weight"], W[pre + "group_norm.bias"], NB=NI, rows=N, eps=1e-6, silu=False)
        q = ops.gemm(xn, W[pre + "to_q.weight"], bias=W[pre + "to_q.bias"])
        k = ops.gemm(xn, W[pre + "to_k.weight"], bias=W[pre + "to_k.bias"])
        o = torch.empty_like(x)
        for f in range(NI):
            sl = slice(f * N, (f + 1) * N)
            s = ops.gemm(q[sl], k[sl], out_scale=1.0 / math.sqrt(C))          # [N, N] scores, scaled before the fp16 rounding
            ops.softmax_rows_(s)
            vt = ops.gemm(W[pre + "to_v.weight"], xn[sl])                      # V^T = W_v xn^T  [C, N]; the bias of V passes
            ops.gemm(s, vt, bias=W[pre + "to_v.bias"], out=o[sl])              # through the softmax (rows sum to 1)
        return ops.gemm(o, W[pre + "to_out.0.weight"], bias=W[pre + "to_out.0.bias"], residual=x)

    # ------------------------------------------------------------------------------------------ encoder
    def _encode_rows(self, x, N, H, Wd):
        """x: [N H W, CPAD] fp16 rows in [-1, 1] -> moment rows [N h w, CPAD] ([mean | logvar] in the first 2 C_lat columns)."""
        W, cfg = self.W, self.config
        boc, lpb = cfg.block_out_channels, cfg.layers_per_block
        if H % (1 << (len(boc) - 1)) or Wd % (1 << (len(boc) - 1)):
            raise ValueError(f"frame height / width must be multiples of {1 << (len(boc) - 1)}")
        hd = ops.conv3x3(x.view(N, H, Wd, -1), W["encoder.conv_in.weight"], bias=W["encoder.conv_in.bias"])
        for i in range(len(boc)):
            for j in range(lpb):
                hd = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}.", hd, N, H, Wd)
            if i < len(boc) - 1:
                p = f"encoder.down_blocks.{i}.downsamplers.0.conv."
                planes = ops.space_to_depth2(hd.view(N, H, Wd, -1))
                H, Wd = H // 2, Wd // 2
                hd = ops.conv3x3_s2_pad_after(planes, W[p + "weight"], bias=W[p + "bias"])
        hd = self._resnet("encoder.mid_block.resnets.0.", hd, N, H, Wd)
        hd = self._attn("encoder.mid_block.attentions.0.", hd, N, H * Wd)
        hd = self._resnet("encoder.mid_block.resnets.1.", hd, N, H, Wd)
        hd = ops.groupnorm(hd, W["encoder.conv_norm_out.weight"], W["encoder.conv_norm_out.bias"], NB=N, rows=H * Wd, eps=1e-6,
                           silu=True)
        mom = torch.zeros((N * H * Wd, CPAD), dtype=torch.float16, device=self.device)
        ops.conv3x3(hd.view(N, H, Wd, -1), W["encoder.conv_out.weight"], bias=W["encoder.conv_out.bias"],
                    out=mom[:, : 2 * cfg.latent_channels])
        out = torch.empty((N * H * Wd, 8 * ((2 * cfg.latent_channels + 7) // 8)), dtype=torch.float16, device=self.device)
        ops.gemm(mom, W["quant_conv.weight"], bias=W["quant_conv.bias"], out=out[:, : 2 * cfg.latent_channels])
        return out, H, Wd

    @torch.no_grad()
    def encode(self, x):
        """x: (N, 3, H, W) in [-1, 1] -> object with ``.latent_dist`` (``.sample(generator)`` / ``.mode()`` -> (N, C, h, w),
        unscaled, as diffusers returns them)."""
        N, C, H, Wd = x.shape
        x = x.to(device=self.device, dtype=torch.float16).contiguous()
        # (N, C, H, W) -> (C, N, H, W): the layout the packing kernel reads (a copy, not arithmetic)
        rows = ops.pack_latents([x.permute(1, 0, 2, 3).contiguous()], Cpad=CPAD).view(N * H * Wd, CPAD)
        mom, h, w = self._encode_rows(rows, N, H, Wd)
        return SimpleNamespace(latent_dist=_Posterior(self, mom, N, h, w))

    @torch.no_grad()
    def encode_frames_u8(self, frames_u8, generator=None, sample: bool = True, noise=None):
        """uint8 frames (F, H, W, 3) -> latents (1, C, F, h, w) already multiplied by the scaling factor: the smoother's
        ``get_latent_image`` (stable_diffusion.py:821-834) and the inversion's encode (ddim_inversion.py:20-31) in one call.
        ``noise`` (F, C, h, w) fp16: the posterior noise of these frames when the caller drew it (a frame shard passes its
        slice of the clip's noise, so that sharded and unsharded encodes agree bit for bit)."""
        F_, H, Wd, _ = frames_u8.shape
        rows = ops.u8_to_frames(frames_u8.to(self.device).contiguous(), CPAD)
        mom, h, w = self._encode_rows(rows, F_, H, Wd)
        C = self.config.latent_channels
        if noise is not None:
            noise = noise.to(self.device, torch.float16).contiguous()
        elif sample:
            noise = torch.randn((F_, C, h, w), generator=generator, device=self.device, dtype=torch.float16).contiguous()
        return ops.vae_sample(mom, noise, C, F_, h * w, self.config.scaling_factor).view(1, C, F_, h, w)

    # ------------------------------------------------------------------------------------------ decoder
    def _decode_rows(self, z_rows, NB, F, h, w):
        """z_rows: [NB F h w, CPAD] (latents / scaling factor) -> pixel rows [NB F H W, 8] (channels 0..2)."""
        W, cfg = self.W, self.config
        boc, lpb = cfg.block_out_channels, cfg.layers_per_block
        NI = NB * F
        hd = ops.conv3x3(z_rows.view(NI, h, w, -1), W["decoder.conv_in.weight"], bias=W["decoder.conv_in.bias"])
        hd = self._st_block("decoder.mid_block.resnets.0.", hd, NB, F, h, w)
        for j in range(1, lpb):
            hd = self._attn("decoder.mid_block.attentions.0.", hd, NI, h * w)
            hd = self._st_block(f"decoder.mid_block.resnets.{j}.", hd, NB, F, h, w)
        for i in range(len(boc)):
            for j in range(lpb + 1):
                hd = self._st_block(f"decoder.up_blocks.{i}.resnets.{j}.", hd, NB, F, h, w)
            if i < len(boc) - 1:
                p = f"decoder.up_blocks.{i}.upsamplers.0.conv."
                up = ops.upsample2x(hd.view(NI, h, w, -1))
                h, w = 2 * h, 2 * w
                hd = ops.conv3x3(up, W[p + "weight"], bias=W[p + "bias"])
        hd = ops.groupnorm(hd, W["decoder.conv_norm_out.weight"], W["decoder.conv_norm_out.bias"], NB=NI, rows=h * w, eps=1e-6,
                           silu=True)
        co = cfg.out_channels
        px = torch.zeros((NI * h * w, CPAD), dtype=torch.float16, device=self.device)
        ops.conv3x3(hd.view(NI, h, w, -1), W["decoder.conv_out.weight"], bias=W["decoder.conv_out.bias"], out=px[:, :co])
        out = torch.empty((NI * h * w, 8), dtype=torch.float16, device=self.device)
        ops.conv_temporal3(px, W["decoder.time_conv_out.weight"], NB=NB, F=F, HW=h * w, bias=W["decoder.time_conv_out.bias"],
                           out=out[:, :co])
        return out, h, w

    @torch.no_grad()
    def decode(self, z, num_frames: int = 1, **kwargs):
        """z: (N, C, h, w) with N = clips x num_frames (already divided by the scaling factor) -> ``.sample`` (N, 3, H, W)."""
        N, C, h, w = z.shape
        if N % num_frames:
            raise ValueError("the batch must hold whole clips of num_frames frames")
        z = z.to(device=self.device, dtype=torch.float16).contiguous()
        rows = ops.pack_latents([z.permute(1, 0, 2, 3).contiguous()], Cpad=CPAD).view(N * h * w, CPAD)
        px, H, Wd = self._decode_rows(rows, N // num_frames, num_frames, h, w)
        out = px[:, : self.config.out_channels].reshape(N, H, Wd, -1).permute(0, 3, 1, 2).contiguous()
        return SimpleNamespace(sample=out)

    @torch.no_grad()
    def decode_latents_u8(self, latents, decode_chunk_size: int = 16):
        """latents (1, C, F, h, w) (scaled) -> uint8 frames (F, H, W, 3): ``get_images_from_latents`` (stable_diffusion.py:
        793-819) in one call -- 1 / scaling, decode ``decode_chunk_size`` frames at a time (each chunk is one clip to the
        temporal layers, :803-811), (x / 2 + 0.5).clamp(0, 1), round(255 x)."""
        _, C, F_, h, w = latents.shape
        lat = latents.to(self.device, torch.float16).contiguous()
        z = ops.axpby(lat, lat, 1.0 / self.config.scaling_factor, 0.0)
        outs = []
        for k in range(0, F_, decode_chunk_size):
            n = min(decode_chunk_size, F_ - k)
            rows = ops.pack_latents([z[0, :, k:k + n].contiguous()], Cpad=CPAD).view(n * h * w, CPAD)
            px, H, Wd = self._decode_rows(rows, 1, n, h, w)
            outs.append(ops.frames_to_u8(px, n * H * Wd).view(n, H, Wd, 3))
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)


def random_state_dict(cfg=None, seed: int = 55, device="cuda"):
    """Seeded random weights with diffusers' key names and shapes (no checkpoint exists offline)."""
    cfg = dict(VAE_CONFIG, **(cfg or {}))
    boc, lpb, lc = cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"]
    g = torch.Generator(device=device).manual_seed(seed)
    shapes: Dict[str, tuple] = {}

    def resnet(pre, cin, cout):
        shapes[pre + "norm1.weight"], shapes[pre + "norm1.bias"] = (cin,), (cin,)
        shapes[pre + "conv1.weight"], shapes[pre + "conv1.bias"] = (cout, cin, 3, 3), (cout,)
        shapes[pre + "norm2.weight"], shapes[pre + "norm2.bias"] = (cout,), (cout,)
        shapes[pre + "conv2.weight"], shapes[pre + "conv2.bias"] = (cout, cout, 3, 3), (cout,)
        if cin != cout:
            shapes[pre + "conv_shortcut.weight"], shapes[pre + "conv_shortcut.bias"] = (cout, cin, 1, 1), (cout,)

    def attn(pre, c):
        shapes[pre + "group_norm.weight"], shapes[pre + "group_norm.bias"] = (c,), (c,)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            shapes[pre + n + ".weight"], shapes[pre + n + ".bias"] = (c, c), (c,)

    def st_block(pre, cin, cout):
        resnet(pre + "spatial_res_block.", cin, cout)
        t = pre + "temporal_res_block."
        for n in ("norm1", "norm2"):
            shapes[t + n + ".weight"], shapes[t + n + ".bias"] = (cout,), (cout,)
        for n in ("conv1", "conv2"):
            shapes[t + n + ".weight"], shapes[t + n + ".bias"] = (cout, cout, 3, 1, 1), (cout,)
        shapes[pre + "time_mixer.mix_factor"] = (1,)

    shapes["encoder.conv_in.weight"], shapes["encoder.conv_in.bias"] = (boc[0], cfg["in_channels"], 3, 3), (boc[0],)
    cin = boc[0]
    for i, c in enumerate(boc):
        for j in range(lpb):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}.", cin, c)
            cin = c
        if i < len(boc) - 1:
            shapes[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (c, c, 3, 3)
            shapes[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (c,)
    resnet("encoder.mid_block.resnets.0.", boc[-1], boc[-1])
    attn("encoder.mid_block.attentions.0.", boc[-1])
    resnet("encoder.mid_block.resnets.1.", boc[-1], boc[-1])
    shapes["encoder.conv_norm_out.weight"], shapes["encoder.conv_norm_out.bias"] = (boc[-1],), (boc[-1],)
    shapes["encoder.conv_out.weight"], shapes["encoder.conv_out.bias"] = (2 * lc, boc[-1], 3, 3), (2 * lc,)
    shapes["quant_conv.weight"], shapes["quant_conv.bias"] = (2 * lc, 2 * lc, 1, 1), (2 * lc,)
    shapes["decoder.conv_in.weight"], shapes["decoder.conv_in.bias"] = (boc[-1], lc, 3, 3), (boc[-1],)
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
            shapes[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (c, c, 3, 3)
            shapes[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (c,)
    shapes["decoder.conv_norm_out.weight"], shapes["decoder.conv_norm_out.bias"] = (boc[0],), (boc[0],)
    shapes["decoder.conv_out.weight"], shapes["decoder.conv_out.bias"] = (cfg["out_channels"], boc[0], 3, 3), (cfg["out_channels"],)
    shapes["decoder.time_conv_out.weight"] = (cfg["out_channels"], cfg["out_channels"], 3, 1, 1)
    shapes["decoder.time_conv_out.bias"] = (cfg["out_channels"],)
    out = {}
    for k, shp in shapes.items():
        if k.endswith("mix_factor"):
            t = torch.randn(shp, device=device, generator=g)
        elif k.endswith("weight") and len(shp) == 1:
            t = torch.ones(shp, device=device)
        elif k.endswith("bias"):
            t = 0.02 * torch.randn(shp, device=device, generator=g)
        else:
            fan = 1
            for d in shp[1:]:
                fan *= d
            t = torch.randn(shp, device=device, generator=g) * fan ** -0.5
        out[k] = t.half()
    return out
