"""ORACLE (test infrastructure, not product code): plain-PyTorch CPU restatement of the loops around the UNet --
``video_style_transfer`` (backbones/video_diffusion_sd/pipelines/stable_diffusion.py:631-780), DDIM sampling step
(diffusers ``DDIMScheduler.step`` as configured for SD-1.5), ``ddim_loop`` /
``ddim_loop_plus`` / ``next_step`` (inversion_tools/ddim_inversion.py:88-204), ``load_mask`` (src/util.py:133-144).

Parity status: PINNED -- ``oracle/gen_golden_extra.py`` runs the reference's own ``video_style_transfer`` and
``ddim_loop(_plus)`` (imported on the test-only shim) on the same synthetic inputs and commits the resulting latents
under ``tests/golden/``; ``tests/test_oracle_cpu.py`` replays them through this file.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import unet_oracle as uo


class DDIMOracle:
    """SD-1.5's scheduler_config.json: beta_schedule scaled_linear [0.00085, 0.012], T = 1000, set_alpha_to_one False,
    steps_offset 1, leading spacing.  AnimateDiff (animatediff-v2.yaml:16-21 -> DDIMScheduler(**kwargs)): "linear" betas and
    the library default set_alpha_to_one True, i.e. the last DDIM step lands on x0 exactly."""

    def __init__(self, T=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", set_alpha_to_one=False):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, T, dtype=torch.float32) ** 2
        else:  # "linear": backbones/animatediff/animatediff-v2.yaml:16-21
            betas = torch.linspace(beta_start, beta_end, T, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.T = T

    def set_timesteps(self, n):
        self.n = n
        ratio = self.T // n
        self.timesteps = [int(t) for t in (np.arange(0, n) * ratio).round()[::-1] + 1]

    def alpha(self, t):
        return self.alphas_cumprod[t] if t >= 0 else self.final_alpha_cumprod

    def step(self, eps, t, x):
        """x_{t - T/n} from x_t (eta = 0, no clipping)."""
        a_t, a_p = self.alpha(t), self.alpha(t - self.T // self.n)
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps

    def next_step(self, eps, t, x):
        """ddim_inversion.py:190-204: x_t from x_{t - T/n}."""
        cur = min(t - self.T // self.n, 999)
        a_t, a_n = self.alpha(cur), self.alpha(t)
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_n ** 0.5 * x0 + (1 - a_n) ** 0.5 * eps


def load_mask_values(pixels_u8: np.ndarray) -> torch.Tensor:
    """src/util.py:138-143 on already-decoded PNG arrays (F, H, W) uint8: ``uint8 * 255`` wraps modulo 256, then
    ``clip(0, 1)`` -> 1 wherever the pixel is non-zero.  Returns (1, F, H, W) uint8."""
    wrapped = (pixels_u8.astype(np.uint8) * np.uint8(255)).astype(np.uint8)
    return torch.from_numpy(wrapped).unsqueeze(0).clip(0, 1)


def resized_mask(mask_1fhw: torch.Tensor, h: int, w: int, dtype=torch.float32) -> torch.Tensor:
    """stable_diffusion.py:689-691 -> (1, 1, F, h, w)."""
    return F.interpolate(mask_1fhw.to(dtype), size=(h, w), mode="bilinear", align_corners=False)[None, :]


def video_style_transfer(unet_fn, latents, traj_c, traj_s, mask_1fhw, ctx3, n=50, record=None, animatediff=False):
    """stable_diffusion.py:681-766.  ``unet_fn(x, t, ctx, idx) -> eps`` evaluates the patched three-branch UNet;
    ``traj_*[k]`` = inversion latent k (k = 1..n); ``mask_1fhw`` (1, F, H, W) in {0, 1} or None.
    ``animatediff``: the AnimationPipeline flavour (backbones/animatediff/pipelines/pipeline_animation.py:501-584):
    trajectory index hard-coded ``50 - i`` (:505-506), late AdaIN from ``i >= 0.8 n`` (:515), linear betas."""
    sch = DDIMOracle(beta_schedule="linear" if animatediff else "scaled_linear", set_alpha_to_one=bool(animatediff))
    sch.set_timesteps(n)
    z = latents.clone()
    for i, t in enumerate(sch.timesteps):
        k = 50 - i if animatediff else n - i
        zc, zs = traj_c[k], traj_s[k]
        if mask_1fhw is not None and i <= 0.9 * n:
            m = resized_mask(mask_1fhw, z.shape[-2], z.shape[-1], z.dtype)
            z = (1 - m) * z + m * zc
        if (i >= 0.8 * n if animatediff else i > 0.8 * n) and i <= 0.9 * n:
            m = resized_mask(mask_1fhw, z.shape[-2], z.shape[-1], z.dtype) if mask_1fhw is not None else 0.0
            z = (1.0 - m) * uo.latent_adain(z, zs) + m * zc
        eps = unet_fn(torch.cat([zc, zs, z]), t, ctx3, i)
        z = sch.step(eps[2:3], t, z)
        if record is not None and i in record:
            record[i] = z.clone()
    return z


def ddim_loop(unet_fn, latent, ctx, n, plus=False):
    """ddim_inversion.py:88-113 (plus=False) / :117-167 (Easy-Inv, plus=True).  Returns [x_0 .. x_n]."""
    sch = DDIMOracle()
    sch.set_timesteps(n)
    all_latent = [latent]
    latent = latent.clone()
    last_latent = None
    for i in range(n):
        t = sch.timesteps[n - i - 1]
        eps = unet_fn(latent, t, ctx, None)
        if plus and (0.05 + 0.2) * 50 > i > 0.05 * 50 and i > 0:
            latent = 0.5 * latent + 0.5 * last_latent
        last_latent = latent
        latent = sch.next_step(eps, t, latent)
        all_latent.append(latent)
    return all_latent


# ----------------------------------------------------------------------------------------------- synthetic inputs
def synthetic_inputs(seed: int, F_: int, hw: int, n: int = 50, mask_px: int = 64):
    """Deterministic small inputs shared by gen_golden_extra.py and the tests (SURVEY.md 8(d) recipe, reduced)."""
    g = torch.Generator().manual_seed(seed)
    sch = DDIMOracle()
    sch.set_timesteps(n)
    z0_c = torch.randn(1, 4, F_, hw, hw, generator=g)
    z0_s = torch.randn(1, 4, 1, hw, hw, generator=g).repeat(1, 1, F_, 1, 1) + 0.02 * torch.randn(1, 4, F_, hw, hw, generator=g)
    eps = torch.randn(1, 4, F_, hw, hw, generator=g)
    traj_c, traj_s = [z0_c], [z0_s]
    for t in sch.timesteps[::-1]:
        a = sch.alpha(t)
        traj_c.append(a ** 0.5 * z0_c + (1 - a) ** 0.5 * eps)
        traj_s.append(a ** 0.5 * z0_s + (1 - a) ** 0.5 * eps)
    yy, xx = np.mgrid[0:mask_px, 0:mask_px]
    frames = []
    for f in range(F_):
        dist = np.sqrt((xx - (mask_px / 2 + 0.75 * f)) ** 2 + (yy - mask_px / 2) ** 2)
        # anti-aliased rim (grey levels 1..254) like the shipped examples/masks/mallard-fly.png
        frames.append(np.clip((mask_px / 4 - dist) * 64 + 128, 0, 255).astype(np.uint8))
    mask_px_u8 = np.stack(frames)
    return traj_c, traj_s, mask_px_u8
