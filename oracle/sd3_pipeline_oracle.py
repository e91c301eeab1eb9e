"""ORACLE (test infrastructure): CPU restatement of the SD3 sampling loops of
backbones/video_diffusion_sd3/pipelines/custom_pipeline.py -- ``generate_eta_values`` (:18-44), ``reconstruction``
(:46-124) and ``video_style_transfer`` (:126-371) -- plus the stand-in pipeline parts the goldens are generated with.
Parity status: PINNED -- ``oracle/gen_golden_sd3_pipeline.py`` runs the reference's own three methods on an instance whose
third-party members (scheduler, transformer, text encoders, VAE) are the stand-ins below and commits the trajectories
under ``tests/golden/sd3_pipeline.pt``.

Reference defect handled here (and in the product): ``video_style_transfer`` reads an undefined name
``ddim_inv_latents_at_t`` at :316 (late AdaIN window, every 50-step run reaches it at i = 40).  The SD and AnimateDiff
loops blend the content inversion latent of the step there (stable_diffusion.py:704), and so do the golden generator (by
defining the name to that tensor), this oracle and the product.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------------
# stand-ins for the third-party members of the pipeline
# ------------------------------------------------------------------------------------------------------------
class _Config(dict):
    __getattr__ = dict.__getitem__


class FakeFlowMatchScheduler:
    """FlowMatchEulerDiscreteScheduler as the loops use it [diffusers 0.35.1, third party -- restated]: ``set_timesteps``
    fills ``sigmas`` (descending, shifted by 3, a trailing 0) and ``timesteps = 1000 sigma``; ``step`` is the Euler update
    ``x + (sigma_next - sigma) v`` in fp32, returned in the dtype of ``v``, advancing an internal step index."""

    def __init__(self):
        self.config = _Config(num_train_timesteps=1000, use_dynamic_shifting=False)
        self.order = 1

    def set_timesteps(self, n, device=None, **kw):
        s = torch.linspace(1.0, 0.001, n)
        s = 3.0 * s / (1.0 + 2.0 * s)
        self.sigmas = torch.cat([s, torch.zeros(1)])
        self.timesteps = self.sigmas[:-1] * 1000.0
        self._step = 0

    def step(self, model_output, t, sample, return_dict=False):
        ds = self.sigmas[self._step + 1] - self.sigmas[self._step]
        self._step += 1
        return ((sample.float() + ds.to(sample.device) * model_output.float()).to(model_output.dtype),)


class FakeTransformer:
    """A fixed velocity field standing in for the MMDiT: channel mixing through tanh, a timestep term, and a coupling of
    every image to the image one branch earlier (so that the edit branch depends on the style branch, as it does through
    the attention shift).  Records the ``idx`` the loop passes through ``joint_attention_kwargs``."""

    def __init__(self, channels=16, seed=13, device="cpu", dtype=torch.float32):
        g = torch.Generator().manual_seed(seed)
        self.mix = (torch.randn(channels, channels, generator=g) * channels ** -0.5).to(device, dtype)
        self.config = _Config(in_channels=channels, patch_size=2)
        self.calls = []

    def __call__(self, hidden_states, timestep, encoder_hidden_states=None, pooled_projections=None, return_dict=False,
                 joint_attention_kwargs=None, **kw):
        self.calls.append((None if joint_attention_kwargs is None else joint_attention_kwargs.get("idx"), float(timestep[0])))
        x = hidden_states
        third = max(x.shape[0] // 3, 1)
        v = torch.tanh(torch.einsum("oc,bchw->bohw", self.mix.to(x.dtype), x))
        v = v + 0.1 * torch.sin(timestep.to(x.dtype) / 1000.0 * 3.0)[:, None, None, None] + 0.05 * torch.roll(x, third, 0)
        return (v,)


def fake_encode_prompt(device="cpu", dtype=torch.float32):
    def encode_prompt(**kw):
        e = torch.zeros(1, 4, 8, device=device, dtype=dtype)
        return e, e, e[:, 0], e[:, 0]
    return encode_prompt


def synthetic_inputs(seed=5, frames=3, channels=16, hw=8, n=10, mask_px=64):
    """Content / style trajectories for k = 50 - n + 1 .. 50 (the loop reads ``50 - i``), the clean content latents, and
    per-frame uint8 masks."""
    import numpy as np
    g = torch.Generator().manual_seed(seed)
    z0_c = torch.randn(frames, channels, hw, hw, generator=g)
    z0_s = torch.randn(1, channels, hw, hw, generator=g).repeat(frames, 1, 1, 1) + 0.02 * torch.randn(frames, channels, hw, hw, generator=g)
    eps = torch.randn(frames, channels, hw, hw, generator=g)
    traj_c, traj_s = {0: z0_c}, {0: z0_s}
    for k in range(50 - n + 1, 51):
        s = k / 50.0
        traj_c[k] = (1 - s) * z0_c + s * eps
        traj_s[k] = (1 - s) * z0_s + s * eps
    yy, xx = np.mgrid[0:mask_px, 0:mask_px]
    mask = np.stack([(((xx - (mask_px / 2 + 3 * f)) ** 2 + (yy - mask_px / 2) ** 2) < (mask_px / 4) ** 2).astype(np.uint8) * 255
                     for f in range(frames)])
    return traj_c, traj_s, mask


# ------------------------------------------------------------------------------------------------------------
# the loops
# ------------------------------------------------------------------------------------------------------------
def generate_eta_values(timesteps, start_step, end_step, eta, eta_trend):
    """custom_pipeline.py:18-44."""
    assert 0 <= start_step < end_step <= len(timesteps)
    out = [0.0] * len(timesteps)
    if eta_trend == "constant":
        for i in range(start_step, end_step):
            out[i] = eta
    elif eta_trend in ("linear_increase", "linear_decrease"):
        total = timesteps[start_step] - timesteps[end_step - 1]
        for i in range(start_step, end_step):
            num = (timesteps[start_step] - timesteps[i]) if eta_trend == "linear_increase" else (timesteps[i] - timesteps[end_step - 1])
            out[i] = eta * num / total
    else:
        raise NotImplementedError(eta_trend)
    return out


def latent_adain(cnt, sty):
    """sd3 pnp_utils.py:304-316 on (F, C, h, w): per-(frame, channel) plane statistics."""
    sm, ss = sty.mean(dim=[2, 3], keepdim=True), sty.std(dim=[2, 3], keepdim=True)
    return (F.instance_norm(cnt) * ss + sm).to(cnt.dtype)


def resized_mask(mask_1fhw, h, w, dtype):
    """custom_pipeline.py:300-303: bilinear to the latent size, frames to the batch axis -> (F, 1, h, w)."""
    m = F.interpolate(mask_1fhw.to(dtype), size=(h, w), mode="bilinear", align_corners=False)
    return m.permute(1, 0, 2, 3).contiguous()


def reconstruction(scheduler, transformer, img_latents, inversed_latents, eta_base, eta_trend, start_step, end_step, n=50):
    """custom_pipeline.py:46-118 with guidance_scale 1 (the loop; the VAE decode after it is third party)."""
    scheduler.set_timesteps(n)
    ts = scheduler.timesteps
    etas = generate_eta_values(ts, start_step, end_step, eta_base, eta_trend)
    z, target = inversed_latents, img_latents.clone().float()
    for i, t in enumerate(ts):
        v = transformer(hidden_states=z, timestep=t.expand(z.shape[0]), return_dict=False)[0].float()
        z = z.float()
        tv = -(target - z) / (t / scheduler.config.num_train_timesteps)
        z = scheduler.step(v + etas[i] * (tv - v), t, z)[0]
    return z


def video_style_transfer(scheduler, transformer, latents, img_latents, traj_c, traj_s, mask_1fhw, n, eta_base, eta_trend,
                         start_step, end_step, record=None):
    """custom_pipeline.py:282-352.  ``traj_*[k]``: inversion latents (F, C, h, w); the loop reads k = 50 - i."""
    scheduler.set_timesteps(n)
    ts = scheduler.timesteps
    etas = generate_eta_values(ts, start_step, end_step, eta_base, eta_trend)
    z, target = latents, img_latents.clone()
    for i, t in enumerate(ts):
        zc, zs = traj_c[50 - i].to(z.dtype), traj_s[50 - i].to(z.dtype)
        if mask_1fhw is not None and i <= 0.9 * n:
            m = resized_mask(mask_1fhw, z.shape[-2], z.shape[-1], z.dtype)
            z = (1 - m) * z + m * zc
        if 0.8 * n <= i <= 0.9 * n:
            m = resized_mask(mask_1fhw, z.shape[-2], z.shape[-1], z.dtype) if mask_1fhw is not None else 0.0
            z = (1.0 - m) * latent_adain(z, zs) + m * zc       # `ddim_inv_latents_at_t` of :316, see the module docstring
        x = torch.cat([zc, zs, z])
        v = transformer(hidden_states=x, timestep=t.expand(x.shape[0]), return_dict=False,
                        joint_attention_kwargs={"idx": i})[0].chunk(3)[2]
        tv = -(target - z) / (t / scheduler.config.num_train_timesteps)
        z = scheduler.step(v + etas[i] * (tv - v), t, z)[0]
        if record is not None:
            record.append(z.clone())
    return z
