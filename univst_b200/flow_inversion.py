"""Mirror of the rectified-flow inversion loops of inversion_tools/flow_inversion.py for the SD3 backbone:
``rf_inversion`` (:123-188, Euler steps along a velocity interpolated between the model's and the straight line to a
fixed Gaussian target) and ``rf_solver`` (:191-264, second-order RF-Solver with a midpoint evaluation).  Same arguments and
file side effects (``ddim_latents_{k}.pt``).  ``pipeline`` is duck-typed like the reference's: ``encode_prompt``,
``scheduler.set_timesteps`` / ``.sigmas``, ``transformer(...)`` (third-party MMDiT, steered through the processors of
``univst_b200.sd3``); the latent arithmetic runs in ``univst_axpby_f16``.
"""
from __future__ import annotations

import os

import torch

from . import ops


def _save(latents, inversion_path, k):
    if inversion_path is not None:
        torch.save(latents.detach().clone(), os.path.join(inversion_path, f"ddim_latents_{k}.pt"))


def _schedule(pipeline, num_inference_steps):
    pipeline.scheduler.set_timesteps(num_inference_steps, device=pipeline.device)
    return [float(s) for s in torch.flip(pipeline.scheduler.sigmas, dims=[0])]   # t goes from 0.0 to 1.0 (:141-142)


def _velocity(pipeline, x, t, embeds, pooled, idx, **ft):
    t_vec = torch.full((x.shape[0],), t * 1000, dtype=x.dtype, device=x.device)
    return pipeline.transformer(hidden_states=x, timestep=t_vec, encoder_hidden_states=embeds, pooled_projections=pooled,
                                idx=idx, return_dict=False, **ft)[0].to(torch.float16).contiguous()


@torch.no_grad()
def rf_inversion(pipeline, image_latents, prompt="", gamma=0.5, num_inference_steps=50, inversion_path=None, ft_indices=None,
                 ft_timesteps=None, ft_path=None, target_noise=None):
    """flow_inversion.py:123-188.  ``target_noise``: the Gaussian target (:151 draws it with torch.randn_like; pass it for
    reproducibility across devices)."""
    embeds, _, pooled, _ = pipeline.encode_prompt(prompt=prompt, prompt_2=prompt, prompt_3=prompt)
    ts = _schedule(pipeline, num_inference_steps)
    x = image_latents.to(pipeline.device, torch.float16).contiguous()
    _save(x, inversion_path, 0)
    noise = (torch.randn_like(x) if target_noise is None else target_noise.to(x.device, torch.float16)).contiguous()
    for idx, (t_curr, t_prev) in enumerate(zip(ts[:-1], ts[1:])):
        v = _velocity(pipeline, x, t_curr, embeds, pooled, idx, ft_indices=ft_indices, ft_timesteps=ft_timesteps, ft_path=ft_path)
        # x + dt (gamma (noise - x) / (1 - t) + (1 - gamma) v)     (:170-175)
        dt = t_prev - t_curr
        a = dt * gamma / (1.0 - t_curr)
        x = ops.axpby(ops.axpby(x, noise, 1.0 - a, a), v, 1.0, dt * (1.0 - gamma))
        _save(x, inversion_path, idx + 1)
    return x


@torch.no_grad()
def rf_solver(pipeline, image_latents, prompt="", num_inference_steps=50, inversion_path=None, ft_indices=None,
              ft_timesteps=None, ft_path=None):
    """flow_inversion.py:191-264.  x + dt v + dt^2 / 2 * (v_mid - v) / (dt / 2) == x + dt v_mid (:250-253)."""
    embeds, _, pooled, _ = pipeline.encode_prompt(prompt=prompt, prompt_2=prompt, prompt_3=prompt)
    ts = _schedule(pipeline, num_inference_steps)
    x = image_latents.to(pipeline.device, torch.float16).contiguous()
    _save(x, inversion_path, 0)
    for idx, (t_curr, t_prev) in enumerate(zip(ts[:-1], ts[1:])):
        dt = t_prev - t_curr
        v = _velocity(pipeline, x, t_curr, embeds, pooled, idx, ft_indices=ft_indices, ft_timesteps=ft_timesteps, ft_path=ft_path)
        x_mid = ops.axpby(x, v, 1.0, dt / 2)
        v_mid = _velocity(pipeline, x_mid, t_curr + dt / 2, embeds, pooled, idx)
        x = ops.axpby(x, v_mid, 1.0, dt)
        _save(x, inversion_path, idx + 1)
    return x
