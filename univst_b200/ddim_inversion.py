"""Mirror of the inversion loops of inversion_tools/ddim_inversion.py: ``ddim_inversion`` (:68-85), ``ddim_loop``
(:88-113), ``ddim_loop_plus`` (Easy-Inv, :117-167), ``next_step`` (:190-204).  Same arguments and file side effects
(``ddim_latents_{k}.pt``, the feature-map dump of the UNet); the arithmetic is the B200 UNet + ``univst_ddim_step``."""
from __future__ import annotations

import os

import torch

from . import ops


def _context(pipeline, prompt, prompt_embeds):
    if prompt_embeds is not None:
        return prompt_embeds.to(pipeline.device, torch.float16)
    return pipeline._encode_prompt(prompt)


def _save(latent, inversion_path, k):
    if inversion_path is not None:
        torch.save(latent.detach().clone(), os.path.join(inversion_path, f"ddim_latents_{k}.pt"))


def next_step(pipeline, t, latent, ddim_scheduler, branch=0):
    """ddim_inversion.py:190-204 applied to the eps the UNet just produced (kept channels-last on the device)."""
    a_cur, a_next = ddim_scheduler.inversion_alphas(t)
    return ops.ddim_step(latent, pipeline.unet.last_eps_rows, branch, a_cur, a_next)


@torch.no_grad()
def ddim_loop(pipeline, ddim_scheduler, latent, num_inv_steps, prompt, inversion_path, ft_indices=None,
              ft_timesteps=None, ft_path=None, prompt_embeds=None, plus=False):
    ctx = _context(pipeline, prompt, prompt_embeds)
    latent = latent.to(pipeline.device, torch.float16).contiguous()
    all_latent = [latent]
    _save(latent, inversion_path, 0)
    latent = latent.clone()
    last_latent = None
    timesteps = [int(t) for t in ddim_scheduler.timesteps]
    for i in range(num_inv_steps):
        t = timesteps[len(timesteps) - i - 1]
        pipeline.unet(latent, t, encoder_hidden_states=ctx, ft_indices=ft_indices, ft_timesteps=ft_timesteps,
                      ft_path=ft_path)
        if plus and (0.05 + 0.2) * 50 > i > 0.05 * 50 and i > 0:
            # Easy-Inv (:142-145): after eps was predicted, pull the latent halfway back to the previous one
            latent = ops.axpby(latent, last_latent, 0.5, 0.5)
        last_latent = latent
        latent = next_step(pipeline, t, latent, ddim_scheduler)
        _save(latent, inversion_path, i + 1)
        all_latent.append(latent)
    return all_latent


def ddim_loop_plus(pipeline, ddim_scheduler, latent, num_inv_steps, prompt, inversion_path, ft_indices=None,
                   ft_timesteps=None, ft_path=None, prompt_embeds=None):
    return ddim_loop(pipeline, ddim_scheduler, latent, num_inv_steps, prompt, inversion_path, ft_indices, ft_timesteps,
                     ft_path, prompt_embeds, plus=True)


def ddim_inversion(pipeline, ddim_scheduler, video_latent, num_inv_steps, prompt="", inversion_path=None,
                   ft_indices=None, ft_timesteps=None, ft_path=None, is_opt=False, prompt_embeds=None):
    """ddim_inversion.py:68-85."""
    fn = ddim_loop_plus if is_opt else ddim_loop
    return fn(pipeline, ddim_scheduler, video_latent, num_inv_steps, prompt, inversion_path, ft_indices=ft_indices,
              ft_timesteps=ft_timesteps, ft_path=ft_path, prompt_embeds=prompt_embeds)
