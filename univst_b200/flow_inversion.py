"""Mirror of the rectified-flow inversion loops of inversion_tools/flow_inversion.py for the SD3 backbone:
``rf_inversion`` (:123-188, Euler steps along a velocity interpolated between the model's and the straight line to a
fixed Gaussian target) and ``rf_solver`` (:191-264, second-order RF-Solver with a midpoint evaluation), plus the two front
doors that wrap them, ``content_inversion_reconstruction`` (:16-69) and ``style_inversion_reconstruction`` (:72-120).  Same
arguments and file side effects (``ddim_latents_{k}.pt``, the reconstructed clip).  ``pipeline`` is duck-typed like the reference's: ``encode_prompt``,
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


def _encode_frames(pipe, pixel_values):
    """flow_inversion.py:29-30 / :80-81: sampled posterior, shifted and scaled; SD3 latents stay frame-major (F, C, h, w)."""
    lat = pipe.vae.encode(pixel_values).latent_dist.sample()
    return (lat - pipe.vae.config.shift_factor) * pipe.vae.config.scaling_factor


def _invert_and_reconstruct(pipe, img_latents, inversion_path, reconstruction_path, name, time_steps, weight_dtype,
                            is_rf_solver, **ft):
    import numpy as np
    from .util import write_video
    if is_rf_solver:
        inv = rf_solver(pipe, img_latents, prompt="", num_inference_steps=time_steps, inversion_path=inversion_path, **ft)
    else:
        inv = rf_inversion(pipe, img_latents, prompt="", gamma=0.0, num_inference_steps=time_steps,
                           inversion_path=inversion_path, **ft)
    images = pipe.reconstruction(prompt="", img_latents=img_latents, inversed_latents=inv, eta_base=0.85,
                                 eta_trend="constant", start_step=25, end_step=39, guidance_scale=1.0, DTYPE=weight_dtype,
                                 num_inference_steps=time_steps)
    frames = [np.asarray(im) if not torch.is_tensor(im) else im.detach().cpu().numpy() for im in images]
    write_video(os.path.join(reconstruction_path, name), frames, fps=8)        # export_to_video(..., fps=8), :69 / :120
    return inv


@torch.no_grad()
def content_inversion_reconstruction(pipe, content_path, inversion_path, reconstruction_path, num_frames, height, width,
                                     time_steps, weight_dtype=torch.float16, ft_indices=None, ft_timesteps=None, ft_path=None,
                                     is_rf_solver=False):
    """flow_inversion.py:16-69: frames folder (``%05d.png``) or ``.mp4`` -> VAE posterior sample -> RF inversion (gamma 0)
    or RF-Solver -> reconstruction (eta 0.85 on steps 25..38) -> ``content_video.mp4``.  Returns the inverted latents."""
    from .util import load_video_frames
    if content_path.endswith(".mp4"):
        try:
            import decord
        except ImportError as e:
            raise ImportError("reading an .mp4 needs decord (requirements.txt of the reference); pass a folder of "
                              "%05d.png frames instead") from e
        decord.bridge.set_bridge("torch")
        vr = decord.VideoReader(content_path, width=width, height=height)
        pixel_values = (vr.get_batch(list(range(len(vr)))[:num_frames]) / 127.5 - 1.0).permute(0, 3, 1, 2)
    else:
        pixel_values = load_video_frames(content_path, num_frames, image_size=(width, height))
    img_latents = _encode_frames(pipe, pixel_values.to(weight_dtype).to(pipe.device))
    return _invert_and_reconstruct(pipe, img_latents, inversion_path, reconstruction_path, "content_video.mp4", time_steps,
                                   weight_dtype, is_rf_solver, ft_indices=ft_indices, ft_timesteps=ft_timesteps, ft_path=ft_path)


@torch.no_grad()
def style_inversion_reconstruction(pipe, style_path, inversion_path, reconstruction_path, num_frames, height, width,
                                   time_steps, weight_dtype=torch.float16, is_rf_solver=False):
    """flow_inversion.py:72-120: one style image, resized, ``2 x / 255 - 1``, repeated over the frames."""
    import numpy as np
    from PIL import Image
    img = Image.open(style_path).convert("RGB").resize((width, height))
    x = torch.from_numpy(np.array(img, dtype=np.uint8)).permute(2, 0, 1).float().div(255)     # transforms.ToTensor()
    pixel_values = (2.0 * x - 1.0).repeat(num_frames, 1, 1, 1)
    img_latents = _encode_frames(pipe, pixel_values.to(weight_dtype).to(pipe.device))
    return _invert_and_reconstruct(pipe, img_latents, inversion_path, reconstruction_path, "style_video.mp4", time_steps,
                                   weight_dtype, is_rf_solver)
