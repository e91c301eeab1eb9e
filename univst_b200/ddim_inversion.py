"""Mirror of inversion_tools/ddim_inversion.py: the inversion loops ``ddim_inversion`` (:68-85), ``ddim_loop`` (:88-113),
``ddim_loop_plus`` (Easy-Inv, :117-167), ``next_step`` (:190-204) and the two front doors that wrap them,
``content_inversion_reconstruction`` (:16-42) and ``style_inversion_reconstruction`` (:45-66).  Same arguments and file
side effects (``ddim_latents_{k}.pt``, the feature-map dump of the UNet, the reconstructed clip); the arithmetic is the
B200 UNet + ``univst_ddim_step`` (+ the VAE mirror when ``pipe.vae`` is ``univst_b200.vae.AutoencoderKLTemporalDecoder``)."""
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
    """ddim_inversion.py:190-204.  Two call forms: ``next_step(pipeline, t, latent, scheduler)`` -- what the loops below
    use: the eps the UNet just produced is read where the forward left it (channels-last, on the device) -- and the
    reference's own ``next_step(model_output, timestep, sample, scheduler)`` on a noise tensor shaped like the sample:
    ``x_next = sqrt(a') (x - sqrt(1 - a) eps) / sqrt(a) + sqrt(1 - a') eps`` as one fused multiply-add pass."""
    if torch.is_tensor(pipeline):
        eps, sample = pipeline, latent
        a_cur, a_next = ddim_scheduler.inversion_alphas(t)
        wx = (a_next / a_cur) ** 0.5
        we = (1.0 - a_next) ** 0.5 - wx * (1.0 - a_cur) ** 0.5
        return ops.axpby(sample.to(torch.float16).contiguous(), eps.to(sample.device, torch.float16).contiguous(), wx, we)
    a_cur, a_next = ddim_scheduler.inversion_alphas(t)
    return ops.ddim_step(latent, pipeline.unet.last_eps_rows, branch, a_cur, a_next)


@torch.no_grad()
def init_prompt(pipeline, prompt):
    """ddim_inversion.py:171-187: ``cat([embedding of "", embedding of prompt])`` from the pipeline's own (third-party)
    tokenizer / text encoder; (2, 77, D)."""
    tok, enc = pipeline.tokenizer, pipeline.text_encoder
    un = tok([""], padding="max_length", max_length=tok.model_max_length, return_tensors="pt")
    tx = tok([prompt], padding="max_length", max_length=tok.model_max_length, truncation=True, return_tensors="pt")
    dev = pipeline.device
    return torch.cat([enc(un.input_ids.to(dev))[0], enc(tx.input_ids.to(dev))[0]])


def get_noise_pred_single(pipeline, latents, t, context, ft_indices=None, ft_timesteps=None, ft_path=None):
    """ddim_inversion.py:207-212."""
    return pipeline.unet(latents, t, encoder_hidden_states=context, ft_indices=ft_indices, ft_timesteps=ft_timesteps,
                         ft_path=ft_path)["sample"]


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


def _encode_clip(pipe, pixel_values, num_frames):
    """ddim_inversion.py:29-31 / :53-55: sampled posterior of every frame -> (1, C, F, h, w) x scaling factor."""
    lat = pipe.vae.encode(pixel_values).latent_dist.sample()
    lat = lat.view(-1, num_frames, *lat.shape[1:]).permute(0, 2, 1, 3, 4)
    return lat * pipe.vae.config.scaling_factor


def _invert_and_reconstruct(pipe, scheduler, latents, inversion_path, reconstruction_path, name, num_frames, time_steps,
                            weight_dtype, is_opt, fps_kw, prompt_embeds, **ft):
    from .util import save_videos_grid
    kw = {} if prompt_embeds is None else {"prompt_embeds": prompt_embeds}
    inv = ddim_inversion(pipe, scheduler, video_latent=latents, num_inv_steps=time_steps, prompt="",
                         inversion_path=inversion_path, is_opt=is_opt, **ft, **kw)[-1].to(weight_dtype)
    sample = pipe.reconstruction("", latents=inv, video_length=num_frames, guidance_scale=1.0, **kw).images
    sample = torch.as_tensor(sample).permute(0, 4, 1, 2, 3).contiguous()
    save_videos_grid(sample, os.path.join(reconstruction_path, name), **fps_kw)
    return inv


@torch.no_grad()
def content_inversion_reconstruction(pipe, ddim_inv_scheduler, content_path, inversion_path, reconstruction_path,
                                     num_frames, height, width, time_steps, weight_dtype=torch.float16, ft_indices=None,
                                     ft_timesteps=None, ft_path=None, is_opt=True, prompt_embeds=None):
    """ddim_inversion.py:16-42: frames folder (``%05d.png``) or ``.mp4`` -> VAE posterior sample -> inversion (files
    ``ddim_latents_{k}.pt`` under ``inversion_path``, feature dump under ``ft_path``) -> reconstruction from the last
    latent -> ``content_video.mp4`` under ``reconstruction_path`` (PNG frames when imageio is missing, util.py).
    Returns the inverted latent.  ``prompt_embeds``: the empty-prompt embedding when the pipeline has no text encoder."""
    from .util import load_video_frames
    if content_path.endswith(".mp4"):
        try:
            import decord
        except ImportError as e:
            raise ImportError("reading an .mp4 needs decord (requirements.txt of the reference); pass a folder of "
                              "%05d.png frames instead") from e
        decord.bridge.set_bridge("torch")
        vr = decord.VideoReader(content_path, width=width, height=height)
        video = vr.get_batch(list(range(len(vr)))[:num_frames])
        pixel_values = (video / 127.5 - 1.0).permute(0, 3, 1, 2)
    else:
        pixel_values = load_video_frames(content_path, num_frames, image_size=(width, height))
    latents = _encode_clip(pipe, pixel_values.to(weight_dtype).to(pipe.device), num_frames)
    return _invert_and_reconstruct(pipe, ddim_inv_scheduler, latents, inversion_path, reconstruction_path,
                                   "content_video.mp4", num_frames, time_steps, weight_dtype, is_opt, {}, prompt_embeds,
                                   ft_indices=ft_indices, ft_timesteps=ft_timesteps, ft_path=ft_path)


@torch.no_grad()
def style_inversion_reconstruction(pipe, ddim_inv_scheduler, style_path, inversion_path, reconstruction_path, num_frames,
                                   height, width, time_steps, weight_dtype=torch.float16, is_opt=True, prompt_embeds=None):
    """ddim_inversion.py:45-66: one style image, resized, ``2 x / 255 - 1``, repeated over the frames (every frame gets its
    own posterior sample, :53) -> inversion -> reconstruction -> ``style_video.mp4`` (fps 8)."""
    import numpy as np
    from PIL import Image
    img = Image.open(style_path).convert("RGB").resize((width, height))
    x = torch.from_numpy(np.array(img, dtype=np.uint8)).permute(2, 0, 1).float().div(255)     # transforms.ToTensor()
    pixel_values = (2.0 * x - 1.0).repeat(num_frames, 1, 1, 1)
    latents = _encode_clip(pipe, pixel_values.to(weight_dtype).to(pipe.device), num_frames)
    return _invert_and_reconstruct(pipe, ddim_inv_scheduler, latents, inversion_path, reconstruction_path,
                                   "style_video.mp4", num_frames, time_steps, weight_dtype, is_opt, {"fps": 8},
                                   prompt_embeds)
