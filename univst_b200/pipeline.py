"""Mirror of ``SpatioTemporalStableDiffusionPipeline`` (backbones/video_diffusion_sd/pipelines/stable_diffusion.py:45)
for the denoising loops on the hot path: ``video_style_transfer`` (:631-780) and ``reconstruction`` (:479-628).

Same method names and keyword arguments.  What differs, deliberately:
* the per-step ``torch.load`` of two latents and the re-decoding of 32 mask PNGs (:683-689) are hoisted out of the
  loop: trajectories and the resized mask are loaded once and stay in HBM;
* scheduler scalars are Python floats, so the loop never synchronises with the device;
* while the AdaIN-guided shift is inactive (``idx > eta2 * 50``) the content and style branches cannot influence the
  edit branch (every op is per-sample; their predictions are discarded at :712), so with ``skip_dead_branches=True``
  only the edit branch is evaluated on those steps -- bit-identical edit output, 32 % fewer FLOPs;
* the VAE / CLIP encoders are third-party networks whose weights are not available offline: ``prompt_embeds`` may be
  passed directly and decoding happens only when a ``vae`` was supplied (``output.latents`` is always returned).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Union

import torch

from . import ops
from .pnp_utils import register_time
from .scheduler import DDIMScheduler
from .util import load_ddim_latents_at_t, load_mask


class SpatioTemporalStableDiffusionPipeline:
    def __init__(self, unet, scheduler: Optional[DDIMScheduler] = None, vae=None, text_encoder=None, tokenizer=None):
        self.unet = unet
        self.scheduler = (scheduler or DDIMScheduler.sd15()).patch_for_pipeline()   # stable_diffusion.py:68-93
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.device = unet.device

    # ------------------------------------------------------------------------------------------ helpers
    def _encode_prompt(self, prompt, prompt_embeds=None):
        if prompt_embeds is not None:
            return prompt_embeds.to(self.device, torch.float16)
        if self.text_encoder is None or self.tokenizer is None:
            raise ValueError("no text encoder was given: pass prompt_embeds=(1, 77, D)")
        ids = self.tokenizer([prompt], padding="max_length", max_length=self.tokenizer.model_max_length,
                             truncation=True, return_tensors="pt").input_ids
        return self.text_encoder(ids.to(self.device))[0].to(torch.float16)

    def _trajectory(self, path_or_list, n):
        """Latents x_0 .. x_n of an inversion: a directory of ``ddim_latents_{k}.pt`` or a list of tensors."""
        if isinstance(path_or_list, (list, tuple)):
            lat = list(path_or_list)
        else:
            lat = [load_ddim_latents_at_t(k, path_or_list) for k in range(1, n + 1)]
            lat = [None] + lat
        return [None if z is None else z.to(self.device, torch.float16).contiguous() for z in lat]

    def _mask(self, mask, F, h, w):
        if mask is None or (isinstance(mask, str) and not mask):
            return None
        m = load_mask(mask, n_frames=F) if isinstance(mask, str) else mask
        m = m.reshape(-1, m.shape[-2], m.shape[-1])
        return ops.mask_resize((m != 0).to(torch.uint8).to(self.device).contiguous(), h, w)

    # ------------------------------------------------------------------------------------------ backbone flavour
    @staticmethod
    def _traj_index(i, n):
        return n - i   # load_ddim_latents_at_t(num_inference_steps - i, ...), :683-684

    @staticmethod
    def _late_adain(i, n):
        return i > 0.8 * n and i <= 0.9 * n   # :694

    # ------------------------------------------------------------------------------------------ hot loop
    @torch.no_grad()
    def video_style_transfer(self, prompt: Union[str, List[str]], num_inference_steps: int = 50, latents=None,
                             content_inv_path=None, style_inv_path=None, mask_path=None, prompt_embeds=None,
                             output_type="tensor", skip_dead_branches: bool = True, callback=None,
                             inv_prompt_embeds=None, smoother=None, flow_fn=None, smoother_generator=None, **kwargs):
        """stable_diffusion.py:631-780.  ``content_inv_path`` / ``style_inv_path``: directory with the inversion's
        ``ddim_latents_{k}.pt`` files (reference format) or an in-memory list [x_0 .. x_n]; ``mask_path``: directory
        of ``%05d.png`` masks or a (F, H, W) tensor (non-zero = keep content).  ``prompt_embeds`` /
        ``inv_prompt_embeds``: embeddings of ``prompt`` (edit branch) and of the empty prompt the inversion branches are
        conditioned on (:659-668); the latter defaults to ``prompt_embeds`` only when ``prompt`` itself is empty.
        ``skip_dead_branches`` (default on; exact): the content / style branches are evaluated only
        where they can still influence the edit branch -- not at all once the shift window is closed, and up to the
        last patched attn1 projection while it is open.  The edit latents are bit-identical to the full evaluation.
        ``smoother="pixel"`` switches the sliding-window smoother of :713-758 on (the reference hard-codes
        ``smoother = None``, :715): on steps 20..24 the predicted x0 is decoded to uint8 frames (``self.vae``), every key
        frame is replaced in place by the mean of itself and its +-2 flow-warped neighbours, pixels inside the mask keep
        their value, the frames are re-encoded and the noise prediction is recomputed from the stabilised x0
        (``return_to_timestep``, :782-791).  ``flow_fn(key_frame, now_frame) -> (fwd, bwd)`` supplies the optical flows
        ([H, W, 2] fp32 CUDA; the reference asks torchvision's RAFT-large, whose weights are not available offline);
        ``smoother_generator`` seeds the posterior sample of the re-encoding (default: a fixed seed, so that the ranks of a
        frame-sharded run, which all evaluate the smoother on the whole clip, stay in lockstep)."""
        n = num_inference_steps
        emb = self._encode_prompt(prompt, prompt_embeds)
        if inv_prompt_embeds is None and prompt_embeds is not None and prompt not in ("", [""]) \
                and (self.text_encoder is None or self.tokenizer is None):
            raise ValueError("a non-empty prompt needs inv_prompt_embeds (the empty-prompt embedding of the content / "
                             "style branches, stable_diffusion.py:659-668) or a text encoder")
        emb_inv = self._encode_prompt("", inv_prompt_embeds if inv_prompt_embeds is not None
                                      else (prompt_embeds if prompt in ("", [""]) else None))
        ctx = torch.cat([emb_inv, emb_inv, emb])  # :668
        self.scheduler.set_timesteps(n)
        timesteps = [int(t) for t in self.scheduler.timesteps]
        z = latents.to(self.device, torch.float16).contiguous().clone()
        _, C, F, h, w = z.shape
        zc_traj, zs_traj = self._trajectory(content_inv_path, n), self._trajectory(style_inv_path, n)
        m = self._mask(mask_path, F, h, w)
        keep_u8 = None
        if smoother is not None:
            if smoother != "pixel":
                raise ValueError("smoother must be None or 'pixel' (stable_diffusion.py:719,755)")
            if self.vae is None or flow_fn is None:
                raise ValueError("the pixel smoother needs a VAE (univst_b200.vae) and flow_fn(key_frame, now_frame)")
            if mask_path is None or (isinstance(mask_path, str) and not mask_path):
                raise ValueError("the pixel smoother needs the mask (stable_diffusion.py:751 reads it unconditionally)")
            mm = load_mask(mask_path, n_frames=F) if isinstance(mask_path, str) else mask_path
            keep_u8 = (mm.reshape(-1, mm.shape[-2], mm.shape[-1]) != 0).to(torch.uint8).to(self.device).contiguous()
            if smoother_generator is None:
                smoother_generator = torch.Generator(device=self.device).manual_seed(0)
        for i, t in enumerate(timesteps):
            zc, zs = zc_traj[self._traj_index(i, n)], zs_traj[self._traj_index(i, n)]
            if m is not None and i <= 0.9 * n:  # localized latent blending, :687-692
                z = ops.latent_blend(z, zc, m)
            if self._late_adain(i, n):  # late latent AdaIN, :694-702
                za = ops.latent_adain(z, zs)
                z = ops.latent_blend(za, zc, m) if m is not None else za
            register_time(self, i)  # :707
            a1 = self.unet.up_blocks[1].attentions[1].transformer_blocks[0].attn1
            if skip_dead_branches and not self.unet.shift_live(a1):
                self.unet(z, t, encoder_hidden_states=ctx[2:3])
            else:
                self.unet.truncate_dead_branches = bool(skip_dead_branches)
                try:
                    self.unet(torch.cat([zc, zs, z]), t, encoder_hidden_states=ctx)
                finally:
                    self.unet.truncate_dead_branches = False
            a_t, a_prev = self.scheduler.step_alphas(t)
            if smoother is not None and 20 <= i < 25:   # sliding-window smoothing, :716-758
                from .flow_warp import sliding_window_smooth
                x0 = torch.empty_like(z)
                ops.ddim_step(z, self.unet.last_eps_rows, self.unet.last_edit_branch, a_t, a_prev, x0_out=x0)   # :718
                frames = self._smoother_decode(x0)                                                               # :721
                est = sliding_window_smooth(frames, keep_mask=keep_u8, flow_fn=flow_fn)                          # :725-751
                x0s = self._smoother_encode(est, smoother_generator)                                             # :753
                eps = ops.axpby(z, x0s, 1.0 / (1.0 - a_t) ** 0.5, -(a_t ** 0.5) / (1.0 - a_t) ** 0.5)           # :782-791
                x0r = ops.axpby(z, eps, 1.0 / a_t ** 0.5, -((1.0 - a_t) ** 0.5) / a_t ** 0.5)                  # the step's own x0
                z = ops.axpby(x0r, eps, a_prev ** 0.5, (1.0 - a_prev) ** 0.5)                                    # :761
            else:
                z = ops.ddim_step(z, self.unet.last_eps_rows, self.unet.last_edit_branch, a_t, a_prev)  # :761, eta = 0
            if callback is not None:
                callback(i, t, z)
        images = self.decode_latents(z) if self.vae is not None else None
        return SimpleNamespace(images=images, latents=z, nsfw_content_detected=None)

    @torch.no_grad()
    def reconstruction(self, prompt, latents=None, video_length=None, num_inference_steps: int = 50,
                       guidance_scale: float = 1.0, prompt_embeds=None, **kwargs):
        """stable_diffusion.py:479-628 with guidance_scale = 1 (how ddim_inversion.py:40 calls it): plain DDIM sampling."""
        if guidance_scale != 1.0:
            raise NotImplementedError("classifier-free guidance is not used on the UniVST path")
        ctx = self._encode_prompt(prompt, prompt_embeds)
        self.scheduler.set_timesteps(num_inference_steps)
        z = latents.to(self.device, torch.float16).contiguous().clone()
        for t in [int(t) for t in self.scheduler.timesteps]:
            self.unet(z, t, encoder_hidden_states=ctx)
            a_t, a_prev = self.scheduler.step_alphas(t)
            z = ops.ddim_step(z, self.unet.last_eps_rows, 0, a_t, a_prev)
        images = self.decode_latents(z) if self.vae is not None else None
        return SimpleNamespace(images=images, latents=z)

    # ------------------------------------------------------------------------------------------ smoother legs
    def _smoother_xr(self, F, W):
        """The cross-rank plumbing of a frame-sharded UNet (xrank transport) when the smoother's VAE legs can be spread over
        its ranks: decode chunks (16 frames = one clip to the temporal decoder, :803-811) round-robin, encode frames evenly."""
        shard, xr = getattr(self.unet, "_shard", None), getattr(self.unet, "_xr", None)
        if shard is None or xr is None or F % xr.world or (W * 3 // 2) % 8 or (W * 3) % 2 or not hasattr(self.vae, "decode_latents_u8"):
            return None
        return xr

    def _smoother_decode(self, x0, decode_chunk_size: int = 16):
        """``get_images_from_latents`` (stable_diffusion.py:793-819).  Frames sharded over GPUs: the 16-frame chunks are
        independent clips to the temporal decoder, so chunk c is decoded by rank c mod P and its uint8 frames are stored
        into every rank's full-clip buffer over NVLink (one multicast store per 16 bytes; the kernel's tail synchronises) --
        the same bits as decoding everything on every rank."""
        _, C, F, h, w = x0.shape
        H, W = 8 * h, 8 * w
        xr = self._smoother_xr(F, W)
        if xr is None:
            return self.vae.decode_latents_u8(x0, decode_chunk_size)
        cols = W * 3 // 2                                              # a uint8 frame row as fp16 pairs
        key = ("vae_frames", F, H, W)
        full, ptrs = xr.buffer(key, (F * H, cols))
        mc = xr.multicast(key)
        nchunks = (F + decode_chunk_size - 1) // decode_chunk_size
        for r0 in range(0, nchunks, xr.world):                         # every rank issues the same number of pushes
            c = r0 + xr.rank
            pushes = []
            if c < nchunks:
                f0, n = c * decode_chunk_size, min(decode_chunk_size, F - c * decode_chunk_size)
                u8 = self.vae.decode_latents_u8(x0[:, :, f0:f0 + n].contiguous(), decode_chunk_size)
                src = u8.reshape(n * H, W * 3).view(torch.float16)
                off = f0 * H * cols * 2
                pushes = [dict(src=src, src_blk_rows=n * H, dst=[p + off for p in ptrs], ld_dst=cols, dst_blk_rows=n * H, nblk=1,
                               rows=n * H, mc=mc + off if mc else 0)]
            ops.xrank_push(xr, pushes)
        return full.view(torch.uint8).view(F, H, W, 3)

    def _smoother_encode(self, frames_u8, generator):
        """``get_latent_image`` (:821-834).  The KL encoder is per frame: under frame sharding every rank encodes its own
        frames with its slice of the clip's posterior noise (drawn identically on every rank) and the latents are stored into
        every rank's buffer -- bit-identical to encoding the whole clip on every rank."""
        F, H, W, _ = frames_u8.shape
        xr = self._smoother_xr(F, W)
        if xr is None:
            return self.vae.encode_frames_u8(frames_u8, generator=generator)
        C, h, w = self.vae.config.latent_channels, H // 8, W // 8
        noise = torch.randn((F, C, h, w), generator=generator, device=self.device, dtype=torch.float16)
        Fl = F // xr.world
        mine = slice(xr.rank * Fl, (xr.rank + 1) * Fl)
        lat = self.vae.encode_frames_u8(frames_u8[mine].contiguous(), noise=noise[mine])          # (1, C, Fl, h, w)
        key = ("vae_latents", C, F, h * w)
        full, ptrs = xr.buffer(key, (C, F * h * w))
        mc = xr.multicast(key)
        off = xr.rank * Fl * h * w * 2
        ops.xrank_push(xr, [dict(src=lat.view(C, Fl * h * w), src_blk_rows=C, dst=[p + off for p in ptrs], ld_dst=F * h * w,
                                 dst_blk_rows=C, nblk=1, rows=C, mc=mc + off if mc else 0)])
        return full.view(1, C, F, h, w).clone()

    def decode_latents(self, latents, decode_chunk_size: int = 16):
        """stable_diffusion.py:369-394: 1 / 0.18215, decode ``decode_chunk_size`` frames at a time (the temporal decoder sees
        each chunk as one clip), (x / 2 + 0.5).clamp(0, 1) -> (b, f, H, W, 3) float32 on the host.  ``self.vae``:
        ``univst_b200.vae.AutoencoderKLTemporalDecoder`` or any object with diffusers' ``decode(z, num_frames=)``."""
        lat = (1 / 0.18215 * latents).permute(0, 2, 1, 3, 4).flatten(0, 1)
        chunks = []
        for k in range(0, lat.shape[0], decode_chunk_size):
            part = lat[k:k + decode_chunk_size].to(self.vae.dtype)
            chunks.append(self.vae.decode(part, num_frames=part.shape[0]).sample)
        video = torch.cat(chunks, dim=0)
        video = video.view(latents.shape[0], -1, *video.shape[1:]).permute(0, 1, 3, 4, 2)
        return ((video / 2 + 0.5).clamp(0, 1)).float().cpu()
