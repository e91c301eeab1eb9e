"""Mirror of the SD3 sampling loops, ``CustomStableDiffusion3Pipeline`` of
backbones/video_diffusion_sd3/pipelines/custom_pipeline.py: ``generate_eta_values`` (:18-44), ``reconstruction`` (:46-124)
and ``video_style_transfer`` (:126-371).  Same method names, arguments and outputs; the third-party members -- MMDiT
``transformer`` (steered through the processors of :mod:`univst_b200.sd3`), ``scheduler``
(FlowMatchEulerDiscreteScheduler), ``encode_prompt``, ``vae`` / ``image_processor`` -- are taken from the diffusers
pipeline object this class wraps (``host``), exactly the members the reference's subclass inherits.

Per step the reference runs ~12 elementwise torch kernels on the (F, 16, h, w) latents; here:

* mask blend                       -> ``univst_latent_blend_fc_f16`` (mask resized once per clip, not 1-2x per step);
* late latent AdaIN (per-plane)     -> ``univst_latent_adain_f16`` with F * C one-frame channels;
* velocity interpolation towards the clean content latents + Euler step, folded:
  ``z + ds (v + eta (-(x0 - z) / t - v)) = (1 + a) z - a x0 + ds (1 - eta) v`` with ``a = ds eta / t``
                                   -> two ``univst_axpby_f16`` (one when eta = 0), fp32 inside, fp16 latents.

Reference defect: ``video_style_transfer`` reads the undefined name ``ddim_inv_latents_at_t`` at :316 (every 50-step run
reaches it at i = 40).  The content inversion latent of the step is blended there, as in the SD and AnimateDiff loops
(stable_diffusion.py:704) -- the same choice the golden generator and the oracle make.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch

from . import ops
from .util import load_ddim_latents_at_t, load_mask


def calculate_shift(image_seq_len, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.15):
    """custom_pipeline.py:375-386."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    return image_seq_len * m + (base_shift - m * base_seq_len)


class CustomStableDiffusion3Pipeline:
    def __init__(self, host):
        self.host = host
        self.transformer, self.scheduler = host.transformer, host.scheduler

    @property
    def device(self):
        return torch.device(getattr(self.host, "_execution_device", None) or getattr(self.host, "device", "cuda"))

    def __getattr__(self, name):
        # everything else the reference's subclass inherits (encode_prompt, vae, image_processor, ...) lives on the host
        if name == "host":
            raise AttributeError(name)
        return getattr(self.host, name)

    # ------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def generate_eta_values(timesteps, start_step, end_step, eta, eta_trend):
        """custom_pipeline.py:18-44."""
        assert start_step < end_step and start_step >= 0 and end_step <= len(timesteps), "Invalid start_step and end_step"
        ts = [float(t) for t in timesteps]
        eta_values = [0.0] * len(ts)
        if eta_trend == "constant":
            for i in range(start_step, end_step):
                eta_values[i] = eta
        elif eta_trend == "linear_increase":
            total_time = ts[start_step] - ts[end_step - 1]
            for i in range(start_step, end_step):
                eta_values[i] = eta * (ts[start_step] - ts[i]) / total_time
        elif eta_trend == "linear_decrease":
            total_time = ts[start_step] - ts[end_step - 1]
            for i in range(start_step, end_step):
                eta_values[i] = eta * (ts[i] - ts[end_step - 1]) / total_time
        else:
            raise NotImplementedError(f"Unsupported eta_trend: {eta_trend}")
        return eta_values

    def _timesteps(self, n, **kw):
        """retrieve_timesteps (diffusers) with the default spacing: the scheduler's own timesteps and sigmas."""
        self.scheduler.set_timesteps(n, device=self.device, **kw)
        ts = [float(t) for t in self.scheduler.timesteps]
        sig = [float(s) for s in self.scheduler.sigmas]
        if len(sig) != len(ts) + 1:
            raise ValueError("scheduler.sigmas must hold one more entry than scheduler.timesteps (FlowMatchEuler layout)")
        return ts, sig

    def _trajectory(self, src, k):
        z = src[k] if isinstance(src, (list, tuple, dict)) else load_ddim_latents_at_t(k, src)
        return z.to(self.device, torch.float16).contiguous()

    def _mask(self, mask, F, h, w):
        if mask is None or (isinstance(mask, str) and not mask):
            return None
        m = load_mask(mask, n_frames=16) if isinstance(mask, str) else mask   # src/util.py:133: always 16 frames
        m = m.reshape(-1, m.shape[-2], m.shape[-1])
        if m.shape[0] != F:
            raise ValueError(f"{m.shape[0]} mask frames for {F} latent frames")
        return ops.mask_resize((m != 0).to(torch.uint8).to(self.device).contiguous(), h, w)

    def _velocity(self, x, t, embeds, pooled, **kw):
        t_vec = torch.full((x.shape[0],), t, dtype=torch.float32, device=x.device)   # `t.expand(B)` of the scheduler's fp32 timesteps
        v = self.transformer(hidden_states=x, timestep=t_vec, encoder_hidden_states=embeds, pooled_projections=pooled,
                             return_dict=False, **kw)[0]
        return v.to(torch.float16).contiguous()

    def _step(self, z, v, target, eta, t, ds):
        """Interpolated velocity (:336-341 / :106-113) + FlowMatchEuler step (x + (sigma_next - sigma) v)."""
        if eta != 0.0:
            a = ds * eta / (t / float(self.scheduler.config.num_train_timesteps))
            z = ops.axpby(z, target, 1.0 + a, -a)
        return ops.axpby(z, v, 1.0, ds * (1.0 - eta))

    def _decode(self, latents, output_type):
        vae = self.host.vae
        latents = (latents / vae.config.scaling_factor) + vae.config.shift_factor
        image = vae.decode(latents, return_dict=False)[0]
        return self.host.image_processor.postprocess(image, output_type=output_type)

    # ------------------------------------------------------------------------------------------ loops
    @torch.no_grad()
    def reconstruction(self, img_latents, inversed_latents, eta_base, eta_trend, start_step, end_step, guidance_scale=1.0,
                       prompt="", DTYPE=torch.float16, num_inference_steps=50, output_type="pil"):
        """custom_pipeline.py:46-124.  ``output_type="latent"`` returns the latents instead of decoded images."""
        ts, sig = self._timesteps(num_inference_steps)
        embeds, neg, pooled, neg_pooled = self.host.encode_prompt(prompt=prompt, prompt_2=prompt, prompt_3=prompt)
        z = inversed_latents.to(self.device, torch.float16).contiguous()
        target = img_latents.to(self.device, torch.float16).contiguous()
        etas = self.generate_eta_values(ts, start_step, end_step, eta_base, eta_trend)
        cfg = guidance_scale > 1.0
        if cfg:
            embeds, pooled = torch.cat([neg, embeds], dim=0), torch.cat([neg_pooled, pooled], dim=0)
        for i, t in enumerate(ts):
            v = self._velocity(torch.cat([z] * 2) if cfg else z, t, embeds, pooled)
            if cfg:
                vu, vt = v.chunk(2)
                v = ops.axpby(vu.contiguous(), vt.contiguous(), 1.0 - guidance_scale, guidance_scale)
            z = self._step(z, v, target, etas[i], t, sig[i + 1] - sig[i])
        return z if output_type == "latent" else self._decode(z, output_type)

    @torch.no_grad()
    def video_style_transfer(self, prompt=None, prompt_2=None, prompt_3=None, height=None, width=None,
                             num_inference_steps: int = 50, latents=None, prompt_embeds=None, pooled_prompt_embeds=None,
                             output_type: Optional[str] = "pil", return_dict: bool = True, callback_on_step_end=None,
                             max_sequence_length: int = 256, mu=None, content_inv_path=None, style_inv_path=None,
                             mask_path=None, eta_base=0.95, eta_trend="constant", start_step=10, end_step=20,
                             img_latents=None, **kwargs):
        """custom_pipeline.py:126-371.  ``content_inv_path`` / ``style_inv_path``: directory of ``ddim_latents_{k}.pt``
        (reference format) or an in-memory mapping k -> (F, C, h, w); ``mask_path``: directory of ``%05d.png`` or a
        (F, H, W) tensor (non-zero = keep content)."""
        n = num_inference_steps
        if prompt_embeds is None:
            prompt_embeds, _, pooled_prompt_embeds, _ = self.host.encode_prompt(
                prompt=prompt, prompt_2=prompt_2, prompt_3=prompt_3, do_classifier_free_guidance=False, device=self.device,
                num_images_per_prompt=1, max_sequence_length=max_sequence_length)
        z = latents.to(self.device, torch.float16).contiguous().clone()
        F, C, h, w = z.shape
        embeds = prompt_embeds.repeat(3 * F, 1, 1)                     # :228-229
        pooled = pooled_prompt_embeds.repeat(3 * F, 1)
        kw = {}
        cfg_s = self.scheduler.config
        if cfg_s.get("use_dynamic_shifting", None) and mu is None:     # :246-258
            ps = self.transformer.config.patch_size
            mu = calculate_shift((h // ps) * (w // ps), cfg_s.get("base_image_seq_len", 256),
                                 cfg_s.get("max_image_seq_len", 4096), cfg_s.get("base_shift", 0.5), cfg_s.get("max_shift", 1.16))
        if mu is not None:
            kw["mu"] = mu
        ts, sig = self._timesteps(n, **kw)
        target = img_latents.to(self.device, torch.float16).contiguous()
        etas = self.generate_eta_values(ts, start_step, end_step, eta_base, eta_trend)
        m = self._mask(mask_path, F, h, w)                              # constant over the loop: resized once
        for i, t in enumerate(ts):
            zc = self._trajectory(content_inv_path, 50 - i)            # :284-285 (hard-coded 50)
            zs = self._trajectory(style_inv_path, 50 - i)
            if m is not None and i <= 0.9 * n:                          # localized latent blending, :287-293
                z = ops.latent_blend_fc(z, zc, m)
            if i >= 0.8 * n and i <= 0.9 * n:                           # :295-304
                z = ops.plane_adain(z, zs)
                if m is not None:
                    z = ops.latent_blend_fc(z, zc, m)                   # the undefined name of :316, see the module docstring
            v = self._velocity(torch.cat([zc, zs, z]), t, embeds, pooled, joint_attention_kwargs={"idx": i})
            z = self._step(z, v[2 * F:], target, etas[i], t, sig[i + 1] - sig[i])
            if callback_on_step_end is not None:
                out = callback_on_step_end(self, i, t, {"latents": z}) or {}
                z = out.pop("latents", z)
        image = z if output_type == "latent" else self._decode(z, output_type)
        return SimpleNamespace(images=image) if return_dict else (image,)
