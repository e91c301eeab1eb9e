"""Host-side DDIM scheduler scalars, mirroring diffusers' ``DDIMScheduler`` (third-party to the reference: called at
stable_diffusion.py:670,761; semantics in SURVEY.md Appendix B).  The constructor has the library's own defaults;
``DDIMScheduler.sd15()`` is SD-1.5's scheduler_config.json (what ``from_pretrained(.., subfolder="scheduler")`` gives the
SD scripts, run_video_style_transfer_sd.py:45) and ``DDIMScheduler.animatediff_v2()`` the ``noise_scheduler_kwargs`` of
animatediff-v2.yaml:16-21 (which leave ``set_alpha_to_one`` at the library default True: the last DDIM step returns x0).
All per-step coefficients are Python floats, so stepping never synchronises with the device; the arithmetic on the
latents is the ``univst_ddim_step_f16`` kernel."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 timestep_spacing="leading"):
        if prediction_type != "epsilon" or timestep_spacing != "leading":
            raise NotImplementedError("only the epsilon / leading configuration of the reference is mirrored")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                                      clip_sample=clip_sample, prediction_type=prediction_type,
                                      timestep_spacing=timestep_spacing, beta_schedule=beta_schedule,
                                      set_alpha_to_one=set_alpha_to_one)
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self._alphas = [float(a) for a in self.alphas_cumprod]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    @classmethod
    def sd15(cls, **overrides):
        """scheduler/scheduler_config.json of SD-1.5 (and SD-2.1-base: same values)."""
        return cls(**dict(dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
                               set_alpha_to_one=False, steps_offset=1), **overrides))

    @classmethod
    def animatediff_v2(cls, **overrides):
        """DDIMScheduler(**noise_scheduler_kwargs) of backbones/animatediff/animatediff-v2.yaml:16-21."""
        return cls(**dict(dict(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", steps_offset=1,
                               clip_sample=False), **overrides))

    def patch_for_pipeline(self):
        """What the reference pipelines' constructors do to an outdated config (stable_diffusion.py:68-93,
        pipeline_animation.py:71-95): steps_offset := 1, clip_sample := False."""
        self.config.steps_offset = 1
        self.config.clip_sample = False
        return self

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self.config.steps_offset
        self.timesteps = torch.from_numpy(ts)  # kept on the host: indexing it must not touch the device

    def scale_model_input(self, sample, timestep=None):
        return sample

    def alpha(self, t: int) -> float:
        return self._alphas[t] if t >= 0 else float(self.final_alpha_cumprod)

    def step_alphas(self, t: int):
        """(alpha_t, alpha_prev) of DDIMScheduler.step (eta = 0; x0 is not clipped: the pipelines force clip_sample off)."""
        if self.config.clip_sample:
            raise NotImplementedError("clip_sample=True: build the scheduler through a pipeline (which patches it to False "
                                      "like the reference's constructors) or pass clip_sample=False")
        return self.alpha(int(t)), self.alpha(int(t) - self.config.num_train_timesteps // self.num_inference_steps)

    def inversion_alphas(self, t: int):
        """(alpha_cur, alpha_next) of next_step (inversion_tools/ddim_inversion.py:190-196)."""
        cur = min(int(t) - self.config.num_train_timesteps // self.num_inference_steps, 999)
        return self.alpha(cur), self.alpha(int(t))


class FlowMatchEulerDiscreteScheduler:
    """Host-side scalars of diffusers' ``FlowMatchEulerDiscreteScheduler`` as SD3 / SD3.5 configure it (``shift = 3.0``, no
    dynamic shifting) -- third-party to the reference, which only reads ``timesteps``, ``sigmas`` and ``config``
    (custom_pipeline.py:246-262, :336-345; inversion_tools/flow_inversion.py:123-264).  PARITY UNPINNED (restated from the
    published algorithm, no library here to check against): ``sigma_k = shift s / (1 + (shift - 1) s)`` for ``s`` linearly
    spaced between the shifted extremes, ``timesteps = 1000 sigma``, a trailing ``sigma = 0``."""
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 3.0, use_dynamic_shifting: bool = False, **extra):
        self.config = dict(num_train_timesteps=num_train_timesteps, shift=shift, use_dynamic_shifting=use_dynamic_shifting, **extra)
        self.config = type("Cfg", (dict,), {"__getattr__": lambda s_, k: s_[k]})(self.config)
        s = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float64)[::-1] / num_train_timesteps
        s = shift * s / (1 + (shift - 1) * s)
        self.sigma_max, self.sigma_min = float(s[0]), float(s[-1])
        self.timesteps = torch.from_numpy((s * num_train_timesteps).astype(np.float32))
        self.sigmas = torch.from_numpy(np.append(s, 0.0).astype(np.float32))

    def set_timesteps(self, num_inference_steps: int, device=None, mu=None, **kwargs):
        T, shift = self.config["num_train_timesteps"], self.config["shift"]
        ts = np.linspace(self.sigma_max * T, self.sigma_min * T, num_inference_steps, dtype=np.float64)
        s = ts / T
        if self.config["use_dynamic_shifting"]:
            if mu is None:
                raise ValueError("use_dynamic_shifting needs mu")
            s = np.exp(mu) / (np.exp(mu) + (1 / s - 1))
        else:
            s = shift * s / (1 + (shift - 1) * s)
        self.timesteps = torch.from_numpy((s * T).astype(np.float32))          # kept on the host: no device sync when read
        self.sigmas = torch.from_numpy(np.append(s, 0.0).astype(np.float32))
