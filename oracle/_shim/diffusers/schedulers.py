"""Restatement of diffusers.DDIMScheduler (0.35.1) with the library's own constructor defaults
(linear betas 1e-4 .. 0.02, clip_sample=True, set_alpha_to_one=True, steps_offset=0).  SD-1.5's scheduler_config.json
values are NOT the defaults: callers that mirror ``DDIMScheduler.from_pretrained(.., subfolder="scheduler")`` pass
``**SD15_SCHEDULER_CONFIG`` explicitly; AnimateDiff builds ``DDIMScheduler(**noise_scheduler_kwargs)`` from its yaml
(animatediff-v2.yaml:16-21), so everything the yaml does not name -- set_alpha_to_one in particular -- stays default."""
from types import SimpleNamespace

import numpy as np
import torch


# scheduler/scheduler_config.json of runwayml/stable-diffusion-v1-5 (what DDIMScheduler.from_pretrained returns there)
SD15_SCHEDULER_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                             clip_sample=False, set_alpha_to_one=False, steps_offset=1)


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon",
                 timestep_spacing="leading", clip_sample_range=1.0):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                                      clip_sample=clip_sample, prediction_type=prediction_type,
                                      timestep_spacing=timestep_spacing, beta_schedule=beta_schedule,
                                      set_alpha_to_one=set_alpha_to_one, clip_sample_range=clip_sample_range)
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output, timestep, sample, eta=0.0, **kwargs):
        prev_t = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            x0 = x0.clamp(-self.config.clip_sample_range, self.config.clip_sample_range)
            model_output = (sample - a_t ** 0.5 * x0) / (1 - a_t) ** 0.5
        prev = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * model_output
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)


class _Unused:
    def __init__(self, *a, **k):
        raise NotImplementedError("only DDIMScheduler is on the UniVST path")


DPMSolverMultistepScheduler = EulerAncestralDiscreteScheduler = EulerDiscreteScheduler = _Unused
LMSDiscreteScheduler = PNDMScheduler = DDPMScheduler = _Unused


class _UnusedScheduler:  # names imported by backbones/animatediff/pipelines/pipeline_animation.py:23-30, never used
    def __init__(self, *a, **k):
        raise NotImplementedError("only DDIMScheduler is on the UniVST path")


DPMSolverMultistepScheduler = EulerAncestralDiscreteScheduler = EulerDiscreteScheduler = _UnusedScheduler
LMSDiscreteScheduler = PNDMScheduler = _UnusedScheduler
