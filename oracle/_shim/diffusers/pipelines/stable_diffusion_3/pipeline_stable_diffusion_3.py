"""TEST-ONLY restatement of what the reference's CustomStableDiffusion3Pipeline (custom_pipeline.py) inherits / imports
from diffusers 0.35.1's StableDiffusion3Pipeline [third party]: ``retrieve_timesteps`` (default branch: the scheduler's own
``set_timesteps`` and ``timesteps``), the ``interrupt`` property, ``progress_bar`` and ``prepare_latents`` returning the
latents it is given.  Everything model-related is supplied by the stand-in the goldens are generated with."""
import contextlib


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    if timesteps is not None or sigmas is not None:
        raise NotImplementedError("custom timesteps / sigmas are not used by the reference scripts")
    scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
    return scheduler.timesteps, num_inference_steps


class StableDiffusion3Pipeline:
    _interrupt = False

    @property
    def interrupt(self):
        return self._interrupt

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        class _Bar:
            def update(self, n=1):
                pass
        yield _Bar()

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        if latents is None:
            raise NotImplementedError("the reference always passes latents")
        return latents.to(device=device, dtype=dtype)

    def maybe_free_model_hooks(self):
        pass
