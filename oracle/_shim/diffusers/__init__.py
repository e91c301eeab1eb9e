"""TEST-ONLY stand-in for the parts of diffusers 0.35.1 that the reference's model files import.

diffusers is not installed in this image and there is no network.  This package restates (from the published
semantics listed in SURVEY.md Appendix B) just enough of ``Attention``, ``FeedForward``/GEGLU, ``Timesteps``,
``TimestepEmbedding``, ``ConfigMixin``/``ModelMixin``/``BaseOutput`` and ``DDIMScheduler`` for
``/root/reference/backbones/video_diffusion_sd/models/*.py`` to import and run on CPU, so that golden vectors can
be generated from the reference's OWN module code (oracle/gen_golden.py).  It is never imported by the product.
"""
from .modeling_utils import ModelMixin  # noqa: F401
from .schedulers import DDIMScheduler  # noqa: F401
from .pipelines.stable_diffusion_3.pipeline_stable_diffusion_3 import StableDiffusion3Pipeline  # noqa: F401,E402
