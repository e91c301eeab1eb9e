from ..modeling_utils import ModelMixin  # noqa: F401  (diffusers >= 0.20 location, used by backbones/animatediff)
