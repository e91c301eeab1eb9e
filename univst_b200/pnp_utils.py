"""Mirror of the reference's attention-patch protocol (backbones/video_diffusion_sd/pnp_utils.py) for the B200 UNet.

``register_time`` / ``register_spatial_attention_pnp`` have the reference's names, arguments and effects on
``pipe.unet.up_blocks[res].attentions[block].transformer_blocks[0].attn1`` -- they only set attributes; the patched
arithmetic (AdaIN-guided Q/K/V shift + [prev, first] K/V gather) runs inside the fused kernels (unet.py).  The
reference's own two functions work unchanged on the B200 UNet as well: an instance-level ``forward`` override on the
attn1 handle is what marks a layer as patched.
"""
from __future__ import annotations

from . import ops

UP_RES_DICT = {1: [1, 2], 2: [0, 1, 2], 3: [0, 1, 2]}  # pnp_utils.py:8, :104


def register_time(model, t):
    """pnp_utils.py:7-15."""
    for res, blocks in UP_RES_DICT.items():
        for block in blocks:
            tb = model.unet.up_blocks[res].attentions[block].transformer_blocks[0]
            setattr(tb.attn1, "idx", t)
            setattr(tb.attn2, "idx", t)


def register_spatial_attention_pnp(model, eta1=0.0, eta2=0.5):
    """pnp_utils.py:18-111: mark the eight decoder attn1 layers as patched and store the shift window."""
    for res, blocks in UP_RES_DICT.items():
        for block in blocks:
            attn1 = model.unet.up_blocks[res].attentions[block].transformer_blocks[0].attn1
            setattr(attn1, "eta1", eta1)
            setattr(attn1, "eta2", eta2)
            setattr(attn1, "_patched", True)


def latent_adain(cnt_feat, sty_feat, ad=True):
    """pnp_utils.py:128-139 on (1, C, F, h, w) fp16 CUDA latents."""
    return ops.latent_adain(cnt_feat.contiguous(), sty_feat.contiguous())
