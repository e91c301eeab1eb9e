"""univst_b200 -- B200 (sm_100a) kernels + host code for UniVST's three-branch DDIM denoising hot path.

Layout: ``csrc/`` hand-written CUDA behind the C ABI of ``include/univst_b200.h``; ``ops.py`` ctypes wrappers;
``unet.py`` / ``pnp_utils.py`` / ``pipeline.py`` / ``ddim_inversion.py`` / ``mask_propagation.py`` /
``flow_warp.py`` mirror the reference's Python interface for the path (same names and argument meaning).
"""
__version__ = "0.1.0"


def accelerate(pipe, device="cuda"):
    """Swap the UNet of a REFERENCE pipeline object for the B200 one, in place, and return the pipeline.

    ``pipe`` is the reference's ``SpatioTemporalStableDiffusionPipeline`` (or anything with a ``.unet`` holding a
    reference ``UNetPseudo3DConditionModel`` / AnimateDiff ``UNet3DConditionModel`` nn.Module).  Afterwards the reference's
    own scripts run unmodified: its ``register_spatial_attention_pnp(pipe)`` / ``register_time(pipe, i)`` find the
    ``unet.up_blocks[r].attentions[b].transformer_blocks[0].attn1`` handles they patch, and its own
    ``video_style_transfer`` / ``ddim_inversion`` loops call ``pipe.unet(sample, t, encoder_hidden_states=...)`` and read
    ``.sample`` (stable_diffusion.py:597,710; ddim_inversion.py:209).  Weights are copied and packed once.  For the loops
    themselves on the GPU kernels as well, use ``univst_b200.pipeline`` / ``univst_b200.animatediff`` instead."""
    ref = pipe.unet
    keys = ref.state_dict().keys()
    if any("motion_modules" in k for k in keys):
        from .animatediff import UNet3DConditionModel
        new = UNet3DConditionModel(ref.state_dict(), device=device)
    else:
        from .unet import UNetPseudo3DConditionModel
        new = UNetPseudo3DConditionModel.from_reference(ref, device=device)
    pipe.unet = new
    return pipe
