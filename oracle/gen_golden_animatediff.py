"""Generate ``tests/golden/animatediff_tiny.pt`` by running the REFERENCE's own AnimateDiff backbone:
``backbones/animatediff/models/unet.py`` ``UNet3DConditionModel`` constructed with the ``animatediff-v2.yaml`` kwargs
(on the test-only diffusers shim, ``oracle/_shim``) and ``backbones/animatediff/pnp_utils.py``.  Build container
only (needs ``/root/reference``).  Weights: ``animatediff_oracle.seeded_state_dict`` loaded with ``load_state_dict``
(strict) -- the motion modules' ``proj_out`` is non-zero, as after loading ``mm_sd_v15_v2.ckpt``.

    python oracle/gen_golden_animatediff.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

# backbones/animatediff/animatediff-v2.yaml:1-14
UNET_ADDITIONAL_KWARGS = dict(
    use_inflated_groupnorm=True, use_motion_module=True, motion_module_resolutions=[1, 2, 4, 8],
    motion_module_mid_block=True, motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1,
                              attention_block_types=["Temporal_Self", "Temporal_Self"],
                              temporal_position_encoding=True, temporal_attention_dim_div=1, zero_initialize=True))


def main():
    from backbones.animatediff import pnp_utils
    from backbones.animatediff.models.unet import UNet3DConditionModel
    from oracle import animatediff_oracle as ao

    cfg = ao.AD_TINY_CONFIG
    m = UNet3DConditionModel(block_out_channels=cfg["block_out_channels"], attention_head_dim=cfg["attention_head_dim"],
                             cross_attention_dim=cfg["cross_attention_dim"], sample_size=16, **UNET_ADDITIONAL_KWARGS).eval()
    ref_sd = m.state_dict()
    shapes = ao.unet_param_shapes(cfg)
    assert set(shapes) == set(ref_sd), (sorted(set(shapes) ^ set(ref_sd))[:10])
    assert all(tuple(ref_sd[k].shape) == shapes[k] for k in shapes)
    m.load_state_dict(ao.seeded_state_dict(cfg, seed=44))
    g = torch.Generator().manual_seed(2025)
    x = torch.randn(3, 4, 4, 16, 16, generator=g)
    ctx = torch.randn(1, 77, cfg["cross_attention_dim"], generator=g).repeat(3, 1, 1)
    out = {"x": x, "ctx": ctx, "seed": 44, "cases": {}}
    with torch.no_grad():
        out["cases"]["stock_t981"] = m(x, torch.tensor(981), encoder_hidden_states=ctx).sample.clone()
        pipe = types.SimpleNamespace(unet=m)
        pnp_utils.register_spatial_attention_pnp(pipe)
        for idx, t in ((0, 981), (13, 721), (24, 501), (25, 481)):   # 25: the window idx < eta2 * 50 has closed
            pnp_utils.register_time(pipe, idx)
            out["cases"][f"patched_idx{idx}_t{t}"] = m(x, torch.tensor(t), encoder_hidden_states=ctx).sample.clone()
    os.makedirs(OUT, exist_ok=True)
    torch.save(out, os.path.join(OUT, "animatediff_tiny.pt"))
    print("wrote animatediff_tiny.pt", os.path.getsize(os.path.join(OUT, "animatediff_tiny.pt")) // 1024, "KiB")
    gen_style_transfer()


def gen_style_transfer():
    """The reference's own ``AnimationPipeline.video_style_transfer`` (pipelines/pipeline_animation.py:449-603) driven
    with the tiny seeded UNet, on-disk trajectories / masks and the yaml's DDIM scheduler (linear betas)."""
    import tempfile

    import numpy as np
    from PIL import Image
    from backbones.animatediff import pnp_utils
    from backbones.animatediff.models.unet import UNet3DConditionModel
    from backbones.animatediff.pipelines.pipeline_animation import AnimationPipeline as P
    from diffusers import DDIMScheduler
    from oracle import animatediff_oracle as ao
    from oracle import pipeline_oracle as po

    cfg = ao.AD_TINY_CONFIG
    m = UNet3DConditionModel(block_out_channels=cfg["block_out_channels"], attention_head_dim=cfg["attention_head_dim"],
                             cross_attention_dim=cfg["cross_attention_dim"], sample_size=8, **UNET_ADDITIONAL_KWARGS).eval()
    m.load_state_dict(ao.seeded_state_dict(cfg, seed=44))
    F_, hw, n, seed = 16, 8, 50, 78   # 16 mask frames (src/util.py:133) and 50 steps are hard-coded in the reference
    traj_c, traj_s, mask_u8 = po.synthetic_inputs(seed, F_, hw, n)
    g = torch.Generator().manual_seed(seed + 1)
    emb = torch.randn(1, 77, cfg["cross_attention_dim"], generator=g)
    with tempfile.TemporaryDirectory() as tmp:
        cdir, sdir, mdir = (os.path.join(tmp, d) for d in ("c", "s", "m"))
        for d in (cdir, sdir, mdir):
            os.makedirs(d)
        for k in range(1, n + 1):
            torch.save(traj_c[k], os.path.join(cdir, f"ddim_latents_{k}.pt"))
            torch.save(traj_s[k], os.path.join(sdir, f"ddim_latents_{k}.pt"))
        for f in range(F_):
            Image.fromarray(mask_u8[f], mode="L").save(os.path.join(mdir, "%05d.png" % f))
        pipe = P.__new__(P)
        pipe.unet = m
        # animatediff-v2.yaml:16-21 (run_video_style_transfer_animatediff.py builds DDIMScheduler(**noise_scheduler_kwargs))
        # nothing else is named there, so set_alpha_to_one keeps the library default True: the last step returns x0
        pipe.scheduler = DDIMScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="linear", steps_offset=1,
                                       clip_sample=False)
        assert float(pipe.scheduler.final_alpha_cumprod) == 1.0
        pipe._encode_prompt = lambda *a, **k: emb
        final = []
        pipe.decode_latents = lambda lat: (final.append(lat.clone()), np.zeros((1, 3, 1, 1, 1), np.float32))[1]
        pnp_utils.register_spatial_attention_pnp(pipe)
        z_T = pnp_utils.latent_adain(traj_c[n], traj_s[n])
        rec = {}
        with torch.no_grad():
            pipe.video_style_transfer("", latents=z_T, num_inference_steps=n, content_inv_path=cdir, style_inv_path=sdir,
                                      mask_path=mdir, callback=lambda i, t, lat: rec.__setitem__(i, lat.clone()))
    keep = {i: rec[i] for i in (0, 24, 25, 40, 46, 49)}
    torch.save({"seed": seed, "F": F_, "hw": hw, "n": n, "emb": emb, "z_T": z_T, "final": final[0], "steps": keep},
               os.path.join(OUT, "style_transfer_animatediff_tiny.pt"))
    print("wrote style_transfer_animatediff_tiny.pt")


if __name__ == "__main__":
    main()
