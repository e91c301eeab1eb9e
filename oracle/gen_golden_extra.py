"""Golden vectors for the loops around the UNet, produced by the REFERENCE's own ``video_style_transfer``
(stable_diffusion.py:631-780) and ``ddim_loop`` / ``ddim_loop_plus`` (ddim_inversion.py:88-167), imported on the
test-only shim and driven with the tiny seeded UNet.  Called from oracle/gen_golden.py (build container only)."""
import os
import tempfile
import types

import numpy as np
import torch


def _tiny_reference_unet():
    from backbones.video_diffusion_sd.models.unet_3d_condition import UNetPseudo3DConditionModel
    from oracle import unet_oracle as uo
    cfg = uo.TINY_CONFIG
    m = UNetPseudo3DConditionModel(block_out_channels=cfg["block_out_channels"], attention_head_dim=cfg["attention_head_dim"],
                                   cross_attention_dim=cfg["cross_attention_dim"], sample_size=8).eval()
    m.load_state_dict(uo.seeded_state_dict(cfg, seed=33))
    return m, cfg


def gen_style_transfer(save):
    from PIL import Image
    from backbones.video_diffusion_sd import pnp_utils
    from backbones.video_diffusion_sd.pipelines.stable_diffusion import SpatioTemporalStableDiffusionPipeline as P
    from diffusers import DDIMScheduler
    from diffusers.schedulers import SD15_SCHEDULER_CONFIG
    from oracle import pipeline_oracle as po

    m, cfg = _tiny_reference_unet()
    F_, hw, n, seed = 16, 8, 50, 77  # the reference hard-codes 16 mask frames (src/util.py:133) and 50 steps (eta2 * 50)
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
        pipe.unet, pipe.scheduler = m, DDIMScheduler(**SD15_SCHEDULER_CONFIG)   # from_pretrained(.., subfolder='scheduler')
        pipe._encode_prompt = lambda *a, **k: emb
        final = []
        pipe.decode_latents = lambda lat: (final.append(lat.clone()), np.zeros((1, 1, 1, 1, 3), np.float32))[1]
        pnp_utils.register_spatial_attention_pnp(pipe)
        z_T = pnp_utils.latent_adain(traj_c[n], traj_s[n])
        rec = {}
        with torch.no_grad():
            pipe.video_style_transfer("", latents=z_T, num_inference_steps=n, content_inv_path=cdir, style_inv_path=sdir,
                                      mask_path=mdir, callback=lambda i, t, lat: rec.__setitem__(i, lat.clone()))
    keep = {i: rec[i] for i in (0, 25, 26, 41, 46, 49)}
    save("style_transfer_tiny.pt", {"seed": seed, "F": F_, "hw": hw, "n": n, "emb": emb, "z_T": z_T, "final": final[0],
                                    "steps": keep})


def gen_inversion(save):
    import inversion_tools.ddim_inversion as di
    from diffusers import DDIMScheduler
    from diffusers.schedulers import SD15_SCHEDULER_CONFIG
    from oracle import pipeline_oracle as po

    m, cfg = _tiny_reference_unet()
    F_, hw, n, seed = 4, 8, 10, 5  # BASELINE.json configs[0] scaled down: 4 frames, 10 steps, CPU
    traj_c, _, _ = po.synthetic_inputs(seed, F_, hw, 50)
    g = torch.Generator().manual_seed(seed + 1)
    emb = torch.randn(1, 77, cfg["cross_attention_dim"], generator=g)
    di.init_prompt = lambda pipeline, prompt: torch.cat([emb, emb])  # CLIP is a third-party network: fixed embeddings
    sch = DDIMScheduler(**SD15_SCHEDULER_CONFIG)
    sch.set_timesteps(n)
    pipe = types.SimpleNamespace(unet=m)
    out = {"seed": seed, "F": F_, "hw": hw, "n": n, "emb": emb}
    with tempfile.TemporaryDirectory() as tmp:
        with torch.no_grad():
            lat = di.ddim_loop(pipe, sch, traj_c[0], n, "", tmp, ft_indices=[2], ft_timesteps=[301], ft_path=tmp)
            out["files"] = sorted(os.listdir(tmp))
            out["feature"] = torch.load(os.path.join(tmp, "inversion_feature_map_2_block_301_step.pt"))
            lat_plus = di.ddim_loop_plus(pipe, sch, traj_c[0], n, "", None)
    out["ddim_loop"] = torch.stack(lat[1:])
    out["ddim_loop_plus"] = torch.stack(lat_plus[1:])
    save("ddim_inversion_tiny.pt", out)


def gen_maskprop(save):
    """Reference mask_propogation (src/mask_propagation.py:72-99) on synthetic features, CPU fp32."""
    import src.mask_propagation as mp
    from oracle import maskprop_oracle as mo
    out = {}
    for name, sep, h in (("smooth", False, 16), ("separated", True, 16)):
        feats = mo.synthetic_features(11, 3, h, h, 64, separated=sep)
        feat_src = torch.cat([feats[0].reshape(h * h, -1).T, feats[1].reshape(h * h, -1).T[:, ::3]], dim=-1).contiguous()
        feat_tar = feats[2].reshape(h * h, -1).contiguous()
        g = torch.Generator().manual_seed(3)
        labels = (torch.rand(feat_src.shape[1], generator=g) > 0.6).long()
        segs = torch.stack([(labels == 0).float(), (labels == 1).float()])
        args = types.SimpleNamespace(temperature=0.2, topk=15, sample_ratio=0.3)
        torch.manual_seed(0)
        segs_tar, feat_s, segs_s = mp.mask_propogation(feat_src, feat_tar, segs, args)
        out[name] = {"segs_tar": segs_tar, "n_sample": feat_s.shape[1], "seed": 11, "h": h, "C": 64, "sep": sep,
                     "feat_sample": feat_s, "segs_sample": segs_s}
    save("maskprop.pt", out)


def gen_flow_warp(save):
    """Reference warp helpers (src/cal_optica_flow.py:20-46, cv2.remap inside) on real example frames + analytic flows."""
    import cv2
    import src.cal_optica_flow as cf
    from oracle import flowwarp_oracle as fo
    frames = np.stack([cv2.cvtColor(cv2.imread(f"/root/reference/examples/contents/mallard-fly/{i:05d}.png"), cv2.COLOR_BGR2RGB)
                       [128:256, 192:320] for i in range(4)])
    fwd, bwd = fo.synthetic_flow(128, 128, 0), fo.synthetic_flow(128, 128, 1, backward_of=fo.synthetic_flow(128, 128, 0))
    occ = cf.compute_occlusion_mask(fwd, bwd, threshold=1.5)
    warped = cf.warp_image_with_flow(frames[1], fwd)
    masked = cf.apply_mask(warped, occ, frames[0])
    save("flow_warp.pt", {"frames": torch.from_numpy(frames), "occ": torch.from_numpy(occ), "warped": torch.from_numpy(warped),
                          "masked": torch.from_numpy(masked)})


def main(save):
    gen_style_transfer(save)
    gen_inversion(save)
    gen_maskprop(save)
    gen_flow_warp(save)
