"""GPU parity of the loops around the UNet (video_style_transfer, ddim_inversion) through the product's mirror of
the reference interface, against golden latents produced by the REFERENCE's own pipeline code (oracle/gen_golden*.py).

Tolerance: the product computes in fp16 (like the reference on GPU), the goldens are fp32; after 50 DDIM steps we
require a relative L2 error <= 3e-2 on the final latents (PSNR >= 30 dB w.r.t. the latent range); the 10-step
inversion must stay within 1e-2.
"""
import os

import numpy as np
import pytest
import torch

from oracle import pipeline_oracle as po
from oracle import unet_oracle as uo

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def pipe(cuda_lib):
    from univst_b200.pipeline import SpatioTemporalStableDiffusionPipeline
    from univst_b200.unet import UNetPseudo3DConditionModel
    sd = uo.seeded_state_dict(uo.TINY_CONFIG, seed=33)
    return SpatioTemporalStableDiffusionPipeline(UNetPseudo3DConditionModel(sd, uo.TINY_CONFIG))


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def _psnr(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return (10 * torch.log10((b.max() - b.min()) ** 2 / ((a - b) ** 2).mean())).item()


def test_video_style_transfer_matches_reference_golden(pipe, tmp_path):
    from PIL import Image
    from univst_b200 import pnp_utils
    g = torch.load(os.path.join(GOLDEN, "style_transfer_tiny.pt"), weights_only=True)
    n = g["n"]
    traj_c, traj_s, mask_u8 = po.synthetic_inputs(g["seed"], g["F"], g["hw"], n)
    # reference on-disk formats: ddim_latents_{k}.pt (fp16) and %05d.png masks
    cdir, sdir, mdir = (tmp_path / d for d in ("c", "s", "m"))
    for d in (cdir, sdir, mdir):
        d.mkdir()
    for k in range(1, n + 1):
        torch.save(traj_c[k].half(), cdir / f"ddim_latents_{k}.pt")
        torch.save(traj_s[k].half(), sdir / f"ddim_latents_{k}.pt")
    for f in range(g["F"]):
        Image.fromarray(mask_u8[f], mode="L").save(mdir / ("%05d.png" % f))
    pnp_utils.register_spatial_attention_pnp(pipe)
    z_T = pnp_utils.latent_adain(traj_c[n].cuda().half(), traj_s[n].cuda().half())
    assert _rel(z_T, g["z_T"]) < 2e-3
    rec = {}
    out = pipe.video_style_transfer("", num_inference_steps=n, latents=z_T, content_inv_path=str(cdir),
                                    style_inv_path=str(sdir), mask_path=str(mdir), prompt_embeds=g["emb"],
                                    callback=lambda i, t, z: rec.__setitem__(i, z.clone()))
    for i, ref in g["steps"].items():
        print(f"step {i}: rel={_rel(rec[i], ref):.3e}")
    rel, psnr = _rel(out.latents, g["final"]), _psnr(out.latents, g["final"])
    print(f"final latents after {n} steps: rel={rel:.3e} psnr={psnr:.1f} dB")
    assert torch.isfinite(out.latents).all() and rel <= 3e-2 and psnr >= 30.0

    # exact dead-branch skipping: identical edit latents, fewer launches
    from univst_b200 import ops
    lists = dict(content_inv_path=[t.half() for t in traj_c], style_inv_path=[t.half() for t in traj_s],
                 mask_path=torch.from_numpy(mask_u8), prompt_embeds=g["emb"])
    full = pipe.video_style_transfer("", num_inference_steps=n, latents=z_T, skip_dead_branches=False, **lists).latents
    batches = []
    fwd = pipe.unet.forward
    pipe.unet.forward = lambda x, *a, **k: (batches.append(x.shape[0]), fwd(x, *a, **k))[1]
    skip = pipe.video_style_transfer("", num_inference_steps=n, latents=z_T, skip_dead_branches=True, **lists).latents
    del pipe.unet.forward
    assert torch.equal(skip, out.latents), "in-memory trajectories must give the same result as the on-disk format"
    assert torch.equal(skip, full), "skipping the dead content/style branches must not change the edit branch"
    assert batches == [3] * 26 + [1] * 24  # shift window idx 0..25 (pnp_utils.py:47), edit branch only afterwards


def test_ddim_inversion_matches_reference_golden(pipe, tmp_path):
    from univst_b200 import ddim_inversion as di
    from univst_b200.scheduler import DDIMScheduler
    g = torch.load(os.path.join(GOLDEN, "ddim_inversion_tiny.pt"), weights_only=True)
    traj_c, _, _ = po.synthetic_inputs(g["seed"], g["F"], g["hw"], 50)
    for tr in pipe.unet._all_transformers():  # inversion runs the stock attention
        tr.transformer_blocks[0].attn1.__dict__.pop("_patched", None)
    sch = DDIMScheduler.sd15()
    sch.set_timesteps(g["n"])
    lat = di.ddim_inversion(pipe, sch, traj_c[0], g["n"], "", inversion_path=str(tmp_path), ft_indices=[2],
                            ft_timesteps=[301], ft_path=str(tmp_path), prompt_embeds=g["emb"])
    assert sorted(os.listdir(tmp_path)) == g["files"]
    rel = _rel(torch.stack(lat[1:]), g["ddim_loop"])
    feat = torch.load(tmp_path / "inversion_feature_map_2_block_301_step.pt", weights_only=True)
    saved = torch.load(tmp_path / "ddim_latents_10.pt", weights_only=True)
    assert torch.equal(saved, lat[10]) and saved.dtype == torch.float16 and tuple(saved.shape) == (1, 4, g["F"], g["hw"], g["hw"])
    rel_f = _rel(feat, g["feature"])
    lat_p = di.ddim_inversion(pipe, sch, traj_c[0], g["n"], "", is_opt=True, prompt_embeds=g["emb"])
    rel_p = _rel(torch.stack(lat_p[1:]), g["ddim_loop_plus"])
    print(f"ddim_loop rel={rel:.3e}  feature rel={rel_f:.3e}  ddim_loop_plus rel={rel_p:.3e}")
    assert rel <= 1e-2 and rel_p <= 1e-2 and rel_f <= 1e-2 and tuple(feat.shape) == tuple(g["feature"].shape)


def test_video_style_transfer_non_square_clip(pipe):
    """A 3-frame 192 x 320 clip (latent 24 x 40: no power-of-two side anywhere in the UNet) through the 10-step loop,
    against the oracle loop + oracle UNet evaluated in fp32 on the same inputs."""
    from univst_b200 import pnp_utils
    F_, h, w, n = 3, 24, 40, 10
    gen = torch.Generator().manual_seed(77)
    sch = po.DDIMOracle()
    sch.set_timesteps(n)
    z0_c = torch.randn(1, 4, F_, h, w, generator=gen)
    z0_s = torch.randn(1, 4, 1, h, w, generator=gen).repeat(1, 1, F_, 1, 1) + 0.02 * torch.randn(1, 4, F_, h, w, generator=gen)
    eps = torch.randn(1, 4, F_, h, w, generator=gen)
    traj_c, traj_s = [z0_c], [z0_s]
    for t in sch.timesteps[::-1]:
        a = sch.alpha(t)
        traj_c.append(a ** 0.5 * z0_c + (1 - a) ** 0.5 * eps)
        traj_s.append(a ** 0.5 * z0_s + (1 - a) ** 0.5 * eps)
    yy, xx = np.mgrid[0:8 * h, 0:8 * w]
    mask_u8 = np.stack([(((xx - (4 * w + 5 * f)) ** 2 + (yy - 4 * h) ** 2) < (2.5 * h) ** 2).astype(np.uint8) * 255
                        for f in range(F_)])
    emb = torch.randn(1, 77, uo.TINY_CONFIG["cross_attention_dim"], generator=gen)
    pnp_utils.register_spatial_attention_pnp(pipe)
    z_T = pnp_utils.latent_adain(traj_c[n].cuda().half(), traj_s[n].cuda().half())
    out = pipe.video_style_transfer("", num_inference_steps=n, latents=z_T, content_inv_path=[t.half() for t in traj_c],
                                    style_inv_path=[t.half() for t in traj_s], mask_path=torch.from_numpy(mask_u8),
                                    prompt_embeds=emb).latents
    sd32 = {k: v.cuda() for k, v in uo.seeded_state_dict(uo.TINY_CONFIG, seed=33).items()}
    with torch.no_grad():
        unet_fn = lambda x, t, ctx, i: uo.unet_forward(sd32, uo.TINY_CONFIG, x, t, ctx, patched=True, idx=i)
        truth = po.video_style_transfer(unet_fn, uo.latent_adain(traj_c[n].cuda(), traj_s[n].cuda()),
                                        [t.cuda() for t in traj_c], [t.cuda() for t in traj_s],
                                        po.load_mask_values(mask_u8).cuda(), emb.cuda().repeat(3, 1, 1), n)
    rel, psnr = _rel(out, truth), _psnr(out, truth)
    print(f"non-square clip, {n} steps: rel={rel:.3e} psnr={psnr:.1f} dB")
    assert out.shape == truth.shape and torch.isfinite(out).all() and rel <= 3e-2 and psnr >= 30.0


def test_pixel_smoother_branch(pipe):
    """The sliding-window smoother of stable_diffusion.py:713-758 (hard-coded off in the reference, ``smoother = None``)
    switched on: steps 0..19 are untouched; step 20 must equal the branch restated from its parts -- predicted x0 (:718),
    temporal VAE decode to uint8 (:721), the NumPy / cv2-exact window smoothing of the ORACLE on those frames with the mask
    kept (:725-751), re-encoding (:753), noise recomputed from the stabilised x0 (:782-791), DDIM update (:761).  VAE: the
    tiny seeded AutoencoderKLTemporalDecoder (parity unpinned, tests/test_vae_gpu.py); flows: analytic, frame-independent."""
    from oracle import flowwarp_oracle as fo
    from oracle import vae_oracle as vo
    from univst_b200 import ops
    from univst_b200.vae import AutoencoderKLTemporalDecoder
    g = torch.load(os.path.join(GOLDEN, "style_transfer_tiny.pt"), weights_only=True)
    n = g["n"]
    traj_c, traj_s, mask_u8 = po.synthetic_inputs(g["seed"], g["F"], g["hw"], n)
    H = g["hw"] * 8
    mask_px = torch.from_numpy(mask_u8)
    if mask_px.shape[-1] != H:   # the smoother applies the mask at pixel resolution
        mask_px = torch.nn.functional.interpolate(mask_px[None].float(), size=(H, H), mode="nearest")[0].to(torch.uint8)
    vae = AutoencoderKLTemporalDecoder(vo.seeded_state_dict(vo.TINY_VAE_CONFIG, seed=55), vo.TINY_VAE_CONFIG)
    fwd = fo.synthetic_flow(H, H, 3)
    bwd = fo.synthetic_flow(H, H, 0, backward_of=fwd)
    fwd_c, bwd_c = torch.from_numpy(fwd).cuda(), torch.from_numpy(bwd).cuda()
    from univst_b200 import pnp_utils
    pnp_utils.register_spatial_attention_pnp(pipe)
    z_T = pnp_utils.latent_adain(traj_c[n].cuda().half(), traj_s[n].cuda().half())
    kw = dict(num_inference_steps=n, latents=z_T, content_inv_path=[t.half() for t in traj_c],
              style_inv_path=[t.half() for t in traj_s], mask_path=mask_px, prompt_embeds=g["emb"])
    plain = {}
    pipe.video_style_transfer("", callback=lambda i, t, z: plain.__setitem__(i, z.clone()), **kw)
    rec, eps20 = {}, {}

    def cb(i, t, z):
        rec[i] = z.clone()
        if i == 20:
            eps20["rows"], eps20["branch"] = pipe.unet.last_eps_rows.clone(), pipe.unet.last_edit_branch
    pipe.vae = vae
    try:
        out = pipe.video_style_transfer("", smoother="pixel", flow_fn=lambda a, b: (fwd_c, bwd_c), callback=cb, **kw)
    finally:
        pipe.vae = None
    assert torch.isfinite(out.latents).all()
    assert all(torch.equal(rec[i], plain[i]) for i in range(20)) and not torch.equal(rec[20], plain[20])
    # step 20 from its parts
    t20 = int(pipe.scheduler.timesteps[20])
    a_t, a_prev = pipe.scheduler.step_alphas(t20)
    z = rec[19]
    from univst_b200 import ops as _ops
    m = pipe._mask(mask_px, g["F"], g["hw"], g["hw"])
    z = _ops.latent_blend(z, traj_c[n - 20].cuda().half().contiguous(), m)      # the step's mask blend (:687-692), i = 20 <= 45
    x0 = torch.empty_like(z)
    _ops.ddim_step(z, eps20["rows"], eps20["branch"], a_t, a_prev, x0_out=x0)
    frames = vae.decode_latents_u8(x0)
    keep = (mask_px != 0).numpy().astype(np.uint8)
    est = fo.sliding_window_smooth(frames.cpu().numpy(), lambda k, nw: (fwd, bwd), keep_mask=keep)
    x0s = vae.encode_frames_u8(torch.from_numpy(est).cuda(), generator=torch.Generator(device="cuda").manual_seed(0))
    eps = (z.float() - a_t ** 0.5 * x0s.float()) / (1 - a_t) ** 0.5
    want = a_prev ** 0.5 * ((z.float() - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5) + (1 - a_prev) ** 0.5 * eps
    rel = _rel(rec[20], want)
    print(f"smoothed step 20 vs the branch restated from its parts: rel {rel:.3e}")
    assert rel < 2e-3
