"""GPU parity of the AnimateDiff backbone (through the C ABI): the temporal-attention kernel against an fp32 torch
softmax, and the full B200 ``UNet3DConditionModel`` forward against the goldens of the reference's own module.

Tolerances as for the SD backbone (tests/test_unet_gpu.py): rel-L2 <= 1e-2, max-abs <= 0.06, and no worse than twice
the error of the same oracle code run in torch fp16 on the GPU."""
import os
from types import SimpleNamespace

import pytest
import torch

from oracle import animatediff_oracle as ao

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _errs(y, ref):
    y, ref = y.float().cpu(), ref.float().cpu()
    return ((y - ref).norm() / ref.norm()).item(), (y - ref).abs().max().item()


@pytest.mark.parametrize("B,F,N,H,d", [(3, 16, 64, 8, 40), (1, 4, 256, 8, 8), (2, 24, 16, 8, 80), (3, 16, 4, 8, 160),
                                       (1, 1, 32, 4, 16), (2, 7, 33, 8, 16)])
def test_temporal_attention_kernel(cuda_lib, B, F, N, H, d):
    from univst_b200 import ops
    torch.manual_seed(B * 1000 + F * 10 + d)
    C = H * d
    qkv = (torch.randn(B * F * N, 3 * C, device="cuda") * 1.3).half()
    o = ops.temporal_attention(qkv, B=B, F=F, N=N, H=H, d=d)
    x = qkv.float().view(B, F, N, 3, H, d).permute(3, 0, 2, 4, 1, 5)   # (3, B, N, H, F, d)
    p = torch.softmax(x[0] @ x[1].transpose(-1, -2) * d ** -0.5, dim=-1)
    ref = (p @ x[2]).permute(0, 3, 1, 2, 4).reshape(B * F * N, C)      # (B, F, N, H, d)
    rel, mx = _errs(o, ref)
    print(f"temporal attention B={B} F={F} N={N} H={H} d={d}: rel={rel:.3e} max={mx:.3e}")
    assert rel < 2e-3 and mx < 2e-2   # fp16 probabilities and output rounding


def test_temporal_attention_rejects_bad_shapes(cuda_lib):
    from univst_b200 import ops
    qkv = torch.zeros(33 * 4, 3 * 64, device="cuda", dtype=torch.float16)
    with pytest.raises(RuntimeError):
        ops.temporal_attention(qkv, B=1, F=33, N=4, H=8, d=8)   # more than 32 frames


@pytest.fixture(scope="module")
def setup(cuda_lib):
    from univst_b200.animatediff import UNet3DConditionModel
    g = torch.load(os.path.join(GOLDEN, "animatediff_tiny.pt"), weights_only=True)
    sd = ao.seeded_state_dict(ao.AD_TINY_CONFIG, seed=g["seed"])
    unet = UNet3DConditionModel(sd, ao.AD_TINY_CONFIG)
    sd16 = {k: v.cuda().half() for k, v in sd.items()}
    return g, sd16, unet


@pytest.mark.parametrize("case", ["stock_t981", "patched_idx0_t981", "patched_idx13_t721", "patched_idx24_t501",
                                  "patched_idx25_t481"])
def test_animatediff_unet_forward_matches_reference(setup, case):
    from univst_b200 import pnp_utils
    g, sd16, unet = setup
    t = int(case.split("_t")[-1])
    patched = case.startswith("patched")
    idx = int(case.split("idx")[1].split("_")[0]) if patched else None
    pipe = SimpleNamespace(unet=unet)
    for tr in unet._all_transformers():
        tr.transformer_blocks[0].attn1.__dict__.pop("_patched", None)
    if patched:
        pnp_utils.register_spatial_attention_pnp(pipe)
        pnp_utils.register_time(pipe, idx)
    x, ctx = g["x"], g["ctx"]
    y = unet(x.cuda().half(), torch.tensor(t), encoder_hidden_states=ctx.cuda().half()).sample
    torch.cuda.synchronize()
    golden = g["cases"][case]   # produced by the reference's own UNet3DConditionModel
    with torch.no_grad():
        y16 = ao.unet_forward(sd16, ao.AD_TINY_CONFIG, x.cuda().half(), t, ctx.cuda().half(), patched=patched, idx=idx)
    rel, mx = _errs(y, golden)
    rel16, mx16 = _errs(y16, golden)
    print(f"animatediff {case}: ours rel={rel:.3e} max={mx:.3e} | torch-fp16 eager rel={rel16:.3e} max={mx16:.3e}")
    assert y.shape == golden.shape and torch.isfinite(y).all()
    assert rel <= 1e-2 and mx <= 0.06
    assert rel <= 2.0 * rel16 + 1e-3


def test_animatediff_dead_branch_skip_is_bit_identical(setup):
    """Once the shift window has closed (idx >= 25) the edit branch alone gives the same bits as the 3-branch call."""
    from univst_b200 import pnp_utils
    g, sd16, unet = setup
    pipe = SimpleNamespace(unet=unet)
    pnp_utils.register_spatial_attention_pnp(pipe)
    pnp_utils.register_time(pipe, 30)
    a1 = unet.up_blocks[1].attentions[1].transformer_blocks[0].attn1
    assert not unet.shift_live(a1)
    x, ctx = g["x"].cuda().half(), g["ctx"].cuda().half()
    y3 = unet(x, 381, encoder_hidden_states=ctx).sample[2:3].clone()
    y1 = unet(x[2:3].contiguous(), 381, encoder_hidden_states=ctx[2:3]).sample
    assert torch.equal(y3, y1)


def test_animation_pipeline_matches_reference_golden(cuda_lib, tmp_path):
    """Our AnimationPipeline.video_style_transfer (through the C ABI, fp16) against the latents of the reference's own
    AnimationPipeline (fp32 golden): rel-L2 <= 3e-2 and PSNR >= 30 dB after 50 steps, as for the SD pipeline."""
    from PIL import Image
    from oracle import pipeline_oracle as po
    from univst_b200 import pnp_utils
    from univst_b200.animatediff import AnimationPipeline, UNet3DConditionModel
    from univst_b200.scheduler import DDIMScheduler
    g = torch.load(os.path.join(GOLDEN, "style_transfer_animatediff_tiny.pt"), weights_only=True)
    n = g["n"]
    unet = UNet3DConditionModel(ao.seeded_state_dict(ao.AD_TINY_CONFIG, seed=44), ao.AD_TINY_CONFIG)
    pipe = AnimationPipeline(unet, DDIMScheduler.animatediff_v2())   # animatediff-v2.yaml:16-21
    traj_c, traj_s, mask_u8 = po.synthetic_inputs(g["seed"], g["F"], g["hw"], n)
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
    rec = {}
    out = pipe.video_style_transfer("", num_inference_steps=n, latents=z_T, content_inv_path=str(cdir),
                                    style_inv_path=str(sdir), mask_path=str(mdir), prompt_embeds=g["emb"],
                                    skip_dead_branches=False, callback=lambda i, t, z: rec.__setitem__(i, z.clone()))
    for i, ref in g["steps"].items():
        print(f"animatediff step {i}: rel={_errs(rec[i], ref)[0]:.3e}")
    a, b = out.latents.float().cpu(), g["final"].float()
    rel = ((a - b).norm() / b.norm()).item()
    psnr = (10 * torch.log10((b.max() - b.min()) ** 2 / ((a - b) ** 2).mean())).item()
    print(f"animatediff final latents after {n} steps: rel={rel:.3e} psnr={psnr:.1f} dB")
    assert torch.isfinite(out.latents).all() and rel <= 3e-2 and psnr >= 30.0
    # exact dead-branch skipping: the window is idx < 25 here (25 three-branch calls, then the edit branch alone)
    batches = []
    fwd = pipe.unet.forward
    pipe.unet.forward = lambda x, *a_, **k_: (batches.append(x.shape[0]), fwd(x, *a_, **k_))[1]
    skip = pipe.video_style_transfer("", num_inference_steps=n, latents=z_T, content_inv_path=str(cdir),
                                     style_inv_path=str(sdir), mask_path=str(mdir), prompt_embeds=g["emb"],
                                     skip_dead_branches=True).latents
    del pipe.unet.forward
    assert torch.equal(skip, out.latents) and batches == [3] * 25 + [1] * 25
