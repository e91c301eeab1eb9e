"""CPU: the AnimateDiff oracle against the golden vectors produced by the reference's own ``UNet3DConditionModel``
(oracle/gen_golden_animatediff.py), and host-side properties of the B200 mirror that need no GPU."""
import os

import pytest
import torch

from oracle import animatediff_oracle as ao

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["stock_t981", "patched_idx0_t981", "patched_idx13_t721", "patched_idx24_t501", "patched_idx25_t481"]


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(GOLDEN, "animatediff_tiny.pt"), weights_only=True)


@pytest.mark.parametrize("case", CASES)
def test_animatediff_oracle_matches_reference_module(golden, case):
    sd = ao.seeded_state_dict(ao.AD_TINY_CONFIG, seed=golden["seed"])
    patched = case.startswith("patched")
    idx = int(case.split("idx")[1].split("_")[0]) if patched else None
    with torch.no_grad():
        y = ao.unet_forward(sd, ao.AD_TINY_CONFIG, golden["x"], int(case.split("_t")[-1]), golden["ctx"], patched=patched,
                            idx=idx)
    ref = golden["cases"][case]
    rel = ((y - ref).norm() / ref.norm()).item()
    assert y.shape == ref.shape and rel < 1e-4, rel   # fp32 vs fp32: summation-order noise only


def test_shift_window_is_half_open(golden):
    """backbones/animatediff/pnp_utils.py:45: ``idx < eta2 * 50`` -- idx 24 shifts, idx 25 does not (the SD backbone
    still shifts at 25).  The goldens must tell the two apart, otherwise the cases above prove nothing about it."""
    assert ao.shift_params(24)[0] and not ao.shift_params(25)[0]
    sd = ao.seeded_state_dict(ao.AD_TINY_CONFIG, seed=golden["seed"])
    with torch.no_grad():
        on = ao.unet_forward(sd, ao.AD_TINY_CONFIG, golden["x"], 481, golden["ctx"], patched=True, idx=24)
    off = golden["cases"]["patched_idx25_t481"]
    assert ((on - off).norm() / off.norm()).item() > 1e-2


def test_motion_modules_are_live_in_the_goldens(golden):
    """With the motion modules removed the output must change: the goldens exercise temporal attention."""
    sd = ao.seeded_state_dict(ao.AD_TINY_CONFIG, seed=golden["seed"])
    dead = {k: (torch.zeros_like(v) if "temporal_transformer.proj_out" in k else v) for k, v in sd.items()}
    with torch.no_grad():
        y = ao.unet_forward(dead, ao.AD_TINY_CONFIG, golden["x"], 981, golden["ctx"])
    ref = golden["cases"]["stock_t981"]
    assert ((y - ref).norm() / ref.norm()).item() > 1e-2


def test_positional_encoding_folds_through_the_projection():
    """The identity the B200 mirror relies on: W (n + pe_f) = W n + W pe_f (to_q / to_k / to_v have no bias)."""
    from univst_b200.animatediff import positional_encoding
    torch.manual_seed(0)
    C, F = 64, 5
    pe = positional_encoding(C, 24)
    assert torch.equal(pe, ao.positional_encoding(C, 24))
    w, n = torch.randn(3 * C, C), torch.randn(F, 7, C)
    lhs = torch.nn.functional.linear(n + pe[:F, None], w)
    rhs = torch.nn.functional.linear(n, w) + (pe[:F] @ w.T)[:, None]
    assert torch.allclose(lhs, rhs, atol=1e-4)


def test_pipeline_flavour_matches_reference_conditions():
    """pipeline_animation.py:505-506 (index 50 - i) and :515 (i >= 0.8 n) against stable_diffusion.py:683, :694."""
    from univst_b200.animatediff import AnimationPipeline
    from univst_b200.pipeline import SpatioTemporalStableDiffusionPipeline as SD
    assert AnimationPipeline._traj_index(7, 50) == 43 and SD._traj_index(7, 50) == 43
    assert AnimationPipeline._late_adain(40, 50) and not SD._late_adain(40, 50)
    assert AnimationPipeline._late_adain(45, 50) and not AnimationPipeline._late_adain(46, 50)


def test_animation_pipeline_oracle_matches_reference_pipeline():
    """The reference's own AnimationPipeline.video_style_transfer (50 steps, 16 frames, linear-beta DDIM, mask blend,
    late AdaIN from step 40, shift window idx < 25) vs the oracle loop: fp32, 2e-4 absolute after 50 steps."""
    from oracle import pipeline_oracle as po
    from oracle import unet_oracle as uo
    g = torch.load(os.path.join(GOLDEN, "style_transfer_animatediff_tiny.pt"), weights_only=True)
    sd = ao.seeded_state_dict(ao.AD_TINY_CONFIG, seed=44)
    traj_c, traj_s, mask_u8 = po.synthetic_inputs(g["seed"], g["F"], g["hw"], g["n"])
    z_T = uo.latent_adain(traj_c[g["n"]], traj_s[g["n"]])
    assert torch.allclose(z_T, g["z_T"], atol=1e-6)

    def unet_fn(x, t, ctx, idx):
        with torch.no_grad():
            return ao.unet_forward(sd, ao.AD_TINY_CONFIG, x, t, ctx, patched=idx is not None, idx=idx)

    rec = {i: None for i in g["steps"]}
    z = po.video_style_transfer(unet_fn, z_T, traj_c, traj_s, po.load_mask_values(mask_u8), g["emb"].repeat(3, 1, 1),
                                g["n"], rec, animatediff=True)
    for i, ref in g["steps"].items():
        assert (rec[i] - ref).abs().max().item() < 2e-4, i
    assert (z - g["final"]).abs().max().item() < 2e-4


@pytest.mark.parametrize("case", ["stock_t981", "patched_idx13_t721", "patched_idx24_t501", "patched_idx25_t481"])
def test_animatediff_unet_host_logic_on_cpu(monkeypatch, golden, case):
    """Host side of the AnimateDiff UNet mirror -- per-frame GroupNorm spans, per-frame attn1 tables, the half-open shift window,
    motion-module packing (fused QKV, the positional encoding folded through the projection as an epilogue row vector,
    tile-interleaved GEGLU) -- on the CPU, with every kernel replaced by a plain-torch definition (tests/_torch_ops.py):
    must give the output of the REFERENCE's own UNet3DConditionModel (golden) up to the fp16 storage between layers."""
    import _torch_ops
    from types import SimpleNamespace
    from univst_b200 import pnp_utils
    from univst_b200.animatediff import UNet3DConditionModel
    _torch_ops.install(monkeypatch)
    sd = ao.seeded_state_dict(ao.AD_TINY_CONFIG, seed=golden["seed"])
    unet = UNet3DConditionModel(sd, ao.AD_TINY_CONFIG, device="cpu")
    if case.startswith("patched"):
        pipe = SimpleNamespace(unet=unet)
        pnp_utils.register_spatial_attention_pnp(pipe)
        pnp_utils.register_time(pipe, int(case.split("idx")[1].split("_")[0]))
    y = unet(golden["x"].half(), torch.tensor(int(case.split("_t")[-1])), encoder_hidden_states=golden["ctx"].half()).sample
    ref = golden["cases"][case]
    rel = ((y.float() - ref).norm() / ref.norm()).item()
    assert y.shape == ref.shape and rel < 5e-3, rel


def test_full_animatediff_stack_on_cpu_matches_reference_pipeline(monkeypatch):
    """The whole AnimateDiff product stack -- AnimationPipeline.video_style_transfer (trajectory index 50 - i, late AdaIN from
    0.8 n, linear betas) driving the UNet3DConditionModel mirror with its 21 motion modules, dead-branch skipping on -- on the
    CPU with every kernel replaced by a torch definition (tests/_torch_ops.py), against the latents of the REFERENCE's own
    AnimationPipeline (golden)."""
    import _torch_ops
    from oracle import pipeline_oracle as po
    from univst_b200 import pnp_utils
    from univst_b200.animatediff import AnimationPipeline, UNet3DConditionModel
    from univst_b200.scheduler import DDIMScheduler
    _torch_ops.install(monkeypatch)
    g = torch.load(os.path.join(GOLDEN, "style_transfer_animatediff_tiny.pt"), weights_only=True)
    n = g["n"]
    unet = UNet3DConditionModel(ao.seeded_state_dict(ao.AD_TINY_CONFIG, seed=44), ao.AD_TINY_CONFIG, device="cpu")
    pipe = AnimationPipeline(unet, DDIMScheduler.animatediff_v2())   # animatediff-v2.yaml:16-21
    pipe.device = torch.device("cpu")
    traj_c, traj_s, mask_u8 = po.synthetic_inputs(g["seed"], g["F"], g["hw"], n)
    pnp_utils.register_spatial_attention_pnp(pipe)
    z_T = pnp_utils.latent_adain(traj_c[n].half(), traj_s[n].half())
    rec, batches = {}, []
    fwd = unet.forward
    unet.forward = lambda x, *a, **k: (batches.append(x.shape[0]), fwd(x, *a, **k))[1]
    out = pipe.video_style_transfer("", num_inference_steps=n, latents=z_T, content_inv_path=[t.half() for t in traj_c],
                                    style_inv_path=[t.half() for t in traj_s], mask_path=torch.from_numpy(mask_u8),
                                    prompt_embeds=g["emb"], skip_dead_branches=True,
                                    callback=lambda i, t, z: rec.__setitem__(i, z.clone())).latents
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
    for i, ref in g["steps"].items():
        assert rel(rec[i], ref) < 1e-2, (i, rel(rec[i], ref))
    assert rel(out, g["final"]) < 1e-2
    assert batches == [3] * 25 + [1] * 25      # shift window idx < 25 (backbones/animatediff/pnp_utils.py:45)
