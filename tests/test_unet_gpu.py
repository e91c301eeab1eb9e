"""GPU parity of the full B200 UNet forward (through the C ABI) against the CPU oracle, on the golden inputs.

Tolerance (stated by north_star as "within a stated fp16 tolerance"): the oracle is evaluated in fp32; the reference
itself runs in fp16.  We require  ||ours - oracle32||_2 / ||oracle32||_2 <= 1e-2  and max-abs error <= 0.06 on outputs
of magnitude ~3, and additionally that our error is no worse than 2x the error of the same oracle code run in fp16 on
the GPU with PyTorch's own kernels (the "reference precision").
"""
import os

import pytest
import torch

from oracle import unet_oracle as uo

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def setup(cuda_lib):
    from univst_b200.unet import UNetPseudo3DConditionModel
    g = torch.load(os.path.join(GOLDEN, "unet_tiny.pt"), weights_only=True)
    sd = uo.seeded_state_dict(uo.TINY_CONFIG, seed=g["seed"])
    unet = UNetPseudo3DConditionModel(sd, uo.TINY_CONFIG)
    sd16 = {k: v.cuda().half() for k, v in sd.items()}
    return g, sd, sd16, unet


def _errs(y, ref):
    y, ref = y.float().cpu(), ref.float().cpu()
    return ((y - ref).norm() / ref.norm()).item(), (y - ref).abs().max().item()


@pytest.mark.parametrize("case", ["stock_t981", "patched_idx0_t981", "patched_idx13_t721", "patched_idx25_t481",
                                  "patched_idx26_t461"])
def test_unet_forward_matches_oracle(setup, case):
    from univst_b200 import pnp_utils
    from types import SimpleNamespace
    g, sd, sd16, unet = setup
    t = int(case.split("_t")[-1])
    patched = case.startswith("patched")
    idx = int(case.split("idx")[1].split("_")[0]) if patched else None
    pipe = SimpleNamespace(unet=unet)
    for tr in unet._all_transformers():  # reset patch state between cases
        tr.transformer_blocks[0].attn1.__dict__.pop("_patched", None)
    if patched:
        pnp_utils.register_spatial_attention_pnp(pipe)
        pnp_utils.register_time(pipe, idx)
    x, ctx = g["x"], g["ctx"]
    y = unet(x.cuda().half(), torch.tensor(t), encoder_hidden_states=ctx.cuda().half()).sample
    torch.cuda.synchronize()
    golden = g["cases"][case]  # produced by the reference's own module code
    with torch.no_grad():
        y16 = uo.unet_forward(sd16, uo.TINY_CONFIG, x.cuda().half(), t, ctx.cuda().half(), patched=patched, idx=idx)
    rel, mx = _errs(y, golden)
    rel16, mx16 = _errs(y16, golden)
    print(f"{case}: ours rel={rel:.3e} max={mx:.3e} | torch-fp16 eager rel={rel16:.3e} max={mx16:.3e}")
    assert y.shape == golden.shape and torch.isfinite(y).all()
    assert rel <= 1e-2 and mx <= 0.06
    assert rel <= 2.0 * rel16 + 1e-3


@pytest.mark.parametrize("case", ["stock_t501", "patched_idx10_t781"])
def test_unet_forward_sd21_layout(cuda_lib, case):
    """BASELINE.json configs[2] backbone layout (SD-2.1: head dim 64, Linear projections) on the tiny golden."""
    from types import SimpleNamespace
    from univst_b200 import pnp_utils
    from univst_b200.unet import UNetPseudo3DConditionModel
    g = torch.load(os.path.join(GOLDEN, "unet_tiny_sd21.pt"), weights_only=True)
    unet = UNetPseudo3DConditionModel(uo.seeded_state_dict(uo.TINY_SD21_CONFIG, seed=g["seed"]), uo.TINY_SD21_CONFIG)
    if case.startswith("patched"):
        pipe = SimpleNamespace(unet=unet)
        pnp_utils.register_spatial_attention_pnp(pipe)
        pnp_utils.register_time(pipe, 10)
    y = unet(g["x"].cuda().half(), int(case.split("_t")[-1]), encoder_hidden_states=g["ctx"].cuda().half()).sample
    rel, mx = _errs(y, g["cases"][case])
    print(f"sd21 {case}: rel={rel:.3e} max={mx:.3e}")
    assert rel <= 1e-2 and mx <= 0.06


def test_reference_patch_protocol_is_honoured(setup):
    """An instance-level ``forward`` override (what the reference's register_spatial_attention_pnp installs) marks the
    layer as patched; unpatched layers ignore idx."""
    g, sd, sd16, unet = setup
    a1 = unet.up_blocks[2].attentions[1].transformer_blocks[0].attn1
    a1.__dict__.pop("_patched", None)
    assert not a1.patched
    a1.forward = lambda *a, **k: None
    assert a1.patched
    del a1.forward
    assert not a1.patched


def test_feature_dump_matches_oracle(setup, tmp_path):
    """ft_indices / ft_timesteps / ft_path side effect (unet_3d_condition.py:430-436): (F, h, w, C) of branch 0."""
    g, sd, sd16, unet = setup
    for tr in unet._all_transformers():
        tr.transformer_blocks[0].attn1.__dict__.pop("_patched", None)
    x, ctx = g["x"], g["ctx"]
    unet(x.cuda().half(), 301, encoder_hidden_states=ctx.cuda().half(), ft_indices=[2], ft_timesteps=[301],
         ft_path=str(tmp_path))
    f = torch.load(tmp_path / "inversion_feature_map_2_block_301_step.pt", weights_only=True)
    feats = {}
    with torch.no_grad():
        uo.unet_forward(sd, uo.TINY_CONFIG, x, 301, ctx, features=feats)
    ref = feats[2][0].permute(1, 2, 3, 0)
    rel, mx = _errs(f, ref)
    print(f"feature dump: rel={rel:.3e} max={mx:.3e}")
    assert f.shape == ref.shape and rel <= 1e-2


@pytest.mark.parametrize("B,F,h,w,patched", [
    (1, 1, 16, 16, False),    # one frame: prev = self = first
    (1, 3, 8, 8, False),      # the inversion form (batch 1, stock attention); 1 token at the deepest level
    (3, 1, 16, 16, True),     # three branches of a single frame through the shift
    (3, 3, 24, 40, True),     # non-square latent: 960 / 240 / 60 / 15 tokens -- ragged tiles at every level
    (3, 5, 40, 8, True),      # odd frame count, tall latent
    (3, 2, 8, 136, True),     # rows wider than one 128-pixel tile (1088-pixel frames)
])
def test_unet_forward_edge_shapes(setup, B, F, h, w, patched):
    """Shapes the goldens do not hold, against the (golden-pinned) oracle in fp32 on the same seeded inputs."""
    from univst_b200 import pnp_utils
    from types import SimpleNamespace
    g, sd, sd16, unet = setup
    pipe = SimpleNamespace(unet=unet)
    for tr in unet._all_transformers():
        tr.transformer_blocks[0].attn1.__dict__.pop("_patched", None)
    idx = 7 if patched else None
    if patched:
        pnp_utils.register_spatial_attention_pnp(pipe)
        pnp_utils.register_time(pipe, idx)
    gen = torch.Generator().manual_seed(1000 * B + 100 * F + h + w)
    x = torch.randn(B, 4, F, h, w, generator=gen)
    ctx = torch.randn(B, 77, uo.TINY_CONFIG["cross_attention_dim"], generator=gen)
    t = 621
    y = unet(x.cuda().half(), torch.tensor(t), encoder_hidden_states=ctx.cuda().half()).sample
    torch.cuda.synchronize()
    sd32 = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        truth = uo.unet_forward(sd32, uo.TINY_CONFIG, x.cuda(), t, ctx.cuda(), patched=patched, idx=idx)
        y16 = uo.unet_forward(sd16, uo.TINY_CONFIG, x.cuda().half(), t, ctx.cuda().half(), patched=patched, idx=idx)
    rel, mx = _errs(y, truth)
    rel16, mx16 = _errs(y16, truth)
    print(f"B{B} F{F} {h}x{w} patched={patched}: ours rel={rel:.3e} max={mx:.3e} | torch-fp16 rel={rel16:.3e} max={mx16:.3e}")
    assert y.shape == truth.shape and torch.isfinite(y).all()
    assert rel <= 1e-2 and rel <= 2.0 * rel16 + 1e-3


def test_cuda_graph_forward_and_truncation_are_bit_identical(setup):
    """The forward replayed as a CUDA graph (step scalars -- timestep, shift parameters -- read from device memory) equals
    the eager launch sequence bit for bit, across DDIM steps that share one captured graph (idx 3 and 17: different beta)
    and across a plan change (idx 30: shift window closed -> another graph); and the edit branch of a call that drops the
    dead content / style branches after the last patched projection equals the edit branch of the full call."""
    from univst_b200 import pnp_utils
    from types import SimpleNamespace
    g, sd, sd16, unet = setup
    pipe = SimpleNamespace(unet=unet)
    pnp_utils.register_spatial_attention_pnp(pipe)
    x, ctx = g["x"].cuda().half(), g["ctx"].cuda().half()
    try:
        for idx, t in ((3, 921), (17, 641), (30, 381), (3, 921)):
            pnp_utils.register_time(pipe, idx)
            unet.use_cuda_graphs = False
            eager = unet(x, t, encoder_hidden_states=ctx).sample.clone()
            unet.truncate_dead_branches = True
            cut = unet(x, t, encoder_hidden_states=ctx).sample.clone()
            live = unet.shift_live(unet.up_blocks[1].attentions[1].transformer_blocks[0].attn1)
            assert cut.shape[0] == (1 if live else 3) and unet.last_edit_branch == cut.shape[0] - 1
            assert torch.equal(cut[-1], eager[2]), f"idx {idx}: truncated edit branch differs"
            unet.use_cuda_graphs = True
            cut_g = unet(x, t, encoder_hidden_states=ctx).sample.clone()
            unet.truncate_dead_branches = False
            graphed = unet(x, t, encoder_hidden_states=ctx).sample.clone()
            assert torch.equal(graphed, eager), f"idx {idx}: graphed forward differs"
            assert torch.equal(cut_g, cut), f"idx {idx}: graphed truncated forward differs"
        assert len(unet._graphs) == 4   # (open, closed) x (full, truncated): idx 3 and 17 share their graphs
    finally:
        unet.use_cuda_graphs = False
        unet.truncate_dead_branches = False
        unet.drop_cuda_graphs()
