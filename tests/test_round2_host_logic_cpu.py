"""Host logic of the round-2 mirrors on the CPU, with every kernel replaced by its plain-torch definition (tests/_torch_ops.py):
the VAE (weight packing incl. the padded tiny-channel convolutions and the tap-major temporal weights, channels-last layouts,
chunked decode, posterior sample, pixel conversions) and the SD3 / SD3.5 MMDiT (patch embedding, cropped positional table,
adaLN chunk order of the joint / dual / context_pre_only blocks, processor protocol, feature dump, unpatchify) against their
fp32 oracles -- a packing or indexing regression is caught before any GPU time is spent.  Both networks are third-party
restatements (parity unpinned, see the oracles' headers); the kernels themselves are only trusted on the GPU."""
import os

import pytest
import torch

import _torch_ops


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_vae_host_logic_on_cpu(monkeypatch):
    from oracle import vae_oracle as vo
    from univst_b200.vae import AutoencoderKLTemporalDecoder
    _torch_ops.install(monkeypatch)
    cfg = vo.TINY_VAE_CONFIG
    sd = vo.seeded_state_dict(cfg, seed=55)
    vae = AutoencoderKLTemporalDecoder(sd, cfg, device="cpu")
    g = torch.Generator().manual_seed(1)
    Fr, h, w = 3, 4, 6
    z = torch.randn(Fr, 4, h, w, generator=g)
    with torch.no_grad():
        ref = vo.decode(sd, cfg, z, Fr)
    out = vae.decode(z.half(), num_frames=Fr).sample
    assert out.shape == ref.shape and _rel(out, ref) < 1e-2
    lat = (z * cfg["scaling_factor"]).permute(1, 0, 2, 3).unsqueeze(0).half()
    u8 = vae.decode_latents_u8(lat)
    assert u8.shape == (Fr, 8 * h, 8 * w, 3) and u8.dtype == torch.uint8
    u8c = vae.decode_latents_u8(lat, decode_chunk_size=2)            # chunks are separate clips to the temporal layers
    assert torch.equal(u8c[:2], vae.decode_latents_u8(lat[:, :, :2]))
    x = torch.rand(2, 3, 32, 48, generator=g) * 2 - 1
    with torch.no_grad():
        mom = vo.encode_moments(sd, cfg, x)
    post = vae.encode(x.half()).latent_dist
    assert _rel(post.mode(), mom[:, :4]) < 1e-2
    img = (x.permute(0, 2, 3, 1) * 127.5 + 127.5).round().clamp(0, 255).to(torch.uint8)
    lat = vae.encode_frames_u8(img, sample=False)
    with torch.no_grad():
        want = vo.sample_latents(vo.encode_moments(sd, cfg, (img.double() / 127.5 - 1.0).float().permute(0, 3, 1, 2)), None,
                                 cfg["scaling_factor"])
    assert lat.shape == want.shape and _rel(lat, want) < 1e-2


@pytest.mark.parametrize("case", ["stock_joint_attention", "shift_idx5"])
def test_sd3_transformer_host_logic_on_cpu(monkeypatch, case, tmp_path):
    from oracle import sd3_transformer_oracle as to
    from univst_b200 import sd3
    from univst_b200.sd3_transformer import SD3Transformer2DModel
    _torch_ops.install(monkeypatch)
    cfg = to.TINY_CONFIG
    sd = to.seeded_state_dict(cfg, seed=71)
    model = SD3Transformer2DModel(sd, cfg, device="cpu")
    g = torch.Generator().manual_seed(0)
    BF = 48
    x, enc = torch.randn(BF, 16, 4, 6, generator=g), torch.randn(BF, 5, cfg["joint_attention_dim"], generator=g)
    pooled, t = torch.randn(BF, cfg["pooled_projection_dim"], generator=g), torch.full((BF,), 640.0)
    kw, okw = {}, dict(cross_frame=False)
    if case != "stock_joint_attention":
        sd3.register_spatial_attention_pnp(type("P", (), {"transformer": model})())
        kw, okw = dict(joint_attention_kwargs={"idx": 5}), dict(idx=5)
    with torch.no_grad():
        ref, feats = to.forward(sd, cfg, x, enc, pooled, t, feature_blocks=(0,), **okw)
    out = model(x.half(), encoder_hidden_states=enc.half(), pooled_projections=pooled.half(), timestep=t, idx=3, ft_indices=[0],
                ft_timesteps=[3], ft_path=str(tmp_path), **kw).sample
    assert out.shape == ref.shape and _rel(out, ref) < 1e-2
    f = torch.load(os.path.join(tmp_path, "inversion_feature_map_0_block_3_step.pt"), weights_only=True)
    assert tuple(f.shape) == tuple(feats[0].shape) and _rel(f, feats[0]) < 1e-2


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_source_tables_name_the_right_frames(world):
    """The K/V source tables of a frame shard (SD backbone: unet.kv_source_table_sharded; SD3 processors:
    sd3._source_table_sharded) against the unsharded tables: every local / halo-bank index, mapped back to the global
    (branch, frame) it holds -- bank 1 = last frame of the previous rank, bank 2 = frame 0 of the clip -- must name the
    frame the unsharded table names."""
    from univst_b200 import sd3
    from univst_b200.unet import kv_source_table, kv_source_table_sharded
    B, F = 3, 16
    Fl = F // world
    for rank in range(world):
        NI = B * Fl

        def to_global(idx):
            if idx < NI:                                  # local image (b, fl)
                b, fl = divmod(idx, Fl)
                return b * F + rank * Fl + fl
            if idx < NI + B:                              # bank 1: last frame of the previous rank
                return (idx - NI) * F + rank * Fl - 1
            return (idx - NI - B) * F                     # bank 2: frame 0 of the clip
        for mode in ("prev_first", "prev_self_first", "self"):
            full = kv_source_table(B, F, mode).tolist()
            loc = kv_source_table_sharded(B, Fl, mode, rank).tolist()
            for b in range(B):
                for fl in range(Fl):
                    want = full[b * F + rank * Fl + fl]
                    assert [to_global(i) for i in loc[b * Fl + fl]] == want, (world, rank, mode, b, fl)
        full3 = sd3._source_table(B * F, "cpu", True, text=False).tolist()
        loc3 = sd3._source_table_sharded(B, Fl, rank, "cpu", text=True).tolist()
        for b in range(B):
            for fl in range(Fl):
                me = b * Fl + fl
                assert [to_global(i) for i in loc3[me][:3]] == full3[b * F + rank * Fl + fl], (world, rank, b, fl)
                assert loc3[me][3] == NI + 2 * B + me          # own text tokens: second K/V tensor, behind all first-tensor images
