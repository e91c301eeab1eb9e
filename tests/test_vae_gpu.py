"""GPU checks of the VAE legs (univst_b200/vae.py: KL encoder, temporal decoder, posterior sample, pixel conversions) and of
the kernels added for them, against plain-PyTorch fp32 definitions.

PARITY UNPINNED for the network as a whole (third-party diffusers ``AutoencoderKLTemporalDecoder``, no source / weights
here): the oracle (oracle/vae_oracle.py) is an independent fp32 evaluation of the same restated architecture, so these tests
pin the product to that definition, not to the library.  Tolerances: fp16 storage between layers -> rel-L2 <= 1e-2 through
the whole encoder / decoder (as for the UNet), a few fp16 ulps for single kernels; uint8 pixels may differ by 1 LSB where
the fp16 value sits on a rounding boundary."""
import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle as vo

pytestmark = pytest.mark.gpu


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half().cuda()


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.mark.parametrize("NB,Fr,HW,C,Cout", [(1, 4, 64, 64, 64), (2, 5, 96, 128, 64), (1, 16, 4096, 128, 128), (1, 3, 300, 64, 96),
                                            (1, 1, 256, 64, 64)])
def test_conv_temporal3(cuda_lib, NB, Fr, HW, C, Cout):
    """(3, 1, 1) temporal convolution = three taps along the frame axis, zero frames beyond the clip (also between clips)."""
    from univst_b200 import ops
    x = _rand(NB * Fr * HW, C, seed=1)
    w5 = _rand(Cout, C, 3, 1, 1, scale=(3 * C) ** -0.5, seed=2)
    bias, res = _rand(Cout, seed=3), _rand(NB * Fr * HW, Cout, seed=4)
    w = w5.float().reshape(Cout, C, 3).permute(0, 2, 1).reshape(Cout, 3 * C).half().contiguous()
    out = ops.conv_temporal3(x, w, NB=NB, F=Fr, HW=HW, bias=bias, residual=res)
    x5 = x.float().view(NB, Fr, HW, 1, C).permute(0, 4, 1, 2, 3)                       # (b, c, t, hw, 1)
    ref = F.conv3d(x5, w5.float(), bias.float(), padding=(1, 0, 0)).permute(0, 2, 3, 4, 1).reshape(-1, Cout) + res.float()
    err = (out.float() - ref).abs().max().item()
    print(f"temporal conv {NB}x{Fr}x{HW} {C}->{Cout}: max abs err {err:.3e}")
    assert err < 3e-3 + 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("NB,H,W,C,Cout", [(2, 16, 16, 64, 64), (1, 64, 64, 128, 128), (3, 24, 40, 64, 96)])
def test_conv3x3_stride2_pad_after(cuda_lib, NB, H, W, C, Cout):
    """diffusers Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then a stride-2 3x3 convolution."""
    from univst_b200 import ops
    x, w4 = _rand(NB, H, W, C, seed=1), _rand(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=2)
    bias = _rand(Cout, seed=3)
    w = w4.float().permute(0, 2, 3, 1).reshape(Cout, 9 * C).half().contiguous()
    out = ops.conv3x3_s2_pad_after(ops.space_to_depth2(x), w, bias=bias)
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), w4.float(), bias.float(), stride=2)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    err = (out.float() - ref).abs().max().item()
    assert err < 3e-3 + 2e-3 * ref.abs().max().item(), err


def test_softmax_rows_and_pixel_conversions(cuda_lib):
    from univst_b200 import ops
    s = _rand(300, 4096, scale=3.0, seed=1)
    ref = torch.softmax(s.float() * 0.25, dim=-1)
    out = ops.softmax_rows_(s.clone(), 0.25)
    assert (out.float() - ref).abs().max().item() < 2e-3 and abs(out.float().sum(-1) - 1).max().item() < 5e-3
    px = _rand(5000, 8, scale=0.8, seed=2)
    u8 = ops.frames_to_u8(px, 5000)
    want = (((px[:, :3] / 2 + 0.5).clamp(0, 1)).float() * 255).round().to(torch.uint8)
    assert torch.equal(u8, want)
    img = torch.randint(0, 256, (7, 9, 11, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
    rows = ops.u8_to_frames(img, 64)
    want = (img.double() / 127.5 - 1.0).half().view(-1, 3)
    assert torch.equal(rows[:, :3], want) and not rows[:, 3:].any()


@pytest.fixture(scope="module")
def tiny(cuda_lib):
    from univst_b200.vae import AutoencoderKLTemporalDecoder
    cfg = vo.TINY_VAE_CONFIG
    sd = vo.seeded_state_dict(cfg, seed=55)
    return cfg, sd, AutoencoderKLTemporalDecoder(sd, cfg)


def test_vae_decode_matches_oracle(tiny):
    cfg, sd, vae = tiny
    Fr, h, w = 4, 8, 12
    z = torch.randn(Fr, 4, h, w, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = vo.decode(sd, cfg, z, Fr)
    out = vae.decode(z.cuda().half(), num_frames=Fr).sample
    rel = _rel(out, ref)
    print(f"temporal decoder ({Fr} frames {h}x{w} latents): rel-L2 vs fp32 oracle {rel:.3e}")
    assert out.shape == ref.shape and rel <= 1e-2
    # latents in the pipeline's (1, C, F, h, w) layout -> uint8 frames (get_images_from_latents)
    lat = (z * cfg["scaling_factor"]).permute(1, 0, 2, 3).unsqueeze(0)
    u8 = vae.decode_latents_u8(lat.cuda().half())
    with torch.no_grad():
        want = vo.frames_to_u8(vo.decode(sd, cfg, (1 / cfg["scaling_factor"] * lat.half()).float()[0].permute(1, 0, 2, 3), Fr))
    diff = (u8.cpu().int() - want.int()).abs()
    print(f"uint8 frames: max diff {diff.max().item()} LSB, {float((diff > 0).float().mean()) * 100:.2f} % of the values differ")
    assert u8.shape == want.shape and diff.max().item() <= 2 and float((diff > 1).float().mean()) < 1e-3


def test_vae_encode_matches_oracle(tiny):
    cfg, sd, vae = tiny
    N, H, W = 3, 64, 96
    x = torch.rand(N, 3, H, W, generator=torch.Generator().manual_seed(2)) * 2 - 1
    with torch.no_grad():
        mom = vo.encode_moments(sd, cfg, x)
    post = vae.encode(x.cuda().half()).latent_dist
    mode = post.mode()
    rel = _rel(mode, mom[:, :4])
    print(f"KL encoder ({N} frames {H}x{W}): rel-L2 of the posterior mean vs fp32 oracle {rel:.3e}")
    assert mode.shape == (N, 4, H // 8, W // 8) and rel <= 1e-2
    # sample(): same torch generator call as diffusers' randn_tensor(mean.shape) -> mean + std * noise
    g = torch.Generator(device="cuda").manual_seed(9)
    smp = post.sample(g)
    noise = torch.randn((N, 4, H // 8, W // 8), generator=torch.Generator(device="cuda").manual_seed(9), device="cuda",
                        dtype=torch.float16)
    want = vo.sample_latents(mom, noise.float().cpu(), 1.0)[0].permute(1, 0, 2, 3)
    assert _rel(smp, want) <= 1e-2
    # uint8 frames -> scaled latents in the pipeline layout (get_latent_image / the inversion's encode)
    img = (x.permute(0, 2, 3, 1) * 127.5 + 127.5).round().clamp(0, 255).to(torch.uint8)
    lat = vae.encode_frames_u8(img.cuda(), sample=False)
    with torch.no_grad():
        want = vo.sample_latents(vo.encode_moments(sd, cfg, (img.double() / 127.5 - 1.0).float().permute(0, 3, 1, 2)), None,
                                 cfg["scaling_factor"])
    assert lat.shape == want.shape == (1, 4, N, H // 8, W // 8) and _rel(lat, want) <= 1e-2
