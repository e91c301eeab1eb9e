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


@pytest.mark.parametrize("world,F", [(2, 32), (4, 32), (8, 32), (4, 16), (8, 48), (3, 18)])
def test_sharded_smoother_legs_place_every_frame(monkeypatch, world, F):
    """The smoother's VAE legs under frame sharding (pipeline._smoother_decode / _smoother_encode): 16-frame chunks decoded
    round-robin, frames encoded evenly, results stored into every rank's full-clip buffer.  Emulated in one process -- the
    ranks run one after another, `xrank_push` is a strided block copy into the ranks' arenas by address -- against the
    unsharded legs, incl. the worlds where some ranks have no chunk to decode (4 and 8 ranks on two chunks) and a ragged
    last chunk."""
    from types import SimpleNamespace
    from univst_b200 import ops
    from univst_b200.pipeline import SpatioTemporalStableDiffusionPipeline as Pipe

    C, h, w = 4, 2, 4
    H, W = 8 * h, 8 * w                                            # 16 x 32 frames: a uint8 row is 96 bytes = 48 halves
    g = torch.Generator().manual_seed(world * 100 + F)
    x0 = torch.randn(1, C, F, h, w, generator=g).half()

    class FakeVAE:
        config = SimpleNamespace(latent_channels=C)
        calls = []

        def decode_latents_u8(self, lat, chunk=16):
            # a chunk is ONE clip to the temporal decoder: make every frame depend on its chunk's content and its place in it
            outs = []
            for k in range(0, lat.shape[2], chunk):
                part = lat[0, :, k:k + chunk].float()
                FakeVAE.calls.append(part.shape[1])
                for j in range(part.shape[1]):
                    v = (part[:, j].sum() * 7 + part.sum() * 3 + j).item()
                    base = torch.arange(H * W * 3, dtype=torch.float32).view(H, W, 3)
                    outs.append(((base * 0.37 + v * 11.0) % 251).to(torch.uint8))
            return torch.stack(outs)

        def encode_frames_u8(self, frames, generator=None, noise=None):
            assert noise is not None or generator is not None
            F_ = frames.shape[0]
            if noise is None:
                noise = torch.randn((F_, C, h, w), generator=generator, dtype=torch.float16)   # as vae.encode_frames_u8 draws it
            mean = frames.float().view(F_, h, 8, w, 8, 3).mean(dim=(2, 4, 5))             # per-frame function
            lat = (mean[:, None] / 255.0 + 0.1 * noise.float()).half()                      # (F, C, h, w)
            return lat.permute(1, 0, 2, 3).unsqueeze(0).contiguous()

    arenas, cur = {}, {"rank": 0}

    def fake_buffer(key, shape, dtype=torch.float16):
        for r in range(world):
            arenas.setdefault((key, r), torch.zeros(*shape, dtype=dtype))
        return arenas[(key, cur["rank"])], [arenas[(key, r)].data_ptr() for r in range(world)]

    def fake_push(xr, pushes):
        for p in pushes:
            src = p["src"]
            assert p["nblk"] == 1 and src.dtype == torch.float16 and src.shape[0] == p["rows"]
            for (key, r), a in arenas.items():
                off = p["dst"][r] - a.data_ptr()
                if not (0 <= off < a.numel() * 2):
                    continue                                          # another buffer
                assert off % 2 == 0 and a.shape[1] == p["ld_dst"]
                flat = a.view(-1)
                for row in range(p["rows"]):
                    o = off // 2 + row * p["ld_dst"]
                    flat[o:o + src.shape[1]] = src[row]

    monkeypatch.setattr(ops, "xrank_push", fake_push)
    vae = FakeVAE()
    plain = SimpleNamespace(vae=vae, unet=SimpleNamespace(), device="cpu")
    plain._smoother_xr = lambda F_, W_: Pipe._smoother_xr(plain, F_, W_)
    want_frames = Pipe._smoother_decode(plain, x0)
    FakeVAE.calls.clear()
    want_lat = Pipe._smoother_encode(plain, want_frames, torch.Generator().manual_seed(5))

    ranks = []
    for r in range(world):
        ns = SimpleNamespace(vae=vae, device="cpu",
                             unet=SimpleNamespace(_shard=(object(), r, world),
                                                  _xr=SimpleNamespace(rank=r, world=world, buffer=fake_buffer, multicast=lambda key: 0)))
        ns._smoother_xr = (lambda ns_: lambda F_, W_: Pipe._smoother_xr(ns_, F_, W_))(ns)
        ranks.append(ns)
    if F % world:
        assert ranks[0]._smoother_xr(F, W) is None                   # not evenly shardable: every rank takes the unsharded legs
        return
    outs = []
    for r in range(world):
        cur["rank"] = r
        outs.append(Pipe._smoother_decode(ranks[r], x0))
    nchunks = (F + 15) // 16
    assert sorted(FakeVAE.calls) == sorted(min(16, F - 16 * c) for c in range(nchunks))   # every chunk decoded exactly once
    for r in range(world):
        assert torch.equal(outs[r], want_frames), (world, F, r)
    lats = []
    for r in range(world):
        cur["rank"] = r
        lats.append(Pipe._smoother_encode(ranks[r], outs[r], torch.Generator().manual_seed(5)))
    for r in range(world):                       # clones taken rank by rank: complete once the last rank has pushed
        cur["rank"] = r
        full = arenas[(("vae_latents", C, F, h * w), r)].view(1, C, F, h, w)
        assert torch.equal(full, want_lat), (world, F, r)
    assert torch.equal(lats[-1], want_lat)


def test_video_io_helpers_match_the_reference(monkeypatch, tmp_path):
    """util.load_video_frames / save_videos_grid against src/util.py:34-81 (build container only): the same tensors from a
    folder of frames that need resizing, the same uint8 grid frames for a batch of two clips (rescale on / off); without
    imageio the clip is written through OpenCV and reads back with the right frame count and size."""
    import sys
    import types
    import numpy as np
    from PIL import Image
    from univst_b200 import util
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs the reference checkout (build container only)")
    rec = []
    monkeypatch.setitem(sys.modules, "imageio", types.SimpleNamespace(mimsave=lambda path, frames, fps=None: rec.append((path, frames, fps))))
    monkeypatch.setitem(sys.modules, "decord", types.SimpleNamespace(bridge=types.SimpleNamespace(set_bridge=lambda n: None)))
    monkeypatch.syspath_prepend("/root/reference")
    for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        monkeypatch.delitem(sys.modules, m)
    import src.util as ref
    for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:   # do not leave the reference's modules behind
        monkeypatch.setitem(sys.modules, m, sys.modules.pop(m))
    rng = np.random.default_rng(0)
    for f in range(3):
        Image.fromarray(rng.integers(0, 255, (40, 56, 3)).astype(np.uint8)).save(tmp_path / ("%05d.png" % f))
    want = ref.load_video_frames(str(tmp_path), 3, image_size=(32, 24))
    got = util.load_video_frames(str(tmp_path), 3, image_size=(32, 24))
    assert got.shape == (3, 3, 24, 32) and got.dtype == want.dtype and torch.equal(got, want)
    with pytest.raises(ValueError):
        util.load_video_frames(str(tmp_path), 4, image_size=(32, 24))       # frame 00003.png does not exist
    g = torch.Generator().manual_seed(1)
    for rescale in (False, True):
        vid = torch.rand(2, 3, 4, 24, 32, generator=g) * (2 if rescale else 1) - (1 if rescale else 0)
        ref.save_videos_grid(vid.clone(), str(tmp_path / "a" / "ref.mp4"), rescale=rescale, n_rows=4, fps=6)
        util.save_videos_grid(vid.clone(), str(tmp_path / "b" / "our.mp4"), rescale=rescale, n_rows=4, fps=6)
        (_, fr_ref, fps_ref), (_, fr_our, fps_our) = rec[-2:]
        assert fps_ref == fps_our == 6 and len(fr_ref) == len(fr_our) == 4
        assert all(np.array_equal(a, b) for a, b in zip(fr_ref, fr_our)) and fr_our[0].shape == (28, 70, 3)
    # save_folder: the reference writes through imageio.imsave, ours through PIL -- same files, same pixels
    saved = {}
    sys.modules["imageio"].imsave = lambda p, x: saved.__setitem__(os.path.basename(p), x.copy())
    vid = torch.rand(1, 3, 3, 24, 32, generator=g)
    (tmp_path / "f").mkdir()
    ref.save_folder(vid.clone(), str(tmp_path / "f"))
    util.save_folder(vid.clone(), str(tmp_path / "f"))
    assert sorted(saved) == sorted(os.listdir(tmp_path / "f")) == ["00000.png", "00001.png", "00002.png"]
    for name, x in saved.items():
        assert np.array_equal(np.array(Image.open(tmp_path / "f" / name)), x)
    import random
    util.seed_everything(7)
    a = (random.random(), np.random.rand(), torch.rand(1).item())
    ref.seed_everything(7)
    assert a == (random.random(), np.random.rand(), torch.rand(1).item())
    # no imageio (this image): the OpenCV writer
    monkeypatch.setitem(sys.modules, "imageio", None)
    cv2 = pytest.importorskip("cv2")
    out = util.write_video(str(tmp_path / "c" / "clip.mp4"), [np.full((24, 32, 3), 40 * i, np.uint8) for i in range(5)], fps=8)
    cap, n = cv2.VideoCapture(out), 0
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        n, shape = n + 1, frame.shape
    assert n == 5 and shape == (24, 32, 3)


def test_inversion_helpers_keep_the_reference_call_forms(monkeypatch):
    """ddim_inversion.next_step in the reference's own call form (model_output, timestep, sample, scheduler),
    init_prompt and get_noise_pred_single (inversion_tools/ddim_inversion.py:171-212) against the reference's functions
    (build container only), on the diffusers-shim scheduler and stand-in tokenizer / text encoder / UNet."""
    import sys
    import types
    if not os.path.isdir("/root/reference"):
        pytest.skip("needs the reference checkout (build container only)")
    import _torch_ops
    _torch_ops.install(monkeypatch)
    monkeypatch.setitem(sys.modules, "imageio", types.SimpleNamespace())
    monkeypatch.setitem(sys.modules, "decord", types.SimpleNamespace(bridge=types.SimpleNamespace(set_bridge=lambda n: None)))
    monkeypatch.syspath_prepend("/root/reference")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "oracle", "_shim"))
    shimmed = lambda: [k for k in sys.modules if k.split(".")[0] in ("src", "inversion_tools", "diffusers")]
    for m in shimmed():
        monkeypatch.delitem(sys.modules, m)
    import inversion_tools.ddim_inversion as ref
    from diffusers import DDIMScheduler as ShimScheduler
    from diffusers.schedulers import SD15_SCHEDULER_CONFIG
    for m in shimmed():                       # the test-only shim must not outlive this test: monkeypatch removes at teardown
        mod = sys.modules.pop(m)              # what it is asked to set now
        monkeypatch.setitem(sys.modules, m, mod)
    from univst_b200 import ddim_inversion as ours
    from univst_b200.scheduler import DDIMScheduler

    g = torch.Generator().manual_seed(0)
    x, eps = torch.randn(1, 4, 3, 8, 8, generator=g), torch.randn(1, 4, 3, 8, 8, generator=g)
    rs, os_ = ShimScheduler(**SD15_SCHEDULER_CONFIG), DDIMScheduler.sd15()
    for n in (10, 50):
        rs.set_timesteps(n)
        os_.set_timesteps(n)
        for t in (1, 481, 981) if n == 50 else (1, 501, 901):
            want = ref.next_step(eps, t, x, rs)
            got = ours.next_step(eps.half(), t, x.half(), os_)
            assert got.dtype == torch.float16 and (got.float() - want).norm() / want.norm() < 1e-3, (n, t)

    class Tok:
        model_max_length = 7
        def __call__(self, texts, padding=None, max_length=None, truncation=None, return_tensors=None):
            ids = torch.tensor([[len(t) + i for i in range(max_length)] for t in texts])
            return types.SimpleNamespace(input_ids=ids)
    enc = lambda ids: (torch.sin(ids.float())[..., None] * torch.arange(1, 5.0),)
    pipe = types.SimpleNamespace(tokenizer=Tok(), text_encoder=enc, device="cpu")
    assert torch.equal(ours.init_prompt(pipe, "a duck"), ref.init_prompt(pipe, "a duck"))
    assert ours.init_prompt(pipe, "a duck").shape == (2, 7, 4)

    seen = {}
    def unet(lat, t, encoder_hidden_states=None, ft_indices=None, ft_timesteps=None, ft_path=None):
        seen.update(t=t, ft=(ft_indices, ft_timesteps, ft_path))
        return {"sample": lat * 2 + encoder_hidden_states.sum()}
    pipe.unet = unet
    a = ours.get_noise_pred_single(pipe, x, 481, torch.ones(1, 2, 2), ft_indices=[2], ft_timesteps=[301], ft_path="p")
    first = dict(seen)
    b = ref.get_noise_pred_single(pipe, x, 481, torch.ones(1, 2, 2), ft_indices=[2], ft_timesteps=[301], ft_path="p")
    assert torch.equal(a, b) and first == seen == {"t": 481, "ft": ([2], [301], "p")}
