"""The SD3 inversion front doors -- ``content_inversion_reconstruction`` / ``style_inversion_reconstruction`` of
inversion_tools/flow_inversion.py:16-120 -- against the REFERENCE's own two functions (build container only) on the
stand-in third-party members of oracle/sd3_pipeline_oracle.py plus a stand-in VAE / image processor: frame loading and
normalisation, posterior sample -> (x - shift) x scale, the ``ddim_latents_{k}.pt`` files of the RF-Solver inversion,
the reconstruction (eta 0.85 on steps 25..38) and the frames handed to the video writer.

Reference defect pinned here: with ``is_rf_solver=False`` the front doors pass ``DTYPE=`` to ``rf_inversion`` (:46, :97),
which has no such parameter -> TypeError.  Ours runs that branch; it is compared with the reference's ``rf_inversion`` +
``reconstruction`` called the way the front door evidently meant to.  In a subprocess (test-only shims)."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

SCRIPT = textwrap.dedent("""
    import os, sys, tempfile, types
    import numpy as np, torch
    ROOT, REF = {root!r}, {ref!r}
    sys.path[:0] = [ROOT, REF, os.path.join(ROOT, "oracle", "_shim"), os.path.join(ROOT, "tests")]
    torch.Tensor.cuda = lambda self, *a, **k: self            # the reference hard-codes .cuda() (flow_inversion.py:26,28,78)
    import imageio
    written = []
    imageio.mimsave = lambda path, frames, fps=None: written.append((os.path.basename(path), [np.asarray(f).copy() for f in frames], fps))
    from PIL import Image
    from backbones.video_diffusion_sd3.pipelines import custom_pipeline as cp
    import inversion_tools.flow_inversion as ref_fi
    ref_fi.export_to_video = lambda images, path, fps=None: written.append((os.path.basename(path), [np.asarray(f).copy() for f in images], fps))
    from oracle import sd3_pipeline_oracle as so
    import _torch_ops
    from univst_b200 import flow_inversion as our_fi
    from univst_b200.sd3_pipeline import CustomStableDiffusion3Pipeline as OurPipe

    class MP:
        def setattr(self, obj, name, val): setattr(obj, name, val)
    _torch_ops.install(MP())

    F, HW, C, steps = 4, 64, 16, 40                              # the front doors hard-code the eta window 25..38
    g = torch.Generator().manual_seed(2)
    proj = torch.randn(C, 3, generator=g) * 0.5

    class FakeVAE:                                              # third-party member: 8 x 8 pooling + projection stands in
        config = so._Config(scaling_factor=1.5305, shift_factor=0.0609)
        def encode(self, x):
            mean = torch.einsum("kc,nchw->nkhw", proj, torch.nn.functional.avg_pool2d(x.float(), 8))
            return types.SimpleNamespace(latent_dist=types.SimpleNamespace(
                sample=lambda generator=None: (mean + 0.05 * torch.randn(mean.shape)).to(x.dtype)))
        def decode(self, z, return_dict=True):
            px = torch.einsum("kc,nkhw->nchw", proj, z.float()).clamp(-1, 1)
            return (torch.nn.functional.interpolate(px, scale_factor=8, mode="nearest"),)
    class FakeProcessor:
        @staticmethod
        def postprocess(x, output_type="pil"):
            u8 = ((x.float() / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).numpy()
            return [Image.fromarray(f) for f in u8]

    def members(dtype):
        tr = so.FakeTransformer(channels=C, seed=13)
        class T:                                                 # the stand-in field evaluated in fp32 whatever comes in
            config = tr.config
            def __call__(self, hidden_states, **kw):
                kw.setdefault("joint_attention_kwargs", {{"idx": kw.pop("idx", None)}})
                for k in ("ft_indices", "ft_timesteps", "ft_path"):
                    kw.pop(k, None)
                return (tr(hidden_states.float(), **kw)[0].to(dtype),)
        return T(), so.FakeFlowMatchScheduler(), so.fake_encode_prompt(dtype=dtype)

    ref_pipe = cp.CustomStableDiffusion3Pipeline.__new__(cp.CustomStableDiffusion3Pipeline)
    ref_pipe.transformer, ref_pipe.scheduler, ref_pipe.encode_prompt = members(torch.float32)
    ref_pipe.check_inputs = lambda *a, **k: None
    ref_pipe.default_sample_size, ref_pipe.vae_scale_factor = HW // 8, 8
    ref_pipe._execution_device = ref_pipe.device = torch.device("cpu")
    ref_pipe.vae, ref_pipe.image_processor = FakeVAE(), FakeProcessor()
    tr16, sch16, enc16 = members(torch.float16)
    our_pipe = OurPipe(types.SimpleNamespace(transformer=tr16, scheduler=sch16, encode_prompt=enc16, device="cpu",
                                             vae=FakeVAE(), image_processor=FakeProcessor()))

    rng = np.random.default_rng(1)
    yy, xx = np.mgrid[0:72, 0:80]
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
    with tempfile.TemporaryDirectory() as tmp:
        fdir = os.path.join(tmp, "frames"); os.makedirs(fdir)
        for f in range(F):
            img = np.stack([127 + 100 * np.sin((xx + 3 * f) / 11.0), 127 + 100 * np.cos(yy / 9.0), 40 + 2 * xx], -1)
            Image.fromarray((img + rng.integers(0, 9, img.shape)).clip(0, 255).astype(np.uint8)).save(os.path.join(fdir, "%05d.png" % f))
        style = os.path.join(tmp, "style.png")
        Image.fromarray(rng.integers(0, 255, (50, 90, 3)).astype(np.uint8)).save(style)

        def run(side, kind, solver):
            inv, rec = os.path.join(tmp, kind, side + str(solver), "inv"), os.path.join(tmp, kind, side + str(solver), "rec")
            os.makedirs(inv); os.makedirs(rec)
            torch.manual_seed(11)
            mod, pipe, dt = (ref_fi, ref_pipe, torch.float32) if side == "ref" else (our_fi, our_pipe, torch.float16)
            if kind == "content":
                mod.content_inversion_reconstruction(pipe, fdir, inv, rec, F, HW, HW, steps, dt, None, None, None, is_rf_solver=solver)
            else:
                mod.style_inversion_reconstruction(pipe, style, inv, rec, F, HW, HW, steps, dt, is_rf_solver=solver)
            return inv

        for kind in ("content", "style"):
            # the reference's rf_inversion branch cannot run (DTYPE= is not a parameter of rf_inversion)
            try:
                run("ref", kind, False)
                raise SystemExit("the reference's rf_inversion branch was expected to raise TypeError")
            except TypeError as e:
                assert "DTYPE" in str(e), e
            a, b = run("ref", kind, True), run("our", kind, True)
            names = sorted(os.listdir(a))
            assert names == sorted(os.listdir(b)) and len(names) == steps + 1, (names, sorted(os.listdir(b)))
            worst = 0.0
            for name in names:
                x, y = torch.load(os.path.join(b, name)), torch.load(os.path.join(a, name))
                assert x.shape == y.shape == (F, C, HW // 8, HW // 8)
                worst = max(worst, rel(x, y))
            (n_ref, fr_ref, fps_ref), (n_our, fr_our, fps_our) = written[-2:]
            assert n_ref == n_our == kind + "_video.mp4" and fps_ref == fps_our == 8 and len(fr_ref) == len(fr_our) == F
            d = np.abs(np.stack(fr_ref).astype(int) - np.stack(fr_our).astype(int))
            print("SD3FRONTDOOR", kind, "rf_solver: latents", worst, "frames mean abs", d.mean(), "max", d.max())
            assert worst < 5e-3 and d.mean() < 1.0 and fr_our[0].shape == (HW, HW, 3) and fr_our[0].dtype == np.uint8

            # the rf_inversion branch: ours vs the reference's own rf_inversion + reconstruction on the same latents
            inv_dir = run("our", kind, False)
            x0 = torch.load(os.path.join(a, "ddim_latents_0.pt"))                       # the reference front door's encode
            assert rel(torch.load(os.path.join(inv_dir, "ddim_latents_0.pt")), x0) < 1e-3
            with tempfile.TemporaryDirectory() as t2:
                torch.manual_seed(11)
                inv_ref = ref_fi.rf_inversion(ref_pipe, x0.clone(), prompt="", gamma=0.0, num_inference_steps=steps, inversion_path=t2)
                worst = max(rel(torch.load(os.path.join(inv_dir, n)), torch.load(os.path.join(t2, n))) for n in sorted(os.listdir(t2)))
                assert sorted(os.listdir(t2)) == sorted(os.listdir(inv_dir))
            imgs = ref_pipe.reconstruction(prompt="", img_latents=x0, inversed_latents=inv_ref, eta_base=0.85, eta_trend="constant",
                                           start_step=25, end_step=39, guidance_scale=1.0, DTYPE=torch.float32,
                                           num_inference_steps=steps)
            d = np.abs(np.stack([np.asarray(i) for i in imgs]).astype(int) - np.stack(written[-1][1]).astype(int))
            print("SD3FRONTDOOR", kind, "rf_inversion: latents", worst, "frames mean abs", d.mean(), "max", d.max())
            assert worst < 5e-3 and d.mean() < 1.0
    print("SD3FRONTDOOR OK")
""")


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")
def test_sd3_inversion_front_doors_match_the_reference(tmp_path):
    script = tmp_path / "sd3_frontdoor.py"
    script.write_text(SCRIPT.format(root=ROOT, ref=REF))
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0 and "SD3FRONTDOOR OK" in out.stdout, out.stdout[-1500:] + out.stderr[-3000:]
    print("\n".join(l for l in out.stdout.splitlines() if l.startswith("SD3FRONTDOOR")))
