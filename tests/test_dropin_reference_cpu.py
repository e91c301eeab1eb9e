"""Drop-in check with the REFERENCE's own code in the driver's seat (build container only: needs /root/reference; skipped
where it does not exist, e.g. on the GPU box): the reference's ``SpatioTemporalStableDiffusionPipeline.video_style_transfer``
(stable_diffusion.py:631-780), its own ``register_spatial_attention_pnp`` / ``register_time`` (pnp_utils.py:7-111) and its
own scheduler call sequence run UNMODIFIED on a pipeline whose UNet was swapped by ``univst_b200.accelerate(pipe)``.  The
kernels behind the swapped UNet are replaced by their plain-torch definitions (tests/_torch_ops.py) because this host has
no GPU: what is pinned here is the seam -- module tree, patch protocol, call signatures, ``.sample`` / ``["sample"]`` --
not the kernels (those: tests/test_*_gpu.py).  Runs in a subprocess so that the test-only diffusers shim never leaks into
the other tests' imports."""
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
    torch.cuda.get_device_name = lambda *a, **k: "cpu-shim"
    from PIL import Image
    from backbones.video_diffusion_sd import pnp_utils as ref_pnp                        # the reference's own patch code
    from backbones.video_diffusion_sd.models.unet_3d_condition import UNetPseudo3DConditionModel as RefUNet
    from backbones.video_diffusion_sd.pipelines.stable_diffusion import SpatioTemporalStableDiffusionPipeline as RefPipe
    from diffusers import DDIMScheduler
    from diffusers.schedulers import SD15_SCHEDULER_CONFIG
    from oracle import pipeline_oracle as po, unet_oracle as uo
    import univst_b200, _torch_ops

    class MP:                                      # pytest-monkeypatch stand-in for _torch_ops.install
        def setattr(self, obj, name, val): setattr(obj, name, val)
    _torch_ops.install(MP())

    g = torch.load(os.path.join(ROOT, "tests", "golden", "style_transfer_tiny.pt"), weights_only=True)
    cfg, n = uo.TINY_CONFIG, g["n"]
    m = RefUNet(block_out_channels=cfg["block_out_channels"], attention_head_dim=cfg["attention_head_dim"],
                cross_attention_dim=cfg["cross_attention_dim"], sample_size=8).eval()
    m.load_state_dict(uo.seeded_state_dict(cfg, seed=33))
    traj_c, traj_s, mask_u8 = po.synthetic_inputs(g["seed"], g["F"], g["hw"], n)
    with tempfile.TemporaryDirectory() as tmp:
        cdir, sdir, mdir = (os.path.join(tmp, d) for d in ("c", "s", "m"))
        for d in (cdir, sdir, mdir):
            os.makedirs(d)
        for k in range(1, n + 1):
            torch.save(traj_c[k], os.path.join(cdir, f"ddim_latents_{{k}}.pt"))
            torch.save(traj_s[k], os.path.join(sdir, f"ddim_latents_{{k}}.pt"))
        for f in range(g["F"]):
            Image.fromarray(mask_u8[f], mode="L").save(os.path.join(mdir, "%05d.png" % f))
        pipe = RefPipe.__new__(RefPipe)                       # the reference pipeline object, third-party members stubbed
        pipe.unet, pipe.scheduler = m, DDIMScheduler(**SD15_SCHEDULER_CONFIG)
        pipe._encode_prompt = lambda *a, **k: g["emb"]
        final = []
        pipe.decode_latents = lambda lat: (final.append(lat.clone()), np.zeros((1, 1, 1, 1, 3), np.float32))[1]
        assert univst_b200.accelerate(pipe, device="cpu") is pipe and type(pipe.unet).__module__ == "univst_b200.unet"
        ref_pnp.register_spatial_attention_pnp(pipe)          # reference code patching OUR handles
        z_T = ref_pnp.latent_adain(traj_c[n], traj_s[n])
        rec = {{}}
        with torch.no_grad():
            pipe.video_style_transfer("", latents=z_T, num_inference_steps=n, content_inv_path=cdir, style_inv_path=sdir,
                                      mask_path=mdir, callback=lambda i, t, lat: rec.__setitem__(i, lat.clone()))
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
    worst = max(rel(rec[i], ref) for i, ref in g["steps"].items())
    print("DROPIN steps", worst, "final", rel(final[0], g["final"]))
    assert worst < 1e-2 and rel(final[0], g["final"]) < 1e-2
    # the reference's ddim_inversion reads the output by item: unet(...)["sample"] (ddim_inversion.py:209-211)
    out = pipe.unet(torch.cat([traj_c[5], traj_s[5], z_T]).half(), 481, encoder_hidden_states=g["emb"].repeat(3, 1, 1))
    assert out["sample"] is out.sample
    print("DROPIN OK")
""")


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")
def test_reference_pipeline_drives_the_accelerated_unet(tmp_path):
    script = tmp_path / "dropin.py"
    script.write_text(SCRIPT.format(root=ROOT, ref=REF))
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0 and "DROPIN OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
