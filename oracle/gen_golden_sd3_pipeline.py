"""Generate ``tests/golden/sd3_pipeline.pt`` by running the REFERENCE's own SD3 pipeline loops
(backbones/video_diffusion_sd3/pipelines/custom_pipeline.py: ``generate_eta_values``, ``reconstruction``,
``video_style_transfer``) on an instance whose third-party members are the stand-ins of oracle/sd3_pipeline_oracle.py
and whose inputs come through the reference's own loaders (``ddim_latents_{k}.pt`` files, PNG masks).

``video_style_transfer`` reads the undefined name ``ddim_inv_latents_at_t`` (:316); it is defined here, step by step, to
the content inversion latent the loop has just loaded (what the SD / AnimateDiff loops blend at the same place).
Build container only.

    python oracle/gen_golden_sd3_pipeline.py
"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)


def main():
    from PIL import Image
    from backbones.video_diffusion_sd3.pipelines import custom_pipeline as cp
    from backbones.video_diffusion_sd3.pnp_utils import latent_adain
    from oracle import sd3_pipeline_oracle as so

    n, frames, channels, hw = 10, 16, 16, 8   # load_mask reads 16 frames (src/util.py:133)
    traj_c, traj_s, mask = so.synthetic_inputs(seed=5, frames=frames, channels=channels, hw=hw, n=n)
    tmp = tempfile.mkdtemp()
    cdir, sdir, mdir = (os.path.join(tmp, d) for d in ("c", "s", "m"))
    for d in (cdir, sdir, mdir):
        os.makedirs(d)
    for k, t in traj_c.items():
        torch.save(t, os.path.join(cdir, f"ddim_latents_{k}.pt"))
        torch.save(traj_s[k], os.path.join(sdir, f"ddim_latents_{k}.pt"))
    for f in range(frames):
        Image.fromarray(mask[f], mode="L").save(os.path.join(mdir, "%05d.png" % f))

    pipe = cp.CustomStableDiffusion3Pipeline.__new__(cp.CustomStableDiffusion3Pipeline)
    pipe.scheduler = so.FakeFlowMatchScheduler()
    pipe.transformer = so.FakeTransformer(channels=channels, seed=13)
    pipe.encode_prompt = so.fake_encode_prompt()
    pipe.check_inputs = lambda *a, **k: None
    pipe.default_sample_size, pipe.vae_scale_factor = hw, 8
    pipe._execution_device = pipe.device = torch.device("cpu")
    pipe.vae = type("V", (), {"config": so._Config(scaling_factor=1.0, shift_factor=0.0), "decode": staticmethod(lambda x: (x,))})()
    pipe.image_processor = type("P", (), {"postprocess": staticmethod(lambda x, output_type=None: x)})()

    real_load = cp.load_ddim_latents_at_t

    def load_and_define(k, path):   # the name :316 reads, see the module docstring
        t = real_load(k, path)
        if path == cdir:
            cp.ddim_inv_latents_at_t = t
        return t
    cp.load_ddim_latents_at_t = load_and_define

    out = {"n": n, "frames": frames, "channels": channels, "hw": hw, "input_seed": 5, "transformer_seed": 13, "cases": {}}
    pipe.scheduler.set_timesteps(n)
    ts = pipe.scheduler.timesteps
    out["eta_values"] = {trend: [float(e) for e in pipe.generate_eta_values(ts, 2, 7, 0.85, trend)]
                         for trend in ("constant", "linear_increase", "linear_decrease")}
    with torch.no_grad():
        z_T = latent_adain(traj_c[50], traj_s[50])
        out["z_T"] = z_T.clone()
        for name, mpath in (("masked", mdir), ("unmasked", None)):
            rec = []
            pipe.transformer.calls.clear()
            res = pipe.video_style_transfer("", latents=z_T.clone(), img_latents=traj_c[0].clone(), num_inference_steps=n,
                                            content_inv_path=cdir, style_inv_path=sdir, mask_path=mpath, eta_base=0.85,
                                            eta_trend="constant", start_step=5, end_step=8, output_type="latent",
                                            callback_on_step_end=lambda p, i, t, kw: (rec.append(kw["latents"].clone()), {})[1])
            out["cases"][name] = {"final": res.images.clone(), "steps": {i: rec[i] for i in (0, 5, 8)},
                                  "idx_seen": [c[0] for c in pipe.transformer.calls]}
        rc = pipe.reconstruction(traj_c[0].clone(), traj_c[50].clone(), 0.9, "linear_decrease", 0, 6, prompt="",
                                 DTYPE=torch.float32, num_inference_steps=n)
        out["cases"]["reconstruction"] = {"final": rc.clone()}
    path = os.path.join(ROOT, "tests", "golden", "sd3_pipeline.pt")
    torch.save(out, path)
    print("wrote sd3_pipeline.pt", os.path.getsize(path) // 1024, "KiB;",
          {k: float(v["final"].abs().mean()) for k, v in out["cases"].items()})


if __name__ == "__main__":
    main()
