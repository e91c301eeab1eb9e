"""Generate ``tests/golden/sd3_processors.pt`` by running the REFERENCE's own SD3 attention processors
(backbones/video_diffusion_sd3/pnp_utils.py: CrossFrameProcessor, AttentionShiftProcessor) on a stand-in ``attn``
module that exposes exactly what they touch (to_q/k/v, add_q/k/v_proj, norm_q/k, norm_added_q/k, to_out, to_add_out,
heads, context_pre_only).  The RMS norms restate diffusers' RMSNorm [third party].  ``thresh2`` (read at :185 but never
set in the reference) is set to ``eta2``.  Build container only.

    python oracle/gen_golden_sd3.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)


class RMSNorm(torch.nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.weight, self.eps = torch.nn.Parameter(torch.ones(dim)), eps

    def forward(self, x):
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps) * self.weight


def main():
    from backbones.video_diffusion_sd3 import pnp_utils as ref
    from oracle import sd3_oracle as so

    heads, d, N, L, Fr = 2, 64, 16, 5, so.CLIP_LENGTH
    C = heads * d
    attn = torch.nn.Module()
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_add_out"):
        setattr(attn, n, torch.nn.Linear(C, C))
    attn.to_out = torch.nn.ModuleList([torch.nn.Linear(C, C), torch.nn.Dropout(0.0)])
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        setattr(attn, n, RMSNorm(d))
    attn.heads, attn.context_pre_only = heads, False
    w = so.seeded_attn_weights(C, heads, seed=3)
    attn.load_state_dict(w)
    hidden, enc = so.synthetic_inputs(41, N, L, C)
    keep = so.GOLDEN_IMAGES   # the committed outputs cover frames 0, 1, 2, 15 of every branch (fixture size)
    out = {"input_seed": 41, "N": N, "L": L, "heads": heads, "seed": 3, "keep": keep, "cases": {}}
    with torch.no_grad():
        h, e = ref.CrossFrameProcessor()(attn, hidden[:Fr].clone(), enc[:Fr].clone())
        out["cases"]["cross_frame"] = (h.clone(), e.clone())   # one branch, all 16 frames
        proc = ref.AttentionShiftProcessor(0.0, 0.6)
        proc.thresh2 = proc.eta2   # never set in the reference (AttributeError at :185); the evident intent
        for idx in (0, 15, 30, 31):
            h, e = proc(attn, hidden.clone(), enc.clone(), idx=idx)
            out["cases"][f"shift_idx{idx}"] = (h[keep].clone(), e[keep].clone())
        g = torch.Generator().manual_seed(42)
        cnt, sty = torch.randn(2, heads, N, d, generator=g) * 1.3 + 0.2, torch.randn(2, heads, N, d, generator=g) * 0.6 - 0.1
        out["adain"] = {"cnt": cnt, "sty": sty, "out": ref.attention_adain(cnt, sty)}
    path = os.path.join(ROOT, "tests", "golden", "sd3_processors.pt")
    torch.save(out, path)
    print("wrote sd3_processors.pt", os.path.getsize(path) // 1024, "KiB")
    gen_rf_loops()


def gen_rf_loops():
    """The reference's own rf_inversion / rf_solver (inversion_tools/flow_inversion.py) on the stand-in pipeline."""
    import tempfile

    import inversion_tools.flow_inversion as fi
    from oracle import rf_oracle as ro
    n = 10
    g = torch.Generator().manual_seed(21)
    x0 = torch.randn(4, 16, 8, 8, generator=g)
    out = {"n": n, "x0_seed": 21, "gamma": 0.5}
    with tempfile.TemporaryDirectory() as tmp:
        pipe = ro.FakePipeline()
        torch.manual_seed(77)                       # rf_inversion draws its target with torch.randn_like (:151)
        noise = torch.randn_like(x0)
        torch.manual_seed(77)
        fi.rf_inversion(pipe, x0.clone(), "", gamma=0.5, num_inference_steps=n, inversion_path=tmp)
        out["rf_inversion"] = torch.stack([torch.load(os.path.join(tmp, f"ddim_latents_{k}.pt")) for k in range(n + 1)])
        out["noise"] = noise
        out["files"] = sorted(os.listdir(tmp))
    with tempfile.TemporaryDirectory() as tmp:
        pipe = ro.FakePipeline()
        fi.rf_solver(pipe, x0.clone(), "", num_inference_steps=n, inversion_path=tmp)
        out["rf_solver"] = torch.stack([torch.load(os.path.join(tmp, f"ddim_latents_{k}.pt")) for k in range(n + 1)])
        out["solver_calls"] = len(pipe.calls)
    path = os.path.join(ROOT, "tests", "golden", "rf_inversion.pt")
    torch.save(out, path)
    print("wrote rf_inversion.pt", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
