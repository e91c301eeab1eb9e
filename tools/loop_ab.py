"""In-situ A/B of attention variants inside the real 50-step three-branch loop (the GPU sits at its power cap there,
so kernel rankings can differ from isolated timings).  usage: loop_ab.py [variants ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from univst_b200 import ops, pnp_utils
from univst_b200.pipeline import SpatioTemporalStableDiffusionPipeline
from univst_b200.unet import SD15_CONFIG, UNetPseudo3DConditionModel
from univst_b200.weights import random_state_dict

variants = [int(a) for a in sys.argv[1:]] or [9, 16, 19]
dev = torch.device("cuda", 0)
unet = UNetPseudo3DConditionModel(random_state_dict(SD15_CONFIG, seed=33, device=dev), SD15_CONFIG, device=dev)
pipe = SpatioTemporalStableDiffusionPipeline(unet)
pnp_utils.register_spatial_attention_pnp(pipe)
clip = bench.synthetic_clip()
c = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in clip.items()}


def stylize():
    z_T = ops.latent_adain(c["traj_c"][50], c["traj_s"][50])
    return pipe.video_style_transfer("", num_inference_steps=50, latents=z_T, content_inv_path=c["traj_c"],
                                     style_inv_path=c["traj_s"], mask_path=c["mask"], prompt_embeds=c["ctx"]).latents


stylize()
torch.cuda.synchronize()
for rnd in range(2):
    for var in variants:
        ops.attention_tune(var, 1, -1)
        ops.profile_start({"sc_attention", "gemm", "conv3x3"})
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = stylize()
        e1.record()
        torch.cuda.synchronize()
        prof = ops.profile_stop()
        a2 = [m for m, meta in prof["sc_attention"] if meta[3] == 4096 and meta[4] == 8192]
        a3 = [m for m, meta in prof["sc_attention"] if meta[3] == 4096 and meta[4] == 12288]
        tot = {k: sum(m for m, _ in v) for k, v in prof.items()}
        print(f"round {rnd} variant {var:2d}: clip {e0.elapsed_time(e1):8.1f} ms | 64x64 KV=2N {sum(a2) / len(a2):.3f} ms "
              f"(first 10: {sum(a2[:10]) / 10:.3f}, last 10: {sum(a2[-10:]) / 10:.3f}) KV=3N {sum(a3) / len(a3):.3f} ms | "
              f"sums: attn {tot['sc_attention']:.0f} gemm {tot['gemm']:.0f} conv {tot['conv3x3']:.0f} ms", flush=True)
ops.attention_tune(-1, -1, -1)
