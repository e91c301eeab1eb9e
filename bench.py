#!/usr/bin/env python
"""Benchmark of the hot path: stylized frames/sec of the three-branch 50-step DDIM denoising loop (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one complete pass of the hot path over one synthetic clip: 50 DDIM steps of the three-branch
[content, style, edit] SD-1.5 UNet on 16 frames of 512x512 (latents 64x64), mask blending and late latent AdaIN
included (configs[1] of BASELINE.json).  ``value`` = frames / second with all inputs resident in HBM; ``e2e`` = the
same through the public pipeline call with pinned HOST buffers (trajectories, mask, prompt embeddings copied in and the
stylized latents read back inside the timed region).

N = 1: one clip on one GPU.  N > 1 (one process per GPU under torchrun): ONE clip whose 16 frames are sharded over the N
ranks -- strong scaling, the split BASELINE.json's north_star names: per attn1 layer the boundary frame's K/V and the
clip's first frame's K/V are stored into the peers' banks over NVLink, every cross-frame GroupNorm exchanges 3 x 32 x 2
partial sums, the noise prediction is stored into every rank's buffer (csrc/xrank.cu: no collective-library call), and the
forward is a replayed CUDA graph.  The result is asserted against the single-GPU evaluation of the same clip.  The
clip-parallel replica rate (each rank its own clip, no communication) is reported as a side number.

The ``roofline`` object describes the dominant kernel (fused sparse-causal attention at the 64x64 level, patched
KV = 2N): algorithmic FLOPs per launch / mean launch duration measured with CUDA events inside the timed region.
``cpu_baseline`` / ``--impl reference``: the reference's algorithm has no CPU entry point of its own and needs
diffusers + model weights that do not exist offline, so the CPU arm is the oracle port (oracle/unet_oracle.py, pinned
to golden vectors of the reference's module code) timed on the host cores on a bounded sample of the same workload.
``config.torch_eager_fp16``: the same oracle code in torch fp16 eager on the GPU (cuDNN / cuBLAS / flash SDPA -- the
library kernels the reference's modules would run), a second reported baseline.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

F_FRAMES, LAT, STEPS_DDIM = 16, 64, 50
# live work of one image (one frame of one branch) in one UNet call, and of a clip
GF_IMAGE = 975.1e9
FLOP_PER_CLIP_REFERENCE = 50 * 48 * GF_IMAGE                       # 2.340 PF: all three branches at every step
# what dead-branch skipping leaves (exact): 26 live steps of 48 images minus the part of the last patched
# transformer + conv_out that runs on the edit branch only (56.87 GF x 32 images), 24 steps of the edit branch alone
FLOP_PER_CLIP_EXECUTED = 26 * (48 * GF_IMAGE - 32 * 56.87e9) + 24 * 16 * GF_IMAGE


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"tflops": p["bf16_tflops_sustained"], "tflops_burst": p["bf16_tflops"], "hbm": p["hbm_gbs"], "src": "measured"}
    except Exception:
        return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            self.t.join(timeout=2)
        return False

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower() == "active" for r in self.rows)]
        # the loop keeps the GPU busy throughout, so every sample is "under load"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synthetic_clip(seed_offset: int = 0, F_FRAMES: int = F_FRAMES, ctx_dim: int = 768):
    """content / style inversion trajectories x_k = sqrt(a_k) x_0 + sqrt(1 - a_k) eps held in
    memory, a moving-disc mask, a fixed (77, ctx_dim) context.  Returned on the host (pinned when CUDA is present)."""
    from univst_b200.scheduler import DDIMScheduler
    sch = DDIMScheduler.sd15()
    sch.set_timesteps(STEPS_DDIM)
    g = lambda s: torch.Generator().manual_seed(s + seed_offset)
    z0_c = torch.randn(1, 4, F_FRAMES, LAT, LAT, generator=g(1234))
    z0_s = torch.randn(1, 4, 1, LAT, LAT, generator=g(4321)).repeat(1, 1, F_FRAMES, 1, 1) \
        + 0.02 * torch.randn(1, 4, F_FRAMES, LAT, LAT, generator=g(4322))
    eps = torch.randn(1, 4, F_FRAMES, LAT, LAT, generator=g(99))
    ts = [int(t) for t in sch.timesteps][::-1]  # 1, 21, ..., 981  -> trajectory index k = 1..50
    traj_c, traj_s = [z0_c], [z0_s]
    for t in ts:
        a = sch.alpha(t)
        traj_c.append(a ** 0.5 * z0_c + (1 - a) ** 0.5 * eps)
        traj_s.append(a ** 0.5 * z0_s + (1 - a) ** 0.5 * eps)
    yy, xx = torch.meshgrid(torch.arange(512), torch.arange(512), indexing="ij")
    mask = torch.stack([(((xx - (256 + 6 * f)) ** 2 + (yy - 256) ** 2) <= 128 ** 2) for f in range(F_FRAMES)]).to(torch.uint8) * 255
    ctx = torch.randn(1, 77, ctx_dim, generator=g(7))
    pin = (lambda t: t.pin_memory()) if torch.cuda.is_available() else (lambda t: t)
    return {"traj_c": [pin(t.half()) for t in traj_c], "traj_s": [pin(t.half()) for t in traj_s], "mask": pin(mask),
            "ctx": pin(ctx.half())}


def h2d_bytes(clip):
    n = sum(t.numel() * t.element_size() for t in clip["traj_c"][1:] + clip["traj_s"][1:])
    return n + clip["mask"].numel() + clip["ctx"].numel() * 2


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
CPU_SAMPLE_FRAMES = 4   # >= 3 so that [previous, first] name two frames other than the query's own for most of the sample


def cpu_reference_step(sample_frames: int, threads: int):
    """One three-branch DDIM step (mask blend -> UNet -> DDIM update) of the oracle port on the host cores, fp32,
    at full SD-1.5 width and 64x64 latents but only ``sample_frames`` frames (cost is linear in frames)."""
    from oracle import unet_oracle as uo
    torch.set_num_threads(threads)
    cfg = uo.SD15_CONFIG
    sd = uo.seeded_state_dict(cfg, seed=33)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 4, sample_frames, LAT, LAT, generator=g)
    ctx = torch.randn(1, 77, 768, generator=g).repeat(3, 1, 1)
    m = (torch.rand(1, 1, sample_frames, LAT, LAT, generator=g) > 0.5).float()

    def step():
        t0 = time.perf_counter()
        with torch.no_grad():
            z = (1 - m) * x[2:3] + m * x[0:1]
            eps = uo.unet_forward(sd, cfg, torch.cat([x[0:1], x[1:2], z]), 481, ctx, patched=True, idx=5)
            a_t, a_p = 0.3, 0.35
            x0 = (z - (1 - a_t) ** 0.5 * eps[2:3]) / a_t ** 0.5
            _ = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps[2:3]
        return time.perf_counter() - t0

    return step


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step = cpu_reference_step(CPU_SAMPLE_FRAMES, threads)
    for _ in range(min(args.warmup, 1)):
        step()
    k = max(1, min(args.steps, 2))  # each CPU step is tens of seconds: bound the run
    dt = sum(step() for _ in range(k)) / k
    fps = CPU_SAMPLE_FRAMES / (STEPS_DDIM * dt)
    sample = (f"{k} three-branch DDIM step(s), 3 x {CPU_SAMPLE_FRAMES} frames at 64x64 latents, full SD-1.5 width, fp32; all "
              "three branches at every step (the reference's own work, no dead-branch skipping)")
    line = {"impl": "reference", "metric": "stylized frames/sec (16x512x512, 50 DDIM steps)", "value": fps,
            "unit": "frames/s", "n_gpus": args.gpus, "steps": k, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SD-v1.5 three-branch localized transfer, 16x512x512, 50 steps",
                       "note": "CPU oracle port; value extrapolated: frames_sample / (50 x seconds per DDIM step)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ GPU library baseline
def torch_eager_forward_baseline(unet, pipe, dev):
    """SURVEY.md 8(d) "library kernel to beat": one three-branch UNet call (3 x 16 x 64 x 64, shift window open) of the
    oracle's restatement of the reference forward in torch fp16 eager on this GPU (cuDNN convolutions, cuBLAS linears,
    flash SDPA) next to this library's forward on the same inputs.  A baseline leg: not on the product path."""
    from oracle import unet_oracle as uo
    from univst_b200 import pnp_utils
    from univst_b200.weights import random_state_dict
    cfg = uo.SD15_CONFIG
    sd16 = random_state_dict(cfg, seed=33, device=dev)   # the same seeded fp16 weights the library's UNet was packed from
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.randn(3, 4, F_FRAMES, LAT, LAT, device=dev, generator=g).half()
    ctx = torch.randn(3, 77, cfg["cross_attention_dim"], device=dev, generator=g).half()
    pnp_utils.register_time(pipe, 5)
    torch.backends.cudnn.benchmark = True

    def med(fn, iters=5):
        for _ in range(2):
            out = fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return out, statistics.median(ts)

    with torch.no_grad():
        ours, ms_ours = med(lambda: unet(x, 881, encoder_hidden_states=ctx).sample)
        eager, ms_eager = med(lambda: uo.unet_forward(sd16, cfg, x, 881, ctx, patched=True, idx=5))
    rel = float((ours.float() - eager.float()).norm() / eager.float().norm())
    return {"ms_per_forward": ms_eager, "ms_per_forward_univst_b200": ms_ours, "speedup": ms_eager / ms_ours,
            "rel_l2_between": rel, "note": "three-branch call, 3 x 16 frames at 64x64 latents, shift window open; the same "
            "seeded weights; oracle code in torch " + torch.__version__ + " fp16 eager (cuDNN / cuBLAS / flash SDPA)"}


def attention_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture -- only while the kernel source is the
    one that capture was taken from (profiles/attention_traffic.json records its hash)."""
    try:
        with open(os.path.join(ROOT, "profiles", "attention_traffic.json")) as f:
            t = json.load(f)
        h = hashlib.sha256()
        for name in ("attention_tc.cu", "ptx.cuh"):
            with open(os.path.join(ROOT, "univst_b200", "csrc", name), "rb") as f:
                h.update(f.read())
        if t.get("kernel_source_sha256_16") == h.hexdigest()[:16]:
            return t["dram_bytes_per_launch"], t.get("source")
    except Exception:
        pass
    return None, None


def dominant_attention(prof, pk, clocks_mhz):
    """roofline object of the fused SC-attention launches at the 64x64 level with the patched KV = [prev, first]."""
    cand = [(m_, meta) for m_, meta in prof.get("sc_attention", []) if meta[3] == LAT * LAT and meta[4] == 2 * LAT * LAT]
    if not cand:
        return None
    ni = max(meta[0] for _, meta in cand)
    dom = [(m_, meta) for m_, meta in cand if meta[0] == ni]
    NI, H, d, N, Nkv = dom[0][1]
    flops = 4.0 * N * Nkv * H * d * NI
    avg_ms = sum(m_ for m_, _ in dom) / len(dom)
    ach = flops / (avg_ms * 1e-3) / 1e12
    traffic, tsrc = attention_traffic() if NI == 3 * F_FRAMES else (None, None)
    clk = clocks_mhz or 1900.0
    # secondary limit of this kernel: one exponential per (query, key, head) on the MUFU pipe, 16 / clock / SM,
    # 4 d = 160 useful FLOP per exponential at head dim 40
    mufu_bound = 16 * 148 * clk * 1e6 * 4 * d / 1e12
    return {"bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
            "traffic": traffic, "traffic_source": tsrc, "mufu_bound_tflops": mufu_bound, "frac_of_mufu_bound": ach / mufu_bound,
            "kernel": f"attention_tc_split_kernel<6,1> (2 query tiles x 128 keys, two softmax threads per row; N={N}, "
                      f"Nkv={Nkv}, H={H}, d={d}, {NI} images)",
            "flop_per_launch": flops, "launches_timed": len(dom), "avg_ms": avg_ms,
            "peak_source": pk["src"] + " (sustained bf16 dense)"}


# ------------------------------------------------------------------------------------------------ BASELINE configs[2]
def run_sd21_smoother(dev, world, rank, args, timed, rel):
    """BASELINE.json configs[2]: SD-2.1 shapes (head dim 64, 1024-wide context, Linear proj_in / proj_out), 32 frames of
    512 x 512 sharded over the ranks, sliding-window flow smoothing on (steps 20..24, stable_diffusion.py:713-758): predicted
    x0 -> temporal VAE decode -> +-2-frame window of flow-warped neighbours -> mask -> VAE encode -> noise recomputed.
    The UNet is frame-sharded; of the smoother legs the temporal decoder is spread by 16-frame chunks (one chunk = one clip to
    its temporal layers, :803-811) and the per-frame encoder by frames, both gathered through peer memory; the window pass,
    sequential in the key frame (:731-747), is evaluated by every rank.
    Flows: synthetic (2.5, -1.25) px translation + sinusoidal field (RAFT's weights are not available offline); VAE: SVD VAE
    shapes with seeded random weights (parity unpinned, univst_b200/vae.py).  Checked against the 1-GPU run of the same clip."""
    import numpy as np
    from univst_b200 import ops, pnp_utils
    from univst_b200.pipeline import SpatioTemporalStableDiffusionPipeline
    from univst_b200.unet import SD21_CONFIG, UNetPseudo3DConditionModel
    from univst_b200.vae import AutoencoderKLTemporalDecoder, random_state_dict as vae_state_dict
    from univst_b200.weights import random_state_dict
    Fr = 32
    unet = UNetPseudo3DConditionModel(random_state_dict(SD21_CONFIG, seed=35, device=dev), SD21_CONFIG, device=dev)
    vae = AutoencoderKLTemporalDecoder(vae_state_dict(seed=55, device=dev), device=dev)
    pipe = SpatioTemporalStableDiffusionPipeline(unet, vae=vae)
    pnp_utils.register_spatial_attention_pnp(pipe)
    c = synthetic_clip(0, F_FRAMES=Fr, ctx_dim=1024)
    c = {k: ([t.to(dev) for t in v] if isinstance(v, list) else v.to(dev)) for k, v in c.items()}
    yy, xx = np.mgrid[0:512, 0:512].astype(np.float32)
    fwd = torch.from_numpy(np.stack([2.5 + 1.5 * np.sin(yy / 37.0), -1.25 + 1.5 * np.cos(xx / 29.0)], -1).astype(np.float32)).to(dev)
    bwd = (-fwd).clone()
    bwd[200:232, 300:332] += 4.0          # a patch that violates forward / backward consistency -> occluded
    flow_fn = lambda key_frame, now_frame: (fwd, bwd)

    def stylize(smoother="pixel"):
        z_T = ops.latent_adain(c["traj_c"][STEPS_DDIM], c["traj_s"][STEPS_DDIM])
        return pipe.video_style_transfer("", num_inference_steps=STEPS_DDIM, latents=z_T, content_inv_path=c["traj_c"],
                                         style_inv_path=c["traj_s"], mask_path=c["mask"], prompt_embeds=c["ctx"],
                                         smoother=smoother, flow_fn=flow_fn).latents

    stylize()
    ref, ms_1 = timed(stylize)
    _, ms_1_plain = timed(lambda: stylize(None))
    res = {"workload": "SD-v2.1 shapes, 32x512x512, 50 steps, sliding-window flow smoothing on steps 20..24",
           "one_gpu": {"ms_per_clip": ms_1, "frames_per_s": Fr / (ms_1 / 1e3), "ms_per_clip_without_smoother": ms_1_plain}}
    if world > 1 and Fr % world == 0:
        import torch.distributed as dist
        ok = torch.ones(1, device=dev)
        try:
            unet.set_frame_sharding(transport="xrank")
        except Exception:  # noqa: BLE001
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not bool(ok.item() > 0):
            unet.set_frame_sharding(transport="nccl", push_halo=False)
        unet.use_cuda_graphs = bool(ok.item() > 0) and not args.no_cuda_graphs
        stylize()
        out, ms_n = timed(stylize, 2)
        if unet._xr is not None:
            unet._xr.check()
        r = rel(out, ref)
        res["frame_sharded"] = {"n_gpus": world, "frames_per_gpu": Fr // world, "ms_per_clip": ms_n / 2,
                                "frames_per_s": Fr / (ms_n / 2e3), "speedup_vs_1gpu_same_box": ms_1 / (ms_n / 2),
                                "rel_l2_vs_1gpu": r, "cuda_graphs": bool(unet.use_cuda_graphs),
                                "note": "UNet frame-sharded; VAE decode by 16-frame chunks round-robin over the ranks, encode by frames, both "
                                        "gathered through peer memory; the (sequential) warp pass on every rank"}
        assert r < 1e-2, f"configs[2]: frame-sharded result differs from one GPU: {res}"
        unet.set_frame_sharding_off()
    return res


# ------------------------------------------------------------------------------------------------ BASELINE configs[4]
def run_sd35_rectified_flow(dev, world, rank, args, timed, rel):
    """BASELINE.json configs[4]: SD-3.5-medium-shaped MMDiT (24 joint blocks, 13 of them with the second, image-only attention;
    1536 wide, 24 heads of 64; 2.24 B parameters, seeded random), 16 frames of 1024 x 1024 (latents 16 x 128 x 128 -> 4096 image
    tokens + 333 text tokens per frame), 28 rectified-flow steps of the three-branch loop (custom_pipeline.py:126-371) with the
    reference's AttentionShiftProcessor in every attention.  N > 1: the 16 frames sharded over the ranks (halo banks for the
    [first, previous] K/V of the cross-frame attention), checked against the 1-GPU run.  The MMDiT blocks are third-party
    diffusers code restated here: parity unpinned outside the processors (univst_b200/sd3_transformer.py)."""
    from types import SimpleNamespace
    from univst_b200 import sd3
    from univst_b200.scheduler import FlowMatchEulerDiscreteScheduler
    from univst_b200.sd3_pipeline import CustomStableDiffusion3Pipeline
    from univst_b200.sd3_transformer import SD35_MEDIUM_CONFIG, SD3Transformer2DModel, random_state_dict
    Fr, lat, steps, L = 16, 128, 28, 333
    tr = SD3Transformer2DModel(random_state_dict(seed=71, device=dev), device=dev)
    host = SimpleNamespace(transformer=tr, scheduler=FlowMatchEulerDiscreteScheduler(shift=3.0), device=dev, vae=None)
    pipe = CustomStableDiffusion3Pipeline(host)
    sd3.register_spatial_attention_pnp(pipe, eta1=0.0, eta2=0.6)
    g = torch.Generator(device=dev).manual_seed(5)
    rn = lambda *shape: torch.randn(*shape, device=dev, generator=g).half()
    x0_c, x0_s, noise = rn(Fr, 16, lat, lat), rn(1, 16, lat, lat).repeat(Fr, 1, 1, 1), rn(Fr, 16, lat, lat)
    # inversion trajectories held in memory, indexed like the reference's files (ddim_latents_{50 - i}.pt, hard-coded 50)
    traj_c = {50 - i: ((1 - s_) * x0_c + s_ * noise).half() for i, s_ in enumerate(torch.linspace(1, 0.02, steps).tolist())}
    traj_s = {50 - i: ((1 - s_) * x0_s + s_ * noise).half() for i, s_ in enumerate(torch.linspace(1, 0.02, steps).tolist())}
    yy, xx = torch.meshgrid(torch.arange(512, device=dev), torch.arange(512, device=dev), indexing="ij")
    mask = torch.stack([(((xx - (256 + 6 * f)) ** 2 + (yy - 256) ** 2) <= 128 ** 2) for f in range(Fr)]).to(torch.uint8) * 255
    emb, pooled = rn(1, L, SD35_MEDIUM_CONFIG["joint_attention_dim"]), rn(1, SD35_MEDIUM_CONFIG["pooled_projection_dim"])

    def stylize():
        return pipe.video_style_transfer(num_inference_steps=steps, latents=traj_c[50].clone(), prompt_embeds=emb,
                                         pooled_prompt_embeds=pooled, output_type="latent", content_inv_path=traj_c,
                                         style_inv_path=traj_s, mask_path=mask, img_latents=x0_c, start_step=5, end_step=12).images

    ref, ms_1 = timed(stylize)      # (the first pass doubles as warm-up of a ~1 min clip: timed once)
    res = {"workload": "SD-3.5-medium shapes, 16x1024x1024, 28 rectified-flow steps, three-branch loop",
           "one_gpu": {"ms_per_clip": ms_1, "frames_per_s": Fr / (ms_1 / 1e3), "finite": bool(torch.isfinite(ref).all()),
                       "note": "first pass after model build (includes one-off workspace allocation)"}}
    if world > 1 and Fr % world == 0:
        tr.set_frame_sharding()
        stylize()
        out, ms_n = timed(stylize)
        tr._xr.check()
        r = rel(out, ref)
        res["frame_sharded"] = {"n_gpus": world, "frames_per_gpu": Fr // world, "ms_per_clip": ms_n,
                                "frames_per_s": Fr / (ms_n / 1e3), "speedup_vs_1gpu_same_box": ms_1 / ms_n, "rel_l2_vs_1gpu": r}
        assert r < 1e-2, f"configs[4]: frame-sharded result differs from one GPU: {res}"
        tr.set_frame_sharding_off()
    return res


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    from univst_b200 import ops, pnp_utils
    from univst_b200.pipeline import SpatioTemporalStableDiffusionPipeline
    from univst_b200.unet import SD15_CONFIG, UNetPseudo3DConditionModel
    from univst_b200.weights import random_state_dict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    if world > 1 and F_FRAMES % world:
        raise SystemExit(f"{F_FRAMES} frames do not shard over {world} GPUs")

    unet = UNetPseudo3DConditionModel(random_state_dict(SD15_CONFIG, seed=33, device=dev), SD15_CONFIG, device=dev)
    pipe = SpatioTemporalStableDiffusionPipeline(unet)
    pnp_utils.register_spatial_attention_pnp(pipe)
    # N > 1: every rank works on rank 0's clip (frames sharded); the replica side number uses a clip per rank
    clip = synthetic_clip(seed_offset=0)

    def to_dev(c=None):
        c = c or clip
        return {"traj_c": [t.to(dev, non_blocking=True) for t in c["traj_c"]],
                "traj_s": [t.to(dev, non_blocking=True) for t in c["traj_s"]],
                "mask": c["mask"].to(dev, non_blocking=True), "ctx": c["ctx"].to(dev, non_blocking=True)}

    def stylize(c, skip=True, p=None):
        p = p or pipe
        z_T = ops.latent_adain(c["traj_c"][STEPS_DDIM], c["traj_s"][STEPS_DDIM])  # run_video_style_transfer_sd.py:57
        return p.video_style_transfer("", num_inference_steps=STEPS_DDIM, latents=z_T, content_inv_path=c["traj_c"],
                                      style_inv_path=c["traj_s"], mask_path=c["mask"], prompt_embeds=c["ctx"],
                                      skip_dead_branches=skip).latents

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k=1):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            out = fn()
        b.record()
        barrier()
        return out, a.elapsed_time(b)

    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
    resident = to_dev()
    extra, fs = {}, None

    if world > 1:
        # ---- single-GPU evaluation of the same clip on every rank: the parity reference and the same-box 1-GPU time
        stylize(resident)
        ref0, ms_single = timed(lambda: stylize(resident))
        # ---- frames sharded over the ranks: peer-memory transport + CUDA graphs when symmetric memory works on every rank
        ok = torch.ones(1, device=dev)
        try:
            unet.set_frame_sharding(transport="xrank")
        except Exception as e:  # noqa: BLE001
            print(f"[rank {rank}] xrank transport unavailable: {e!r}", file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        xrank = bool(ok.item() > 0)
        if not xrank:
            unet.set_frame_sharding(transport="nccl", push_halo=False)
        unet.use_cuda_graphs = xrank and not args.no_cuda_graphs
        fs = {"transport": "peer-memory stores + device-side flags (csrc/xrank.cu), no collective-library call" if xrank
              else "NCCL send/recv + broadcast + all-reduce + all-gather (symmetric memory unavailable on this box)",
              "cuda_graphs": bool(unet.use_cuda_graphs), "split_k_of_few_tile_gemms": bool(getattr(unet, "_split_k", False)),
              "ms_per_clip_1gpu_same_box": ms_single}

    for _ in range(args.warmup):
        out = stylize(resident)
    barrier()
    assert torch.isfinite(out).all(), "non-finite latents"

    # ---- timed region 1: inputs resident in HBM
    xr_before = unet._xr.wait_stats() if unet._xr is not None else None
    n0 = ops.launch_count
    profiling = not unet.use_cuda_graphs     # per-kernel CUDA events cannot be recorded inside a replayed graph
    if profiling:
        ops.profile_start({"sc_attention"})
    with ClockSampler(local) as clocks:
        out, ms = timed(lambda: stylize(resident), args.steps)
    prof = ops.profile_stop() if profiling else {}
    launches = ops.launch_count - n0
    if xr_before is not None:   # device-side counters of the control block: how long this rank sat in cross-rank waits
        xr_after = unet._xr.wait_stats()
        fs["cross_rank_syncs_per_clip"] = (xr_after[0] - xr_before[0]) // args.steps
        fs["cross_rank_wait_ms_per_clip_rank0"] = (xr_after[1] - xr_before[1]) / args.steps

    # ---- timed region 2: end to end through the public call with host buffers
    host_out, ms_e2e = timed(lambda: stylize(to_dev()).cpu(), args.steps)

    if world > 1:
        if unet._xr is not None:
            unet._xr.check()
        fs["rel_l2_vs_single_gpu"] = rel(out, ref0)
        assert fs["rel_l2_vs_single_gpu"] < 5e-3, f"frame-sharded result differs from the single-GPU one: {fs}"
        # one eager pass with per-kernel events for the roofline object (not part of the timed region)
        if unet.use_cuda_graphs:
            unet.use_cuda_graphs = False
            ops.profile_start({"sc_attention"})
            _, fs["ms_per_clip_without_cuda_graphs"] = timed(lambda: stylize(resident))
            prof = ops.profile_stop()
        unet.set_frame_sharding_off()
        unet.use_cuda_graphs = False
        # ---- side number: clip-parallel replicas (each rank its own clip, no communication; weak scaling)
        own = to_dev(synthetic_clip(seed_offset=1000 * rank))
        stylize(own)
        _, ms_rep = timed(lambda: stylize(own), 2)
        fs["clip_parallel_replicas"] = {"ms_per_clip_per_gpu": ms_rep / 2}
    else:
        # ---- extra: the reference's own amount of work (all three branches at every step); must give the same latents
        out_full, ms_full = timed(lambda: stylize(resident, skip=False))
        extra["all_branches_every_step"] = {"frames_per_s": F_FRAMES / (ms_full / 1e3), "ms_per_clip": ms_full,
                                            "edit_latents_bit_identical": bool(torch.equal(out_full, out)),
                                            "flop_per_clip": FLOP_PER_CLIP_REFERENCE}
        # ---- extra: the forward as a replayed CUDA graph (one cudaGraphLaunch per UNet call)
        if not args.no_cuda_graphs:
            unet.use_cuda_graphs = True
            out_g = stylize(resident)
            _, ms_graph = timed(lambda: stylize(resident), 2)
            unet.use_cuda_graphs = False
            unet.drop_cuda_graphs()
            extra["cuda_graphs"] = {"frames_per_s": F_FRAMES / (ms_graph / 2e3), "ms_per_clip": ms_graph / 2,
                                    "edit_latents_bit_identical": bool(torch.equal(out_g, out))}
        if not args.no_extras:
            extra["torch_eager_fp16"] = torch_eager_forward_baseline(unet, pipe, dev)

    # ---- extra (N = 1): 50-step DDIM content inversion of the clip (the stage that produces the trajectories the loop
    # consumes: stock sparse-causal attention in all 16 layers, one branch; inversion_tools/ddim_inversion.py:88-113)
    if world == 1 and not args.no_extras:
        from univst_b200 import ddim_inversion as di
        for tr in unet._all_transformers():   # inversion runs the UNPATCHED model (run_content_inversion_sd.py)
            tr.transformer_blocks[0].attn1.__dict__.pop("_patched", None)
        pipe.scheduler.set_timesteps(STEPS_DDIM)
        z0 = resident["traj_c"][0]
        di.ddim_inversion(pipe, pipe.scheduler, z0, 2, "", None, prompt_embeds=resident["ctx"])   # warm-up
        inv, ms_inv = timed(lambda: di.ddim_inversion(pipe, pipe.scheduler, z0, STEPS_DDIM, "", None, prompt_embeds=resident["ctx"]))
        if bool(torch.isfinite(inv[-1]).all()):
            extra["ddim_inversion_50_steps"] = {"frames_per_s": F_FRAMES / (ms_inv / 1e3), "ms_per_clip": ms_inv,
                                                "note": "content inversion of the same clip, stock [prev, self, first] "
                                                        "attention in all layers, supplementary"}
        pnp_utils.register_spatial_attention_pnp(pipe)

    # ---- extra: the same clip through the AnimateDiff-v2 backbone (BASELINE.json configs[3] architecture: per-frame
    # self-attention + 21 motion modules), AnimationPipeline loop; N > 1: its frames sharded over the ranks (the motion
    # modules exchange frames <-> pixels through peer memory), checked bit-identical against one GPU
    if not args.no_extras:
        from univst_b200.animatediff import AD_SD15_CONFIG, AnimationPipeline, UNet3DConditionModel
        from univst_b200.scheduler import DDIMScheduler
        xr_ok = fs is not None and fs["transport"].startswith("peer")
        del unet, pipe
        torch.cuda.empty_cache()
        unet_ad = UNet3DConditionModel(random_state_dict(AD_SD15_CONFIG, seed=34, device=dev, animatediff=True), AD_SD15_CONFIG, device=dev)
        pipe_ad = AnimationPipeline(unet_ad, DDIMScheduler.animatediff_v2())
        pnp_utils.register_spatial_attention_pnp(pipe_ad)
        stylize(resident, p=pipe_ad)   # warm-up
        out_ad, ms_ad = timed(lambda: stylize(resident, p=pipe_ad))
        ad = {"frames_per_s": F_FRAMES / (ms_ad / 1e3), "ms_per_clip": ms_ad, "finite": bool(torch.isfinite(out_ad).all()),
              "note": "same clip and loop, AnimateDiff-v2 UNet (21 motion modules, per-frame attention), one GPU, supplementary"}
        if world > 1:
            unet_ad.set_frame_sharding(transport="xrank" if xr_ok else "nccl")
            unet_ad.use_cuda_graphs = xr_ok and not args.no_cuda_graphs
            stylize(resident, p=pipe_ad)
            out_ads, ms_ads = timed(lambda: stylize(resident, p=pipe_ad), 2)
            if unet_ad._xr is not None:
                unet_ad._xr.check()
            # everything is frame-local or exchanged verbatim in this backbone: bit-identical to one GPU -- unless the K
            # loops of the few-tile GEMMs are split over the idle SMs (from 4 ranks up), which changes fp32 summation order
            split_k = bool(getattr(unet_ad, "_split_k", False))
            ad["frame_sharded"] = {"frames_per_s": F_FRAMES / (ms_ads / 2e3), "ms_per_clip": ms_ads / 2, "n_gpus": world,
                                   "speedup_vs_1gpu_same_box": ms_ad / (ms_ads / 2),
                                   "bit_identical_to_1gpu": bool(torch.equal(out_ads, out_ad)),
                                   "rel_l2_vs_1gpu": rel(out_ads, out_ad), "split_k": split_k,
                                   "cuda_graphs": bool(unet_ad.use_cuda_graphs)}
            ad["frame_sharded"]["parity_ok"] = bool(ad["frame_sharded"]["rel_l2_vs_1gpu"] < 5e-3
                                                    and (split_k or ad["frame_sharded"]["bit_identical_to_1gpu"]))
            assert ad["frame_sharded"]["parity_ok"], f"AnimateDiff frame-sharded result differs from one GPU: {ad['frame_sharded']}"
            unet_ad.set_frame_sharding_off()   # (also switches the split-K of few-tile GEMMs back off)
        extra["animatediff_v2_backbone"] = ad

    # ---- extra: BASELINE.json configs[2] (SD-2.1 shapes, 32 frames, smoother on) at the world size it names (4 GPUs); one
    # GPU with --config2
    if (world == 4 and not args.no_extras) or args.config2:
        torch.cuda.empty_cache()
        try:   # a supplementary pass must not take the headline line down with it (failures are symmetric across ranks)
            extra["sd21_32_frames_flow_smoothing"] = run_sd21_smoother(dev, world, rank, args, timed, rel)
        except Exception as e:  # noqa: BLE001
            extra["sd21_32_frames_flow_smoothing"] = {"error": repr(e)[:400]}

    # ---- extra: BASELINE.json configs[4] (SD-3.5-medium shapes, rectified flow) at the world size it names (8 GPUs); any N
    # with --config5
    if (world == 8 and not args.no_extras) or args.config5:
        if not args.no_extras:
            unet_ad = pipe_ad = None   # noqa: F841  (free the AnimateDiff model before building the 2.2 B-parameter MMDiT)
        torch.cuda.empty_cache()
        try:
            extra["sd35_medium_rectified_flow"] = run_sd35_rectified_flow(dev, world, rank, args, timed, rel)
        except Exception as e:  # noqa: BLE001
            extra["sd35_medium_rectified_flow"] = {"error": repr(e)[:400]}

    # dominant-kernel time: max over ranks (rank 0 holds the clip's first frames, whose [previous, first] sources collapse
    # to one deduplicated source -- half the keys -- so its launches are not representative)
    roof_local = dominant_attention(prof, peaks(), None)
    t = torch.tensor([ms, ms_e2e, roof_local["avg_ms"] if roof_local else 0.0]
                     + ([fs["clip_parallel_replicas"]["ms_per_clip_per_gpu"]] if fs else []), device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tl = t.tolist()
    ms, ms_e2e, dom_ms = tl[0], tl[1], tl[2]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    ms_per_step = ms / args.steps
    fps = F_FRAMES / (ms_per_step / 1e3)                 # one clip per step at every N (N > 1: its frames are sharded)
    fps_e2e = F_FRAMES / (ms_e2e / args.steps / 1e3)
    clk = clocks.summary()
    roof = dominant_attention(prof, pk, clk["sm_mhz"])
    if roof and world > 1 and dom_ms > 0:
        roof["avg_ms"] = dom_ms
        roof["achieved"] = roof["flop_per_launch"] / (dom_ms * 1e-3) / 1e12
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["frac_of_mufu_bound"] = roof["achieved"] / roof["mufu_bound_tflops"]
        roof["note"] = "launch duration = max over ranks of the per-rank mean (rank 0's sources are deduplicated)"
    attn_ms = sum(m_ for m_, _ in prof.get("sc_attention", [])) / (args.steps if profiling else 1)
    config = {"workload": "SD-v1.5 three-branch localized transfer, 16x512x512, 50 steps, one clip"
                          + (f", frames sharded over {world} GPUs" if world > 1 else ""),
              "weights": "random init (seed 33), SD-1.5 UNet shapes", "l2": "working set per UNet call ~5 GiB >> 126 MB L2",
              "skip_dead_branches": True,
              "parallelism": f"frame-sharded x{world} (one clip, {F_FRAMES // world} frames per GPU)" if world > 1 else "1 GPU",
              "flop_per_clip_executed": FLOP_PER_CLIP_EXECUTED, "flop_per_clip_reference_equivalent": FLOP_PER_CLIP_REFERENCE,
              "whole_loop_tensor_frac": FLOP_PER_CLIP_EXECUTED / (ms_per_step / 1e3) / 1e12 / (pk["tflops"] * world),
              "sc_attention_ms_per_clip" + ("" if world == 1 else "_per_gpu"): attn_ms}
    config.update(extra)
    if fs is not None:
        fs.pop("clip_parallel_replicas")
        fs["speedup_vs_1gpu_same_box"] = fs["ms_per_clip_1gpu_same_box"] / ms_per_step
        fs["clip_parallel_replicas"] = {"frames_per_s": world * F_FRAMES / (tl[3] / 1e3), "scaling": "weak",
                                        "note": "side number: each rank stylizes its own clip, no communication"}
        config["frame_sharding"] = fs
    line = {
        "metric": "stylized frames/sec (16x512x512, 50 DDIM steps)", "value": fps, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": config,
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes(clip), "d2h_bytes_per_step": host_out.numel() * 2},
        "gpu_launches": launches, "clocks": clk, "roofline": roof,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        step = cpu_reference_step(CPU_SAMPLE_FRAMES, threads)
        dt = step()
        line["cpu_baseline"] = {"value": CPU_SAMPLE_FRAMES / (STEPS_DDIM * dt), "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"1 three-branch DDIM step, 3 x {CPU_SAMPLE_FRAMES} frames at 64x64 latents, full SD-1.5 "
                                          "width, fp32 oracle port, all three branches (the reference's own work)"}
    print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


_OUT = sys.stdout   # where the JSON line goes (main() re-points it at a private duplicate of fd 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the supplementary passes (inversion, AnimateDiff, torch eager)")
    ap.add_argument("--no-animatediff", dest="no_extras", action="store_true")
    ap.add_argument("--no-cuda-graphs", action="store_true")
    ap.add_argument("--config5", action="store_true", help="also run BASELINE configs[4] (SD-3.5 MMDiT, 16x1024x1024, 28 steps) at this N")
    ap.add_argument("--config2", action="store_true", help="also run BASELINE configs[2] (SD-2.1, 32 frames, smoother) at this N")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: libraries that write banners to fd 1 from C (NCCL prints its version there
    # under torchrun) are sent to stderr, and the line itself goes to a duplicate of the original descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        # the product path is the sm_100a library and nothing else: no CPU fallback, no number without a GPU
        sys.exit("bench.py: no CUDA device -- the hot path runs only on the sm_100a kernels of univst_b200 "
                 "(the CPU oracle is timed by --impl reference)")
    run_ours(args)


if __name__ == "__main__":
    main()
