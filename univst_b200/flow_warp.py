"""Mirror of the reference's flow-warp helpers (src/cal_optica_flow.py) and of the sliding-window smoother that calls
them (pipelines/stable_diffusion.py:725-751), on CUDA uint8 frames.  The optical flows are inputs: RAFT is a
third-party network (torchvision ``raft_large``) whose weights cannot be downloaded offline; pass any estimator as
``flow_fn(key_frame, now_frame) -> (fwd, bwd)`` or precomputed flows.  ``raft_flow_fn(model)`` wraps a torchvision RAFT the
way the reference drives it."""
from __future__ import annotations

import torch

from . import ops


def sliding_window_smooth(frames: torch.Tensor, flow_of=None, keep_mask: torch.Tensor = None, r: int = 2,
                          threshold: float = 1.5, flow_fn=None) -> torch.Tensor:
    """frames: [F, H, W, 3] uint8 CUDA.  ``flow_of(key, now) -> (fwd, bwd)`` ([H, W, 2] fp32 CUDA each) by frame INDEX,
    or ``flow_fn(key_frame, now_frame) -> (fwd, bwd)`` on the current, partially smoothed uint8 frames -- what the
    reference's ``get_warp(key_frame, now_frame, key_frame, now_frame)`` hands to RAFT (stable_diffusion.py:744,
    cal_optica_flow.py:75-79).  ``keep_mask`` [F, H, W] uint8: non-zero pixels keep their original value
    (stable_diffusion.py:751).  Key frames are processed in ascending order in place, exactly like the reference
    (neighbours with a smaller index are already smoothed)."""
    if (flow_of is None) == (flow_fn is None):
        raise ValueError("pass exactly one of flow_of(key, now) and flow_fn(key_frame, now_frame)")
    est = frames.clone()
    F = est.shape[0]
    for key in range(F):
        nbs = [key + b for b in range(-r, r + 1) if b != 0 and 0 <= key + b < F]
        flows = [flow_of(key, n) if flow_fn is None else flow_fn(est[key], est[n]) for n in nbs]
        ops.flow_warp_key_(est, key, nbs, [f[0].contiguous() for f in flows], [f[1].contiguous() for f in flows], threshold)
    if keep_mask is not None:
        est = ops.mask_select(keep_mask.contiguous(), frames.contiguous(), est)
    return est


def raft_flow_fn(model):
    """``flow_fn`` for ``sliding_window_smooth`` / ``video_style_transfer(smoother="pixel", flow_fn=...)`` from a
    torchvision RAFT (``torchvision.models.optical_flow.raft_large``, third-party: its weights are the caller's business).
    Follows the reference's ``get_warp`` (src/cal_optica_flow.py:49-72): frames as float RGB / 255 -- NOT the [-1, 1] range
    torchvision documents; the reference feeds [0, 1] and so does this (:11-13) --, forward = model(key, now)[-1],
    backward = model(now, key)[-1] (the last refinement iteration), each returned as (H, W, 2) fp32.  The reference builds
    the model anew for every pair (58 constructions per smoothing step); here it is built once by the caller."""
    model.eval()

    @torch.no_grad()
    def flow_fn(key_frame: torch.Tensor, now_frame: torch.Tensor):
        dev = next(model.parameters()).device
        prep = lambda im: (im.to(dev).permute(2, 0, 1).float() / 255.0).unsqueeze(0)
        a, b = prep(key_frame), prep(now_frame)
        fwd = model(a, b)[-1]
        bwd = model(b, a)[-1]
        out = lambda f: f.squeeze(0).permute(1, 2, 0).float().contiguous().to(key_frame.device)
        return out(fwd), out(bwd)

    return flow_fn
