"""Mirror of the reference's flow-warp helpers (src/cal_optica_flow.py) and of the sliding-window smoother that calls
them (pipelines/stable_diffusion.py:725-751), on CUDA uint8 frames.  The optical flows are inputs: RAFT is a
third-party network (torchvision ``raft_large``) whose weights cannot be downloaded offline; pass any estimator as
``flow_fn(key_frame, now_frame) -> (fwd, bwd)`` or precomputed flows."""
from __future__ import annotations

import torch

from . import ops


def sliding_window_smooth(frames: torch.Tensor, flow_of, keep_mask: torch.Tensor = None, r: int = 2,
                          threshold: float = 1.5) -> torch.Tensor:
    """frames: [F, H, W, 3] uint8 CUDA.  ``flow_of(key, now) -> (fwd, bwd)`` ([H, W, 2] fp32 CUDA each; it may look at
    the current, partially smoothed ``frames`` like the reference's RAFT calls do).  ``keep_mask`` [F, H, W] uint8:
    non-zero pixels keep their original value (stable_diffusion.py:751).  Key frames are processed in ascending
    order in place, exactly like the reference (neighbours with a smaller index are already smoothed)."""
    est = frames.clone()
    F = est.shape[0]
    for key in range(F):
        nbs = [key + b for b in range(-r, r + 1) if b != 0 and 0 <= key + b < F]
        flows = [flow_of(key, n) for n in nbs]
        ops.flow_warp_key_(est, key, nbs, [f[0].contiguous() for f in flows], [f[1].contiguous() for f in flows], threshold)
    if keep_mask is not None:
        est = ops.mask_select(keep_mask.contiguous(), frames.contiguous(), est)
    return est
