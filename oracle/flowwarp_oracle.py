"""ORACLE (test infrastructure, not product code): NumPy restatement of the reference's flow-warp helpers
(src/cal_optica_flow.py:20-46) and of the sliding-window loop that calls them
(backbones/video_diffusion_sd/pipelines/stable_diffusion.py:725-751).

The one third-party piece is ``cv2.remap(INTER_LINEAR, BORDER_CONSTANT)`` on uint8 images (opencv-python 4.9.0.80 in
the reference's requirements.txt:61).  Its published algorithm is restated in :func:`remap_bilinear_u8`: map
coordinates are rounded to 1/32 pixel (``cvRound(x * 32)``), the four bilinear weights are the exact products
``(32 - fx)(32 - fy) * 32 ...`` (sum 2^15), the result is ``(sum + 2^14) >> 15``; out-of-image taps read 0.
Parity status: PINNED -- ``tests/test_oracle_cpu.py`` checks this restatement bit-for-bit against ``cv2.remap`` itself
(cv2 4.13 is in the image) and against golden outputs of the reference's own functions (tests/golden/flow_warp.npz).
RAFT (the flow estimator) is a third-party network whose weights are unavailable offline: flows are inputs here.
"""
from __future__ import annotations

import numpy as np


def compute_occlusion_mask(forward_flow, backward_flow, threshold=1.0):
    """cal_optica_flow.py:20-29 (same fp32 operation order: ((p + fwd) + bwd) - p)."""
    h, w, _ = forward_flow.shape
    gx, gy = np.meshgrid(np.arange(w), np.arange(h))
    coords2 = np.stack([gx, gy], axis=-1).astype(np.float32)
    back = (coords2 + forward_flow) + backward_flow
    error = np.linalg.norm(back - coords2, axis=-1)
    return (error > threshold).astype(np.uint8) * 255


def remap_bilinear_u8(image, map_x, map_y):
    """cv2.remap(image, map_x, map_y, INTER_LINEAR, borderMode=BORDER_CONSTANT (0)) for uint8 HxWxC images."""
    h, w = image.shape[:2]
    sx = np.rint(map_x.astype(np.float32) * np.float32(32)).astype(np.int64)  # cvRound: half to even
    sy = np.rint(map_y.astype(np.float32) * np.float32(32)).astype(np.int64)
    ix, iy, fx, fy = sx >> 5, sy >> 5, sx & 31, sy & 31
    # OpenCV clamps the integer coordinates to the short range before use
    ix, iy = np.clip(ix, -32768, 32767), np.clip(iy, -32768, 32767)
    img = image.astype(np.int64)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        v = img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return v * ok[..., None]

    w00 = ((32 - fx) * (32 - fy) * 32)[..., None]
    w01 = (fx * (32 - fy) * 32)[..., None]
    w10 = ((32 - fx) * fy * 32)[..., None]
    w11 = (fx * fy * 32)[..., None]
    acc = tap(iy, ix) * w00 + tap(iy, ix + 1) * w01 + tap(iy + 1, ix) * w10 + tap(iy + 1, ix + 1) * w11
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def warp_image_with_flow(image, flow):
    """cal_optica_flow.py:31-41."""
    h, w, _ = flow.shape
    gx, gy = np.meshgrid(np.arange(w), np.arange(h))
    coords = np.stack([gx, gy], axis=-1).astype(np.float32) + flow
    return remap_bilinear_u8(image, coords[..., 0].astype(np.float32), coords[..., 1].astype(np.float32))


def apply_mask(image, mask, original_image):
    """cal_optica_flow.py:43-46."""
    m = np.repeat(mask[:, :, np.newaxis], 3, axis=2) / 255.0
    return (image * (1 - m) + original_image * m).astype(np.uint8)


def get_warp(key_frame, now_frame, fwd, bwd):
    """get_warp (cal_optica_flow.py:51-99) with the two RAFT flows given: fwd = flow(key -> now), bwd = flow(now -> key)."""
    occ = compute_occlusion_mask(fwd, bwd, threshold=1.5)
    return apply_mask(warp_image_with_flow(now_frame, fwd), occ, key_frame)


def sliding_window_smooth(frames, flow_of, keep_mask=None, r=2):
    """stable_diffusion.py:725-751.  frames: (F, H, W, 3) uint8; ``flow_of(key, now) -> (fwd, bwd)``;
    keep_mask: (F, H, W) in {0, 1} (1 = keep the original pixel).  In place, ascending key order: neighbours with a
    smaller index have already been smoothed when they are read (Gauss-Seidel), results truncate to uint8."""
    est = frames.copy()
    F_ = est.shape[0]
    for key in range(F_):
        key_frame = est[key].copy()
        acc = np.zeros(key_frame.shape, np.float32)
        weight = 0
        for bias in range(-r, r + 1):
            now = key + bias
            if 0 <= now < F_:
                if bias == 0:
                    acc = acc + est[now].astype(np.float32)
                else:
                    fwd, bwd = flow_of(key, now)
                    acc = acc + get_warp(key_frame, est[now].copy(), fwd, bwd).astype(np.float32)
                weight += 1
        est[key] = acc / weight  # float32 -> uint8 assignment truncates
    if keep_mask is not None:
        m = keep_mask[..., None].astype(np.float64)
        est = (frames * m + (1 - m) * est).astype(np.uint8)
    return est


def synthetic_flow(h, w, seed, backward_of=None):
    """Deterministic test flows: a translation + smooth sinusoidal field, with exact-integer and half-pixel displacement
    patches (exercise the 1/32-px rounding) and vectors that leave the image.  ``backward_of``: build the backward flow
    of a given forward flow (= -fwd) except for a patch that violates consistency (-> occluded)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    if backward_of is not None:
        f = (-backward_of).astype(np.float32)
        f[h // 4:h // 2, w // 4:w // 2] += np.float32(4.0)
        f[::9, ::7] += np.float32(1.4)   # borderline consistency errors around the 1.5 px threshold
        return f
    rng = np.random.default_rng(seed)
    f = np.stack([2.5 + 1.5 * np.sin(yy / 9.0) + 0.3 * rng.standard_normal((h, w)),
                  -1.25 + 1.5 * np.cos(xx / 7.0) + 0.3 * rng.standard_normal((h, w))], axis=-1).astype(np.float32)
    f[:16, :16] = np.round(f[:16, :16])
    f[16:32, :16] = np.round(f[16:32, :16] * 2) / 2
    f[-8:, -8:] += np.float32(40.0)
    return f
