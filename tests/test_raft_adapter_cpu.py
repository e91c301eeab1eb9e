"""``flow_warp.raft_flow_fn`` against the reference's own ``get_warp`` (src/cal_optica_flow.py:49-89) with the same
torchvision RAFT -- random weights (the pretrained ones cannot be downloaded), constructor patched inside the reference
module so that its ``raft_large(weights=DEFAULT)`` returns that model.  The flows the reference hands to its
``compute_occlusion_mask`` are recorded and must equal the adapter's bit for bit; the reference's warped + masked frame
must equal the NumPy / cv2 oracle on the adapter's flows.  Build container only (needs /root/reference)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference checkout (build container only)")
def test_raft_adapter_feeds_the_model_like_the_reference(monkeypatch):
    tvof = pytest.importorskip("torchvision.models.optical_flow")
    pytest.importorskip("cv2")
    monkeypatch.syspath_prepend(REF)
    import src.cal_optica_flow as ref
    from univst_b200.flow_warp import raft_flow_fn
    from oracle import flowwarp_oracle as fo

    torch.manual_seed(3)
    model = tvof.raft_small(weights=None).eval()      # same forward signature / output list as raft_large, seconds on a CPU
    monkeypatch.setattr(ref, "raft_large", lambda weights=None: model)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    seen = {}
    real_occ = ref.compute_occlusion_mask

    def spy(fwd, bwd, threshold=1.0):
        seen["fwd"], seen["bwd"], seen["thr"] = fwd.copy(), bwd.copy(), threshold
        return real_occ(fwd, bwd, threshold=threshold)

    monkeypatch.setattr(ref, "compute_occlusion_mask", spy)
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:128, 0:160]
    base = (127 + 90 * np.sin(xx / 9.0)[..., None] * np.cos(yy / 7.0)[..., None] * np.array([1.0, 0.6, -0.8])).clip(0, 255)
    key = (base + rng.integers(0, 12, base.shape)).clip(0, 255).astype(np.uint8)
    now = np.roll(key, (2, -3), axis=(0, 1))
    want = ref.get_warp(key, now, key, now)                                  # the call of stable_diffusion.py:744

    fwd, bwd = raft_flow_fn(model)(torch.from_numpy(key), torch.from_numpy(now))
    assert fwd.shape == (128, 160, 2) and fwd.dtype == torch.float32 and fwd.is_contiguous()
    assert np.array_equal(fwd.numpy(), seen["fwd"]) and np.array_equal(bwd.numpy(), seen["bwd"]) and seen["thr"] == 1.5
    occ = fo.compute_occlusion_mask(fwd.numpy(), bwd.numpy(), 1.5)
    got = fo.apply_mask(fo.warp_image_with_flow(now, fwd.numpy()), occ, key)
    assert np.array_equal(got, want)
