"""On-disk formats of the reference stages (src/util.py): ``ddim_latents_{k}.pt`` and ``%05d.png`` masks."""
from __future__ import annotations

import os

import numpy as np
import torch


def load_ddim_latents_at_t(t, ddim_latents_path, is_x0=False):
    """src/util.py:123-130."""
    name = f"ddim_x0_{t}.pt" if is_x0 else f"ddim_latents_{t}.pt"
    path = os.path.join(ddim_latents_path, name)
    assert os.path.exists(path), f"Missing latents at t {t} path {path}"
    return torch.load(path, weights_only=True)


def load_mask(mask_path="", n_frames=16):
    """src/util.py:133-144 -> (1, F, H, W) uint8 in {0, 1}.  The reference computes ``uint8 * 255`` (wraps modulo 256)
    and clips to [0, 1]: every non-zero pixel becomes 1."""
    from PIL import Image
    files = sorted(f"{mask_path}/%05d.png" % i for i in range(n_frames))
    imgs = np.stack([np.array(Image.open(f)) for f in files])
    return torch.from_numpy((imgs != 0).astype(np.uint8)).unsqueeze(0)
