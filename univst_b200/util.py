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


def load_video_frames(frames_path, n_frames, image_size=(512, 512)):
    """src/util.py:63-81: ``%05d.png`` frames -> (F, 3, H, W) float32 in [-1, 1] (``x / 127.5 - 1``).  ``image_size`` is
    (width, height), PIL's order; each file is resized to it on loading (PIL's default filter, like the reference's
    ``load_image``, :104) and converted to RGB."""
    from PIL import Image, ImageOps
    frames = []
    for i in range(n_frames):
        path = f"{frames_path}/%05d.png" % i
        if not os.path.isfile(path):
            raise ValueError(f"Incorrect path or URL. URLs must start with `http://` or `https://`, and {path} is not a valid path.")
        img = ImageOps.exif_transpose(Image.open(path).resize(tuple(image_size))).convert("RGB")
        if img.size != tuple(image_size):
            raise ValueError("Frame size does not match config.image_size")
        frames.append(torch.from_numpy(np.array(img) / 127.5 - 1.0).permute(2, 0, 1).float())
    return torch.stack(frames)


def _grid_frames(videos: torch.Tensor, rescale: bool, n_rows: int):
    """(b, c, t, h, w) in [0, 1] -> t uint8 (H, W, c) images, the batch tiled ``n_rows`` per row with torchvision's
    2-pixel padding (src/util.py:35-44; one clip -> the frame itself); ``(x * 255)`` truncated like the reference."""
    import torchvision
    outs = []
    for x in videos.permute(2, 0, 1, 3, 4):
        x = torchvision.utils.make_grid(x, nrow=n_rows).permute(1, 2, 0)
        if rescale:
            x = (x + 1.0) / 2.0
        outs.append((x * 255).numpy().astype(np.uint8))
    return outs


def write_video(path: str, frames, fps=8):
    """``frames``: uint8 RGB (H, W, 3) arrays -> ``path``.  imageio when it is importable (what the reference's
    ``save_videos_grid`` and diffusers' ``export_to_video`` use), else OpenCV's ``mp4v`` writer (opencv-python is one of the
    reference's requirements), else PNG frames under ``<path without extension>/%05d.png`` (the layout of the reference's
    ``save_folder``, src/util.py:22-31)."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    try:
        import imageio
        imageio.mimsave(path, frames, fps=fps)
        return path
    except ImportError:
        pass
    try:
        import cv2
        h, w = frames[0].shape[:2]
        wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), float(fps), (w, h))
        if wr.isOpened():
            for x in frames:
                wr.write(cv2.cvtColor(np.ascontiguousarray(x), cv2.COLOR_RGB2BGR) if x.ndim == 3 and x.shape[-1] == 3
                         else cv2.cvtColor(np.ascontiguousarray(x.squeeze(-1) if x.ndim == 3 else x), cv2.COLOR_GRAY2BGR))
            wr.release()
            return path
    except ImportError:
        pass
    from PIL import Image
    folder = os.path.splitext(path)[0]
    os.makedirs(folder, exist_ok=True)
    for i, x in enumerate(frames):
        Image.fromarray(x.squeeze(-1) if x.ndim == 3 and x.shape[-1] == 1 else x).save(os.path.join(folder, "%05d.png" % i))
    return folder


def save_folder(videos: torch.Tensor, path: str, rescale=False, n_rows=4, fps=8):
    """src/util.py:22-31: (b, c, t, h, w) in [0, 1] -> ``path/%05d.png``, one grid image per frame (what the run scripts
    write the stylized clip with, run_video_style_transfer_sd.py:69).  Returns the frames."""
    from PIL import Image
    frames = _grid_frames(videos, rescale, n_rows)
    for i, x in enumerate(frames):
        Image.fromarray(x.squeeze(-1) if x.shape[-1] == 1 else x).save(os.path.join(path, "%05d.png" % i))
    return frames


def seed_everything(seed=42):
    """src/util.py:16-19."""
    import random
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def save_videos_grid(videos: torch.Tensor, path: str, rescale=False, n_rows=4, fps=8):
    """src/util.py:34-47: (b, c, t, h, w) in [0, 1] -> one video of the batch tiled per frame.  Returns the frames."""
    frames = _grid_frames(videos, rescale, n_rows)
    write_video(path, frames, fps=fps)
    return frames
