"""Golden vectors for the mask-propagation DRIVER: the reference's own ``video_mask_propogation`` (src/mask_propagation.py:
15-69) -- first-mask resize + one-hot, anchor queue, per-frame ``mask_propogation``, bilinear upsampling, per-class min-max
normalisation, argmax, ``!= 0 -> 255``, PNG output -- run here on the CPU on synthetic decoder features and (a) the
shipped anti-aliased ``examples/masks/mallard-fly.png`` (256 one-hot classes) and (b) a clean two-class mask.

Build container only (needs /root/reference).  The reference hard-codes ``.cuda()`` / ``.to("cuda")`` (mask_propagation.py:
104, :138) and reads ``src/palette.txt`` relative to the working directory: both are accommodated from the outside
(device calls mapped to the CPU, cwd = the reference checkout); ``imageio`` is absent here and replaced by a PIL writer.

    PYTHONPATH=/root/repo python oracle/gen_golden_maskprop_video.py
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _import_reference():
    from PIL import Image
    imageio = types.ModuleType("imageio")
    imageio.imwrite = lambda path, arr: Image.fromarray(np.asarray(arr)).save(path)
    sys.modules.setdefault("imageio", imageio)
    sys.path.insert(0, REF)
    import src.mask_propagation as mp
    return mp


def _cpu_only():
    """Map the reference's hard-coded device moves onto the CPU (no GPU in the build container)."""
    torch.Tensor.cuda = lambda self, *a, **k: self
    _to = torch.Tensor.to

    def to(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return _to(self, *a, **k)
    torch.Tensor.to = to


def run_case(mp, mask_u8, feats, n_last_frames, seed):
    from PIL import Image
    with tempfile.TemporaryDirectory() as tmp:
        mpath, fpath = os.path.join(tmp, "mask.png"), os.path.join(tmp, "feat.pt")
        Image.fromarray(mask_u8, mode="L").save(mpath)
        torch.save(feats, fpath)
        args = types.SimpleNamespace(temperature=0.2, n_last_frames=n_last_frames, topk=15, sample_ratio=0.3,
                                     num_frames=feats.shape[0], mask_path=mpath, backbone="sd", feature_path=fpath,
                                     output_path=os.path.join(tmp, "out"))
        torch.manual_seed(seed)
        cwd = os.getcwd()
        os.chdir(REF)
        try:
            mp.video_mask_propogation(args)
        finally:
            os.chdir(cwd)
        odir = os.path.join(args.output_path, "sd", "mask")
        names = sorted(os.listdir(odir))
        masks = np.stack([np.asarray(Image.open(os.path.join(odir, n))) for n in names])
    return names, masks


def main():
    from PIL import Image
    from oracle import maskprop_oracle as mo
    mp = _import_reference()
    _cpu_only()
    out = {}
    shipped = np.asarray(Image.open(os.path.join(REF, "examples/masks/mallard-fly.png")))
    yy, xx = np.mgrid[0:256, 0:320]
    clean = (((xx - 150) ** 2 + (yy - 120) ** 2) <= 70 ** 2).astype(np.uint8)        # labels {0, 1}, non-square frame
    for name, mask, (nF, h, w, C), nlast in (("shipped_antialiased", shipped, (6, 16, 16, 32), 3),
                                             ("clean_two_class", clean, (7, 16, 20, 32), 9)):
        feats = mo.synthetic_features(21, nF, max(h, w), max(h, w), C)[:, :h, :w].contiguous().half()
        names, masks = run_case(mp, mask, feats, nlast, seed=5)
        assert masks.dtype == np.uint8 and set(np.unique(masks[1:])) <= {0, 255} and np.array_equal(masks[0], mask)
        out[name] = {"first_mask": torch.from_numpy(mask.copy()), "features": feats, "n_last_frames": nlast, "seed": 5,
                     "names": names, "shape": list(masks.shape),   # frame 0 = the first mask itself (:30), asserted here
                     "masks_bits": torch.from_numpy(np.packbits(masks[1:] != 0))}
        print(name, masks.shape, "foreground fraction per frame", [round(float((m != 0).mean()), 3) for m in masks])
    torch.save(out, os.path.join(OUT, "maskprop_video.pt"))
    print("wrote maskprop_video.pt", os.path.getsize(os.path.join(OUT, "maskprop_video.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
