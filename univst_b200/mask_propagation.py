"""Mirror of src/mask_propagation.py: ``mask_propogation`` (sic, :72-99) and ``video_mask_propogation`` (:15-69).

The similarity / top-k / label-transport arithmetic is the fused ``univst_maskprop_f32`` kernel; the anchor-queue
bookkeeping and the ``torch.randperm`` subsampling stay host-side with the reference's exact RNG call sequence (same
global CPU generator, same order), so a seeded run picks the same anchor points."""
from __future__ import annotations

import os
import queue
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

_DEVICE = "cuda"   # where the features live (the kernel is CUDA-only; CPU host-logic tests re-point this)
DEFAULT_ARGS = SimpleNamespace(temperature=0.2, n_last_frames=9, topk=15, sample_ratio=0.3, num_frames=16)


def mask_propogation(feat_src, feat_tar, segs, args=DEFAULT_ARGS):
    """feat_src [C, M], feat_tar [N, C], segs [Ccls, M] (fp32 CUDA) -> (segs_tar [Ccls, N], feat_sample [C, n],
    segs_sample [Ccls, n]) -- mask_propagation.py:72-99."""
    feat_tar_ori = feat_tar.T
    segs_tar = ops.maskprop(feat_tar.contiguous(), feat_src.contiguous(), segs.contiguous(), args.temperature, args.topk)
    fore_index = torch.where(segs_tar[0, :] != 0)[0]
    back_index = torch.where(segs_tar[0, :] == 0)[0]
    fore_nums, back_nums = len(fore_index), len(back_index)
    perm = torch.randperm(len(fore_index))[: int(len(fore_index) * fore_nums / (fore_nums + back_nums) * args.sample_ratio)]
    fore_sample = fore_index[perm.to(fore_index.device)]
    perm = torch.randperm(len(back_index))[: int(len(back_index) * back_nums / (fore_nums + back_nums) * args.sample_ratio)]
    back_sample = back_index[perm.to(back_index.device)]
    all_index = torch.cat([fore_sample, back_sample])
    return segs_tar, feat_tar_ori[:, all_index], segs_tar[:, all_index]


def to_one_hot(y, n_dims=None):
    """mask_propagation.py:126-138: (1, h, w) integer labels -> (1, n_dims, h, w) one-hot."""
    n_dims = int(y.max() + 1) if n_dims is None else n_dims
    _, h, w = y.shape
    idx = y.long().view(-1, 1)
    one_hot = torch.zeros(idx.shape[0], n_dims, device=y.device).scatter_(1, idx, 1)
    return one_hot.view(h, w, n_dims).permute(2, 0, 1).unsqueeze(0)


def norm_mask(mask):
    """mask_propagation.py:114-123."""
    for c in range(mask.shape[0]):
        m = mask[c]
        if m.max() > 0:
            m = m - m.min()
            mask[c] = m / m.max()
    return mask


def read_feature(path_or_tensor, frame_index=None, return_h_w=False):
    """mask_propagation.py:102-111 (one frame of the dumped feature map as [h w, C] fp32) -- or, without a frame index, the
    whole (F, h, w, C) map on the device, which is what the driver below reads ONCE instead of once per frame (:104)."""
    data = path_or_tensor if torch.is_tensor(path_or_tensor) else torch.load(path_or_tensor, weights_only=True)
    data = data.to(_DEVICE).float()
    if frame_index is None:
        return data
    data = data[frame_index]
    _h, _w, _ = data.shape
    data = data.reshape(_h * _w, -1).contiguous()
    return (data, _h, _w) if return_h_w else data


@torch.no_grad()
def video_mask_propogation(args_or_mask, features=None, args=None, output_path=None):
    """mask_propagation.py:15-69.  Two call forms:

    * the reference's: ``video_mask_propogation(args)`` with ``args.mask_path`` (first-frame label PNG),
      ``args.feature_path`` (``inversion_feature_map_2_block_301_step.pt``), ``args.output_path``, ``args.backbone`` and
      the hyper-parameters -- writes ``<output_path>/<backbone>/<mask name>/%05d.png`` like the reference (:18-20, :30, :69);
    * in memory: ``video_mask_propogation(first_mask (H, W) uint8, features (F, h, w, C), args, output_path)``.

    Returns the list of (H, W) uint8 masks (frame 0 = the input)."""
    from PIL import Image
    if not isinstance(args_or_mask, np.ndarray) and hasattr(args_or_mask, "mask_path"):
        args = args_or_mask
        name = args.mask_path.split("/")[-1].split(".")[0]
        output_path = os.path.join(args.output_path, args.backbone, name)
        first_mask = np.asarray(Image.open(args.mask_path))
        features = args.feature_path
    else:
        first_mask = np.asarray(args_or_mask)
        args = args if args is not None else DEFAULT_ARGS
    feats = read_feature(features)
    nF, h, w, C = feats.shape
    ori_h, ori_w = first_mask.shape
    seg0 = np.array(Image.fromarray(first_mask).resize((w, h), 0))
    first_seg = to_one_hot(torch.from_numpy(seg0).float().unsqueeze(0).to(feats.device))
    Ccls = first_seg.shape[1]
    que = queue.Queue(args.n_last_frames)
    feat_first = feats[0].reshape(h * w, C).T.contiguous()
    masks = [first_mask.astype(np.uint8)]
    for cnt in range(1, min(args.num_frames, nF)):
        feat_src = torch.cat([feat_first] + [p[0] for p in list(que.queue)], dim=-1)
        segs_src = torch.cat([first_seg.squeeze(0).flatten(1)] + [p[1] for p in list(que.queue)], dim=-1)
        feat_tgt = feats[cnt].reshape(h * w, C).contiguous()
        final_mask, feat_s, segs_s = mask_propogation(feat_src, feat_tgt, segs_src, args)
        if que.qsize() == args.n_last_frames:
            que.get()
        que.put([feat_s, segs_s])
        up = F.interpolate(final_mask.reshape(1, Ccls, h, w), size=(ori_h, ori_w), mode="bilinear", align_corners=False)[0]
        lab = torch.max(norm_mask(up), dim=0)[1]
        m = lab.cpu().numpy().astype(np.uint8)
        m[m != 0] = 255
        masks.append(m)
    if output_path is not None:
        os.makedirs(output_path, exist_ok=True)
        for i, m in enumerate(masks):
            Image.fromarray(m).save(os.path.join(output_path, "%05d.png" % i))
    return masks
