"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's point-matching mask propagation
core, src/mask_propagation.py:72-83 (fp32, torch CPU).  Parity status: PINNED -- tests/test_oracle_cpu.py checks it
against golden outputs of the reference's own ``mask_propogation`` (tests/golden/maskprop.pt, made by
oracle/gen_golden_extra.py)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def affinity_topk(feat_src, feat_tar, temperature=0.2, topk=15):
    """mask_propagation.py:75-81 -> (aff (M, N) with everything below each column's topk-th largest zeroed and the
    columns normalised, thresholds (N,))."""
    src = F.normalize(feat_src, dim=0, p=2)
    tar = F.normalize(feat_tar, dim=1, p=2)
    aff = torch.exp(tar @ src / temperature).transpose(1, 0)
    thr = torch.topk(aff, topk, dim=0).values.min(dim=0).values
    aff = aff.clone()
    aff[aff < thr] = 0
    return aff / torch.sum(aff, keepdim=True, axis=0), thr


def mask_propogation_core(feat_src, feat_tar, segs, temperature=0.2, topk=15):
    """segs_tar = segs @ aff (mask_propagation.py:83)."""
    aff, thr = affinity_topk(feat_src, feat_tar, temperature, topk)
    return torch.mm(segs, aff), aff, thr


def synthetic_features(seed: int, F_: int, h: int, w: int, C: int, separated: bool = False):
    """Smooth random feature maps advected by 1 px / frame (SURVEY.md 8(d) mask-prop recipe, reduced).  With
    ``separated=True`` every point gets a strong private component, so similarities are well separated and the kept
    index set is insensitive to fp32 summation order (used for the bit-exact index test)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(1, C, h + F_, w + F_, generator=g)
    base = F.avg_pool2d(base, 3, 1, 1) * 3
    feats = torch.stack([base[0, :, f:f + h, f:f + w] for f in range(F_)]).permute(0, 2, 3, 1).contiguous()
    if separated:
        feats = feats + 2.0 * torch.randn(F_, h, w, C, generator=g)
    return feats
