"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's SD3 / SD3.5 joint-attention
processors, backbones/video_diffusion_sd3/pnp_utils.py: ``CrossFrameProcessor`` (:9-132, used for the inversions) and
``AttentionShiftProcessor`` (:135-271, used for the three-branch transfer), plus ``attention_adain`` (:287-300).

Only ``tests/`` may import this.  Parity status: PINNED at the processor level -- ``oracle/gen_golden_sd3.py`` runs the
reference's own processor classes on a stand-in ``attn`` module (the projections / RMS norms they touch; the MMDiT
around them is third-party diffusers code that is not in this image) and commits input / output vectors under
``tests/golden/sd3_processors.pt``.  One emulated reference bug: ``AttentionShiftProcessor`` reads ``self.thresh2``
(:185), which is never set -- the reference raises AttributeError inside the shift window; the goldens set
``thresh2 = eta2``, the evident intent (beta then runs 0.9 -> 0.1 over the window, as in the SD / AnimateDiff patches).

Semantics that differ from the SD patch (oracle/unet_oracle.py): Q/K/V are (B*F, heads, N, d) when shifted, so
``attention_adain`` takes its statistics over the tokens per (frame, head, channel) and ``F.instance_norm`` of a 4-D
tensor normalises over (tokens, head_dim) jointly per (frame, head); RMS-norm on q / k per head before the shift; K/V of
the image tokens are gathered from [first, previous, self] frames (:26) and the text tokens are appended to Q, K and V.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

CLIP_LENGTH = 16  # hard-coded in the reference processors (:25, :153)
GOLDEN_IMAGES = [b * CLIP_LENGTH + f for b in range(3) for f in (0, 1, 2, 15)]  # images whose golden outputs are committed


def synthetic_inputs(seed: int, N: int, L: int, C: int):
    """(hidden (48, N, C), enc (48, L, C)) of the goldens: three branches x 16 frames."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(3 * CLIP_LENGTH, N, C, generator=g), torch.randn(3 * CLIP_LENGTH, L, C, generator=g)


def rms_norm(x, weight, eps=1e-6):
    """diffusers RMSNorm(dim_head, eps=1e-6) as used for norm_q / norm_k / norm_added_q / norm_added_k [3p]."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * weight


def attention_adain(cnt, sty):
    """pnp_utils.py:287-300 on (F, heads, N, d): style mean / unbiased std over the tokens; content instance-normalised
    over (N, d) per (frame, head)."""
    sty_mean = sty.mean(dim=[-2], keepdim=True)
    sty_std = sty.std(dim=[-2], keepdim=True)
    return (F.instance_norm(cnt) * sty_std + sty_mean).to(cnt.dtype)


def shift_params(idx: int, eta1: float = 0.0, eta2: float = 0.6):
    """pnp_utils.py:182-187 with thresh2 := eta2 (see the module docstring)."""
    active = idx >= eta1 * 50 and idx <= eta2 * 50
    beta = (0.9 - 0.1) / (eta1 * 50 - eta2 * 50) * (idx - eta2 * 50) + 0.1
    return active, 0.8, beta, 2.0


def joint_attention(w, hidden, enc, heads: int, idx=None, eta1=0.0, eta2=0.6, cross_frame=True):
    """Both processors: ``idx is None`` -> CrossFrameProcessor; else AttentionShiftProcessor at step ``idx``.
    hidden (B*16, N, C), enc (B*16, L, C); w: dict of the attn module's tensors.  Returns (hidden_out, enc_out).
    ``enc is None`` (the self-attention ``attn2`` of SD3.5's dual-attention blocks, :92,:120,:134): image tokens only,
    returns hidden_out.  No ``to_add_out`` in ``w`` = a ``context_pre_only`` attention (:127): the text half of the
    attention output is returned unprojected.  ``cross_frame=False``: the image K/V are the frame's own (diffusers'
    stock JointAttnProcessor2_0, third-party)."""
    BF, N, C = hidden.shape
    d = C // heads
    split = lambda t: t.view(BF, -1, heads, d).transpose(1, 2)
    q = split(F.linear(hidden, w["to_q.weight"], w["to_q.bias"]))
    k = split(F.linear(hidden, w["to_k.weight"], w["to_k.bias"]))
    v = split(F.linear(hidden, w["to_v.weight"], w["to_v.bias"]))
    q, k = rms_norm(q, w["norm_q.weight"]), rms_norm(k, w["norm_k.weight"])
    if idx is not None:
        chunk = BF // 3
        active, alpha, beta, gamma = shift_params(idx, eta1, eta2)
        if active:
            q, k, v = q.clone(), k.clone(), v.clone()
            q[2 * chunk:] = alpha * q[:chunk] + (1 - alpha) * q[2 * chunk:]
            k[2 * chunk:] = beta * attention_adain(k[2 * chunk:], k[chunk:2 * chunk]) + (1 - beta) * k[chunk:2 * chunk]
            v[2 * chunk:] = beta * attention_adain(v[2 * chunk:], v[chunk:2 * chunk]) + (1 - beta) * v[chunk:2 * chunk]
            q[2 * chunk:] = gamma * q[2 * chunk:]
    Fr = CLIP_LENGTH
    B = BF // Fr
    k5, v5 = k.view(B, Fr, heads, N, d), v.view(B, Fr, heads, N, d)
    first = torch.zeros(Fr, dtype=torch.long)
    prev = (torch.arange(Fr) - 1).clip(0, Fr - 1)
    me = torch.arange(Fr)
    srcs = (first, prev, me) if cross_frame else (me,)
    k = torch.cat([k5[:, s] for s in srcs], dim=-2).reshape(BF, heads, len(srcs) * N, d)
    v = torch.cat([v5[:, s] for s in srcs], dim=-2).reshape(BF, heads, len(srcs) * N, d)
    if enc is None:
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(BF, -1, C)
        return F.linear(o, w["to_out.0.weight"], w["to_out.0.bias"])
    eq = rms_norm(split(F.linear(enc, w["add_q_proj.weight"], w["add_q_proj.bias"])), w["norm_added_q.weight"])
    ek = rms_norm(split(F.linear(enc, w["add_k_proj.weight"], w["add_k_proj.bias"])), w["norm_added_k.weight"])
    ev = split(F.linear(enc, w["add_v_proj.weight"], w["add_v_proj.bias"]))
    q, k, v = torch.cat([q, eq], 2), torch.cat([k, ek], 2), torch.cat([v, ev], 2)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(BF, -1, C)
    ho, eo = o[:, :N], o[:, N:]
    if "to_add_out.weight" not in w:
        return F.linear(ho, w["to_out.0.weight"], w["to_out.0.bias"]), eo
    return (F.linear(ho, w["to_out.0.weight"], w["to_out.0.bias"]),
            F.linear(eo, w["to_add_out.weight"], w["to_add_out.bias"]))


def attn_param_shapes(C: int, heads: int):
    d = C // heads
    s = {}
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
        s[n + ".weight"], s[n + ".bias"] = (C, C), (C,)
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        s[n + ".weight"] = (d,)
    return s


def seeded_attn_weights(C: int, heads: int, seed: int = 3):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in attn_param_shapes(C, heads).items():
        if k.startswith("norm"):
            out[k] = 1.0 + 0.2 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            out[k] = 0.1 * torch.randn(shp, generator=g)
        else:
            out[k] = torch.randn(shp, generator=g) * shp[1] ** -0.5
    return out
