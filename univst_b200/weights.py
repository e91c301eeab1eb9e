"""Random-initialised weights of the reference UNet architecture (no checkpoint can be downloaded offline).

``unet_param_shapes`` lists the ``state_dict`` keys / shapes of the reference ``UNetPseudo3DConditionModel``
(= SD ``UNet2DConditionModel`` names + the never-loaded ``*_temporal*`` keys, models/unet_3d_condition.py:493-509);
``random_state_dict`` fills them on the device: weights ~ N(0, 1/fan_in), biases ~ N(0, 0.02^2), norm scales 1, the
temporal parts at the reference's constructor values (Dirac conv_temporal, zero attn_temporal output projection).
"""
from __future__ import annotations

from typing import Dict

import torch


def unet_param_shapes(cfg) -> Dict[str, tuple]:
    """Key -> shape of the reference module tree (verified against the reference constructor in gen_golden.py)."""
    boc, lpb = cfg["block_out_channels"], cfg["layers_per_block"]
    D, temb_dim, nlev = cfg["cross_attention_dim"], boc[0] * 4, len(boc)
    s: Dict[str, tuple] = {}

    def conv(pre, cin, cout, k):
        s[pre + "weight"], s[pre + "bias"] = (cout, cin, k, k), (cout,)
        if k > 1:
            s[pre + "conv_temporal.weight"], s[pre + "conv_temporal.bias"] = (cout, cout, k), (cout,)

    def lin(pre, cin, cout, bias=True):
        s[pre + "weight"] = (cout, cin)
        if bias:
            s[pre + "bias"] = (cout,)

    def norm(pre, c):
        s[pre + "weight"], s[pre + "bias"] = (c,), (c,)

    def res(pre, cin, cout):
        norm(pre + "norm1.", cin), conv(pre + "conv1.", cin, cout, 3), lin(pre + "time_emb_proj.", temb_dim, cout)
        norm(pre + "norm2.", cout), conv(pre + "conv2.", cout, cout, 3)
        if cin != cout:
            conv(pre + "conv_shortcut.", cin, cout, 1)

    def attn(pre, c, kv):
        lin(pre + "to_q.", c, c, False), lin(pre + "to_k.", kv, c, False), lin(pre + "to_v.", kv, c, False)
        lin(pre + "to_out.0.", c, c)

    def tr(pre, c):
        norm(pre + "norm.", c)
        if cfg["use_linear_projection"]:
            lin(pre + "proj_in.", c, c), lin(pre + "proj_out.", c, c)
        else:
            conv(pre + "proj_in.", c, c, 1), conv(pre + "proj_out.", c, c, 1)
        b = pre + "transformer_blocks.0."
        attn(b + "attn1.", c, c), norm(b + "norm1.", c), attn(b + "attn2.", c, D), norm(b + "norm2.", c)
        attn(b + "attn_temporal.", c, c), norm(b + "norm_temporal.", c)
        lin(b + "ff.net.0.proj.", c, 8 * c), lin(b + "ff.net.2.", 4 * c, c), norm(b + "norm3.", c)

    conv("conv_in.", cfg["in_channels"], boc[0], 3)
    lin("time_embedding.linear_1.", boc[0], temb_dim), lin("time_embedding.linear_2.", temb_dim, temb_dim)
    skip_ch = [boc[0]]
    cout = boc[0]
    for i in range(nlev):
        cin, cout = cout, boc[i]
        for j in range(lpb):
            res(f"down_blocks.{i}.resnets.{j}.", cin if j == 0 else cout, cout)
            if i < nlev - 1:
                tr(f"down_blocks.{i}.attentions.{j}.", cout)
            skip_ch.append(cout)
        if i < nlev - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv.", cout, cout, 3)
            skip_ch.append(cout)
    res("mid_block.resnets.0.", boc[-1], boc[-1]), tr("mid_block.attentions.0.", boc[-1])
    res("mid_block.resnets.1.", boc[-1], boc[-1])
    rev = list(reversed(boc))
    x_ch = boc[-1]
    for i in range(nlev):
        cout = rev[i]
        for j in range(lpb + 1):
            res(f"up_blocks.{i}.resnets.{j}.", x_ch + skip_ch.pop(), cout)
            x_ch = cout
            if i > 0:
                tr(f"up_blocks.{i}.attentions.{j}.", cout)
        if i < nlev - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv.", cout, cout, 3)
    norm("conv_norm_out.", boc[0]), conv("conv_out.", boc[0], cfg["out_channels"], 3)
    return s


def animatediff_param_shapes(cfg) -> Dict[str, tuple]:
    """Key -> shape of the reference AnimateDiff ``UNet3DConditionModel`` with the animatediff-v2.yaml kwargs
    (backbones/animatediff/models/unet.py:41): the SD tree without ``*_temporal*`` keys plus one motion module
    (models/motion_module.py:52) per (resnet, attention) pair -- 21 in all."""
    s = {k: v for k, v in unet_param_shapes(cfg).items() if "_temporal" not in k}
    boc, lpb, nlev = cfg["block_out_channels"], cfg["layers_per_block"], len(cfg["block_out_channels"])

    def mm(pre, c):
        t = pre + "temporal_transformer."
        s[t + "norm.weight"], s[t + "norm.bias"] = (c,), (c,)
        s[t + "proj_in.weight"], s[t + "proj_in.bias"] = (c, c), (c,)
        s[t + "proj_out.weight"], s[t + "proj_out.bias"] = (c, c), (c,)
        b = t + "transformer_blocks.0."
        for i in range(2):
            a = b + f"attention_blocks.{i}."
            s[a + "to_q.weight"] = s[a + "to_k.weight"] = s[a + "to_v.weight"] = s[a + "to_out.0.weight"] = (c, c)
            s[a + "to_out.0.bias"] = (c,)
            s[b + f"norms.{i}.weight"], s[b + f"norms.{i}.bias"] = (c,), (c,)
        s[b + "ff.net.0.proj.weight"], s[b + "ff.net.0.proj.bias"] = (8 * c, c), (8 * c,)
        s[b + "ff.net.2.weight"], s[b + "ff.net.2.bias"] = (c, 4 * c), (c,)
        s[b + "ff_norm.weight"], s[b + "ff_norm.bias"] = (c,), (c,)

    for i in range(nlev):
        for j in range(lpb):
            mm(f"down_blocks.{i}.motion_modules.{j}.", boc[i])
    mm("mid_block.motion_modules.0.", boc[-1])
    rev = list(reversed(boc))
    for i in range(nlev):
        for j in range(lpb + 1):
            mm(f"up_blocks.{i}.motion_modules.{j}.", rev[i])
    return s


def random_state_dict(cfg, seed: int = 33, device="cuda", dtype=torch.float16, animatediff: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for key, shape in (animatediff_param_shapes(cfg) if animatediff else unet_param_shapes(cfg)).items():
        if "conv_temporal.weight" in key:
            t = torch.zeros(shape, device=device)
            torch.nn.init.dirac_(t)
        elif "conv_temporal.bias" in key or "attn_temporal.to_out.0.weight" in key:
            t = torch.zeros(shape, device=device)
        elif key.endswith("weight") and len(shape) == 1:
            t = torch.ones(shape, device=device)
        elif key.endswith("bias"):
            t = 0.02 * torch.randn(shape, device=device, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, device=device, generator=g) * fan_in ** -0.5
        out[key] = t.to(dtype)
    return out
