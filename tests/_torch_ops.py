"""TEST INFRASTRUCTURE: plain-torch definitions of what every kernel behind ``univst_b200.ops`` computes, with the same call
signatures, so that the product's HOST logic (weight packing, buffer layouts, source tables, epilogue arguments, loop control)
can be exercised on a CPU-only machine against the reference goldens.  ``install(monkeypatch)`` swaps them into
``univst_b200.ops``; the kernels themselves are checked on the GPU (``-m gpu``).  fp32 math, fp16 storage between ops."""
import math

import torch
import torch.nn.functional as F


def _f(t):
    return t.float()


def pack_latents(zs, Cpad=64):
    z = torch.stack([z.reshape(z.shape[-4:]) for z in zs])            # (B, C, F, h, w)
    B, C, Fr, h, w = z.shape
    out = torch.zeros(B * Fr, h, w, Cpad, dtype=torch.float16)
    out[..., :C] = z.permute(0, 2, 3, 4, 1).reshape(B * Fr, h, w, C)
    return out


def unpack_latents(x, B, C_, Fr, h, w):
    return x[:, :C_].reshape(B, Fr, h, w, C_).permute(0, 4, 1, 2, 3).contiguous()


def timestep_embedding(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    e = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(e), torch.sin(e)], dim=-1).half()


def _epilogue(acc, bias, rowvec, rows_per_group, residual, bias2, geglu, out_scale, act):
    M, N = acc.shape
    if geglu:   # tile-interleaved columns: [BN/2 values | BN/2 gates] per BN-wide tile (pack.pack_geglu)
        bn = 256 if N % 256 == 0 else 128
        y = acc + (_f(bias) if bias is not None else 0.0)
        y = y.view(M, N // bn, 2, bn // 2)
        return (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(M, N // 2).half()
    y = acc
    if bias is not None:
        y = y + _f(bias)
    if rowvec is not None:
        y = y + _f(rowvec)[torch.arange(M) // rows_per_group, :N]
    if residual is not None:
        y = y + _f(residual)
    y = y * out_scale
    if act in (True, 1, "silu"):
        y = F.silu(y.half().float())
    elif act in (2, "gelu_tanh"):
        y = F.gelu(y.half().float(), approximate="tanh")
    if bias2 is not None:
        y = y.half().float() + _f(bias2)
    return y.half()


def gemm(a, w, *, a2=None, out=None, bias=None, rowvec=None, rows_per_group=1, residual=None, bias2=None, geglu=False,
         out_scale=1.0, act=False):
    x = _f(a) if a2 is None else torch.cat([_f(a), _f(a2)], dim=1)
    y = _epilogue(x @ _f(w).T, bias, rowvec, rows_per_group, residual, bias2, geglu, out_scale, act)
    if out is not None:
        out.copy_(y)
        return out
    return y


def conv3x3(x, w, *, x2=None, stride=1, out=None, bias=None, rowvec=None, rows_per_group=1, residual=None, out_scale=1.0):
    if stride == 2:   # parity planes [4 = (row parity, col parity)][NB, H/2, W/2, C] -> the full-resolution image
        _, NB, h2, w2, C = x.shape
        full = torch.empty(NB, 2 * h2, 2 * w2, C)
        for hp in range(2):
            for wp in range(2):
                full[:, hp::2, wp::2] = _f(x[hp * 2 + wp])
        x = full
    else:
        x = _f(x) if x2 is None else torch.cat([_f(x), _f(x2)], dim=-1)
    Cout = w.shape[0]
    w4 = _f(w).view(Cout, 3, 3, -1).permute(0, 3, 1, 2)                # tap-major [Cout, 3, 3, Cin] -> [Cout, Cin, 3, 3]
    acc = F.conv2d(x.permute(0, 3, 1, 2), w4, stride=stride, padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    y = _epilogue(acc, bias, rowvec, rows_per_group, residual, None, False, out_scale, False)
    if out is not None:
        out.zero_()
        out[:, :Cout] = y
        return out
    return y


def groupnorm(x1, gamma, beta, *, NB, rows, groups=32, eps=1e-5, silu=False, x2=None, out=None):
    x = _f(x1) if x2 is None else torch.cat([_f(x1), _f(x2)], dim=-1)
    C = x.shape[-1]
    y = F.group_norm(x.reshape(NB, rows, C).permute(0, 2, 1), groups, _f(gamma), _f(beta), eps).permute(0, 2, 1).reshape(NB * rows, C)
    return (F.silu(y) if silu else y).half()


def groupnorm_sharded(x1, gamma, beta, *, NB, rows, group, world, groups=32, eps=1e-5, silu=False, x2=None, reduce_fn=None):
    """GroupNorm over the rows of ALL ranks: local sums -> all-reduce -> apply with the global count (frame-sharded UNet)."""
    import torch.distributed as dist
    x = _f(x1) if x2 is None else torch.cat([_f(x1), _f(x2)], dim=-1)
    C = x.shape[-1]
    xg = x.reshape(NB, rows, groups, C // groups)
    sums = torch.stack([xg.sum(dim=(1, 3)), (xg * xg).sum(dim=(1, 3))], dim=-1)      # [NB, groups, 2]
    dist.all_reduce(sums, group=group)
    cnt = rows * world * (C // groups)
    mean = sums[..., 0] / cnt
    var = sums[..., 1] / cnt - mean * mean
    y = (xg - mean[:, None, :, None]) * torch.rsqrt(var + eps)[:, None, :, None]
    y = y.reshape(NB * rows, C) * _f(gamma) + _f(beta)
    return (F.silu(y) if silu else y).half()


def layernorm(x, gamma, beta, eps=1e-5, out=None):
    return F.layer_norm(_f(x), (x.shape[-1],), _f(gamma), _f(beta), eps).half()


def sc_attention(q, k, v, kv_src, *, NI, NIkv, H, d, N, Nkv, out=None):
    heads = lambda t, img, n: _f(t[img * n:(img + 1) * n]).view(n, H, d).transpose(0, 1)   # (H, n, d)
    res = torch.empty(NI * N, H * d, dtype=torch.float16)
    table = kv_src.view(NI, -1).tolist()
    for i in range(NI):
        kk = torch.cat([heads(k, s, Nkv) for s in table[i]], dim=1)
        vv = torch.cat([heads(v, s, Nkv) for s in table[i]], dim=1)
        res[i * N:(i + 1) * N] = F.scaled_dot_product_attention(heads(q, i, N), kk, vv).transpose(0, 1).reshape(N, H * d).half()
    return res


def cross_attention(q, k, v, kv_src, *, NI, NIkv, H, d, N, Nkv, out=None):
    return sc_attention(q, k, v, kv_src.view(NI, 1), NI=NI, NIkv=NIkv, H=H, d=d, N=N, Nkv=Nkv)


def temporal_attention(qkv, *, B, F, N, H, d, out=None):
    """Attention over the F frames of every (branch, pixel): rows are (b, f, pixel)."""
    C = H * d
    x = _f(qkv[:, :3 * C]).view(B, F, N, 3, H, d).permute(3, 0, 2, 4, 1, 5)      # (3, B, N, H, F, d)
    o = F_sdpa(x[0], x[1], x[2])                                                  # (B, N, H, F, d)
    return o.permute(0, 3, 1, 2, 4).reshape(B * F * N, C).half()


def F_sdpa(q, k, v):
    return F.scaled_dot_product_attention(q, k, v)


def attn_shift_(qkv, Fr, N, C_, alpha, beta, gamma):
    """pnp_utils.py:47-57 on the fused [3 F N, 3 C] buffer, in place (branches: content, style, edit)."""
    def adain(cnt, sty):   # pnp_utils.py:114-125 on (frames, tokens, channels)
        return F.instance_norm(cnt) * sty.std(dim=[1], keepdim=True) + sty.mean(dim=[1], keepdim=True)
    q, k, v = (_f(qkv[:, i * C_:(i + 1) * C_]).view(3, Fr, N, C_) for i in range(3))
    q2 = gamma * (alpha * q[0] + (1 - alpha) * q[2])
    k2 = beta * adain(k[2], k[1]) + (1 - beta) * k[1]
    v2 = beta * adain(v[2], v[1]) + (1 - beta) * v[1]
    for i, t in enumerate((q2, k2, v2)):
        qkv[2 * Fr * N:, i * C_:(i + 1) * C_] = t.reshape(Fr * N, C_).half()
    return qkv


def upsample2x(x):
    return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).contiguous()


def space_to_depth2(x):
    return torch.stack([x[:, hp::2, wp::2] for hp in range(2) for wp in range(2)]).contiguous()


def mask_resize(mask_u8, h, w):
    return F.interpolate((mask_u8 != 0)[None].float(), size=(h, w), mode="bilinear", align_corners=False)[0].half()


def latent_blend(a, b, mask, out=None):
    return ((1 - _f(mask)) * _f(a) + _f(mask) * _f(b)).half()


def latent_adain(cnt, sty, out=None):
    """pnp_utils.py:128-139 on (.., C, F, h, w): instance norm over (F, h, w) per channel, style statistics per (channel, frame)."""
    c, s = _f(cnt).reshape(1, *cnt.shape[-4:]), _f(sty).reshape(1, *sty.shape[-4:])
    sm, ss = s.mean(dim=[0, 3, 4], keepdim=True), s.std(dim=[0, 3, 4], keepdim=True)
    return (F.instance_norm(c) * ss + sm).reshape(cnt.shape).half()


def ddim_step(z, eps_rows, branch, alpha_t, alpha_prev, out=None, x0_out=None):
    C_, Fz, h, w = z.shape[-4:]
    e = _f(eps_rows[branch * Fz * h * w:(branch + 1) * Fz * h * w, :C_]).view(Fz, h, w, C_).permute(3, 0, 1, 2)
    x0 = (_f(z) - (1 - alpha_t) ** 0.5 * e) / alpha_t ** 0.5
    return (alpha_prev ** 0.5 * x0 + (1 - alpha_prev) ** 0.5 * e).half()


def axpby(a, b, wa, wb, out=None):
    return (wa * _f(a) + wb * _f(b)).half()


# ------------------------------------------------------------------------------------------------ round 2: VAE / SD3 / xrank-free helpers
def conv3x3_s2_pad_after(planes, w, *, bias=None, out=None):
    _, NB, h2, w2, C = planes.shape
    full = torch.empty(NB, 2 * h2, 2 * w2, C)
    for hp in range(2):
        for wp in range(2):
            full[:, hp::2, wp::2] = _f(planes[hp * 2 + wp])
    Cout = w.shape[0]
    w4 = _f(w).view(Cout, 3, 3, -1).permute(0, 3, 1, 2)
    y = F.conv2d(F.pad(full.permute(0, 3, 1, 2), (0, 1, 0, 1)), w4, None if bias is None else _f(bias), stride=2)
    y = y.permute(0, 2, 3, 1).reshape(-1, Cout).half()
    if out is not None:
        out.copy_(y)
        return out
    return y


def conv_temporal3(x, w, *, NB, F, HW, bias=None, residual=None, out=None):
    C, Cout = x.shape[1], w.shape[0]
    x5 = _f(x).view(NB, F, HW, 1, C).permute(0, 4, 1, 2, 3)
    w5 = _f(w).view(Cout, 3, -1)[:, :, :C].permute(0, 2, 1).reshape(Cout, C, 3, 1, 1)
    y = torch.nn.functional.conv3d(x5, w5, None if bias is None else _f(bias), padding=(1, 0, 0))
    y = y.permute(0, 2, 3, 4, 1).reshape(-1, Cout)
    if residual is not None:
        y = y + _f(residual)
    y = y.half()
    if out is not None:
        out.copy_(y)
        return out
    return y


def softmax_rows_(x, scale=1.0):
    x.copy_(torch.softmax(_f(x) * scale, dim=-1).half())
    return x


def frames_to_u8(x, pixels):
    f = (x[:pixels, :3] / 2 + 0.5).clamp(0, 1).float()
    return (f * 255).round().to(torch.uint8)


def u8_to_frames(img, cpad=64):
    out = torch.zeros(img.numel() // 3, cpad, dtype=torch.float16)
    out[:, :3] = (img.reshape(-1, 3).double() / 127.5 - 1.0).half()
    return out


def vae_sample(moments, noise, C_, Fr, HW, scaling):
    m = _f(moments[:, :2 * C_]).view(Fr, HW, 2 * C_)
    z = m[..., :C_]
    if noise is not None:
        z = z + torch.exp(0.5 * m[..., C_:].clamp(-30, 20)) * _f(noise).view(Fr, C_, HW).transpose(1, 2)
    return (scaling * z).permute(2, 0, 1).reshape(1, C_, Fr, HW).half()


def layernorm_modulate(x, scale, shift, rows_per_sample, eps=1e-6, out=None):
    y = F.layer_norm(_f(x), (x.shape[-1],), eps=eps)
    return (y * (1 + _f(scale).repeat_interleave(rows_per_sample, 0)) + _f(shift).repeat_interleave(rows_per_sample, 0)).half()


def gated_add(x, y, gate, rows_per_sample, out=None):
    return (_f(x) + (_f(gate).repeat_interleave(rows_per_sample, 0) * _f(y)).half().float()).half()


def rmsnorm_heads_(qkv, H, d, wq, wk, eps=1e-6):
    C = H * d
    for blk, w in ((0, wq), (1, wk)):
        if w is None:
            continue
        t = _f(qkv[:, blk * C:(blk + 1) * C]).view(-1, H, d)
        t = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + eps) * _f(w)
        qkv[:, blk * C:(blk + 1) * C] = t.reshape(-1, C).half()
    return qkv


def sd3_attn_shift_(qkv, Fr, N, H, d, alpha, beta, gamma):
    """backbones/video_diffusion_sd3/pnp_utils.py:180-193 + attention_adain :287-300 on the fused [3 F N, 3 H d] buffer."""
    C = H * d

    def adain(cnt, sty):   # (F, H, N, d): style stats over the tokens, content instance norm over (N, d) per (frame, head)
        return F.instance_norm(cnt) * sty.std(dim=[-2], keepdim=True) + sty.mean(dim=[-2], keepdim=True)
    q, k, v = (_f(qkv[:, i * C:(i + 1) * C]).view(3, Fr, N, H, d).permute(0, 1, 3, 2, 4) for i in range(3))
    q2 = gamma * (alpha * q[0] + (1 - alpha) * q[2])
    k2 = beta * adain(k[2], k[1]) + (1 - beta) * k[1]
    v2 = beta * adain(v[2], v[1]) + (1 - beta) * v[1]
    for i, t in enumerate((q2, k2, v2)):
        qkv[2 * Fr * N:, i * C:(i + 1) * C] = t.permute(0, 2, 1, 3).reshape(Fr * N, C).half()
    return qkv


def joint_attention(q, k, v, k2, v2, kv_src, *, NI, NIkv, NIkv2, H, d, N, Nkv, Nkv2, out=None):
    heads = lambda t, img, n: _f(t[img * n:(img + 1) * n]).view(n, H, d).transpose(0, 1)
    res = torch.empty(NI * N, H * d, dtype=torch.float16)
    table = kv_src.view(NI, -1).tolist()
    for i in range(NI):
        ks = [heads(k, s, Nkv) if s < NIkv else heads(k2, s - NIkv, Nkv2) for s in table[i]]
        vs = [heads(v, s, Nkv) if s < NIkv else heads(v2, s - NIkv, Nkv2) for s in table[i]]
        res[i * N:(i + 1) * N] = F.scaled_dot_product_attention(heads(q, i, N), torch.cat(ks, 1), torch.cat(vs, 1)) \
            .transpose(0, 1).reshape(N, H * d).half()
    return res


ROUND2 = ("conv3x3_s2_pad_after", "conv_temporal3", "softmax_rows_", "frames_to_u8", "u8_to_frames", "vae_sample",
          "layernorm_modulate", "gated_add", "rmsnorm_heads_", "sd3_attn_shift_", "joint_attention")


def install(monkeypatch=None):
    """Swap the definitions into ``univst_b200.ops`` (through ``monkeypatch`` in a test, directly in a worker script)."""
    from univst_b200 import ops
    for name in ("pack_latents", "unpack_latents", "timestep_embedding", "gemm", "conv3x3", "groupnorm", "groupnorm_sharded",
                 "layernorm", "sc_attention", "cross_attention", "temporal_attention", "attn_shift_", "upsample2x",
                 "space_to_depth2", "mask_resize", "latent_blend", "latent_adain", "ddim_step", "axpby") + ROUND2:
        if monkeypatch is not None:
            monkeypatch.setattr(ops, name, globals()[name])
        else:
            setattr(ops, name, globals()[name])
