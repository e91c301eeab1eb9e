"""GPU parity tests of the individual sm_100a kernels against plain PyTorch fp32 references of the same op
(the floating-point kernels' checker; the end-to-end oracle comparisons live in test_unet_gpu.py).

Tolerances are stated per test: inputs are fp16, the references are evaluated in fp32 from the same fp16 values,
outputs are fp16 -> the bound is a few fp16 ulps of the output magnitude plus accumulation-order noise.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).half().cuda()


def _close(out, ref, atol, rtol, what):
    out32, ref32 = out.float(), ref.float()
    err = (out32 - ref32).abs()
    bound = atol + rtol * ref32.abs()
    bad = (err > bound).sum().item()
    msg = (f"{what}: max_abs_err={err.max().item():.4e} ref_absmax={ref32.abs().max().item():.3e} "
           f"violations={bad}/{err.numel()}")
    print(msg)
    assert torch.isfinite(out32).all(), f"{what}: non-finite output"
    assert bad == 0, msg


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 128, 128), (300, 320, 320), (1000, 1280, 320), (4096, 640, 2560),
                                   (64, 32, 64), (3, 1280, 320), (200, 160, 96)])
def test_gemm_plain(cuda_lib, M, N, K):
    from univst_b200 import ops
    a, w = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2)
    out = ops.gemm(a, w)
    _close(out, a.float() @ w.float().t(), 2e-3, 2e-3, f"gemm {M}x{N}x{K}")


def test_gemm_epilogue_bias_residual_rowvec(cuda_lib):
    from univst_b200 import ops
    M, N, K, rpg = 768, 320, 640, 256
    a, w = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2)
    bias, res, rv = _rand(N, seed=3), _rand(M, N, seed=4), _rand(M // rpg, N, seed=5)
    out = ops.gemm(a, w, bias=bias, residual=res, rowvec=rv, rows_per_group=rpg, out_scale=0.5)
    ref = (a.float() @ w.float().t() + bias.float() + res.float() + rv.float().repeat_interleave(rpg, 0)) * 0.5
    _close(out, ref, 3e-3, 2e-3, "gemm bias+residual+rowvec")


def test_gemm_two_sources_and_strided_views(cuda_lib):
    from univst_b200 import ops
    M, N, K1, K2 = 512, 256, 128, 192
    big = _rand(M, K1 + 64, seed=1)
    a = big[:, :K1]  # row-strided view
    a2, w = _rand(M, K2, seed=6), _rand(N, K1 + K2, scale=(K1 + K2) ** -0.5, seed=2)
    outbuf = torch.zeros(M, N + 32, dtype=torch.float16, device="cuda")
    ops.gemm(a, w, a2=a2, out=outbuf[:, :N])
    ref = torch.cat([a, a2], 1).float() @ w.float().t()
    _close(outbuf[:, :N], ref, 2e-3, 2e-3, "gemm two sources")
    assert (outbuf[:, N:] == 0).all()


def test_gemm_geglu(cuda_lib):
    from univst_b200 import ops
    from univst_b200.pack import pack_geglu
    M, C = 640, 320
    a, w, b = _rand(M, C, seed=1), _rand(8 * C, C, scale=C ** -0.5, seed=2), _rand(8 * C, seed=3)
    wp, bp = pack_geglu(w, b)
    out = ops.gemm(a, wp, bias=bp, geglu=True)
    proj = (a.float() @ w.float().t() + b.float()).half().float()
    h, gate = proj.chunk(2, -1)
    ref = h * F.gelu(gate).half().float()
    _close(out, ref, 3e-3, 3e-3, "gemm geglu")


def test_gemm_act_bias2(cuda_lib):
    from univst_b200 import ops
    M, N, K = 3, 1280, 320
    a, w, b, b2 = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2), _rand(N, seed=3), _rand(N, seed=4)
    out = ops.gemm(a, w, bias=b, act=True)
    _close(out, F.silu((a.float() @ w.float().t() + b.float()).half().float()), 2e-3, 2e-3, "gemm silu")
    out = ops.gemm(a, w, bias=b, bias2=b2)
    _close(out, (a.float() @ w.float().t() + b.float()).half().float() + b2.float(), 2e-3, 2e-3, "gemm bias2")


# ------------------------------------------------------------------------------------------------ conv
def _conv_ref(x_nhwc, w, stride=1):
    Cout = w.shape[0]
    w4 = w.view(Cout, 3, 3, -1).permute(0, 3, 1, 2).float()
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2).float(), w4, stride=stride, padding=1)
    return y.permute(0, 2, 3, 1).reshape(-1, Cout)


@pytest.mark.parametrize("NB,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (3, 8, 8, 128, 96), (1, 64, 64, 64, 320),
                                             (12, 4, 4, 64, 128), (2, 32, 32, 320, 640), (5, 8, 8, 64, 4)])
def test_conv3x3(cuda_lib, NB, H, W, Cin, Cout):
    from univst_b200 import ops
    x, w = _rand(NB, H, W, Cin, seed=1), _rand(Cout, 9 * Cin, scale=(9 * Cin) ** -0.5, seed=2)
    bias = _rand(Cout, seed=3)
    out = ops.conv3x3(x, w, bias=bias)
    _close(out, _conv_ref(x, w) + bias.float(), 3e-3, 2e-3, f"conv3x3 {NB}x{H}x{W} {Cin}->{Cout}")


def test_conv3x3_two_sources_residual_rowvec(cuda_lib):
    from univst_b200 import ops
    NB, H, W, C1, C2, Cout = 6, 16, 16, 128, 64, 128
    x1, x2 = _rand(NB, H, W, C1, seed=1), _rand(NB, H, W, C2, seed=2)
    w = _rand(Cout, 9 * (C1 + C2), scale=(9 * (C1 + C2)) ** -0.5, seed=3)
    res, rv = _rand(NB * H * W, Cout, seed=4), _rand(3, Cout, seed=5)
    out = ops.conv3x3(x1, w, x2=x2, residual=res, rowvec=rv, rows_per_group=2 * H * W)
    ref = _conv_ref(torch.cat([x1, x2], -1), w) + res.float() + rv.float().repeat_interleave(2 * H * W, 0)
    _close(out, ref, 3e-3, 2e-3, "conv3x3 two sources")


@pytest.mark.parametrize("NB,H,W,C,Cout", [(2, 16, 16, 64, 64), (3, 64, 64, 64, 128), (4, 8, 8, 128, 128)])
def test_conv3x3_stride2(cuda_lib, NB, H, W, C, Cout):
    from univst_b200 import ops
    x, w = _rand(NB, H, W, C, seed=1), _rand(Cout, 9 * C, scale=(9 * C) ** -0.5, seed=2)
    planes = ops.space_to_depth2(x)
    out = ops.conv3x3(planes, w, stride=2)
    _close(out, _conv_ref(x, w, stride=2), 3e-3, 2e-3, f"conv3x3 stride 2 {NB}x{H}x{W}")


@pytest.mark.parametrize("kind,shape", [
    ("conv", (2, 8, 8, 1280, 0, 1280)),      # the 8 x 8 level of a 2-frame shard: 5 tiles x 180 k-blocks -> 29 slices
    ("conv", (2, 16, 16, 640, 640, 1280)),   # two sources (skip concat), 20 tiles x 180 k-blocks -> 7 slices
    ("conv", (6, 8, 8, 1280, 1280, 1280)),   # 15 tiles x 360 k-blocks, residual + per-branch row vector
    ("gemm", (384, 1280, 5120)),             # GEGLU output projection at the deepest level: 15 tiles x 80 k-blocks
    ("gemm", (130, 320, 2560)),              # ragged M (two row tiles, the second almost empty), 2 column tiles
])
def test_split_k_matches_reference_and_unsplit(cuda_lib, kind, shape):
    """Split-K (univst_gemm_tune): the K loop of launches with few output tiles is spread over the idle SMs, the slices'
    fp32 accumulators are added in slice order by the last epilogue warp to arrive.  Against the fp32 reference within the
    usual bound, against the unsplit kernel within fp32 summation-order noise (a few fp16 ulps), and repeatable bit for bit."""
    from univst_b200 import ops
    if kind == "conv":
        NB, H, W, C1, C2, Cout = shape
        x1 = _rand(NB, H, W, C1, seed=1)
        x2 = _rand(NB, H, W, C2, seed=2) if C2 else None
        w = _rand(Cout, 9 * (C1 + C2), scale=(9 * (C1 + C2)) ** -0.5, seed=3)
        bias, res, rv = _rand(Cout, seed=4), _rand(NB * H * W, Cout, seed=5), _rand(NB, Cout, seed=6)
        run = lambda: ops.conv3x3(x1, w, x2=x2, bias=bias, residual=res, rowvec=rv, rows_per_group=H * W)
        xin = x1 if x2 is None else torch.cat([x1, x2], -1)
        ref = _conv_ref(xin, w) + bias.float() + res.float() + rv.float().repeat_interleave(H * W, 0)
    else:
        M, N, K = shape
        a, w = _rand(M, K, seed=1), _rand(N, K, scale=K ** -0.5, seed=2)
        bias, res = _rand(N, seed=3), _rand(M, N, seed=4)
        run = lambda: ops.gemm(a, w, bias=bias, residual=res)
        ref = a.float() @ w.float().t() + bias.float() + res.float()
    unsplit = run()
    try:
        ops.gemm_splitk(74)
        split = run()
        again = run()
    finally:
        ops.gemm_splitk(0)
    _close(split, ref, 3e-3, 2e-3, f"split-K {kind} {shape}")
    _close(split, unsplit, 2e-3, 2e-3, f"split-K vs unsplit {kind} {shape}")
    assert torch.equal(split, again), "split-K must be deterministic (slices are added in slice order)"


@pytest.mark.parametrize("NB,H,W,C1,C2,Cout,stride", [
    (3, 24, 40, 64, 0, 96, 1),      # 128 // 40 = 3 full rows per tile, 8 row blocks per image
    (5, 3, 5, 64, 0, 64, 1),        # 15-pixel images: 8 whole images per tile, ragged image tail
    (2, 12, 20, 128, 64, 128, 1),   # two sources, 6 rows per tile
    (7, 6, 10, 64, 0, 160, 1),      # 60 pixels: 2 images per tile, odd image count
    (2, 72, 128, 64, 0, 64, 1),     # widest supported row (one row per tile), non-power-of-two height
    (1, 6, 160, 64, 0, 64, 1),      # rows wider than a tile (1280-pixel frames): one 128-pixel segment of a row per tile
    (2, 3, 256, 64, 64, 96, 1),     # power-of-two width above 128, two sources
    (2, 4, 144, 64, 0, 64, 2),      # stride 2 from 8 x 288
    (3, 12, 20, 64, 0, 128, 2),     # stride 2 from 24 x 40
    (4, 3, 5, 128, 0, 64, 2),       # stride 2 from 6 x 10
])
def test_conv3x3_any_size(cuda_lib, NB, H, W, C1, C2, Cout, stride):
    """Images whose sides are not powers of two (latents of 192 x 320, 576 x 1024 ... frames): row-block tiles."""
    from univst_b200 import ops
    Cin = C1 + C2
    Hi, Wi = H * stride, W * stride
    x = _rand(NB, Hi, Wi, Cin, seed=1)
    w = _rand(Cout, 9 * Cin, scale=(9 * Cin) ** -0.5, seed=2)
    bias, res, rv = _rand(Cout, seed=3), _rand(NB * H * W, Cout, seed=4), _rand(NB, Cout, seed=5)
    ref = _conv_ref(x, w, stride=stride) + bias.float() + res.float() + rv.float().repeat_interleave(H * W, 0)
    if stride == 2:
        out = ops.conv3x3(ops.space_to_depth2(x), w, stride=2, bias=bias, residual=res, rowvec=rv, rows_per_group=H * W)
    elif C2:
        out = ops.conv3x3(x[..., :C1].contiguous(), w, x2=x[..., C1:].contiguous(), bias=bias, residual=res, rowvec=rv,
                          rows_per_group=H * W)
    else:
        out = ops.conv3x3(x, w, bias=bias, residual=res, rowvec=rv, rows_per_group=H * W)
    _close(out, ref, 3e-3, 2e-3, f"conv3x3 {NB}x{H}x{W} stride {stride} {Cin}->{Cout}")


def test_upsample2x(cuda_lib):
    from univst_b200 import ops
    x = _rand(3, 8, 16, 64, seed=1)
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(ops.upsample2x(x).float(), ref)


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, src, NI, H, d, N, Nkv):
    C = H * d
    q4 = q.float().view(NI, N, H, d).transpose(1, 2)
    NIkv = k.shape[0] // Nkv
    k4 = k.float().view(NIkv, Nkv, H, d)
    v4 = v.float().view(NIkv, Nkv, H, d)
    kk = torch.cat([k4[src[:, j].long()] for j in range(src.shape[1])], 1).transpose(1, 2)
    vv = torch.cat([v4[src[:, j].long()] for j in range(src.shape[1])], 1).transpose(1, 2)
    o = F.scaled_dot_product_attention(q4, kk, vv)
    return o.transpose(1, 2).reshape(NI * N, C)


@pytest.mark.parametrize("B,Fr,H,d,N,mode", [
    (1, 2, 2, 64, 256, "prev_first"), (1, 3, 2, 64, 128, "prev_self_first"), (3, 2, 8, 40, 256, "prev_first"),
    (1, 2, 4, 80, 256, "prev_self_first"), (1, 3, 2, 160, 64, "prev_self_first"), (1, 2, 4, 16, 64, "prev_first"),
    (1, 2, 2, 32, 1024, "self"), (1, 4, 8, 40, 1024, "prev_first")])
def test_sc_attention(cuda_lib, B, Fr, H, d, N, mode):
    from univst_b200 import ops
    from univst_b200.unet import kv_source_table
    NI, C = B * Fr, H * d
    qkv = _rand(NI * N, 3 * C, seed=1)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    src = kv_source_table(B, Fr, mode).cuda()
    out = ops.sc_attention(q, k, v, src, NI=NI, NIkv=NI, H=H, d=d, N=N, Nkv=N)
    ref = _attn_ref(q.contiguous(), k.contiguous(), v.contiguous(), src, NI, H, d, N, N)
    _close(out, ref, 2e-3, 5e-3, f"sc_attention B{B} F{Fr} H{H} d{d} N{N} {mode}")


def test_sc_attention_large_logits(cuda_lib):
    """Peaked softmax (exercises the lazy O rescale): scores spread over ~ +-60."""
    from univst_b200 import ops
    from univst_b200.unet import kv_source_table
    B, Fr, H, d, N = 1, 2, 2, 64, 512
    NI, C = B * Fr, H * d
    qkv = _rand(NI * N, 3 * C, scale=2.5, seed=3)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    src = kv_source_table(B, Fr, "prev_self_first").cuda()
    out = ops.sc_attention(q, k, v, src, NI=NI, NIkv=NI, H=H, d=d, N=N, Nkv=N)
    ref = _attn_ref(q.contiguous(), k.contiguous(), v.contiguous(), src, NI, H, d, N, N)
    _close(out, ref, 5e-3, 1e-2, "sc_attention peaked")


@pytest.mark.parametrize("d", [40, 32, 64])
def test_sc_attention_peaked_small_head_dim(cuda_lib, d):
    """The split-row kernel (d <= 48) keeps one running maximum per half row: peaked scores force its rescale path
    (P pieces already stored, both accumulators) and the merge of two halves with very different maxima."""
    from univst_b200 import ops
    from univst_b200.unet import kv_source_table
    B, Fr, H, N = 1, 3, 2, 384
    NI, C = B * Fr, H * d
    qkv = _rand(NI * N, 3 * C, scale=2.5, seed=5)
    qkv[:, C:2 * C][::7] *= 3.0          # a few keys dominate, scattered over both halves of the tiles
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    src = kv_source_table(B, Fr, "prev_self_first").cuda()
    out = ops.sc_attention(q, k, v, src, NI=NI, NIkv=NI, H=H, d=d, N=N, Nkv=N)
    ref = _attn_ref(q.contiguous(), k.contiguous(), v.contiguous(), src, NI, H, d, N, N)
    _close(out, ref, 5e-3, 1e-2, f"sc_attention peaked d{d}")


@pytest.mark.parametrize("H,d,N,Nkv,nsrc", [(2, 40, 200, 330, 2), (2, 40, 128, 64, 3), (3, 16, 72, 136, 2), (2, 40, 256, 192, 1),
                                            (2, 64, 200, 330, 2), (2, 80, 200, 330, 2)])
def test_sc_attention_ragged(cuda_lib, H, d, N, Nkv, nsrc):
    """Query / key counts that are not multiples of the 128-token tiles: partial query tiles, a ragged last KV tile
    per source, and (Nkv = 64, 192) key halves of a tile that are entirely past the end (their half-row softmax must
    contribute exactly nothing, not NaN)."""
    from univst_b200 import ops
    NI, NIkv, C = 4, 5, H * d
    q = _rand(NI * N, C, seed=1)
    kv = _rand(NIkv * Nkv, 2 * C, seed=2)
    k, v = kv[:, :C], kv[:, C:]
    g = torch.Generator().manual_seed(9)
    src = torch.stack([torch.randperm(NIkv, generator=g)[:nsrc] for _ in range(NI)]).to(torch.int32).cuda()
    out = ops.sc_attention(q, k, v, src, NI=NI, NIkv=NIkv, H=H, d=d, N=N, Nkv=Nkv)
    ref = _attn_ref(q, k.contiguous(), v.contiguous(), src, NI, H, d, N, Nkv)
    assert torch.isfinite(out).all()
    _close(out, ref, 2e-3, 5e-3, f"sc_attention ragged H{H} d{d} N{N} Nkv{Nkv} x{nsrc}")


@pytest.mark.parametrize("H,d,N", [(8, 40, 256), (8, 160, 64), (5, 64, 128)])
def test_cross_attention_77_tokens(cuda_lib, H, d, N):
    from univst_b200 import ops
    NI, C, Nkv = 6, H * d, 77
    q, kv = _rand(NI * N, C, seed=1), _rand(Nkv, 2 * C, seed=2)
    k, v = kv[:, :C], kv[:, C:]
    src = torch.zeros(NI, 1, dtype=torch.int32, device="cuda")
    out = ops.sc_attention(q, k, v, src, NI=NI, NIkv=1, H=H, d=d, N=N, Nkv=Nkv)
    ref = _attn_ref(q, k.contiguous(), v.contiguous(), src, NI, H, d, N, Nkv)
    _close(out, ref, 2e-3, 5e-3, f"cross attention H{H} d{d} N{N}")


@pytest.mark.parametrize("H,d,N,Nkv,NIkv", [(8, 40, 256, 77, 3), (8, 160, 64, 77, 3), (5, 64, 200, 77, 2), (8, 80, 1024, 77, 3),
                                             (4, 16, 72, 33, 1), (2, 8, 128, 80, 2), (2, 40, 130, 1, 1)])
def test_cross_attention_kernel(cuda_lib, H, d, N, Nkv, NIkv):
    """The dedicated short-context kernel (mma.sync fragments, K/V of a branch in shared memory) vs fp32 SDPA."""
    from univst_b200 import ops
    NI, C = 6, H * d
    q, kv = _rand(NI * N, C, seed=1), _rand(NIkv * Nkv, 2 * C, seed=2)
    k, v = kv[:, :C], kv[:, C:]
    src = (torch.arange(NI, dtype=torch.int32) % NIkv).view(NI, 1).cuda()
    out = ops.cross_attention(q, k, v, src, NI=NI, NIkv=NIkv, H=H, d=d, N=N, Nkv=Nkv)
    ref = _attn_ref(q, k.contiguous(), v.contiguous(), src, NI, H, d, N, Nkv)
    assert torch.isfinite(out).all()
    _close(out, ref, 2e-3, 5e-3, f"cross_attention H{H} d{d} N{N} Nkv{Nkv}")


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("NB,rows,C1,C2,silu", [(3, 1024, 320, 0, True), (3, 256, 1280, 640, True), (12, 64, 64, 0, False),
                                                (2, 4096, 320, 0, False), (3, 100, 128, 64, True)])
def test_groupnorm(cuda_lib, NB, rows, C1, C2, silu):
    from univst_b200 import ops
    C = C1 + C2
    x1 = _rand(NB * rows, C1, seed=1) + 0.5
    x2 = _rand(NB * rows, C2, scale=2.0, seed=2) if C2 else None
    g, b = _rand(C, seed=3), _rand(C, seed=4)
    out = ops.groupnorm(x1, g, b, NB=NB, rows=rows, groups=32, eps=1e-5, silu=silu, x2=x2)
    x = torch.cat([x1, x2], 1) if C2 else x1
    ref = F.group_norm(x.float().view(NB, rows, C).transpose(1, 2), 32, g.float(), b.float(), 1e-5)
    if silu:
        ref = F.silu(ref.half().float())
    _close(out, ref.transpose(1, 2).reshape(NB * rows, C), 4e-3, 4e-3, f"groupnorm {NB}x{rows}x{C}")


@pytest.mark.parametrize("rows,C", [(1000, 320), (77, 1280), (4096, 640), (10, 64)])
def test_layernorm(cuda_lib, rows, C):
    from univst_b200 import ops
    x, g, b = _rand(rows, C, seed=1) * 2 + 0.3, _rand(C, seed=2), _rand(C, seed=3)
    _close(ops.layernorm(x, g, b), F.layer_norm(x.float(), (C,), g.float(), b.float(), 1e-5), 4e-3, 4e-3,
           f"layernorm {rows}x{C}")


# ------------------------------------------------------------------------------------------------ AdaIN shift
def _attention_adain_ref(cnt, sty):
    # restatement of pnp_utils.py:114-125 in fp32
    return F.instance_norm(cnt) * sty.std(dim=[1], keepdim=True) + sty.mean(dim=[1], keepdim=True)


@pytest.mark.parametrize("Fr,N,C", [(2, 64, 64), (4, 256, 320), (2, 1024, 1280)])
def test_attn_shift(cuda_lib, Fr, N, C):
    from univst_b200 import ops
    alpha, beta, gamma = 0.65, 0.516, 3.0
    qkv = _rand(3 * Fr * N, 3 * C, seed=1) + 0.25
    ref = qkv.float().view(3, Fr, N, 3, C).clone()
    q, k, v = ref[:, :, :, 0], ref[:, :, :, 1], ref[:, :, :, 2]
    q2 = gamma * (alpha * q[0] + (1 - alpha) * q[2]).half().float()
    k2 = beta * _attention_adain_ref(k[2], k[1]) + (1 - beta) * k[1]
    v2 = beta * _attention_adain_ref(v[2], v[1]) + (1 - beta) * v[1]
    out = ops.attn_shift_(qkv.clone(), Fr, N, C, alpha, beta, gamma).float().view(3, Fr, N, 3, C)
    assert torch.equal(out[:2], ref[:2]), "content / style branches must be untouched"
    _close(out[2, :, :, 0], q2, 4e-3, 3e-3, "shift Q")
    _close(out[2, :, :, 1], k2, 4e-3, 3e-3, "shift K")
    _close(out[2, :, :, 2], v2, 4e-3, 3e-3, "shift V")


# ------------------------------------------------------------------------------------------------ latent ops
def test_latent_ops(cuda_lib):
    from univst_b200 import ops
    C, Fr, h, w = 4, 5, 16, 16
    zc, zs, z = _rand(1, C, Fr, h, w, seed=1), _rand(1, C, Fr, h, w, seed=2) * 0.7 + 0.1, _rand(1, C, Fr, h, w, seed=3)
    x = ops.pack_latents([zc, zs, z], Cpad=64)
    ref = torch.cat([zc, zs, z]).permute(0, 2, 3, 4, 1).reshape(3 * Fr, h, w, C)
    assert torch.equal(x[..., :C], ref) and (x[..., C:] == 0).all()
    back = ops.unpack_latents(x.view(-1, 64), 3, C, Fr, h, w)
    assert torch.equal(back, torch.cat([zc, zs, z]))
    # latent_adain vs a fp32 restatement of pnp_utils.py:128-139
    zf, sf = z.float(), zs.float()
    ref = F.instance_norm(zf) * sf.std(dim=[0, 3, 4], keepdim=True) + sf.mean(dim=[0, 3, 4], keepdim=True)
    _close(ops.latent_adain(z, zs), ref, 4e-3, 4e-3, "latent_adain")
    # mask resize + blend
    m = (torch.rand(Fr, 128, 128) > 0.5).to(torch.uint8).cuda() * 255
    mr = ops.mask_resize(m, h, w)
    mref = F.interpolate((m != 0).float()[None], size=(h, w), mode="bilinear", align_corners=False)[0]
    _close(mr, mref, 1e-3, 0, "mask_resize")
    _close(ops.latent_blend(z, zc, mr), (1 - mr.float()) * zf + mr.float() * zc.float(), 2e-3, 2e-3, "latent_blend")
    # ddim step
    eps = _rand(3 * Fr * h * w, 32, seed=4)
    a_t, a_p = 0.3, 0.45
    e = eps[2 * Fr * h * w:, :C].float().view(Fr, h, w, C).permute(3, 0, 1, 2)[None]
    x0 = (zf - math.sqrt(1 - a_t) * e) / math.sqrt(a_t)
    _close(ops.ddim_step(z, eps, 2, a_t, a_p), math.sqrt(a_p) * x0 + math.sqrt(1 - a_p) * e, 3e-3, 3e-3, "ddim_step")
    t = torch.tensor([981.0, 981.0, 1.0], device="cuda")
    emb = ops.timestep_embedding(t, 320)
    freqs = torch.exp(-math.log(10000.0) * torch.arange(160, device="cuda").float() / 160)
    e_ref = t[:, None] * freqs[None]
    _close(emb, torch.cat([e_ref.cos(), e_ref.sin()], -1), 2e-3, 0, "timestep_embedding")
