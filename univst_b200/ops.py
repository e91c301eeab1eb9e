"""Tensor-facing wrappers over the C ABI (``include/univst_b200.h``).

PyTorch is used for device memory and streams only: every function takes CUDA fp16 tensors, passes raw
pointers / sizes / the current stream through ctypes and returns the output tensor.  Activations are
channels-last: ``[images, H, W, C]`` viewed as ``[tokens, C]``.  Nothing here falls back to PyTorch math.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import Epilogue, check

_vp, _i32, _f32, _i64 = C.c_void_p, C.c_int32, C.c_float, C.c_int64
_lib.register("univst_sc_attention_f16", [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp])
_lib.register("univst_attention_tune", [_i32, _i32, _i32])
_lib.register("univst_joint_attention_f16", [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp])
_lib.register("univst_rmsnorm_heads_f16", [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _f32, _vp])
_lib.register("univst_sd3_shift_workspace_bytes", [_i32, _i32, _i32], _i64)
_lib.register("univst_sd3_attn_shift_f16", [_vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _vp, _vp])
_lib.register("univst_cross_attention_supported", [_i32, _i32])
_lib.register("univst_cross_attention_f16", [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp])
_lib.register("univst_sc_attention_sharded_f16", [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp])
_lib.register("univst_temporal_attention_f16", [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp])
_lib.register("univst_attn_shift_workspace_bytes", [_i32, _i32], _i64)
_lib.register("univst_attn_shift_f16", [_vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _vp, _vp])
_lib.register("univst_groupnorm_workspace_bytes", [_i32, _i32], _i64)
_lib.register("univst_groupnorm_f16", [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _f32, _i32, _vp, _vp, _vp])
_lib.register("univst_groupnorm_stats_f16", [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp])
_lib.register("univst_groupnorm_apply_f16", [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _f32, _i32, _vp, _vp])
_lib.register("univst_layernorm_f16", [_vp, _i32, _i32, _vp, _vp, _f32, _vp, _vp])
_lib.register("univst_upsample2x_f16", [_vp, _i32, _i32, _i32, _i32, _vp, _vp])
_lib.register("univst_space_to_depth2_f16", [_vp, _i32, _i32, _i32, _i32, _vp, _vp])
_lib.register("univst_pack_latents_f16", [C.POINTER(_vp), _i32, _i32, _i32, _i32, _i32, _vp, _vp])
_lib.register("univst_unpack_latents_f16", [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp])
_lib.register("univst_timestep_embedding_f16", [_vp, _i32, _i32, _vp, _vp])
_lib.register("univst_mask_resize_u8", [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp])
_lib.register("univst_latent_blend_f16", [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp])
_lib.register("univst_latent_blend_fc_f16", [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp])
_lib.register("univst_latent_adain_f16", [_vp, _vp, _i32, _i32, _i32, _vp, _vp])
_lib.register("univst_ddim_step_f16", [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _vp])
_lib.register("univst_axpby_f16", [_vp, _vp, _f32, _f32, _i64, _vp, _vp])
_lib.register("univst_halo_push_f16", [_vp, _i32, _i64, C.POINTER(_vp), _i32, _i32, _i64, _i32, _i32, _i32, _vp])
_lib.register("univst_exchange_push_f16", [_i32, _vp, _i32, C.POINTER(_vp), _i32, _i32, _i32, _i32, _i32, _i32, _vp])
_lib.register("univst_conv3x3_s2_pad_after_f16", [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, C.POINTER(Epilogue), _vp])
_lib.register("univst_conv_temporal3_f16", [_vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i32, C.POINTER(Epilogue), _vp])
_lib.register("univst_softmax_rows_f16", [_vp, _i32, _i32, _i32, _f32, _vp])
_lib.register("univst_frames_to_u8", [_vp, _i32, _i64, _vp, _vp])
_lib.register("univst_u8_to_frames_f16", [_vp, _i64, _i32, _vp, _vp])
_lib.register("univst_vae_sample_f16", [_vp, _i32, _vp, _i32, _i32, _i32, _f32, _vp, _vp])
_lib.register("univst_layernorm_modulate_f16", [_vp, _i32, _i32, _vp, _vp, _i32, _i32, _f32, _vp, _vp])
_lib.register("univst_gated_add_f16", [_vp, _vp, _vp, _i32, _i32, _i64, _i32, _vp, _vp])
_lib.register("univst_gemm_set_workspace", [_vp, _i64, _vp])
_lib.register("univst_gemm_tune", [_i32])
_lib.register("univst_xrank_ctl_bytes", [], _i64)
_lib.register("univst_xrank_slot_floats", [], _i32)
_lib.register("univst_xrank_barrier", [_vp, _i32, _i32, _vp])


class Push(C.Structure):
    """Mirror of ``univst_push_t``."""
    _fields_ = [("src", _vp), ("ld_src", _i32), ("src_blk_rows", _i64), ("dst", _vp * 16), ("ld_dst", _i32),
                ("dst_blk_rows", _i64), ("nblk", _i32), ("rows", _i32), ("cols", _i32), ("mc_dst", _vp)]


_lib.register("univst_xrank_push_f16", [C.POINTER(Push), _i32, _vp, _i32, _i32, _vp])
_lib.register("univst_groupnorm_xrank_f16", [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _f32, _i32, _vp, _vp, _vp, _i32, _i32, _vp])
_lib.register("univst_exchange_push_xrank_f16", [_i32, _vp, _i32, C.POINTER(_vp), _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp])
_lib.register("univst_attn_shift_dev_f16", [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp])
_lib.register("univst_set_floats", [_vp, C.POINTER(_f32), _i32, _vp])
_lib.register("univst_maskprop_workspace_bytes", [_i32, _i32, _i32], _i64)
_lib.register("univst_maskprop_f32", [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp, _vp, _vp, _i32, _vp, _vp])
_lib.register("univst_flow_warp_key_u8", [_vp, _i32, _i32, _i32, _i32, _i32, C.POINTER(_i32), C.POINTER(_vp), C.POINTER(_vp), _f32, _vp])
_lib.register("univst_mask_select_u8", [_vp, _vp, _vp, _i64, _vp, _vp])

# number of kernels launched through this module (bench.py reports it as ``gpu_launches``)
launch_count = 0
_LAUNCHES = {
    "groupnorm_stats": 2, "groupnorm_apply": 1, "gemm": 1, "conv3x3": 1, "sc_attention": 1, "attn_shift": 3, "groupnorm": 3, "layernorm": 1, "upsample2x": 1,
    "temporal_attention": 1, "cross_attention": 1, "joint_attention": 1, "rmsnorm_heads": 1, "sd3_attn_shift": 4, "space_to_depth2": 1, "pack_latents": 1, "unpack_latents": 1, "timestep_embedding": 1, "mask_resize": 1,
    "latent_blend": 1, "latent_adain": 1, "ddim_step": 1, "axpby": 1, "exchange_push": 1, "halo_push": 1, "maskprop": 3, "flow_warp_key": 1, "mask_select": 1,
    "xrank_barrier": 1, "xrank_push": 1, "groupnorm_xrank": 3, "set_floats": 1,
    "layernorm_modulate": 1, "gated_add": 1,
    "conv3x3_s2_pad_after": 1, "conv_temporal3": 1, "softmax_rows": 1, "frames_to_u8": 1, "u8_to_frames": 1, "vae_sample": 1,
}


def _count(name):
    global launch_count
    launch_count += _LAUNCHES[name]


# Optional per-kernel device timing (bench.py): CUDA events recorded on the launching stream around selected kernels.
_profile = None


def profile_start(names):
    """Start recording (start, end, meta) CUDA-event triples for the kernels in ``names`` (e.g. {"sc_attention"})."""
    global _profile
    _profile = {n: [] for n in names}


def profile_stop():
    """Stop recording; returns {name: [(milliseconds, meta), ...]} (synchronises)."""
    global _profile
    prof, _profile = _profile, None
    torch.cuda.synchronize()
    return {n: [(a.elapsed_time(b), meta) for a, b, meta in ev] for n, ev in (prof or {}).items()}


class _Timed:
    def __init__(self, name, meta):
        self.rec = _profile.get(name) if _profile is not None else None
        self.meta = meta

    def __enter__(self):
        if self.rec is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if self.rec is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            self.rec.append((self.a, b, self.meta))
        return False


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, name: str, dtype=torch.float16):
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.dtype} {t.device} "
                         f"contiguous={t.is_contiguous()}")


def make_epilogue(bias=None, rowvec=None, rows_per_group=1, residual=None, bias2=None, geglu=False, out_scale=1.0,
                  act=False) -> Epilogue:
    ep = Epilogue()
    ep.bias = _ptr(bias)
    ep.rowvec = _ptr(rowvec)
    ep.rows_per_group = rows_per_group
    ep.rowvec_ld = rowvec.stride(0) if rowvec is not None else 0
    ep.act = {"silu": 1, "gelu_tanh": 2}[act] if isinstance(act, str) else int(act)   # True / 1 = SiLU, 2 = tanh GELU
    ep.residual = _ptr(residual)
    ep.ldr = residual.stride(0) if residual is not None else 0
    ep.bias2 = _ptr(bias2)
    ep.geglu = 1 if geglu else 0
    ep.out_scale = out_scale
    return ep


# split-K of the tensor-core GEMM / conv (univst_gemm_tune): off unless a caller turns it on -- the frame-sharded UNet does
# for shards of a few images, whose deep levels would otherwise occupy a handful of SMs
_splitk_max_tiles = 0
_splitk_ws = {}
SPLITK_WORKSPACE_BYTES = 64 << 20


def gemm_splitk(max_tiles: int):
    """Split the K loop of GEMM / conv launches with at most ``max_tiles`` output tiles over the idle SMs (0 = never).
    Deterministic; the fp32 summation order differs from the unsplit kernel's (last-bit differences in fp16 outputs)."""
    global _splitk_max_tiles
    _splitk_max_tiles = int(max_tiles)
    check(_lib.lib().univst_gemm_tune(_splitk_max_tiles), "univst_gemm_tune")


def _splitk_ready(device):
    """Register the split-K workspace of the current stream (once per device and stream)."""
    if not _splitk_max_tiles:
        return
    st = _stream()
    key = (device, st)
    if key not in _splitk_ws:
        ws = torch.zeros(SPLITK_WORKSPACE_BYTES, dtype=torch.uint8, device=device)
        check(_lib.lib().univst_gemm_set_workspace(ws.data_ptr(), ws.numel(), st), "univst_gemm_set_workspace")
        _splitk_ws[key] = ws


def gemm(a: torch.Tensor, w: torch.Tensor, *, a2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
         bias=None, rowvec=None, rows_per_group=1, residual=None, bias2=None, geglu=False, out_scale=1.0, act=False):
    """``out[M, N_out] = epilogue([a | a2] @ w.T)``; ``a``/``a2``/``residual``/``out`` may be row-strided 2-D views."""
    _lib.require_device()
    M, K1 = a.shape
    K = K1 + (a2.shape[1] if a2 is not None else 0)
    N = w.shape[0]
    assert w.shape[1] == K and w.is_contiguous() and a.stride(1) == 1
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((M, n_out), dtype=torch.float16, device=a.device)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    ep = make_epilogue(bias, rowvec, rows_per_group, residual, bias2, geglu, out_scale, act)
    _splitk_ready(a.device)
    with _Timed("gemm", (M, N, K)):
        check(_lib.lib().univst_gemm_f16(a.data_ptr(), a.stride(0), _ptr(a2), a2.stride(0) if a2 is not None else 0, K1,
                                         w.data_ptr(), M, N, K, out.data_ptr(), out.stride(0), C.byref(ep), _stream()),
              "univst_gemm_f16")
    _count("gemm")
    return out


def conv3x3(x: torch.Tensor, w: torch.Tensor, *, x2: Optional[torch.Tensor] = None, stride: int = 1,
            out: Optional[torch.Tensor] = None, bias=None, rowvec=None, rows_per_group=1, residual=None,
            out_scale=1.0):
    """3x3 conv, padding 1.  ``x``: [NB, H, W, C1] (stride 1) or the parity planes [4, NB, H/2, W/2, C1] from
    :func:`space_to_depth2` (stride 2); ``w``: [Cout, 3, 3, C1 + C2] flattened to [Cout, 9 (C1 + C2)]."""
    _lib.require_device()
    _chk(x, "x")
    if stride == 1:
        NB, H, W, C1 = x.shape
    else:
        _, NB, H, W, C1 = x.shape
    C2 = x2.shape[-1] if x2 is not None else 0
    Cout = w.shape[0]
    assert w.is_contiguous() and w.numel() == Cout * 9 * (C1 + C2)
    if out is None:
        out = torch.empty((NB * H * W, Cout), dtype=torch.float16, device=x.device)
    ep = make_epilogue(bias, rowvec, rows_per_group, residual, None, False, out_scale)
    _splitk_ready(x.device)
    with _Timed("conv3x3", (NB * H * W, Cout, 9 * (C1 + C2))):
        check(_lib.lib().univst_conv3x3_f16(x.data_ptr(), _ptr(x2), NB, H, W, C1, C2, w.data_ptr(), Cout, stride,
                                            out.data_ptr(), out.stride(0), C.byref(ep), _stream()), "univst_conv3x3_f16")
    _count("conv3x3")
    return out


def conv3x3_s2_pad_after(planes: torch.Tensor, w: torch.Tensor, *, bias=None, out: Optional[torch.Tensor] = None):
    """3x3 conv, stride 2, zero row / column AFTER the image (diffusers ``Downsample2D(padding=0)``: F.pad (0, 1, 0, 1) then
    a stride-2 conv).  ``planes``: [4, NB, H/2, W/2, C] from :func:`space_to_depth2`; ``w``: [Cout, 9 C] tap-major."""
    _lib.require_device()
    _chk(planes, "planes")
    _, NB, H, W, C1 = planes.shape
    Cout = w.shape[0]
    assert w.is_contiguous() and w.numel() == Cout * 9 * C1
    if out is None:
        out = torch.empty((NB * H * W, Cout), dtype=torch.float16, device=planes.device)
    ep = make_epilogue(bias)
    _splitk_ready(planes.device)
    check(_lib.lib().univst_conv3x3_s2_pad_after_f16(planes.data_ptr(), NB, H, W, C1, w.data_ptr(), Cout, out.data_ptr(),
                                                     out.stride(0), C.byref(ep), _stream()), "univst_conv3x3_s2_pad_after_f16")
    _count("conv3x3_s2_pad_after")
    return out


def conv_temporal3(x: torch.Tensor, w: torch.Tensor, *, NB: int, F: int, HW: int, bias=None, residual=None,
                   out: Optional[torch.Tensor] = None):
    """(3, 1, 1) temporal convolution over the frames of every pixel.  ``x``: [NB * F * HW, C] rows (clip, frame, pixel);
    ``w``: [Cout, 3 C] (tap t = frame offset t - 1); zero frames beyond the ends of a clip."""
    _lib.require_device()
    _chk(x, "x")
    C1 = x.shape[1]
    Cout = w.shape[0]
    assert x.shape[0] == NB * F * HW and w.is_contiguous() and w.numel() == Cout * 3 * C1
    if out is None:
        out = torch.empty((NB * F * HW, Cout), dtype=torch.float16, device=x.device)
    ep = make_epilogue(bias, residual=residual)
    _splitk_ready(x.device)
    check(_lib.lib().univst_conv_temporal3_f16(x.data_ptr(), NB, F, HW, C1, w.data_ptr(), Cout, out.data_ptr(), out.stride(0),
                                               C.byref(ep), _stream()), "univst_conv_temporal3_f16")
    _count("conv_temporal3")
    return out


def softmax_rows_(x: torch.Tensor, scale: float = 1.0):
    """In-place softmax(scale * x) over the last dimension of a 2-D fp16 tensor (row stride arbitrary)."""
    _lib.require_device()
    assert x.dtype == torch.float16 and x.is_cuda and x.dim() == 2 and x.stride(1) == 1
    check(_lib.lib().univst_softmax_rows_f16(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], float(scale), _stream()),
          "univst_softmax_rows_f16")
    _count("softmax_rows")
    return x


def frames_to_u8(x: torch.Tensor, pixels: int):
    """Decoder output rows [pixels, ld >= 3] fp16 -> uint8 [pixels, 3] (stable_diffusion.py:812-814)."""
    _lib.require_device()
    assert x.dtype == torch.float16 and x.is_cuda and x.stride(1) == 1 and x.shape[0] >= pixels
    out = torch.empty((pixels, 3), dtype=torch.uint8, device=x.device)
    check(_lib.lib().univst_frames_to_u8(x.data_ptr(), x.stride(0), pixels, out.data_ptr(), _stream()), "univst_frames_to_u8")
    _count("frames_to_u8")
    return out


def u8_to_frames(img: torch.Tensor, cpad: int = 64):
    """uint8 [pixels, 3] -> fp16 [pixels, cpad] = image / 127.5 - 1 in channels 0..2, zeros beyond (:826-827)."""
    _lib.require_device()
    _chk(img, "img", torch.uint8)
    pixels = img.numel() // 3
    out = torch.empty((pixels, cpad), dtype=torch.float16, device=img.device)
    check(_lib.lib().univst_u8_to_frames_f16(img.data_ptr(), pixels, cpad, out.data_ptr(), _stream()), "univst_u8_to_frames_f16")
    _count("u8_to_frames")
    return out


def vae_sample(moments: torch.Tensor, noise: Optional[torch.Tensor], C_: int, F: int, HW: int, scaling: float):
    """[mean | logvar] rows [(f) hw, ld] -> (1, C, F, hw) latents: (mean + std * noise) * scaling (``noise`` (F, C, hw) fp16;
    None = the posterior mode)."""
    _lib.require_device()
    assert moments.dtype == torch.float16 and moments.is_cuda and moments.stride(1) == 1 and moments.shape[0] == F * HW
    if noise is not None:
        _chk(noise, "noise")
        assert noise.numel() == F * C_ * HW
    out = torch.empty((1, C_, F, HW), dtype=torch.float16, device=moments.device)
    check(_lib.lib().univst_vae_sample_f16(moments.data_ptr(), moments.stride(0), _ptr(noise), C_, F, HW, float(scaling),
                                           out.data_ptr(), _stream()), "univst_vae_sample_f16")
    _count("vae_sample")
    return out


def sc_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_src: torch.Tensor, *, NI: int, NIkv: int, H: int,
                 d: int, N: int, Nkv: int, out: Optional[torch.Tensor] = None):
    """``q``: [NI*N, >= H*d] (row-strided view), ``k``/``v``: [NIkv*Nkv, .] views sharing one row stride;
    ``kv_src``: int32 CUDA [NI, nsrc] source-image table."""
    _lib.require_device()
    assert q.stride(1) == 1 and k.stride(1) == 1 and v.stride(1) == 1 and k.stride(0) == v.stride(0)
    assert kv_src.dtype == torch.int32 and kv_src.is_cuda and kv_src.is_contiguous() and kv_src.shape[0] == NI
    if out is None:
        out = torch.empty((NI * N, H * d), dtype=torch.float16, device=q.device)
    with _Timed("sc_attention", (NI, H, d, N, Nkv * kv_src.shape[1])):
        check(_lib.lib().univst_sc_attention_f16(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), NI,
                                                 NIkv, H, d, N, Nkv, kv_src.data_ptr(), kv_src.shape[1], out.data_ptr(),
                                                 out.stride(0), _stream()), "univst_sc_attention_f16")
    _count("sc_attention")
    return out


def cross_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_src: torch.Tensor, *, NI: int, NIkv: int, H: int,
                    d: int, N: int, Nkv: int, out: Optional[torch.Tensor] = None):
    """attn2: every query image attends to the short context of its branch (``kv_src``: int32 [NI] or [NI, 1])."""
    _lib.require_device()
    if not _lib.lib().univst_cross_attention_supported(d, Nkv):   # long contexts / odd head dims: the general kernel
        return sc_attention(q, k, v, kv_src.view(NI, 1), NI=NI, NIkv=NIkv, H=H, d=d, N=N, Nkv=Nkv, out=out)
    assert q.stride(1) == 1 and k.stride(1) == 1 and v.stride(1) == 1 and k.stride(0) == v.stride(0)
    assert kv_src.dtype == torch.int32 and kv_src.is_cuda and kv_src.is_contiguous() and kv_src.numel() == NI
    if out is None:
        out = torch.empty((NI * N, H * d), dtype=torch.float16, device=q.device)
    with _Timed("cross_attention", (NI, H, d, N, Nkv)):
        check(_lib.lib().univst_cross_attention_f16(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), NI,
                                                    NIkv, H, d, N, Nkv, kv_src.data_ptr(), out.data_ptr(), out.stride(0),
                                                    _stream()), "univst_cross_attention_f16")
    _count("cross_attention")
    return out


def joint_attention(q, k, v, k2, v2, kv_src, *, NI: int, NIkv: int, NIkv2: int, H: int, d: int, N: int, Nkv: int, Nkv2: int,
                    out: Optional[torch.Tensor] = None):
    """SD3 joint attention: sources < NIkv are images of (k, v) with Nkv tokens, sources >= NIkv images of (k2, v2) with
    Nkv2 tokens (the text tokens).  ``kv_src``: int32 CUDA [NI, nsrc]."""
    _lib.require_device()
    assert all(t.stride(1) == 1 for t in (q, k, v, k2, v2)) and k.stride(0) == v.stride(0) and k2.stride(0) == v2.stride(0)
    assert kv_src.dtype == torch.int32 and kv_src.is_cuda and kv_src.is_contiguous() and kv_src.shape[0] == NI
    if out is None:
        out = torch.empty((NI * N, H * d), dtype=torch.float16, device=q.device)
    with _Timed("sc_attention", (NI, H, d, N, Nkv * (kv_src.shape[1] - 1) + Nkv2)):
        check(_lib.lib().univst_joint_attention_f16(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), NI, NIkv,
                                                    H, d, N, Nkv, k2.data_ptr(), v2.data_ptr(), k2.stride(0), NIkv2, Nkv2,
                                                    kv_src.data_ptr(), kv_src.shape[1], out.data_ptr(), out.stride(0),
                                                    _stream()), "univst_joint_attention_f16")
    _count("joint_attention")
    return out


def rmsnorm_heads_(qkv: torch.Tensor, H: int, d: int, wq, wk, eps: float = 1e-6):
    """In-place per-head RMS norm of the Q (weight ``wq``) and K (``wk``) blocks of a fused [rows, 3 H d] buffer."""
    _lib.require_device()
    _chk(qkv, "qkv")
    check(_lib.lib().univst_rmsnorm_heads_f16(qkv.data_ptr(), qkv.stride(0), qkv.shape[0], H, d, _ptr(wq), _ptr(wk), eps,
                                              _stream()), "univst_rmsnorm_heads_f16")
    _count("rmsnorm_heads")
    return qkv


def sd3_attn_shift_(qkv: torch.Tensor, F: int, N: int, H: int, d: int, alpha: float, beta: float, gamma: float):
    """In-place AdaIN-guided shift of the edit branch of the fused [3 F N, 3 H d] buffer, SD3 semantics."""
    _lib.require_device()
    _chk(qkv, "qkv")
    assert qkv.shape[0] == 3 * F * N
    ws = _workspace(_lib.lib().univst_sd3_shift_workspace_bytes(F, H * d, d), qkv.device)
    check(_lib.lib().univst_sd3_attn_shift_f16(qkv.data_ptr(), qkv.stride(0), F, N, H, d, alpha, beta, gamma, ws.data_ptr(),
                                               _stream()), "univst_sd3_attn_shift_f16")
    _count("sd3_attn_shift")
    return qkv


def sc_attention_sharded(qkv: torch.Tensor, qkv_prev: Optional[torch.Tensor], qkv_first: Optional[torch.Tensor],
                         kv_src: torch.Tensor, *, B: int, Fl: int, H: int, d: int, N: int, out: Optional[torch.Tensor] = None):
    """Frame-sharded attn1: ``qkv`` [B*Fl*N, 3C] is this rank's fused projection, ``qkv_prev`` / ``qkv_first`` the
    previous rank's / rank 0's (peer-mapped views of the same shape, or None where the table never names them)."""
    _lib.require_device()
    C_ = H * d
    NI = B * Fl
    assert qkv.stride(1) == 1 and qkv.shape[0] == NI * N and kv_src.dtype == torch.int32 and kv_src.shape[0] == NI
    if out is None:
        out = torch.empty((NI * N, C_), dtype=torch.float16, device=qkv.device)
    es = qkv.element_size()
    kp = lambda t: t.data_ptr() + C_ * es if t is not None else None
    vp = lambda t: t.data_ptr() + 2 * C_ * es if t is not None else None
    for t in (qkv_prev, qkv_first):
        assert t is None or (t.stride(0) == qkv.stride(0) and t.shape == qkv.shape)
    with _Timed("sc_attention", (NI, H, d, N, N * kv_src.shape[1])):
        check(_lib.lib().univst_sc_attention_sharded_f16(qkv.data_ptr(), qkv.stride(0), kp(qkv), vp(qkv), qkv.stride(0), NI, H,
                                                         d, N, kv_src.data_ptr(), kv_src.shape[1], out.data_ptr(),
                                                         out.stride(0), kp(qkv_prev), vp(qkv_prev), kp(qkv_first),
                                                         vp(qkv_first), B, Fl, _stream()),
              "univst_sc_attention_sharded_f16")
    _count("sc_attention")
    return out


def temporal_attention(qkv: torch.Tensor, *, B: int, F: int, N: int, H: int, d: int, out: Optional[torch.Tensor] = None):
    """Motion-module attention over the frames of every pixel.  ``qkv``: [B*F*N, 3*H*d] (rows: branch, frame, pixel)."""
    _lib.require_device()
    assert qkv.stride(1) == 1 and qkv.shape[0] == B * F * N and qkv.shape[1] >= 3 * H * d
    if out is None:
        out = torch.empty((B * F * N, H * d), dtype=torch.float16, device=qkv.device)
    with _Timed("temporal_attention", (B, F, N, H, d)):
        check(_lib.lib().univst_temporal_attention_f16(qkv.data_ptr(), qkv.stride(0), B, F, N, H, d, out.data_ptr(),
                                                       out.stride(0), _stream()), "univst_temporal_attention_f16")
    _count("temporal_attention")
    return out


def attention_tune(variant: int = -1, dedupe: int = -1, stagger: int = -1):
    """Select the tile / exp2 variant of the fused attention kernel (tuning and A/B timing; -1 = default)."""
    check(_lib.lib().univst_attention_tune(variant, dedupe, stagger), "univst_attention_tune")


_workspaces = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device, torch.cuda.current_stream().cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def attn_shift_(qkv: torch.Tensor, F: int, N: int, C_: int, alpha: float, beta: float, gamma: float):
    """In-place AdaIN-guided shift of the edit branch of the fused [3 F N, 3 C] projection buffer."""
    _lib.require_device()
    _chk(qkv, "qkv")
    assert qkv.shape[0] == 3 * F * N
    ws = _workspace(_lib.lib().univst_attn_shift_workspace_bytes(F, C_), qkv.device)
    check(_lib.lib().univst_attn_shift_f16(qkv.data_ptr(), qkv.stride(0), F, N, C_, alpha, beta, gamma, ws.data_ptr(),
                                           _stream()), "univst_attn_shift_f16")
    _count("attn_shift")
    return qkv


def attn_shift_dev_(qkv: torch.Tensor, F: int, N: int, C_: int, abg: torch.Tensor):
    """:func:`attn_shift_` with (alpha, beta, gamma) read from a float32 CUDA tensor (>= 3 elements) at run time."""
    _lib.require_device()
    _chk(qkv, "qkv")
    assert qkv.shape[0] == 3 * F * N and abg.dtype == torch.float32 and abg.is_cuda and abg.numel() >= 3
    ws = _workspace(_lib.lib().univst_attn_shift_workspace_bytes(F, C_), qkv.device)
    check(_lib.lib().univst_attn_shift_dev_f16(qkv.data_ptr(), qkv.stride(0), F, N, C_, abg.data_ptr(), ws.data_ptr(),
                                               _stream()), "univst_attn_shift_dev_f16")
    _count("attn_shift")
    return qkv


def set_floats(dst: torch.Tensor, values):
    """Write up to 64 Python floats into the float32 CUDA tensor ``dst`` (values travel as launch arguments: no host
    buffer has to stay alive, nothing synchronises)."""
    _lib.require_device()
    n = len(values)
    assert dst.dtype == torch.float32 and dst.is_cuda and dst.is_contiguous() and dst.numel() >= n
    arr = (_f32 * n)(*[float(v) for v in values])
    check(_lib.lib().univst_set_floats(dst.data_ptr(), arr, n, _stream()), "univst_set_floats")
    _count("set_floats")
    return dst


def xrank_barrier(xr):
    """Cross-rank synchronisation on the current stream (``xr``: :class:`univst_b200.xrank.XRank`)."""
    _lib.require_device()
    check(_lib.lib().univst_xrank_barrier(xr.ctl, xr.rank, xr.world, _stream()), "univst_xrank_barrier")
    _count("xrank_barrier")


def xrank_push(xr, pushes):
    """Up to two block copies into peer memory + the cross-rank synchronisation as the tail of the same kernel.
    ``pushes``: list of dicts(src=strided 2-D fp16 view, src_blk_rows, dst=[device pointer or 0 per rank], ld_dst,
    dst_blk_rows, nblk, rows[, mc=multicast address: one switch-replicated store instead of one per peer]); an empty
    list is a plain synchronisation."""
    _lib.require_device()
    n = len(pushes)
    arr = (Push * max(n, 1))()
    for k, p in enumerate(pushes):
        src = p["src"]
        assert src.dtype == torch.float16 and src.is_cuda and src.stride(1) == 1
        arr[k].src, arr[k].ld_src, arr[k].src_blk_rows = src.data_ptr(), src.stride(0), p["src_blk_rows"]
        for r in range(16):
            arr[k].dst[r] = (p["dst"][r] or None) if r < len(p["dst"]) else None
        arr[k].ld_dst, arr[k].dst_blk_rows = p["ld_dst"], p["dst_blk_rows"]
        arr[k].nblk, arr[k].rows, arr[k].cols = p["nblk"], p["rows"], src.shape[1]
        arr[k].mc_dst = p.get("mc") or None
    check(_lib.lib().univst_xrank_push_f16(arr, n, xr.ctl, xr.rank, xr.world, _stream()), "univst_xrank_push_f16")
    _count("xrank_push")


def groupnorm_xrank(x1, gamma, beta, *, NB, rows, xr, groups=32, eps=1e-5, silu=False, x2=None):
    """GroupNorm whose statistics span the rows of all ranks of ``xr``: partial sums exchanged through the ranks' control
    blocks inside the fold kernel (no collective call); every rank normalises with bit-identical statistics."""
    _lib.require_device()
    _chk(x1, "x1")
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    out = torch.empty((NB * rows, C1 + C2), dtype=torch.float16, device=x1.device)
    ws = _workspace(_lib.lib().univst_groupnorm_workspace_bytes(NB, groups), x1.device)
    check(_lib.lib().univst_groupnorm_xrank_f16(x1.data_ptr(), _ptr(x2), C1, C2, NB, rows, groups, gamma.data_ptr(),
                                                beta.data_ptr(), eps, 1 if silu else 0, out.data_ptr(), ws.data_ptr(),
                                                xr.ctl, xr.rank, xr.world, _stream()), "univst_groupnorm_xrank_f16")
    global launch_count
    launch_count += _gn_launches(NB, rows, C1 + C2)
    return out


def _gn_launches(NB: int, rows: int, C_: int) -> int:
    """Kernels one GroupNorm launches (mirrors gn_plan in csrc/norm.cu): statistics spans of at most 8 MB are folded by the
    statistics kernel's last block (statistics + apply); larger ones take the separate fold kernel."""
    return 2 if rows * C_ * 2 <= (8 << 20) and NB < 255 else 3


def groupnorm(x1: torch.Tensor, gamma, beta, *, NB: int, rows: int, groups: int = 32, eps: float = 1e-5,
              silu: bool = False, x2: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    _lib.require_device()
    _chk(x1, "x1")
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    if out is None:
        out = torch.empty((NB * rows, C1 + C2), dtype=torch.float16, device=x1.device)
    ws = _workspace(_lib.lib().univst_groupnorm_workspace_bytes(NB, groups), x1.device)
    check(_lib.lib().univst_groupnorm_f16(x1.data_ptr(), _ptr(x2), C1, C2, NB, rows, groups, gamma.data_ptr(),
                                          beta.data_ptr(), eps, 1 if silu else 0, out.data_ptr(), ws.data_ptr(),
                                          _stream()), "univst_groupnorm_f16")
    global launch_count
    launch_count += _gn_launches(NB, rows, C1 + C2)
    return out


def groupnorm_sharded(x1, gamma, beta, *, NB, rows, group, world, groups=32, eps=1e-5, silu=False, x2=None):
    """GroupNorm whose statistics span the rows of ALL ranks of ``group`` (frame-sharded UNet): local sums ->
    all-reduce of NB x groups x 2 floats -> apply with the global row count."""
    import torch.distributed as dist
    _lib.require_device()
    _chk(x1, "x1")
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    sums = torch.empty((NB, groups, 2), dtype=torch.float32, device=x1.device)
    ws = _workspace(_lib.lib().univst_groupnorm_workspace_bytes(NB, groups), x1.device)
    check(_lib.lib().univst_groupnorm_stats_f16(x1.data_ptr(), _ptr(x2), C1, C2, NB, rows, groups, sums.data_ptr(),
                                                ws.data_ptr(), _stream()), "univst_groupnorm_stats_f16")
    _count("groupnorm_stats")
    dist.all_reduce(sums, group=group)
    out = torch.empty((NB * rows, C1 + C2), dtype=torch.float16, device=x1.device)
    check(_lib.lib().univst_groupnorm_apply_f16(x1.data_ptr(), _ptr(x2), C1, C2, NB, rows, groups, sums.data_ptr(),
                                                rows * world, gamma.data_ptr(), beta.data_ptr(), eps, 1 if silu else 0,
                                                out.data_ptr(), _stream()), "univst_groupnorm_apply_f16")
    _count("groupnorm_apply")
    return out


def layernorm(x: torch.Tensor, gamma, beta, eps: float = 1e-5, out: Optional[torch.Tensor] = None):
    _lib.require_device()
    _chk(x, "x")
    rows, C_ = x.shape
    if out is None:
        out = torch.empty_like(x)
    check(_lib.lib().univst_layernorm_f16(x.data_ptr(), rows, C_, gamma.data_ptr(), beta.data_ptr(), eps, out.data_ptr(),
                                          _stream()), "univst_layernorm_f16")
    _count("layernorm")
    return out


def layernorm_modulate(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, rows_per_sample: int, eps: float = 1e-6,
                       out: Optional[torch.Tensor] = None):
    """adaLN modulation: ``LayerNorm(x, no affine) * (1 + scale[s]) + shift[s]``; ``scale`` / ``shift``: [samples, C] views
    (column slices of the block's modulation GEMM: one row stride, last dim contiguous); sample of a row = row // rows_per_sample."""
    _lib.require_device()
    _chk(x, "x")
    rows, C_ = x.shape
    assert scale.shape == shift.shape and scale.shape[1] == C_ and scale.stride(0) == shift.stride(0)
    assert scale.stride(1) == 1 and shift.stride(1) == 1 and rows == scale.shape[0] * rows_per_sample
    if out is None:
        out = torch.empty_like(x)
    check(_lib.lib().univst_layernorm_modulate_f16(x.data_ptr(), rows, C_, scale.data_ptr(), shift.data_ptr(), scale.stride(0),
                                                   rows_per_sample, eps, out.data_ptr(), _stream()),
          "univst_layernorm_modulate_f16")
    _count("layernorm_modulate")
    return out


def gated_add(x: torch.Tensor, y: torch.Tensor, gate: torch.Tensor, rows_per_sample: int, out: Optional[torch.Tensor] = None):
    """``x + gate[s] * y`` with a per-sample gate row ([samples, C] view, last dim contiguous)."""
    _lib.require_device()
    _chk(x, "x"), _chk(y, "y")
    rows, C_ = x.shape
    assert y.shape == x.shape and gate.shape[1] == C_ and gate.stride(1) == 1 and rows == gate.shape[0] * rows_per_sample
    if out is None:
        out = torch.empty_like(x)
    check(_lib.lib().univst_gated_add_f16(x.data_ptr(), y.data_ptr(), gate.data_ptr(), gate.stride(0), rows_per_sample, rows, C_,
                                          out.data_ptr(), _stream()), "univst_gated_add_f16")
    _count("gated_add")
    return out


def upsample2x(x: torch.Tensor):
    _lib.require_device()
    _chk(x, "x")
    NB, H, W, C_ = x.shape
    out = torch.empty((NB, 2 * H, 2 * W, C_), dtype=torch.float16, device=x.device)
    check(_lib.lib().univst_upsample2x_f16(x.data_ptr(), NB, H, W, C_, out.data_ptr(), _stream()), "univst_upsample2x_f16")
    _count("upsample2x")
    return out


def space_to_depth2(x: torch.Tensor):
    _lib.require_device()
    _chk(x, "x")
    NB, H, W, C_ = x.shape
    out = torch.empty((4, NB, H // 2, W // 2, C_), dtype=torch.float16, device=x.device)
    check(_lib.lib().univst_space_to_depth2_f16(x.data_ptr(), NB, H // 2, W // 2, C_, out.data_ptr(), _stream()),
          "univst_space_to_depth2_f16")
    _count("space_to_depth2")
    return out


def pack_latents(zs, Cpad: int = 64):
    """zs: list of B latents (C, F, h, w) (or (1, C, F, h, w)) -> [(b f), h, w, Cpad] zero-padded channels."""
    _lib.require_device()
    B = len(zs)
    z0 = zs[0]
    C_, F, h, w = z0.shape[-4:]
    for z in zs:
        _chk(z, "latent")
    arr = (_vp * B)(*[z.data_ptr() for z in zs])
    out = torch.empty((B * F, h, w, Cpad), dtype=torch.float16, device=z0.device)
    check(_lib.lib().univst_pack_latents_f16(arr, B, C_, F, h * w, Cpad, out.data_ptr(), _stream()), "univst_pack_latents_f16")
    _count("pack_latents")
    return out


def unpack_latents(x: torch.Tensor, B: int, C_: int, F: int, h: int, w: int):
    """x: [(b f) h w, ld] channels-last rows -> (B, C, F, h, w)."""
    _lib.require_device()
    out = torch.empty((B, C_, F, h, w), dtype=torch.float16, device=x.device)
    check(_lib.lib().univst_unpack_latents_f16(x.data_ptr(), x.stride(0), B, C_, F, h * w, out.data_ptr(), _stream()),
          "univst_unpack_latents_f16")
    _count("unpack_latents")
    return out


def timestep_embedding(t: torch.Tensor, dim: int):
    _lib.require_device()
    _chk(t, "t", torch.float32)
    out = torch.empty((t.numel(), dim), dtype=torch.float16, device=t.device)
    check(_lib.lib().univst_timestep_embedding_f16(t.data_ptr(), t.numel(), dim, out.data_ptr(), _stream()),
          "univst_timestep_embedding_f16")
    _count("timestep_embedding")
    return out


def mask_resize(mask_u8: torch.Tensor, h: int, w: int):
    """mask_u8: [F, Hin, Win] uint8 (non-zero = inside) -> fp16 [F, h, w] bilinear, align_corners=False."""
    _lib.require_device()
    _chk(mask_u8, "mask", torch.uint8)
    F, Hin, Win = mask_u8.shape
    out = torch.empty((F, h, w), dtype=torch.float16, device=mask_u8.device)
    check(_lib.lib().univst_mask_resize_u8(mask_u8.data_ptr(), F, Hin, Win, h, w, out.data_ptr(), _stream()),
          "univst_mask_resize_u8")
    _count("mask_resize")
    return out


def latent_blend(a, b, mask, out=None):
    """(1 - m) * a + m * b on (.., C, F, h, w) latents, m: [F, h, w]."""
    _lib.require_device()
    _chk(a, "a"), _chk(b, "b"), _chk(mask, "mask")
    C_, F, h, w = a.shape[-4:]
    if out is None:
        out = torch.empty_like(a)
    check(_lib.lib().univst_latent_blend_f16(a.data_ptr(), b.data_ptr(), mask.data_ptr(), C_, F, h * w, out.data_ptr(),
                                             _stream()), "univst_latent_blend_f16")
    _count("latent_blend")
    return out


def latent_adain(cnt, sty, out=None):
    _lib.require_device()
    _chk(cnt, "cnt"), _chk(sty, "sty")
    C_, F, h, w = cnt.shape[-4:]
    if out is None:
        out = torch.empty_like(cnt)
    check(_lib.lib().univst_latent_adain_f16(cnt.data_ptr(), sty.data_ptr(), C_, F, h * w, out.data_ptr(), _stream()),
          "univst_latent_adain_f16")
    _count("latent_adain")
    return out


def latent_blend_fc(a, b, mask, out=None):
    """(1 - m) * a + m * b on frame-major (F, C, h, w) latents (the SD3 loop), m: [F, h, w]."""
    _lib.require_device()
    _chk(a, "a"), _chk(b, "b"), _chk(mask, "mask")
    F, C_, h, w = a.shape
    assert b.shape == a.shape and tuple(mask.shape) == (F, h, w)
    if out is None:
        out = torch.empty_like(a)
    check(_lib.lib().univst_latent_blend_fc_f16(a.data_ptr(), b.data_ptr(), mask.data_ptr(), F, C_, h * w, out.data_ptr(),
                                                _stream()), "univst_latent_blend_fc_f16")
    _count("latent_blend")
    return out


def plane_adain(cnt, sty, out=None):
    """SD3 ``latent_adain`` (video_diffusion_sd3/pnp_utils.py:304-316) on (F, C, h, w): instance norm of every content
    plane re-scaled by the unbiased std / mean of the same style plane -- the latent-AdaIN kernel with F * C one-frame
    channels."""
    _lib.require_device()
    _chk(cnt, "cnt"), _chk(sty, "sty")
    F, C_, h, w = cnt.shape
    assert sty.shape == cnt.shape
    if out is None:
        out = torch.empty_like(cnt)
    check(_lib.lib().univst_latent_adain_f16(cnt.data_ptr(), sty.data_ptr(), F * C_, 1, h * w, out.data_ptr(), _stream()),
          "univst_latent_adain_f16")
    _count("latent_adain")
    return out


def ddim_step(z, eps_rows, branch: int, alpha_t: float, alpha_prev: float, out=None, x0_out=None):
    """z: (.., C, F, h, w); eps_rows: channels-last conv_out rows [(b f) h w, ld]."""
    _lib.require_device()
    _chk(z, "z")
    C_, F, h, w = z.shape[-4:]
    if out is None:
        out = torch.empty_like(z)
    check(_lib.lib().univst_ddim_step_f16(z.data_ptr(), eps_rows.data_ptr(), eps_rows.stride(0), branch, C_, F, h * w,
                                          float(alpha_t), float(alpha_prev), out.data_ptr(), _ptr(x0_out), _stream()),
          "univst_ddim_step_f16")
    _count("ddim_step")
    return out


def axpby(a, b, wa: float, wb: float, out=None):
    _lib.require_device()
    _chk(a, "a"), _chk(b, "b")
    if out is None:
        out = torch.empty_like(a)
    check(_lib.lib().univst_axpby_f16(a.data_ptr(), b.data_ptr(), wa, wb, a.numel(), out.data_ptr(), _stream()),
          "univst_axpby_f16")
    _count("axpby")
    return out


def halo_push(src: torch.Tensor, src_blk_rows: int, dst_ptrs, ld_dst: int, dst_blk_rows: int, nblk: int, rows: int):
    """Store ``nblk`` blocks [rows, src.shape[1]] of the strided view ``src`` (block b starts ``b * src_blk_rows`` rows in)
    into every non-zero destination pointer (block b at ``b * dst_blk_rows`` rows of stride ``ld_dst``)."""
    _lib.require_device()
    assert src.dtype == torch.float16 and src.is_cuda and src.stride(1) == 1
    P = len(dst_ptrs)
    arr = (_vp * P)(*[p or None for p in dst_ptrs])
    check(_lib.lib().univst_halo_push_f16(src.data_ptr(), src.stride(0), src_blk_rows, arr, P, ld_dst, dst_blk_rows, nblk, rows,
                                          src.shape[1], _stream()), "univst_halo_push_f16")
    _count("halo_push")


def exchange_push(direction: int, src: torch.Tensor, dst_ptrs, rank: int, P: int, B: int, Fl: int, N: int, xr=None):
    """Frames <-> pixels exchange of the frame-sharded motion modules: store the local rows [B * Fl * N, C] at their place
    in every owner's buffer (``dst_ptrs``: device pointers of the P ranks' symmetric-memory buffers, each [rows, C]).
    direction 0: frames -> pixels, 1: pixels -> frames.  With ``xr`` the cross-rank synchronisation is the tail of the
    kernel; without it the caller issues a barrier."""
    _lib.require_device()
    assert src.dtype == torch.float16 and src.is_cuda and src.stride(1) == 1 and src.shape[0] == B * Fl * N and len(dst_ptrs) == P
    arr = (_vp * P)(*dst_ptrs)
    if xr is None:
        check(_lib.lib().univst_exchange_push_f16(direction, src.data_ptr(), src.stride(0), arr, rank, P, B, Fl, N,
                                                  src.shape[1], _stream()), "univst_exchange_push_f16")
    else:
        check(_lib.lib().univst_exchange_push_xrank_f16(direction, src.data_ptr(), src.stride(0), arr, xr.ctl, rank, P, B, Fl,
                                                        N, src.shape[1], _stream()), "univst_exchange_push_xrank_f16")
    _count("exchange_push")


def maskprop(feat_tar, feat_src, segs, temperature: float = 0.2, topk: int = 15, return_kept: int = 0):
    """feat_tar [N, C], feat_src [C, M], segs [Ccls, M] fp32 CUDA -> segs_tar [Ccls, N] (mask_propagation.py:75-83)."""
    _lib.require_device()
    for t, n in ((feat_tar, "feat_tar"), (feat_src, "feat_src"), (segs, "segs")):
        _chk(t, n, torch.float32)
    N, C_ = feat_tar.shape
    M = feat_src.shape[1]
    Ccls = segs.shape[0]
    assert feat_src.shape[0] == C_ and segs.shape[1] == M
    out = torch.empty((Ccls, N), dtype=torch.float32, device=feat_tar.device)
    thr = torch.empty((N,), dtype=torch.float32, device=feat_tar.device) if return_kept else None
    kept = torch.empty((N, return_kept), dtype=torch.int32, device=feat_tar.device) if return_kept else None
    ws = _workspace(_lib.lib().univst_maskprop_workspace_bytes(N, C_, M), feat_tar.device)
    check(_lib.lib().univst_maskprop_f32(feat_tar.data_ptr(), feat_src.data_ptr(), segs.data_ptr(), N, C_, M, Ccls,
                                         temperature, topk, out.data_ptr(), _ptr(thr), _ptr(kept), return_kept,
                                         ws.data_ptr(), _stream()),
          "univst_maskprop_f32")
    _count("maskprop")
    return (out, thr, kept) if return_kept else out


def flow_warp_key_(frames, key: int, neighbours, fwd_flows, bwd_flows, threshold: float = 1.5):
    """In-place smoothing of frames[key] ([F, H, W, 3] uint8 CUDA) from its neighbours (stable_diffusion.py:731-747)."""
    _lib.require_device()
    _chk(frames, "frames", torch.uint8)
    F, H, W, ch = frames.shape
    assert ch == 3 and len(neighbours) == len(fwd_flows) == len(bwd_flows) <= 4
    for f in list(fwd_flows) + list(bwd_flows):
        _chk(f, "flow", torch.float32)
        assert tuple(f.shape) == (H, W, 2)
    n = len(neighbours)
    idx = (_i32 * max(n, 1))(*neighbours)
    fw = (_vp * max(n, 1))(*[f.data_ptr() for f in fwd_flows])
    bw = (_vp * max(n, 1))(*[f.data_ptr() for f in bwd_flows])
    check(_lib.lib().univst_flow_warp_key_u8(frames.data_ptr(), F, H, W, key, n, idx, fw, bw, threshold, _stream()),
          "univst_flow_warp_key_u8")
    _count("flow_warp_key")
    return frames


def mask_select(keep_mask, orig, est):
    """keep_mask [.., H, W] uint8 (non-zero = keep ``orig``), orig / est [.., H, W, 3] uint8."""
    _lib.require_device()
    _chk(keep_mask, "keep_mask", torch.uint8), _chk(orig, "orig", torch.uint8), _chk(est, "est", torch.uint8)
    out = torch.empty_like(est)
    check(_lib.lib().univst_mask_select_u8(keep_mask.data_ptr(), orig.data_ptr(), est.data_ptr(), keep_mask.numel(),
                                           out.data_ptr(), _stream()), "univst_mask_select_u8")
    _count("mask_select")
    return out
