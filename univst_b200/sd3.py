"""Mirror of the reference's SD3 / SD3.5 joint-attention processors (backbones/video_diffusion_sd3/pnp_utils.py):
``CrossFrameProcessor`` (:9-132, the inversions) and ``AttentionShiftProcessor`` (:135-271, the three-branch transfer),
with the diffusers ``AttnProcessor.__call__(attn, hidden_states, encoder_hidden_states, attention_mask, idx=...)``
protocol -- they are installed with ``transformer.set_attn_processor`` exactly like the reference's
(run_video_style_transfer_sd3.py:55-63, pnp_utils.py:276-284) and read the projections / RMS-norm weights off the
``attn`` module they are called with (packed once per module).

The arithmetic runs in the sm_100a kernels: fused QKV / added-QKV GEMMs with bias, in-place per-head RMS norm, the
AdaIN-guided shift in the (frame, head) layout, and the fused attention kernel with the image K/V gathered from the
[first, previous, self] frames plus the text tokens as a second K/V tensor (no concatenated K/V is ever built); image and
text queries are two launches over the same sources.

Scope: the processors only.  The MMDiT around them (``SD3Transformer2DModel``) is third-party
diffusers code that is not available in this image.  ``thresh2`` (read but never set at pnp_utils.py:185, which makes
the reference raise inside the shift window) is taken to be ``eta2``.
"""
from __future__ import annotations

import torch

from . import ops

CLIP_LENGTH = 16  # hard-coded in the reference processors (pnp_utils.py:25, :153)


class _Packed:
    def __init__(self, attn, device):
        h = lambda t: t.detach().to(device=device, dtype=torch.float16).contiguous()
        cat = lambda names, attr: h(torch.cat([getattr(getattr(attn, n), attr) for n in names], 0))
        self.w_qkv, self.b_qkv = cat(("to_q", "to_k", "to_v"), "weight"), cat(("to_q", "to_k", "to_v"), "bias")
        self.w_add = self.b_add = self.w_add_out = self.b_add_out = None
        if getattr(attn, "add_q_proj", None) is not None:   # absent on the self-attention attn2 of SD3.5's dual blocks
            self.w_add, self.b_add = cat(("add_q_proj", "add_k_proj", "add_v_proj"), "weight"), cat(("add_q_proj", "add_k_proj", "add_v_proj"), "bias")
        self.w_out, self.b_out = h(attn.to_out[0].weight), h(attn.to_out[0].bias)
        if getattr(attn, "to_add_out", None) is not None:   # absent on a context_pre_only attention (the last block)
            self.w_add_out, self.b_add_out = h(attn.to_add_out.weight), h(attn.to_add_out.bias)
        nw = lambda n: h(getattr(attn, n).weight) if getattr(attn, n, None) is not None else None
        self.norm_q, self.norm_k, self.norm_added_q, self.norm_added_k = (nw(n) for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"))
        self.eps = float(getattr(getattr(attn, "norm_q", None), "eps", 1e-6) or 1e-6)
        self.heads = attn.heads


_tables = {}


def _source_table_sharded(B: int, Fl: int, rank: int, device, text: bool = True) -> torch.Tensor:
    """[first, previous, self] (+ own text tokens) of a frame shard: local images 0 .. B Fl - 1, then the halo banks of the
    projection buffer -- B Fl + b = last frame of the previous rank, B Fl + B + b = frame 0 of the clip (rank 0 has both
    locally).  Text K/V live in the second tensor: source (B Fl + 2 B) + image."""
    key = ("shard", B, Fl, rank, str(device), text)
    if key not in _tables:
        NI = B * Fl
        rows = []
        for b in range(B):
            for fl in range(Fl):
                me = b * Fl + fl
                first = b * Fl if rank == 0 else NI + B + b
                prev = me - 1 if fl > 0 else (NI + b if rank > 0 else me)
                rows.append([first, prev, me] + ([NI + 2 * B + me] if text else []))
        _tables[key] = torch.tensor(rows, dtype=torch.int32, device=device)
    return _tables[key]


def _source_table(BF: int, device, cross_frame: bool = True, text: bool = True) -> torch.Tensor:
    """[first, previous, self] frames of the image K/V (pnp_utils.py:26) -- or the image alone for the stock joint attention
    -- + the image's own text tokens (second tensor) where the attention is a joint one."""
    key = (BF, str(device), cross_frame, text)
    if key not in _tables:
        rows = []
        for i in range(BF):
            b, f = divmod(i, CLIP_LENGTH)
            row = [b * CLIP_LENGTH, b * CLIP_LENGTH + max(f - 1, 0), i] if cross_frame else [i]
            rows.append(row + ([BF + i] if text else []))
        _tables[key] = torch.tensor(rows, dtype=torch.int32, device=device)
    return _tables[key]


class CrossFrameProcessor:
    """pnp_utils.py:9-132."""

    cross_frame = True   # image K/V from the [first, previous, self] frames (pnp_utils.py:26)

    def __init__(self):
        self._packed = {}

    def _shift(self, idx):
        return None

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, idx=-1, *args, **kwargs):
        if attention_mask is not None:
            raise NotImplementedError("attention masks are unused on the UniVST path")
        dev = hidden_states.device
        pk = self._packed.get(id(attn))
        if pk is None:
            pk = self._packed[id(attn)] = _Packed(attn, dev)
        BF, N, C = hidden_states.shape
        H, d = pk.heads, C // pk.heads
        # frames sharded over ranks (sd3_transformer.set_frame_sharding): this call sees Fl = 16 / world frames per branch;
        # the K/V of the neighbouring frames that live on other ranks arrive in two halo banks behind the local images
        shard = getattr(attn, "_shard", None) if self.cross_frame else None
        Fl = shard.Fl if shard is not None else CLIP_LENGTH
        if self.cross_frame and BF % Fl:
            raise ValueError(f"the reference processors assume clips of {CLIP_LENGTH} frames (batch {BF})")
        x = hidden_states.to(torch.float16).reshape(BF * N, C).contiguous()
        NIkv = BF
        if shard is None:
            qkv = kv = ops.gemm(x, pk.w_qkv, bias=pk.b_qkv)          # [BF N, 3C]
        else:
            B = BF // Fl
            NIkv = BF + 2 * B
            kv, halo_ptrs, halo_mc = shard.buffers.next(NIkv * N, 3 * C)
            qkv = ops.gemm(x, pk.w_qkv, bias=pk.b_qkv, out=kv[: BF * N])
        if pk.norm_q is not None or pk.norm_k is not None:
            ops.rmsnorm_heads_(qkv, H, d, pk.norm_q, pk.norm_k, pk.eps)
        shift = self._shift(idx)
        if shift is not None:
            if BF != 3 * Fl:
                raise ValueError("the AdaIN-guided shift needs the three-branch batch [content, style, edit]")
            ops.sd3_attn_shift_(qkv, Fl, N, H, d, *shift)
        if shard is not None:
            from .xrank import push_kv_halo
            push_kv_halo(shard.xr, kv, halo_ptrs, halo_mc, BF // Fl, Fl, N, C)
        table_of = (lambda text: _source_table_sharded(BF // Fl, Fl, shard.xr.rank, dev, text)) if shard is not None else \
            (lambda text: _source_table(BF, dev, self.cross_frame, text))
        if encoder_hidden_states is None:
            # the image-only attention (attn2 of SD3.5's dual-attention blocks; pnp_utils.py:92,120,134 skip the text half)
            o = ops.sc_attention(qkv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:], table_of(False), NI=BF, NIkv=NIkv, H=H, d=d, N=N, Nkv=N)
            return ops.gemm(o, pk.w_out, bias=pk.b_out).view(BF, N, C)
        L = encoder_hidden_states.shape[1]
        e = encoder_hidden_states.to(torch.float16).reshape(BF * L, C).contiguous()
        tqkv = ops.gemm(e, pk.w_add, bias=pk.b_add)         # [BF L, 3C]
        if pk.norm_added_q is not None or pk.norm_added_k is not None:
            ops.rmsnorm_heads_(tqkv, H, d, pk.norm_added_q, pk.norm_added_k, pk.eps)
        table = table_of(True)
        kw = dict(NI=BF, NIkv=NIkv, NIkv2=BF, H=H, d=d, Nkv=N, Nkv2=L)
        o_img = ops.joint_attention(qkv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:], tqkv[:, C:2 * C], tqkv[:, 2 * C:], table, N=N, **kw)
        o_txt = ops.joint_attention(tqkv[:, :C], kv[:, C:2 * C], kv[:, 2 * C:], tqkv[:, C:2 * C], tqkv[:, 2 * C:], table, N=L, **kw)
        h_out = ops.gemm(o_img, pk.w_out, bias=pk.b_out).view(BF, N, C)
        if getattr(attn, "context_pre_only", False) or pk.w_add_out is None:
            return h_out, o_txt.view(BF, L, C)
        return h_out, ops.gemm(o_txt, pk.w_add_out, bias=pk.b_add_out).view(BF, L, C)


class JointAttnProcessor(CrossFrameProcessor):
    """diffusers' stock ``JointAttnProcessor2_0`` (third-party): every image attends to its own tokens + its text tokens.  The
    default processor of :class:`univst_b200.sd3_transformer.SD3Transformer2DModel` until the reference's are installed."""
    cross_frame = False


class AttentionShiftProcessor(CrossFrameProcessor):
    """pnp_utils.py:135-271: window ``eta1*50 <= idx <= eta2*50``, alpha 0.8, gamma 2.0, beta 0.9 -> 0.1."""

    def __init__(self, eta1, eta2):
        super().__init__()
        self.eta1, self.eta2 = eta1, eta2
        self.thresh2 = eta2   # read at :185, never set in the reference

    def _shift(self, idx):
        if idx >= self.eta1 * 50 and idx <= self.eta2 * 50:
            beta = (0.9 - 0.1) / (self.eta1 * 50 - self.thresh2 * 50) * (idx - self.eta2 * 50) + 0.1
            return 0.8, beta, 2.0
        return None


def register_spatial_attention_pnp(model, eta1=0.0, eta2=0.6):
    """pnp_utils.py:276-284: install the shift processor on every attention of ``model.transformer``."""
    procs = {name: (AttentionShiftProcessor(eta1, eta2) if "attn" in name else proc)
             for name, proc in model.transformer.attn_processors.items()}
    model.transformer.set_attn_processor(procs)


class FeatureDumpTransformer:
    """``CustomSD3Transformer2DModel`` (backbones/video_diffusion_sd3/models/transformer_3D_model.py:12-113) without
    re-stating the third-party forward: the reference subclasses diffusers' ``SD3Transformer2DModel`` only to accept
    ``idx`` / ``ft_indices`` / ``ft_timesteps`` / ``ft_path`` and to save, after block ``i`` in ``ft_indices`` when ``idx`` in
    ``ft_timesteps``, the image stream as ``inversion_feature_map_{i}_block_{idx}_step.pt`` with shape
    (B, h / 2, w / 2, C) (:77-84 -- the features mask propagation reads).  This wrapper adds the same keywords and the same
    files to a stock transformer through forward hooks on ``transformer_blocks[i]`` (a joint block returns
    ``(encoder_hidden_states, hidden_states)``; the last block's first element may be None)."""

    def __init__(self, transformer):
        self.transformer = transformer
        self.config = transformer.config

    def __getattr__(self, name):   # attn_processors, set_attn_processor, dtype, device, ...
        return getattr(self.transformer, name)

    def __call__(self, hidden_states, *args, idx=0, ft_indices=None, ft_timesteps=None, ft_path=None, **kwargs):
        import os
        hooks = []
        if ft_indices is not None and ft_timesteps and ft_path is not None and idx in ft_timesteps:
            h2, w2 = hidden_states.shape[-2] // 2, hidden_states.shape[-1] // 2

            def saver(i):
                def hook(_module, _inputs, output):
                    hs = output[1] if isinstance(output, (tuple, list)) else output
                    path = os.path.join(ft_path, f"inversion_feature_map_{i}_block_{idx}_step.pt")
                    torch.save(hs.view(hs.shape[0], h2, w2, -1).detach(), path)
                return hook
            for i in ft_indices:
                hooks.append(self.transformer.transformer_blocks[i].register_forward_hook(saver(i)))
        try:
            return self.transformer(hidden_states, *args, **kwargs)
        finally:
            for h in hooks:
                h.remove()
