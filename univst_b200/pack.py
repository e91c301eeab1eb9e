"""One-off weight packing: reference ``state_dict`` layouts -> the layouts the sm_100a kernels consume."""
from __future__ import annotations

import torch


def geglu_tile(n: int) -> int:
    """Tile width the GEMM picks for a GEGLU projection with ``n`` = 8C rows (mirrors pick_bn in gemm_tc.cu)."""
    return 256 if n % 256 == 0 else 128


def pack_geglu(w: torch.Tensor, b: torch.Tensor):
    """diffusers ``GEGLU.proj`` weight [8C, C] (rows [0, 4C) = values, [4C, 8C) = gates) -> rows interleaved per
    output tile so that one accumulator tile holds BN/2 value columns followed by their BN/2 gate columns; the
    ``h * gelu(gate)`` product is then formed in the GEMM epilogue and the [tokens, 8C] projection never exists."""
    n = w.shape[0]
    half = n // 2
    bn = geglu_tile(n)
    hb = bn // 2
    assert half % hb == 0, f"GEGLU inner dim {half} must be a multiple of {hb}"
    idx = []
    for t in range(half // hb):
        idx.append(torch.arange(t * hb, (t + 1) * hb))
        idx.append(half + torch.arange(t * hb, (t + 1) * hb))
    idx = torch.cat(idx).to(w.device)
    return w[idx].contiguous(), b[idx].contiguous()


def pack_conv3x3(w: torch.Tensor, cin_pad: int = 0) -> torch.Tensor:
    """Conv2d weight [Cout, Cin, 3, 3] -> [Cout, 3, 3, Cin(+pad)] flattened to [Cout, 9 Cin] (tap-major K)."""
    cout, cin = w.shape[:2]
    w = w.permute(0, 2, 3, 1)
    if cin_pad > cin:
        w = torch.nn.functional.pad(w, (0, cin_pad - cin))
    return w.reshape(cout, -1).contiguous()


def pad_rows(w: torch.Tensor, rows: int) -> torch.Tensor:
    if w.shape[0] >= rows:
        return w.contiguous()
    return torch.nn.functional.pad(w, (0, 0) * (w.dim() - 1) + (0, rows - w.shape[0])).contiguous()
