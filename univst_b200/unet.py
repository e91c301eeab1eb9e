"""placeholder -- filled in below"""
import torch


def kv_source_table(B: int, F: int, mode: str) -> torch.Tensor:
    """int32 [B*F, nsrc]: K/V source images of image (b, f) for the sparse-causal modes of the reference:
    ``prev_first`` = SparseCausalAttention_index [-1, 'first'] (patched attn1, pnp_utils.py:25),
    ``prev_self_first`` = [-1, 0, 'first'] (stock, models/attention.py:356), ``self`` = plain self-attention."""
    rows = []
    for b in range(B):
        for f in range(F):
            prev, first, me = b * F + max(f - 1, 0), b * F, b * F + f
            rows.append({"prev_first": [prev, first], "prev_self_first": [prev, me, first], "self": [me]}[mode])
    return torch.tensor(rows, dtype=torch.int32)
