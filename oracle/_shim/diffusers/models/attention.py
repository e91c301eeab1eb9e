"""Restatement of diffusers.models.attention{,_processor}: Attention (AttnProcessor2_0), FeedForward(GEGLU)."""
import torch
import torch.nn.functional as F
from torch import nn


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, **kwargs):
        super().__init__()
        inner = heads * dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.group_norm = None
        self.added_kv_proj_dim = None
        self.upcast_attention = upcast_attention
        self.upcast_softmax = False
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        b = hidden_states.shape[0]
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        d = q.shape[-1] // self.heads
        q = q.view(b, -1, self.heads, d).transpose(1, 2)
        k = k.view(b, -1, self.heads, d).transpose(1, 2)
        v = v.view(b, -1, self.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * d).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))


    # --- explicit-softmax helpers used by VersatileAttention.forward (backbones/animatediff/models/motion_module.py:
    # 297-316); restated from the published diffusers 0.35.1 Attention methods of the same names
    def head_to_batch_dim(self, tensor, out_dim=3):
        b, n, dim = tensor.shape
        tensor = tensor.reshape(b, n, self.heads, dim // self.heads).permute(0, 2, 1, 3)
        if out_dim == 3:
            tensor = tensor.reshape(b * self.heads, n, dim // self.heads)
        return tensor

    def batch_to_head_dim(self, tensor):
        bh, n, d = tensor.shape
        tensor = tensor.reshape(bh // self.heads, self.heads, n, d)
        return tensor.permute(0, 2, 1, 3).reshape(bh // self.heads, n, d * self.heads)

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        if attention_mask is None:
            base = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype, device=query.device)
            beta = 0
        else:
            base, beta = attention_mask, 1
        scores = torch.baddbmm(base, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        if self.upcast_softmax:
            scores = scores.float()
        return scores.softmax(dim=-1).to(dtype)


CrossAttention = Attention


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", **kwargs):
        super().__init__()
        assert activation_fn == "geglu"
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out or dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class AdaLayerNorm(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("AdaLayerNorm is unreachable on the UniVST path (num_embeds_ada_norm is None)")
