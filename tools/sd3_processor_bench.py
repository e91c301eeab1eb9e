"""SD3.5-medium-shaped joint attention (BASELINE.json configs[4] layer shape: 24 heads x 64, 4096 image + 333 text tokens,
3 branches x 16 frames): our AttentionShiftProcessor vs the oracle code evaluated with PyTorch's own kernels on the same
GPU (fp32 = truth for the error, fp16 = the library path to beat for the time).  Prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import sd3_oracle as so
from univst_b200 import sd3



def _attn_module(w, heads):
    """Stand-in for the diffusers Attention module: exactly the attributes the processors read."""
    C = w["to_q.weight"].shape[0]
    attn = torch.nn.Module()
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_add_out"):
        setattr(attn, n, torch.nn.Linear(C, C))
    attn.to_out = torch.nn.ModuleList([torch.nn.Linear(C, C), torch.nn.Dropout(0.0)])

    class _Norm(torch.nn.Module):
        def __init__(self, dim):
            super().__init__()
            self.weight, self.eps = torch.nn.Parameter(torch.ones(dim)), 1e-6
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        setattr(attn, n, _Norm(C // heads))
    attn.heads, attn.context_pre_only = heads, False
    attn.load_state_dict(w)
    return attn


heads, d, N, L = 24, 64, 4096, 333
C = heads * d
torch.backends.cuda.matmul.allow_tf32 = False
w = so.seeded_attn_weights(C, heads, 3)
attn = _attn_module(w, heads)
g = torch.Generator().manual_seed(5)
hidden = torch.randn(48, N, C, generator=g).cuda()
enc = torch.randn(48, L, C, generator=g).cuda()
proc = sd3.AttentionShiftProcessor(0.0, 0.6)
idx = 5


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / iters


h16, e16 = hidden.half(), enc.half()
(h, e), ms = timed(lambda: proc(attn, h16, e16, idx=idx))
w16 = {k: v.cuda().half() for k, v in w.items()}
with torch.no_grad():
    (ht, et), ms_torch = timed(lambda: so.joint_attention(w16, h16, e16, heads, idx=idx), iters=2)
    w32 = {k: v.cuda() for k, v in w.items()}
    hr, er = so.joint_attention(w32, hidden, enc, heads, idx=idx)
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
flops = 48 * (4.0 * (N + L) * (3 * N + L) * C + 2.0 * (N + L) * C * C * 4)
print(json.dumps({"shape": f"48 images x ({N} + {L}) tokens, {heads} heads x {d}", "ours_ms": ms, "torch_fp16_eager_ms": ms_torch,
                  "speedup_vs_torch": ms_torch / ms, "tflops_algorithmic": flops / ms / 1e9,
                  "rel_err_hidden_ours": rel(h, hr), "rel_err_hidden_torch_fp16": rel(ht, hr),
                  "rel_err_text_ours": rel(e, er), "rel_err_text_torch_fp16": rel(et, er)}))
