"""A/B of the fused attention variants at the headline shape: each variant is checked against an fp32 torch softmax
on a slice (so a wrong variant cannot win) and timed back to back with CUDA events.

    python tools/attn_variants.py [variants ...]      # e.g. 0 9 16 17 19 (see the switch in attention_tc.cu)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from univst_b200 import ops
from univst_b200.unet import kv_source_table

variants = [int(a) for a in sys.argv[1:]] or [0, 9, 16, 17, 19]
B, F, H, d, N = 3, 16, 8, 40, 4096
C = H * d
NI = B * F
torch.manual_seed(0)
qkv = (torch.randn(NI * N, 3 * C, device="cuda") * 1.5).half()
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]


def reference(img, head, rows, table):
    qq = q.view(NI, N, H, d)[img, rows, head].float()
    srcs = table[img].tolist()
    kk = torch.cat([k.view(NI, N, H, d)[s, :, head] for s in srcs]).float()
    vv = torch.cat([v.view(NI, N, H, d)[s, :, head] for s in srcs]).float()
    p = torch.softmax(qq @ kk.T * d ** -0.5, dim=-1)
    return p @ vv


def timed(fn, iters=6):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


import statistics

for mode in ("prev_first", "prev_self_first"):
    table_cpu = kv_source_table(B, F, mode)
    table = table_cpu.cuda()
    nsrc = table.shape[1]
    flops = 4.0 * N * N * nsrc * C * NI
    o = torch.empty((NI * N, C), dtype=torch.float16, device="cuda")
    run = lambda: ops.sc_attention(q, k, v, table, NI=NI, NIkv=NI, H=H, d=d, N=N, Nkv=N, out=o)
    errs, times = {}, {var: [] for var in variants}
    for var in variants:
        ops.attention_tune(var, 1, -1)
        o.zero_()
        run()
        torch.cuda.synchronize()
        err = 0.0
        for img, head in ((0, 0), (1, 3), (17, 7), (47, 5)):
            rows = torch.arange(1000, 1128, device="cuda")
            ref = reference(img, head, rows, table_cpu)
            got = o.view(NI, N, H, d)[img, rows, head].float()
            err = max(err, ((got - ref).norm() / ref.norm()).item())
        errs[var] = err
    for rnd in range(4):   # interleaved rounds: clock / power drift hits every variant alike
        for var in variants:
            ops.attention_tune(var, 1, -1)
            times[var].append(timed(run, iters=8))
    for var in variants:
        ms = statistics.median(times[var])
        print(f"{mode:16s} variant {var}: median {ms:7.3f} ms (min {min(times[var]):7.3f})  {flops / ms / 1e9:7.1f} TF/s "
              f"(algorithmic)  rel err {errs[var]:.2e} {'OK' if errs[var] < 3e-3 else 'WRONG'}", flush=True)
ops.attention_tune(-1, -1, -1)
