"""Sustained (power-capped steady state) A/B of the attention variants: each variant runs back to back for several
seconds; the average launch time over the last part of the window is reported with the SM clock / power nvidia-smi saw.
Short bursts run at boost clocks and rank the variants differently from the 50-step loop, which sits at the power cap.

    python tools/attn_sustained.py [seconds] [variants ...]
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from univst_b200 import ops
from univst_b200.unet import kv_source_table

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
variants = [int(a) for a in sys.argv[2:]] or [0, 9, 16, 19]
B, F, H, d, N = 3, 16, 8, 40, 4096
C, NI = H * d, B * F
torch.manual_seed(0)
qkv = (torch.randn(NI * N, 3 * C, device="cuda") * 1.5).half()
q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
table = kv_source_table(B, F, "prev_first").cuda()
o = torch.empty((NI * N, C), dtype=torch.float16, device="cuda")
run = lambda: ops.sc_attention(q, k, v, table, NI=NI, NIkv=NI, H=H, d=d, N=N, Nkv=N, out=o)


def smi():
    r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu", "--format=csv,noheader,nounits", "--id=0"],
                       capture_output=True, text=True).stdout.strip()
    return r


for rnd in range(2):
    for var in variants:
        ops.attention_tune(var, 1, -1)
        run()
        torch.cuda.synchronize()
        t_end = time.time() + secs
        times, samples = [], []
        while time.time() < t_end:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40):
                run()
            e1.record()
            samples.append(smi())          # sampled while the next batch is already queued? no: sync below first
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 40)
        tail = times[len(times) // 2:]
        print(f"round {rnd} variant {var:2d}: first {times[0]:.3f} ms, steady {sum(tail) / len(tail):.3f} ms over {len(tail)} batches; "
              f"smi(clock MHz, W, C) {samples[-1]}", flush=True)
ops.attention_tune(-1, -1, -1)
