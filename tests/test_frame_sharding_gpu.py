"""Frame-sharded UNet forward over 2 GPUs == single-GPU forward (needs 2 visible GPUs; skipped otherwise)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_frame_sharded_forward_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631",
                          os.path.join(ROOT, "tools", "check_frame_sharding.py"), "4", "32", "--push", "--xrank"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    # The xrank transport (peer-memory stores + device-side flags, no collective call): same fp16 noise floor against one
    # GPU, every rank holds the same bits, and the replayed CUDA graph equals the eager launch sequence bit for bit.
    for key in ("idx5", "idx30"):
        assert res[key]["xrank_vs_single_rel_l2"] < 5e-3 and res[key]["xrank_ranks_agree"], res
        assert res[key]["xrank_graph_vs_eager_max_abs"] == 0.0, res
    # Identical arithmetic except for the order in which the GroupNorm partial sums are added across ranks; the last-bit
    # differences in the statistics re-round fp16 activations and settle at the fp16 noise floor of the network
    # (measured 1.9e-3, the same as the single-GPU path against the fp32 oracle).
    # The K/V halo pushed into the peers' symmetric memory (univst_halo_push_f16) fills the same banks with the same bytes
    # as the NCCL exchange: bit-identical outputs.
    for key in ("idx5", "idx30"):
        assert res[key]["rel_l2"] < 5e-3 and res[key]["push_vs_nccl_max_abs"] == 0.0, res


@pytest.mark.gpu
def test_animatediff_frame_sharded_forward_matches_single_gpu():
    """AnimateDiff backbone: frames sharded over 2 GPUs, motion modules through the frames <-> pixels all-to-all."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29633",
                          os.path.join(ROOT, "tools", "check_frame_sharding.py"), "4", "32", "--animatediff", "--push", "--xrank"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("idx5", "idx30"):   # xrank transport, eager and graphed: bit-identical to one GPU
        assert res[key]["xrank_vs_single_max_abs"] == 0.0 and res[key]["xrank_graph_vs_eager_max_abs"] == 0.0, res
    # every kernel sees the same operands in the same order as on one GPU (GroupNorm is per frame here): bit-identical
    # ... with the NCCL all-to-all and with the rows pushed into the peers' symmetric memory (univst_exchange_push_f16)
    for key in ("idx5", "idx30"):
        assert res[key]["max_abs"] == 0.0 and res[key]["push_vs_single_max_abs"] == 0.0, res


@pytest.mark.gpu
def test_sd3_frame_sharded_forward_matches_single_gpu():
    """SD3 / SD3.5 MMDiT mirror (BASELINE.json configs[4]): 16 frames per branch over 2 GPUs, the cross-frame attention's
    [first, previous] K/V through halo banks in peer memory -- bit-identical to one GPU (everything else is per image)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29635",
                          os.path.join(ROOT, "tools", "check_sd3_sharding.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("idx5", "idx40"):
        assert res[key]["max_abs_vs_single"] == 0.0 and res[key]["second_call_equal"] and res[key]["ranks_agree"], res
