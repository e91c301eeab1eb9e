"""GPU parity at BASELINE.json's full sizes (configs[1] and the configs[3] backbone): one patched three-branch UNet
call, 3 x 16 frames at 64 x 64 latents, full SD-1.5 width.

The CPU oracle would need minutes per call at this size, so the same oracle code is evaluated on the GPU with PyTorch's
own kernels: in fp32 (the "truth") and in fp16 (the precision the reference runs at).  Bar: our fp16 path is within
rel-L2 1e-2 of the fp32 truth and no worse than twice the torch-fp16 error -- the tolerance tests/test_unet_gpu.py
states for the tiny goldens, now at the size the benchmark runs."""
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm()).item()


def _inputs():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 4, 16, 64, 64, generator=g).cuda()
    ctx = torch.randn(1, 77, 768, generator=g).repeat(3, 1, 1).cuda()
    return x, ctx


@pytest.mark.parametrize("backbone,idx", [("sd", 5), ("sd", 30), ("animatediff", 5), ("sd21", 5)])
def test_full_size_patched_forward(cuda_lib, backbone, idx):
    from univst_b200 import pnp_utils
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if backbone in ("sd", "sd21"):
        from oracle import unet_oracle as orc
        from univst_b200.unet import UNetPseudo3DConditionModel as Net
        cfg = orc.SD15_CONFIG if backbone == "sd" else orc.SD21_CONFIG   # configs[2] backbone: head dim 64, ctx 1024
    else:
        from oracle import animatediff_oracle as orc
        from univst_b200.animatediff import UNet3DConditionModel as Net
        cfg = orc.AD_SD15_CONFIG
    sd = orc.seeded_state_dict(cfg, seed=33)
    unet = Net(sd, cfg)
    pipe = SimpleNamespace(unet=unet)
    pnp_utils.register_spatial_attention_pnp(pipe)
    pnp_utils.register_time(pipe, idx)
    x, ctx = _inputs()
    if cfg["cross_attention_dim"] != ctx.shape[-1]:
        ctx = torch.randn(1, 77, cfg["cross_attention_dim"], generator=torch.Generator().manual_seed(8)).repeat(3, 1, 1).cuda()
    t = 981 - 20 * idx
    y = unet(x.half(), t, encoder_hidden_states=ctx.half()).sample
    torch.cuda.synchronize()
    del unet
    sd32 = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        truth = orc.unet_forward(sd32, cfg, x, t, ctx, patched=True, idx=idx)
    del sd32
    sd16 = {k: v.cuda().half() for k, v in sd.items()}
    with torch.no_grad():
        y16 = orc.unet_forward(sd16, cfg, x.half(), t, ctx.half(), patched=True, idx=idx)
    rel, rel16 = _rel(y, truth), _rel(y16, truth)
    edit, edit16 = _rel(y[2], truth[2]), _rel(y16[2], truth[2])
    print(f"full size {backbone} idx={idx}: ours rel={rel:.3e} (edit branch {edit:.3e}) | torch-fp16 rel={rel16:.3e} (edit {edit16:.3e})")
    assert torch.isfinite(y).all() and y.shape == truth.shape
    assert rel <= 1e-2 and rel <= 2.0 * rel16 + 1e-3
    assert edit <= 1e-2 and edit <= 2.0 * edit16 + 1e-3


def test_full_size_aux_kernels(cuda_lib):
    """Mask propagation (4096 x 15000 x 640) and the 16 x 512 x 512 flow-warp window pass against their oracles."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "aux_kernels_bench.py")], capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    print(json.dumps(res, indent=1))
    assert res["maskprop"]["max_abs_err_vs_oracle"] < 1e-4 and res["flow_warp_window"]["bit_exact_vs_oracle"]
