"""Host-side cost of one three-branch UNet call (Python + ctypes + allocator), measured WITHOUT a GPU: every C entry point is
replaced by a no-op of the same name and prototype (a throw-away shared library built with gcc), tensors live on the CPU and
are never touched.  What is left is exactly the work the host does between kernel launches -- the floor of a call when the
GPU side shrinks (frame sharding over 4-8 GPUs: ~10-20 ms of kernels per rank).

    python tools/host_overhead.py [frames] [--profile]
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from univst_b200 import _lib, ops, pnp_utils
from univst_b200.unet import SD15_CONFIG, UNetPseudo3DConditionModel
from univst_b200.weights import unet_param_shapes


def noop_library():
    special = {"univst_last_error", "univst_device_check", "univst_abi_version"}
    names = sorted(n for n in _lib.PROTOTYPES if n not in special)
    src = "".join(f"long long {n}() {{ return {'4096' if n.endswith('_bytes') else '1' if n.endswith('_supported') else '0'}; }}\n"
                  for n in names)
    src += 'const char* univst_last_error() { return ""; }\nint univst_device_check() { return 0; }\nint univst_abi_version() { return 1; }\n'
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "noop.c"), "w") as f:
        f.write(src)
    so = os.path.join(d, "libnoop.so")
    subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-w", os.path.join(d, "noop.c"), "-o", so], check=True)
    h = C.CDLL(so)
    h.univst_last_error.restype = C.c_char_p
    for name, argtypes in _lib.PROTOTYPES.items():
        fn = getattr(h, name)
        fn.argtypes = argtypes
        fn.restype = _lib._RESTYPES.get(name, C.c_int)
    return h


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 16
    _lib._lib, _lib._device_ok = noop_library(), True
    ops._stream = lambda: 0
    torch.cuda.current_stream = lambda *a, **k: SimpleNamespace(cuda_stream=0)   # the only CUDA runtime query on the path
    ops._chk = lambda *a, **k: None            # the CUDA-placement checks (cheap attribute reads) cannot pass on CPU tensors
    torch.Tensor.is_cuda = property(lambda self: True)   # (the wrappers' own `t.is_cuda` asserts likewise)
    sd = {k: torch.empty(s, dtype=torch.float16) for k, s in unet_param_shapes(SD15_CONFIG).items()}
    for k in sd:   # the temporal parts at their constructor values (the UNet verifies them at pack time)
        if "attn_temporal.to_out.0.weight" in k or "conv_temporal" in k:
            sd[k].zero_()
        if "conv_temporal.weight" in k:
            sd[k] = torch.nn.init.dirac_(torch.zeros(sd[k].shape)).half()
    unet = UNetPseudo3DConditionModel(sd, SD15_CONFIG, device="cpu")
    pipe = SimpleNamespace(unet=unet)
    pnp_utils.register_spatial_attention_pnp(pipe)
    pnp_utils.register_time(pipe, 5)
    x = torch.empty(3, 4, F, 64, 64, dtype=torch.float16)
    ctx = torch.empty(3, 77, 768, dtype=torch.float16)
    for _ in range(3):
        unet(x, 981, encoder_hidden_states=ctx)
    n0, t0 = ops.launch_count, time.perf_counter()
    iters = 20
    for _ in range(iters):
        unet(x, 981, encoder_hidden_states=ctx)
    dt = (time.perf_counter() - t0) / iters
    print(f"host side of one 3 x {F} x 64 x 64 UNet call: {dt * 1e3:.2f} ms for {(ops.launch_count - n0) // iters} launches "
          f"({dt * 1e6 / ((ops.launch_count - n0) / iters):.1f} us per launch)")
    if "--profile" in sys.argv:
        import cProfile
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(10):
            unet(x, 981, encoder_hidden_states=ctx)
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(22)


if __name__ == "__main__":
    main()
