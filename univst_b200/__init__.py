"""univst_b200 -- B200 (sm_100a) kernels + host code for UniVST's three-branch DDIM denoising hot path.

Layout: ``csrc/`` hand-written CUDA behind the C ABI of ``include/univst_b200.h``; ``ops.py`` ctypes wrappers;
``unet.py`` / ``pnp_utils.py`` / ``pipeline.py`` / ``ddim_inversion.py`` / ``mask_propagation.py`` /
``flow_warp.py`` mirror the reference's Python interface for the path (same names and argument meaning).
"""
__version__ = "0.1.0"
