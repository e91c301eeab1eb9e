"""Generate the committed golden vectors under ``tests/golden/`` by running the REFERENCE's own code.

Runs only in the build container (needs ``/root/reference``, read-only).  The reference's model files are imported on
top of the test-only diffusers shim (``oracle/_shim``); ``torch.cuda.get_device_name`` is patched because
``SpatioTemporalTransformerBlock.__init__`` calls it (models/attention.py:238).  Weights are the deterministic
``seeded_state_dict`` (no checkpoint exists offline), loaded into the reference module with ``load_state_dict``.

    python oracle/gen_golden.py            # rewrites tests/golden/*.pt
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "_shim"))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
torch.cuda.get_device_name = lambda *a, **k: "cpu-shim"
OUT = os.path.join(ROOT, "tests", "golden")


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    torch.save(obj, os.path.join(OUT, name))
    print("wrote", name, os.path.getsize(os.path.join(OUT, name)) // 1024, "KiB")


def gen_unet():
    from backbones.video_diffusion_sd import pnp_utils
    from backbones.video_diffusion_sd.models.unet_3d_condition import UNetPseudo3DConditionModel
    from oracle import unet_oracle as uo

    cfg = uo.TINY_CONFIG
    m = UNetPseudo3DConditionModel(block_out_channels=cfg["block_out_channels"], attention_head_dim=cfg["attention_head_dim"],
                                   cross_attention_dim=cfg["cross_attention_dim"], sample_size=16).eval()
    ref_sd = m.state_dict()
    shapes = uo.unet_param_shapes(cfg)
    assert set(shapes) == set(ref_sd) and all(tuple(ref_sd[k].shape) == shapes[k] for k in shapes)
    m.load_state_dict(uo.seeded_state_dict(cfg, seed=33))
    g = torch.Generator().manual_seed(2024)
    x = torch.randn(3, 4, 3, 16, 16, generator=g)
    ctx = torch.randn(1, 77, cfg["cross_attention_dim"], generator=g).repeat(3, 1, 1)
    out = {"x": x, "ctx": ctx, "seed": 33, "cases": {}}
    with torch.no_grad():
        out["cases"]["stock_t981"] = m(x, torch.tensor(981), encoder_hidden_states=ctx).sample.clone()
        pipe = types.SimpleNamespace(unet=m)
        pnp_utils.register_spatial_attention_pnp(pipe)
        for idx, t in ((0, 981), (13, 721), (25, 481), (26, 461)):
            pnp_utils.register_time(pipe, idx)
            out["cases"][f"patched_idx{idx}_t{t}"] = m(x, torch.tensor(t), encoder_hidden_states=ctx).sample.clone()
    save("unet_tiny.pt", out)

    # SD-2.1 layout (use_linear_projection, per-level head counts, head dim 64)
    cfg = uo.TINY_SD21_CONFIG
    m = UNetPseudo3DConditionModel(block_out_channels=cfg["block_out_channels"], attention_head_dim=cfg["attention_head_dim"],
                                   cross_attention_dim=cfg["cross_attention_dim"], use_linear_projection=True,
                                   sample_size=16).eval()
    shapes = uo.unet_param_shapes(cfg)
    ref_sd = m.state_dict()
    assert set(shapes) == set(ref_sd) and all(tuple(ref_sd[k].shape) == shapes[k] for k in shapes)
    m.load_state_dict(uo.seeded_state_dict(cfg, seed=21))
    ctx = torch.randn(1, 77, cfg["cross_attention_dim"], generator=g).repeat(3, 1, 1)
    out = {"x": x, "ctx": ctx, "seed": 21, "cases": {}}
    with torch.no_grad():
        out["cases"]["stock_t501"] = m(x, torch.tensor(501), encoder_hidden_states=ctx).sample.clone()
        pipe = types.SimpleNamespace(unet=m)
        pnp_utils.register_spatial_attention_pnp(pipe)
        pnp_utils.register_time(pipe, 10)
        out["cases"]["patched_idx10_t781"] = m(x, torch.tensor(781), encoder_hidden_states=ctx).sample.clone()
    save("unet_tiny_sd21.pt", out)


def gen_pnp():
    from backbones.video_diffusion_sd import pnp_utils

    g = torch.Generator().manual_seed(7)
    cnt, sty = torch.randn(3, 32, 16, generator=g) * 1.5 + 0.3, torch.randn(3, 32, 16, generator=g) * 0.7 - 0.2
    zc, zs = torch.randn(1, 4, 3, 8, 8, generator=g), torch.randn(1, 4, 3, 8, 8, generator=g) * 0.5 + 0.1
    out = {"cnt": cnt, "sty": sty, "attention_adain": pnp_utils.attention_adain(cnt, sty),
           "zc": zc, "zs": zs, "latent_adain": pnp_utils.latent_adain(zc, zs)}
    # the patched attn1.forward on a stand-in module exposing exactly what it touches (pnp_utils.py:29-43,97-99)
    C, heads, Fr, N = 32, 4, 3, 16
    attn = torch.nn.Module()
    attn.to_q, attn.to_k, attn.to_v = (torch.nn.Linear(C, C, bias=False) for _ in range(3))
    attn.to_out = torch.nn.ModuleList([torch.nn.Linear(C, C), torch.nn.Dropout(0.0)])
    attn.heads, attn.group_norm = heads, None
    torch.manual_seed(11)
    for p in attn.parameters():
        torch.nn.init.normal_(p, std=0.2)
    blk = types.SimpleNamespace(attn1=attn, attn2=torch.nn.Module())
    ups = [types.SimpleNamespace(attentions=[types.SimpleNamespace(transformer_blocks=[blk]) for _ in range(3)])
           for _ in range(4)]
    pipe = types.SimpleNamespace(unet=types.SimpleNamespace(up_blocks=ups))
    pnp_utils.register_spatial_attention_pnp(pipe)
    x = torch.randn(3 * Fr, N, C, generator=g)
    out["attn1"] = {"weights": {k: v.clone() for k, v in attn.state_dict().items()}, "x": x, "heads": heads, "F": Fr,
                    "out": {}}
    with torch.no_grad():
        for idx in (0, 13, 25, 26):
            pnp_utils.register_time(pipe, idx)
            out["attn1"]["out"][idx] = attn.forward(hidden_states=x.clone(), clip_length=Fr).clone()
    save("pnp_utils.pt", out)


def main():
    gen_unet()
    gen_pnp()
    try:
        from oracle import gen_golden_extra
        gen_golden_extra.main(save)
    except ImportError:
        pass


if __name__ == "__main__":
    main()
