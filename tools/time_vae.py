"""Device timing of the VAE legs at full size (SVD VAE shapes, seeded random weights): decode of a 16 x 64 x 64 latent clip to
16 x 512 x 512 uint8 frames and the encode back (stable_diffusion.py:793-834), on one GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from univst_b200 import ops
from univst_b200.vae import AutoencoderKLTemporalDecoder, random_state_dict

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
vae = AutoencoderKLTemporalDecoder(random_state_dict(seed=55))
lat = (0.18215 * torch.randn(1, 4, F, 64, 64, device="cuda")).half()


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


n0 = ops.launch_count
frames, ms_dec = timed(lambda: vae.decode_latents_u8(lat))
n1 = ops.launch_count
back, ms_enc = timed(lambda: vae.encode_frames_u8(frames, generator=torch.Generator(device="cuda").manual_seed(0)))
print(json.dumps({"frames": F, "decode_ms": ms_dec, "encode_ms": ms_enc, "decode_launches": (n1 - n0) // 4,
                  "frames_shape": list(frames.shape), "latents_shape": list(back.shape), "finite": bool(torch.isfinite(back).all()),
                  "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}))
