"""ORACLE (test infrastructure): CPU restatement of the rectified-flow inversion loops, inversion_tools/flow_inversion.py
:123-188 (``rf_inversion``) and :191-264 (``rf_solver``), plus the stand-in pipeline the goldens are generated with.
Parity status: PINNED -- ``oracle/gen_golden_sd3.py`` runs the reference's own two functions on ``FakePipeline`` and
commits the trajectories under ``tests/golden/rf_inversion.pt``."""
from __future__ import annotations

import contextlib

import torch


class FakePipeline:
    """What the loops touch: encode_prompt, scheduler.set_timesteps / .sigmas, transformer(...), progress_bar, device.
    The "transformer" is a fixed channel-mixing velocity field (the real MMDiT is third-party)."""

    def __init__(self, channels=16, seed=13, device="cpu", dtype=torch.float32):
        g = torch.Generator().manual_seed(seed)
        self.mix = (torch.randn(channels, channels, generator=g) * channels ** -0.5).to(device, dtype)
        self.device, self.dtype = torch.device(device), dtype
        self.scheduler = self
        self.calls = []

    def encode_prompt(self, **kw):
        z = torch.zeros(1, 4, 8, device=self.device, dtype=self.dtype)
        return z, z, z[:, 0], z[:, 0]

    def set_timesteps(self, n, device=None):
        self.sigmas = torch.cat([torch.linspace(1.0, 1.0 / n, n), torch.zeros(1)])   # FlowMatchEuler layout: descending, 0 last

    def transformer(self, hidden_states, timestep, encoder_hidden_states, pooled_projections, idx=0, return_dict=False, **kw):
        self.calls.append((idx, float(timestep[0])))
        x = hidden_states
        v = torch.tanh(torch.einsum("oc,bchw->bohw", self.mix.to(x.dtype), x)) + 0.1 * torch.sin(timestep / 1000.0 * 3.0)[:, None, None, None].to(x.dtype)
        return (v,)

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        yield type("P", (), {"update": lambda self: None})()


def rf_inversion(pipe, x, gamma, n, noise):
    pipe.scheduler.set_timesteps(n)
    ts = torch.flip(pipe.scheduler.sigmas, dims=[0])
    out = [x]
    for idx, (tc, tp) in enumerate(zip(ts[:-1], ts[1:])):
        v = pipe.transformer(x, torch.full((x.shape[0],), tc * 1000, dtype=x.dtype), None, None, idx=idx)[0]
        x = x + (tp - tc) * (gamma * (noise - x) / (1.0 - tc) + (1 - gamma) * v)
        out.append(x)
    return out


def rf_solver(pipe, x, n):
    pipe.scheduler.set_timesteps(n)
    ts = torch.flip(pipe.scheduler.sigmas, dims=[0])
    out = [x]
    for idx, (tc, tp) in enumerate(zip(ts[:-1], ts[1:])):
        v = pipe.transformer(x, torch.full((x.shape[0],), 1000 * tc, dtype=x.dtype), None, None, idx=idx)[0]
        xm = x + (tp - tc) / 2 * v
        vm = pipe.transformer(xm, torch.full((x.shape[0],), 1000 * (tc + (tp - tc) / 2), dtype=x.dtype), None, None, idx=idx)[0]
        x = x + (tp - tc) * v + 0.5 * (tp - tc) ** 2 * ((vm - v) / ((tp - tc) / 2))
        out.append(x)
    return out
