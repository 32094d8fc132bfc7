"""Synthetic manifold point clouds named by BASELINE.json.  Generated on the host from a seeded CPU
generator (so every arm of the benchmark sees bit-identical points) and moved to the requested device."""
from __future__ import annotations

import math

import torch


def torus(n: int, R: float = 2.0, r: float = 1.0, seed: int = 0, dtype=torch.float32, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n, generator=g, dtype=torch.float64) * (2 * math.pi)
    v = torch.rand(n, generator=g, dtype=torch.float64) * (2 * math.pi)
    x = torch.stack(((R + r * v.cos()) * u.cos(), (R + r * v.cos()) * u.sin(), r * v.sin()), dim=1)
    return x.to(dtype).contiguous().to(device)


def sphere(n: int, seed: int = 0, dtype=torch.float32, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, generator=g, dtype=torch.float64)
    return (x / x.norm(dim=1, keepdim=True)).to(dtype).contiguous().to(device)


def swiss_roll(n: int, seed: int = 0, dtype=torch.float32, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    t = (1.5 + 3.0 * torch.rand(n, generator=g, dtype=torch.float64)) * math.pi
    h = 21.0 * torch.rand(n, generator=g, dtype=torch.float64)
    return torch.stack((t * t.cos(), h, t * t.sin()), dim=1).to(dtype).contiguous().to(device)


def rmnist_shape(n: int = 70000, d: int = 784, prototypes: int = 70, seed: int = 0, dtype=torch.float32, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    per = (n + prototypes - 1) // prototypes
    p = torch.rand(prototypes, d, generator=g) - 0.5
    q = torch.rand(prototypes, d, generator=g) - 0.5
    theta = (torch.rand(prototypes, per, generator=g) - 0.5) * (math.pi / 2)
    x = p.unsqueeze(1) * theta.cos().unsqueeze(-1) + q.unsqueeze(1) * theta.sin().unsqueeze(-1)
    return x.reshape(-1, d)[:n].to(dtype).contiguous().to(device)
