"""Synthetic manifold point clouds named by BASELINE.json / SURVEY.md 8(d).  TEST INFRASTRUCTURE ONLY
(the product-side generators live in ``manifold_gp_b200/utils/synthetic.py``; these are the CPU twins the
oracle and the CPU baseline use -- same formulas, same seeds).
"""

from __future__ import annotations

import math

import torch


def torus(n: int, R: float = 2.0, r: float = 1.0, seed: int = 0, dtype=torch.float32):
    """cfg-C: u,v ~ U[0,2pi) i.i.d.; x = ((R + r cos v) cos u, (R + r cos v) sin u, r sin v)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(n, generator=g, dtype=torch.float64) * (2 * math.pi)
    v = torch.rand(n, generator=g, dtype=torch.float64) * (2 * math.pi)
    x = torch.stack(((R + r * v.cos()) * u.cos(), (R + r * v.cos()) * u.sin(), r * v.sin()), dim=1)
    return x.to(dtype).contiguous()


def sphere(n: int, seed: int = 0, dtype=torch.float32):
    """cfg-D: normalised N(0, I_3)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, generator=g, dtype=torch.float64)
    return (x / x.norm(dim=1, keepdim=True)).to(dtype).contiguous()


def swiss_roll(n: int, seed: int = 0, dtype=torch.float32):
    """cfg-D: t ~ U[1.5pi, 4.5pi], h ~ U[0, 21] -> (t cos t, h, t sin t)."""
    g = torch.Generator().manual_seed(seed)
    t = (1.5 + 3.0 * torch.rand(n, generator=g, dtype=torch.float64)) * math.pi
    h = 21.0 * torch.rand(n, generator=g, dtype=torch.float64)
    return torch.stack((t * t.cos(), h, t * t.sin()), dim=1).to(dtype).contiguous()


def rmnist_shape(n: int = 70000, d: int = 784, prototypes: int = 70, seed: int = 0, dtype=torch.float32):
    """cfg-B: ``prototypes`` random pairs (p, q) ~ U[-0.5,0.5]^d, each deformed along a smooth 1-parameter
    family p cos(theta) + q sin(theta), theta ~ U[-pi/4, pi/4] -- 1-D manifolds in R^d like rotated digits."""
    g = torch.Generator().manual_seed(seed)
    per = (n + prototypes - 1) // prototypes
    p = torch.rand(prototypes, d, generator=g) - 0.5
    q = torch.rand(prototypes, d, generator=g) - 0.5
    theta = (torch.rand(prototypes, per, generator=g) - 0.5) * (math.pi / 2)
    x = p.unsqueeze(1) * theta.cos().unsqueeze(-1) + q.unsqueeze(1) * theta.sin().unsqueeze(-1)
    return x.reshape(-1, d)[:n].to(dtype).contiguous()


def circle_curve(n: int, seed: int = 0, dtype=torch.float32):
    """Small 1-D closed curve in R^2 (dumbbell-like) for quick tests."""
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(n, generator=g, dtype=torch.float64) * (2 * math.pi)
    rad = 1.0 + 0.3 * (2 * t).cos()
    return torch.stack((rad * t.cos(), rad * t.sin()), dim=1).to(dtype).contiguous()
