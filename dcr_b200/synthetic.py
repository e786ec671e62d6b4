"""Deterministic synthetic inputs shared by tests, smoke() and bench.py.

Descriptors: unit-norm gaussian rows; a fraction of the gallery rows is replaced by noisy copies of queries so that
top-1 has planted ground truth and the scores span [0, 1] as in DCR's use (replicated images).
Images: uint8 [N,256,256,3] low-frequency random fields (bilinear-upsampled 8x8 noise + per-image colour offset);
a fraction of the query images are brightness/shift-augmented copies of gallery images.
All generators use a CPU torch.Generator so the same seed gives the same bytes on every machine.
"""
from __future__ import annotations

import numpy as np
import torch


def _normalize(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=1, keepdim=True).clamp_min(1e-12)


def descriptors(nq: int, ng: int, d: int, seed: int = 0, planted: float = 0.01, noise: float = 0.1):
    """(q f32[nq,d], g f32[ng,d]) CPU tensors, rows unit-norm."""
    gq = torch.Generator().manual_seed(1000 + seed)
    gg = torch.Generator().manual_seed(2000 + seed)
    q = _normalize(torch.randn(nq, d, generator=gq))
    g = _normalize(torch.randn(ng, d, generator=gg))
    n_pl = int(min(ng, max(0, round(planted * ng))))
    if n_pl > 0 and nq > 0:
        rows = torch.randperm(ng, generator=gg)[:n_pl]
        src = torch.randint(0, nq, (n_pl,), generator=gg)
        g[rows] = _normalize(q[src] + noise * torch.randn(n_pl, d, generator=gg))
    return q.contiguous(), g.contiguous()


def images(n: int, seed: int = 0, size: int = 256, copies_of: torch.Tensor | None = None,
           copy_fraction: float = 0.1) -> torch.Tensor:
    """uint8 [n,size,size,3] CPU tensor."""
    gen = torch.Generator().manual_seed(3000 + seed)
    low = torch.rand(n, 3, 8, 8, generator=gen)
    img = torch.nn.functional.interpolate(low, size=(size, size), mode="bilinear", align_corners=False)
    img = img * 0.6 + 0.4 * torch.rand(n, 3, 1, 1, generator=gen)
    img = img + 0.04 * torch.randn(n, 3, size, size, generator=gen)
    if copies_of is not None and copy_fraction > 0 and n > 0:
        n_c = int(round(copy_fraction * n))
        if n_c > 0:
            dst = torch.randperm(n, generator=gen)[:n_c]
            src = torch.randint(0, copies_of.shape[0], (n_c,), generator=gen)
            base = copies_of[src].permute(0, 3, 1, 2).float() / 255.0
            gain = 0.8 + 0.4 * torch.rand(n_c, 1, 1, 1, generator=gen)
            sh = torch.randint(-8, 9, (n_c, 2), generator=gen)
            out = torch.empty_like(base)
            for i in range(n_c):
                out[i] = torch.roll(base[i], shifts=(int(sh[i, 0]), int(sh[i, 1])), dims=(1, 2))
            img[dst] = out * gain
    return (img.clamp(0, 1) * 255.0).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
