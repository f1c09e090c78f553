"""Deterministic synthetic frames (integer arithmetic only, so CPU and GPU runs produce identical bytes).

Image = per-channel sum of three low-frequency triangle waves + hash noise (sum of four uniform bytes, sigma ~ 12)
+ a few flat rectangles, clipped to 8 bits -- natural-image-like coefficient sparsity for the entropy coder
(SURVEY.md section 8d).  Seed = frame index.
"""
from __future__ import annotations

import torch


def _hash32(x):
    x = (x ^ (x >> 16)) * 0x45D9F3B & 0xFFFFFFFF
    x = (x ^ (x >> 16)) * 0x45D9F3B & 0xFFFFFFFF
    return (x ^ (x >> 16)) & 0xFFFFFFFF


def _tri(t, period):
    """triangle wave in 0..255 with the given period (integers)"""
    u = t % period
    half = period // 2
    return (torch.where(u < half, u, period - u) * 255) // half


def frame(seed: int, width: int, height: int, device="cpu") -> torch.Tensor:
    """-> uint8 tensor (height, width, 3)"""
    dev = torch.device(device)
    y = torch.arange(height, dtype=torch.int64, device=dev).view(-1, 1)
    x = torch.arange(width, dtype=torch.int64, device=dev).view(1, -1)
    out = []
    for c in range(3):
        acc = torch.zeros((height, width), dtype=torch.int64, device=dev)
        for k in range(3):
            h = int(_hash32(torch.tensor(seed * 977 + c * 131 + k * 17 + 1)).item())
            fx, fy = 1 + (h & 7), 1 + ((h >> 3) & 7)
            period = 256 << ((h >> 6) & 3)
            phase = (h >> 8) & 1023
            amp = 20 + ((h >> 18) & 31)
            acc += (_tri(fx * x + fy * y + phase, period) - 128) * amp // 64
        n = _hash32(x * 73856093 + y * 19349663 + (seed * 3 + c) * 83492791)
        noise = (n & 255) + ((n >> 8) & 255) + ((n >> 16) & 255) + ((n >> 24) & 255) - 510  # sigma ~ 148
        acc += 128 + (noise * 12) // 148
        for r in range(4):
            h = int(_hash32(torch.tensor(seed * 31 + r * 7 + 5)).item())
            x0, y0 = (h & 0xFFF) % max(width - 8, 1), ((h >> 12) & 0xFFF) % max(height - 8, 1)
            w, hh = 8 + ((h >> 24) & 0x7F) * width // 512, 8 + ((h >> 17) & 0x7F) * height // 512
            level = ((h >> 5) * (c + 3)) & 255
            acc[y0:y0 + hh, x0:x0 + w] = level
        out.append(acc.clamp_(0, 255).to(torch.uint8))
    return torch.stack(out, dim=-1).contiguous()
