"""Integer-only synthetic frames: identical bytes from numpy on the host and torch on the GPU.

noise: v(x, y, plane, frame) = fmix32((x + y*W + plane*W*H + frame*0x9E3779B9) mod 2^32) >> 24
       (murmur3 finaliser) -- adversarial for parity.
scene: three integer triangle waves at different scales plus 1/8-amplitude noise -- smoother content for
       throughput runs (content does not change the kernels' work).
"""
from __future__ import annotations

import numpy as np

_M32 = 0xFFFFFFFF


def _fmix32_np(h):
    h = h.astype(np.uint64)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(_M32)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(_M32)
    h ^= h >> np.uint64(16)
    return h


def noise_plane(w: int, h: int, plane: int = 0, frame: int = 0) -> np.ndarray:
    idx = np.arange(w * h, dtype=np.uint64)
    idx = (idx + np.uint64((plane * w * h) & _M32) + np.uint64((frame * 0x9E3779B9) & _M32)) & np.uint64(_M32)
    return (_fmix32_np(idx) >> np.uint64(24)).astype(np.uint8).reshape(h, w)


def _tri(v, period):
    t = v % (2 * period)
    return np.where(t < period, t, 2 * period - t) * 255 // period


def scene_plane(w: int, h: int, plane: int = 0, frame: int = 0) -> np.ndarray:
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    s = (_tri(x + 3 * frame, max(w // 7, 2)) + _tri(y + 2 * x // 3 + 5 * plane, max(h // 5, 2)) + _tri(x + y, 61)) // 3
    n = noise_plane(w, h, plane, frame).astype(np.int64) >> 3
    return np.clip(s * 7 // 8 + n, 0, 255).astype(np.uint8)


def noise_plane_torch(w: int, h: int, plane: int = 0, frame: int = 0, device="cuda", pitch: int | None = None):
    """Same bytes as noise_plane(), generated on `device`; returns a (h, pitch) uint8 tensor (columns >= w are 0)."""
    import torch
    pitch = pitch or w
    idx = torch.arange(w * h, dtype=torch.int64, device=device)
    idx = (idx + ((plane * w * h) & _M32) + ((frame * 0x9E3779B9) & _M32)) & _M32
    hsh = idx
    hsh = hsh ^ (hsh >> 16)
    hsh = (hsh * 0x85EBCA6B) & _M32
    hsh = hsh ^ (hsh >> 13)
    hsh = (hsh * 0xC2B2AE35) & _M32
    hsh = hsh ^ (hsh >> 16)
    v = (hsh >> 24).to(torch.uint8).view(h, w)
    if pitch == w:
        return v
    out = torch.zeros((h, pitch), dtype=torch.uint8, device=device)
    out[:, :w] = v
    return out
