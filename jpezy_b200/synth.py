"""Synthetic planar RGB images -- numpy twin of csrc/synth.cuh.

Pure 32-bit integer arithmetic so that host and device produce identical pixels without
shipping gigabytes over PCIe.  family: 0 = S-photo, 1 = S-noise, 2 = adversarial flat/ramps.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFF)


def _u32(x):
    return (x & _M).astype(np.uint64)


def _hash(frame, c, idx):
    h = np.uint64(0x6A70657A ^ ((frame * 0x9E3779B1) & 0xFFFFFFFF) ^ ((c * 0x85EBCA77) & 0xFFFFFFFF))
    h = _u32(h + _u32(idx * np.uint64(0xC2B2AE3D)))
    h ^= h >> np.uint64(16)
    h = _u32(h * np.uint64(0x85EBCA6B))
    h ^= h >> np.uint64(13)
    h = _u32(h * np.uint64(0xC2B2AE35))
    h ^= h >> np.uint64(16)
    return h


def _tri(v, P):
    d = (v % P).astype(np.int64) - P // 2
    return np.abs(d)


def plane(family, W, H, frame, c):
    y, x = np.mgrid[0:H, 0:W].astype(np.uint64)
    idx = _u32(y * np.uint64(W) + x)
    if family == 0:
        P1, P2 = 256 + 64 * c, 192 + 32 * c
        base = 20 + (_tri(_u32(x + np.uint64(37 * c + 5 * frame)), P1) * 220) // P1 + \
            (_tri(_u32(y + np.uint64(91 * c + 3 * frame)), P2) * 180) // P2
        v = base + ((_hash(frame, c, idx) >> np.uint64(24)) % np.uint64(17)).astype(np.int64) - 8
    elif family == 1:
        v = 128 + ((_hash(frame, c, idx) >> np.uint64(24)) % np.uint64(255)).astype(np.int64) - 127
    else:
        left = (_u32(((x >> np.uint64(4)) + np.uint64(3) * (y >> np.uint64(4)) + np.uint64(frame)) * np.uint64(7)) & np.uint64(255))
        right = (_u32(x * np.uint64(c + 1) + y) & np.uint64(255))
        v = np.where(x < np.uint64(W // 2), left, right).astype(np.int64)
    return np.clip(v, 0, 255).astype(np.uint8)


def image(family, W, H, frame=0):
    """-> (r, g, b) uint8 arrays of shape (H, W)."""
    return tuple(plane(family, W, H, frame, c) for c in range(3))
