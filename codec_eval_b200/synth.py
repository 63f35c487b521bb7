"""Deterministic synthetic image pairs for tests and benchmarks.

G(seed, W, H): reference generator; J(q, ss): Pillow JPEG encode->decode distortion.
A GPU-free, Pillow-free distortion `cheap_distort` is used where the JPEG codec is
too slow to generate (large benchmark batches).
"""
from __future__ import annotations

import io

import numpy as np


def G(seed: int, w: int, h: int) -> np.ndarray:
    """v[y,x,c] = 128 + 90 sin(x/23 + c) cos(y/31 - c) + 40 ((x//32 + y//32) & 1) + N(0,6); uint8 [h,w,3]."""
    rng = np.random.default_rng(seed)
    y, x, c = np.meshgrid(np.arange(h), np.arange(w), np.arange(3), indexing="ij")
    v = 128.0 + 90.0 * np.sin(x / 23.0 + c) * np.cos(y / 31.0 - c) + 40.0 * (((x // 32) + (y // 32)) & 1)
    v = v + rng.normal(0.0, 6.0, size=(h, w, 3))
    return np.clip(v, 0, 255).astype(np.uint8)


def G_many(seeds, w: int, h: int) -> np.ndarray:
    """[len(seeds), h, w, 3] uint8, image k bit-identical to G(seeds[k], w, h): the deterministic pattern is formed once,
    only the noise differs per seed."""
    y, x, c = np.meshgrid(np.arange(h), np.arange(w), np.arange(3), indexing="ij")
    base = 128.0 + 90.0 * np.sin(x / 23.0 + c) * np.cos(y / 31.0 - c) + 40.0 * (((x // 32) + (y // 32)) & 1)
    out = np.empty((len(seeds), h, w, 3), np.uint8)
    for k, seed in enumerate(seeds):
        v = base + np.random.default_rng(int(seed)).normal(0.0, 6.0, size=(h, w, 3))
        out[k] = np.clip(v, 0, 255).astype(np.uint8)
    return out


def J(img: np.ndarray, quality: int, subsampling: int = 2) -> np.ndarray:
    """Pillow JPEG round trip; subsampling 2 = 4:2:0, 0 = 4:4:4."""
    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(img, "RGB").save(buf, format="JPEG", quality=int(quality), subsampling=int(subsampling))
    buf.seek(0)
    return np.asarray(Image.open(buf).convert("RGB")).copy()


def cheap_distort(img: np.ndarray, strength: int, seed: int = 0) -> np.ndarray:
    """Codec-free distortion: 8x8 block mean blending + quantisation noise, deterministic.

    strength 0..100 (100 = nearly lossless), loosely mimicking a JPEG quality knob.
    """
    h, w, _ = img.shape
    f = img.astype(np.float32)
    hb, wb = (h // 8) * 8, (w // 8) * 8
    blk = f[:hb, :wb].reshape(hb // 8, 8, wb // 8, 8, 3).mean(axis=(1, 3), keepdims=True)
    blk = np.broadcast_to(blk, (hb // 8, 8, wb // 8, 8, 3)).reshape(hb, wb, 3)
    a = (100 - strength) / 250.0
    out = f.copy()
    out[:hb, :wb] = (1 - a) * f[:hb, :wb] + a * blk
    step = 1.0 + (100 - strength) / 6.0
    out = np.round(out / step) * step
    rng = np.random.default_rng(seed + 7919)
    out = out + rng.normal(0, (100 - strength) / 40.0, size=out.shape)
    return np.clip(out, 0, 255).astype(np.uint8)
