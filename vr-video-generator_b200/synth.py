"""Seeded synthetic inputs for tests, golden fixtures and bench.py.

Everything here is bit-reproducible on any host: integer RNG streams plus IEEE double
add/mul/div only (no libm transcendentals, no torch CPU kernels), then one correctly rounded
narrowing to fp16.  Random-init model depth is deliberately NOT used for the warp: it collapses
to limit_step 0/1.
"""
import numpy as np

DPT_H, DPT_W = 518, 924          # Depth-Anything-V2 output for a 16:9 input (dpt.py:205-214)


def frames_noise(n, H, W, seed=0):
    """Incompressible uint8 RGB frames [n,H,W,3]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, size=(n, H, W, 3), dtype=np.uint8)


def frames_gradient(n, H, W, seed=0):
    """Smooth, compressible frames (diagonal ramps + seeded rectangles) so shifts are visible."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    y = np.arange(H, dtype=np.int64)[:, None]
    x = np.arange(W, dtype=np.int64)[None, :]
    out = np.empty((n, H, W, 3), dtype=np.uint8)
    for t in range(n):
        r = (x * 255 // max(W - 1, 1) + 3 * t) % 256
        g = (y * 255 // max(H - 1, 1) + 0 * x) % 256
        b = ((x + y) * 2 + 5 * t) % 256
        f = np.stack([r + 0 * y, g, b], axis=-1).astype(np.uint8)
        for _ in range(6):
            y0, x0 = int(rng.integers(0, H)), int(rng.integers(0, W))
            hh, ww = int(rng.integers(H // 16 + 1, H // 4 + 2)), int(rng.integers(W // 16 + 1, W // 4 + 2))
            f[y0:y0 + hh, x0:x0 + ww] = rng.integers(0, 256, size=3, dtype=np.uint8)
        out[t] = f
    return out


def _catmull_rom(grid, factor):
    """Upsample a 2-D float64 grid by an integer factor with Catmull-Rom cubics (overshoots like
    bicubic does).  Polynomial arithmetic only."""
    def weights(f):
        t = np.arange(f, dtype=np.float64) / f
        return np.stack([((-t + 2) * t - 1) * t / 2, ((3 * t - 5) * t * t + 2) / 2,
                         ((-3 * t + 4) * t + 1) * t / 2, (t - 1) * t * t / 2], axis=0)   # [4,f]

    def along(a, axis):
        a = np.moveaxis(a, axis, 0)
        n = a.shape[0] - 3
        w = weights(factor)
        out = np.zeros((n, factor) + a.shape[1:], dtype=np.float64)
        for k in range(4):
            out = out + a[k:k + n][:, None] * w[k].reshape((1, factor) + (1,) * (a.ndim - 1))
        return np.moveaxis(out.reshape((n * factor,) + a.shape[1:]), 0, axis)

    return along(along(grid, 0), 1)


def depth_stress(n, h=DPT_H, w=DPT_W, seed=0, peak=13.9, drift=2):
    """D-stress: smooth random field, ReLU-clamped, translated `drift` px per frame.  [n,h,w] fp16."""
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    f = 14
    gh, gw = -(-h // f) + 3, -(-(w + drift * n) // f) + 3
    grid = rng.integers(0, int(round(peak * 100)) + 1, size=(gh, gw)).astype(np.float64) / 100.0
    field = np.maximum(_catmull_rom(grid, f), 0.0)
    return np.ascontiguousarray(np.stack([field[:h, drift * t:drift * t + w] for t in range(n)]), dtype=np.float16)


def depth_scene(n, h=DPT_H, w=DPT_W, seed=0, peak=13.9, drift=2):
    """D-scene: vertical ramp 1 -> 7 (far at the top) plus six seeded ellipses at 7.5 .. peak."""
    rng = np.random.Generator(np.random.PCG64(seed + 1299709))
    wide = w + drift * n
    y = np.arange(h, dtype=np.float64)[:, None]
    x = np.arange(wide, dtype=np.float64)[None, :]
    field = 1.0 + 6.0 * y / max(h - 1, 1) + 0.0 * x
    levels = [7.5 + (peak - 7.5) * i / 5 for i in range(6)]
    for lev in levels:
        cy, cx = float(rng.integers(h // 8, 7 * h // 8)), float(rng.integers(wide // 8, 7 * wide // 8))
        ry, rx = float(rng.integers(h // 12 + 1, h // 4 + 2)), float(rng.integers(wide // 16 + 1, wide // 5 + 2))
        inside = ((y - cy) / ry) ** 2 + ((x - cx) / rx) ** 2 <= 1.0
        field = np.where(inside, lev, field)
    return np.ascontiguousarray(np.stack([field[:, drift * t:drift * t + w] for t in range(n)]), dtype=np.float16)


def depth_lowres(kind, n, h=DPT_H, w=DPT_W, seed=0, peak=13.9):
    if kind == "scene":
        return depth_scene(n, h, w, seed, peak)
    if kind == "stress":
        return depth_stress(n, h, w, seed, peak)
    raise ValueError(f"unknown depth kind {kind!r}")
