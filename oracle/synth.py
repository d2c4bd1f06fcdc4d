"""Seeded synthetic inputs shared by tests and bench (pure numpy; SURVEY.md §8(d) "Synthetic inputs").

Frames are NOT white noise (flat SuperPoint response): a large canvas of random overlapping rectangles and
line-ish slabs with random gray levels, separable Gaussian blur (sigma 1), N(0,2^2) noise, clipped to u8.
A stream is a slowly translating window over the canvas, so consecutive frames overlap and the trajectory
revisits earlier views (loop closures).  Not part of the product path.
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20240701


def _blur1(a: np.ndarray) -> np.ndarray:
    k = np.exp(-0.5 * (np.arange(-3, 4) ** 2)).astype(np.float32)
    k /= k.sum()
    p = np.pad(a, ((0, 0), (3, 3)), mode="edge")
    out = np.zeros_like(a)
    for i in range(7):
        out += k[i] * p[:, i:i + a.shape[1]]
    p = np.pad(out, ((3, 3), (0, 0)), mode="edge")
    out2 = np.zeros_like(a)
    for i in range(7):
        out2 += k[i] * p[i:i + a.shape[0], :]
    return out2


def make_canvas(h: int, w: int, seed: int = BASE_SEED, density: float = 1.0) -> np.ndarray:
    """float32 canvas in [0,255] with many corners."""
    rng = np.random.default_rng(seed)
    c = np.full((h, w), 110.0, dtype=np.float32)
    n = int(density * 400 * (h * w) / (480 * 752))
    for _ in range(n):
        rh = int(rng.integers(6, 90)); rw = int(rng.integers(6, 90))
        if rng.random() < 0.3:       # thin slab = line segment
            if rng.random() < 0.5:
                rh = int(rng.integers(2, 5))
            else:
                rw = int(rng.integers(2, 5))
        y = int(rng.integers(-20, h)); x = int(rng.integers(-20, w))
        g = float(rng.integers(10, 246))
        c[max(y, 0):max(y + rh, 0), max(x, 0):max(x + rw, 0)] = g
    return _blur1(c)


def frame_from_canvas(canvas: np.ndarray, y0: int, x0: int, h: int, w: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    f = canvas[y0:y0 + h, x0:x0 + w] + rng.normal(0.0, 2.0, (h, w)).astype(np.float32)
    return np.clip(np.floor(f + 0.5), 0, 255).astype(np.uint8)


def make_frame(h: int = 480, w: int = 752, seed: int = BASE_SEED) -> np.ndarray:
    return frame_from_canvas(make_canvas(h, w, seed), 0, 0, h, w, seed + 7)


def make_pair(h: int = 480, w: int = 752, seed: int = BASE_SEED, shift=(6, 9)):
    """Frame A and frame B = the same scene translated by `shift` (dy,dx) px + fresh noise."""
    cv = make_canvas(h + 64, w + 64, seed)
    a = frame_from_canvas(cv, 16, 16, h, w, seed + 11)
    b = frame_from_canvas(cv, 16 + shift[0], 16 + shift[1], h, w, seed + 12)
    return a, b


class Stream:
    """EuRoC-shaped synthetic keyframe stream: window sliding along a closed path over a canvas.

    frame(t) is deterministic in (seed, t); the path returns to its start after `period` frames so
    frame t and t+period see the same scene (loop closure candidates)."""

    def __init__(self, h: int = 480, w: int = 752, seed: int = BASE_SEED, period: int = 600, margin: int = 400):
        self.h, self.w, self.seed, self.period, self.margin = h, w, seed, period, margin
        self.canvas = make_canvas(h + margin, w + margin, seed, density=1.0)

    def offset(self, t: int):
        ph = 2.0 * np.pi * (t % self.period) / self.period
        m = self.margin / 2.0
        return int(round(m + (m - 1) * np.sin(ph))), int(round(m + (m - 1) * np.cos(ph)))

    def frame(self, t: int) -> np.ndarray:
        y0, x0 = self.offset(t)
        return frame_from_canvas(self.canvas, y0, x0, self.h, self.w, self.seed + 1000 + t)


def vio_points(n: int, h: int, w: int, seed: int, min_dist: float = 30.0) -> np.ndarray:
    """n sub-pixel f32 (x,y) in [8,W-8]x[8,H-8] with >= min_dist spacing (KLT front end: euroc yaml:32-33)."""
    rng = np.random.default_rng(seed)
    pts = []
    tries = 0
    while len(pts) < n and tries < 200000:
        tries += 1
        p = np.array([rng.uniform(8, w - 8), rng.uniform(8, h - 8)], dtype=np.float32)
        if all((p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 >= min_dist ** 2 for q in pts):
            pts.append(p)
        if tries % 5000 == 0:
            min_dist *= 0.9
    return np.stack(pts).astype(np.float32)


def make_bank(n: int, d: int = 512, seed: int = BASE_SEED, dup_frac: float = 0.01):
    """Unit-norm f32 bank [n,d] + a query with planted near-duplicates so top-3 are well separated."""
    rng = np.random.default_rng(seed)
    bank = rng.standard_normal((n, d)).astype(np.float32)
    bank /= np.linalg.norm(bank, axis=1, keepdims=True)
    q = rng.standard_normal(d).astype(np.float32)
    q /= np.linalg.norm(q)
    ndup = max(3, int(n * dup_frac)) if n >= 3 else 0
    idx = rng.choice(n, size=min(ndup, n), replace=False) if n else np.array([], dtype=np.int64)
    for j, i in enumerate(idx):
        v = q + rng.normal(0, 0.02 + 0.01 * j, d).astype(np.float32)
        bank[i] = v / np.linalg.norm(v)
    return bank, q
