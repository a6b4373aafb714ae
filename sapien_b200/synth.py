"""Deterministic synthetic IR stereo pairs and camera presets.

Everything here is integer arithmetic on top of numpy's PCG64 stream, so the same seed gives
bit-identical images in this container and on the GPU box (no float blur whose rounding could
differ between hosts).  Used by tests, bench.py and smoke(); not part of the compute path.
"""
from __future__ import annotations

import numpy as np


def _blur121(a: np.ndarray) -> np.ndarray:
    """Separable [1,2,1]/4 blur, twice (~ sigma 1 px), edge-replicated, integer rounding."""
    a = a.astype(np.int32)
    for _ in range(2):
        p = np.pad(a, ((0, 0), (1, 1)), mode="edge")
        a = (p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:] + 2) >> 2
        p = np.pad(a, ((1, 1), (0, 0)), mode="edge")
        a = (p[:-2, :] + 2 * p[1:-1, :] + p[2:, :] + 2) >> 2
    return a


def make_disparity(h: int, w: int, max_disp: int, rng: np.random.Generator) -> np.ndarray:
    """Ground-truth integer disparity: background plane, 5 boxes, one horizontal ramp."""
    d = np.full((h, w), int(0.15 * max_disp), dtype=np.int32)
    for _ in range(5):
        bh = int(rng.integers(max(h // 8, 2), max(h // 3, 3)))
        bw = int(rng.integers(max(w // 8, 2), max(w // 3, 3)))
        y0 = int(rng.integers(0, h - bh + 1))
        x0 = int(rng.integers(0, w - bw + 1))
        d[y0 : y0 + bh, x0 : x0 + bw] = int(rng.integers(int(0.2 * max_disp), int(0.8 * max_disp) + 1))
    ry0 = int(rng.integers(0, max(h - h // 6, 1)))
    ramp = (np.arange(w, dtype=np.int64) * int(0.5 * max_disp)) // max(w - 1, 1) + int(0.2 * max_disp)
    d[ry0 : ry0 + max(h // 6, 1), :] = ramp[None, :].astype(np.int32)
    return d


def make_pair(h: int, w: int, max_disp: int, seed: int = 0):
    """Returns (left u8 [h,w], right u8 [h,w], disparity int32 [h,w]).

    Left = blurred white noise blended 50/50 with a 4-px-cell random checker (dense texture);
    right(y,x) = left(y, x + d(y,x)) sampled with the disparity of the left pixel it lands on,
    occlusion holes filled with fresh noise.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    noise = rng.integers(0, 256, size=(h, w), dtype=np.int32)
    cells = rng.integers(0, 256, size=((h + 3) // 4, (w + 3) // 4), dtype=np.int32)
    checker = np.repeat(np.repeat(cells, 4, axis=0), 4, axis=1)[:h, :w]
    left = ((_blur121(noise) + checker + 1) >> 1).clip(0, 255).astype(np.uint8)
    disp = make_disparity(h, w, max_disp, rng)
    # forward-warp left -> right: pixel (y,x) of the left image appears at x-d in the right one;
    # nearer surfaces (larger d) win.
    right = rng.integers(0, 256, size=(h, w), dtype=np.int32).astype(np.uint8)
    best = np.full((h, w), -1, dtype=np.int32)
    ys, xs = np.mgrid[0:h, 0:w]
    xr = xs - disp
    ok = xr >= 0
    order = np.argsort(disp[ok], kind="stable")  # far first, near last -> near overwrites
    yy, xx, xt = ys[ok][order], xs[ok][order], xr[ok][order]
    right[yy, xt] = left[yy, xx]
    best[yy, xt] = disp[yy, xx]
    return left, right, disp


def to_rgba(img_u8: np.ndarray) -> np.ndarray:
    """u8 [.., h, w] -> float32 RGBA [.., h, w, 4] with trunc(R*255) == img exactly (core.cu:51)."""
    v = (img_u8.astype(np.float32) + np.float32(0.5)) / np.float32(255.0)
    out = np.empty(img_u8.shape + (4,), dtype=np.float32)
    out[..., 0] = v
    out[..., 1] = v
    out[..., 2] = v
    out[..., 3] = 1.0
    return out


def make_rgb(h: int, w: int, seed: int = 0) -> np.ndarray:
    """Synthetic float32 RGBA colour image [h,w,4] for the RGB point cloud."""
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    out = rng.random((h, w, 4), dtype=np.float32)
    out[..., 3] = 1.0
    return out
