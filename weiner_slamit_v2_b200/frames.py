"""Seeded synthetic textured frames -- pure numpy, so the CPU oracle,
the GPU path, the tests and the bench all see byte-identical inputs on every machine.

frame(i) = 128 + three octaves of cubic-upsampled Gaussian noise (grid pitch 64/16/4 px,
amplitude 40/30/20), clipped to uint8, then W*H/600 random filled grey rectangles.  The
rectangles give FAST plenty of corners at every pyramid level; the noise gives the descriptors
texture.
"""
import numpy as np


def _cubic_matrix(n_out, n_in):
    """(n_out x n_in) Catmull-Rom interpolation matrix sampling n_in knots over n_out pixels."""
    pos = (np.arange(n_out) + 0.5) * (n_in - 1) / n_out
    i0 = np.floor(pos).astype(np.int64)
    t = pos - i0
    w = np.stack([((-t + 2) * t - 1) * t / 2, ((3 * t - 5) * t * t + 2) / 2,
                  ((-3 * t + 4) * t + 1) * t / 2, (t - 1) * t * t / 2], axis=1)
    m = np.zeros((n_out, n_in))
    for k in range(4):
        idx = np.clip(i0 + k - 1, 0, n_in - 1)
        np.add.at(m, (np.arange(n_out), idx), w[:, k])
    return m


_MATS = {}


def _mat(n_out, n_in):
    key = (n_out, n_in)
    if key not in _MATS:
        _MATS[key] = _cubic_matrix(n_out, n_in)
    return _MATS[key]


def synthetic_frame(index, width=640, height=480, seed_base=1000):
    rng = np.random.default_rng(seed_base + int(index))
    img = np.full((height, width), 128.0)
    for pitch, amp in ((64, 40.0), (16, 30.0), (4, 20.0)):
        gh, gw = height // pitch + 3, width // pitch + 3
        g = rng.standard_normal((gh, gw))
        img += amp * (_mat(height, gh) @ g @ _mat(width, gw).T)
    img = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    nrect = width * height // 600
    xs = rng.integers(0, width, nrect); ys = rng.integers(0, height, nrect)
    ws = rng.integers(4, 60, nrect); hs = rng.integers(4, 60, nrect)
    gs = rng.integers(0, 256, nrect)
    for x, y, w, h, g in zip(xs, ys, ws, hs, gs):
        img[y:y + h, x:x + w] = g
    return img


def synthetic_batch(first, count, width=640, height=480, seed_base=1000):
    return np.stack([synthetic_frame(first + i, width, height, seed_base) for i in range(count)])


def low_contrast_frame(index, width=640, height=480):
    """Exercises the minThFAST retry: corners whose contrast sits between 7 and 20 grey levels."""
    rng = np.random.default_rng(5000 + int(index))
    img = np.full((height, width), 100, np.uint8)
    n = width * height // 900
    xs = rng.integers(0, width, n); ys = rng.integers(0, height, n)
    ws = rng.integers(6, 50, n); hs = rng.integers(6, 50, n)
    gs = rng.integers(100, 116, n)
    for x, y, w, h, g in zip(xs, ys, ws, hs, gs):
        img[y:y + h, x:x + w] = g
    # a few strong corners so both thresholds are in play
    for k in range(12):
        x, y = rng.integers(30, width - 60), rng.integers(30, height - 60)
        img[y:y + 25, x:x + 25] = 220
    return img


def plateau_retry_frame(index, width=640, height=480):
    """Cells in which FAST(iniThFAST) finds corners but non-maximum suppression keeps none of them: the end of a 2-px wide
    bar is a plateau of equal scores (a corner must be STRICTLY greater than its 8 neighbours), so the reference's
    `if (vKeysCell.empty())` (ORBextractor.cc:829) fires although the cell has corners at 20, and the faint corner next to
    it (contrast 12) appears through the FAST(minThFAST) retry.  Other cells hold a strong corner, a faint one, or both."""
    rng = np.random.default_rng(7000 + int(index))
    img = np.full((height, width), 60, np.uint8)
    for j, cy in enumerate(range(40, height - 40, 48)):
        for i, cx in enumerate(range(40, width - 40, 48)):
            kind = (i + 2 * j + int(index)) % 4
            ox, oy = rng.integers(-4, 5, 2)
            x, y = cx + ox, cy + oy
            if kind in (0, 1):
                img[y - 14:y, x:x + 2] = 200                  # bar: plateau of equal FAST scores at its end
            if kind in (1, 2):
                img[y + 4:y + 14, x + 6:x + 16] = 180         # strong isolated corners
            if kind in (0, 2, 3):
                img[y + 4:y + 12, x - 14:x - 6] = 72          # faint corners: only FAST(7) sees them
    return img


def stereo_right_frame(left, index, max_disparity=48.0, band=40):
    """The right image of a rectified pair: every band of rows sees the left image shifted by its own (fractional)
    disparity, linearly interpolated, plus a little noise, so ComputeStereoMatches finds row-aligned matches, a cost
    minimum inside the +-5 column window and a non-trivial parabola."""
    rng = np.random.default_rng(9000 + int(index))
    h, w = left.shape
    L = left.astype(np.float64)
    right = np.empty_like(L)
    xs = np.arange(w)
    for y0 in range(0, h, band):
        d = rng.uniform(1.0, max_disparity)
        src = np.clip(xs + d, 0, w - 1)
        i0 = np.floor(src).astype(np.int64); f = src - i0
        i1 = np.minimum(i0 + 1, w - 1)
        right[y0:y0 + band] = L[y0:y0 + band][:, i0] * (1 - f) + L[y0:y0 + band][:, i1] * f
    right += rng.integers(-2, 3, right.shape)
    return np.clip(np.rint(right), 0, 255).astype(np.uint8)
