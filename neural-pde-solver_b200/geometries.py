"""Synthetic geometries and sources that feed the hot path -- host-side mirror of
utils/geometries.py:4-162 and utils/misc.py:80-97 (`gaussian`).

These are the generators BASELINE's configs name ("L-shape and cylinder geometries",
"synthetic sources"); they run once per problem on the host in numpy, draw from
``np.random`` in the same order as the reference (so a seeded run yields the same
problems), and return float64 arrays exactly like the reference.  The filled discs
the reference draws with ``cv2.circle(..., -1)`` are rasterised here by the same
midpoint rule, so no OpenCV is needed (tests/test_pipeline.py checks the rasteriser
against cv2 when it is importable).
"""
import numpy as np


def filled_disc(image_size, cx, cy, radius):
    """(image_size, image_size) float64 array, 1 inside the disc centred on column cx / row cy.

    Midpoint circle with horizontal span filling: the rule behind cv2.circle(img, (cx, cy),
    radius, 255, -1) as used by utils/geometries.py:36-37,43,67-68,72."""
    img = np.zeros((image_size, image_size))

    def span(row, lo, hi):
        if 0 <= row < image_size:
            lo, hi = max(lo, 0), min(hi, image_size - 1)
            if lo <= hi:
                img[row, lo:hi + 1] = 1

    err, dx, dy = 0, int(radius), 0
    plus, minus = 1, 2 * int(radius) - 1
    while dx >= dy:
        span(cy - dy, cx - dx, cx + dx)
        span(cy + dy, cx - dx, cx + dx)
        span(cy - dx, cx - dy, cx + dy)
        span(cy + dx, cx - dy, cx + dy)
        dy += 1
        err += plus
        plus += 2
        if err > 0:
            err -= minus
            dx -= 1
            minus -= 2
    return img


def _cylinder_radii(image_size):
    c = (image_size + 1) // 2
    r1 = (image_size - 7) // 2
    return c, r1, int(r1 / 4)


def cylinders(image_size):
    """Outer cylinder wall (two temperatures, upper / lower half) and a randomly placed hot inner
    cylinder -- utils/geometries.py:26-60.  Draw order: centre (2 ints), v1, v2."""
    c, r1, r2 = _cylinder_radii(image_size)
    outer = 1 - filled_disc(image_size, c, c, r1)
    center = np.random.randint(-r2 * 1, r2 * 1 + 1, size=2) + c
    inner = filled_disc(image_size, int(center[0]), int(center[1]), r2)
    bc_mask = inner + outer
    assert np.all(np.logical_or(bc_mask == 0, bc_mask == 1))

    half = (image_size + 1) // 2
    v1 = np.random.uniform(0.1, 0.3)
    v2 = np.random.uniform(0.4, 0.6)
    outer[:half, :] = outer[:half, :] * v1
    outer[half:, :] = outer[half:, :] * v2
    bc = outer + inner * 1.0
    x = bc + (1 - bc_mask) * (v1 + v2 + 1) / 3
    return x, bc, bc_mask


def centered_cylinders(image_size):
    """utils/geometries.py:62-87: both cylinders centred, one wall temperature."""
    c, r1, r2 = _cylinder_radii(image_size)
    outer = 1 - filled_disc(image_size, c, c, r1)
    inner = filled_disc(image_size, c, c, r2)
    bc_mask = inner + outer
    assert np.all(np.logical_or(bc_mask == 0, bc_mask == 1))
    v = np.random.uniform(0.2, 0.8)
    bc = outer * v + inner * 1.0
    x = bc + (1 - bc_mask) * (v + 1) / 3
    return x, bc, bc_mask


def _lshape_fields(image_size, rows, cols, temperatures):
    """Mask / values / initial field of an L-shaped domain whose removed corner is rows x cols
    (utils/geometries.py:98-122 and :136-160)."""
    bc_mask = np.zeros((image_size, image_size))
    bc_mask[:rows, :cols] = 1
    bc_mask[0, :] = bc_mask[-1, :] = 1
    bc_mask[:, 0] = bc_mask[:, -1] = 1

    bc_values = np.zeros((image_size, image_size))
    bc_values[0, :] = temperatures[0]
    bc_values[:, 0] = temperatures[1]
    bc_values[-1, :] = temperatures[2]
    bc_values[:, -1] = temperatures[3]
    # removed corner: below its diagonal (j / i < cols / rows, i != 0) the left wall's temperature,
    # on and above it the top wall's
    i = np.arange(rows, dtype=np.float64)[:, None]
    j = np.arange(cols, dtype=np.float64)[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        below = (i != 0) & (j / i < cols / rows)
    bc_values[:rows, :cols] = np.where(below, temperatures[1], temperatures[0])

    x = np.ones((image_size, image_size)) * temperatures.mean()
    x = x * (1 - bc_mask) + bc_values
    return x, bc_values, bc_mask


def _wall_temperatures():
    return np.concatenate([np.random.uniform(0.5, 1, size=2), np.random.uniform(0, 0.5, size=2)])


def Lshape(image_size):
    """utils/geometries.py:89-130.  Draw order: corner rows, corner cols, 2 + 2 temperatures, rotation."""
    rows = np.random.randint(image_size // 8 * 3, image_size // 2)
    cols = np.random.randint(image_size // 8 * 3, image_size // 2)
    x, bc_values, bc_mask = _lshape_fields(image_size, rows, cols, _wall_temperatures())
    k = np.random.randint(4)
    if k > 0:
        x, bc_values, bc_mask = (np.rot90(a, k).copy() for a in (x, bc_values, bc_mask))
    return x, bc_values, bc_mask


def centered_Lshape(image_size):
    """utils/geometries.py:132-162: the corner is the upper-left quadrant, no rotation."""
    half = (image_size + 1) // 2
    return _lshape_fields(image_size, half, half, _wall_temperatures())


_GEOMETRIES = {"cylinders": cylinders, "Lshape": Lshape,
               "centered_cylinders": centered_cylinders, "centered_Lshape": centered_Lshape}


def get_geometry(geometry, image_size, batch_size, max_temp):
    """utils/geometries.py:4-24: (x, bc_values, bc_mask), each (batch_size, N, N) float64; x and
    bc_values scaled by max_temp."""
    if geometry not in _GEOMETRIES:
        raise NotImplementedError
    samples = [_GEOMETRIES[geometry](image_size) for _ in range(batch_size)]
    x = np.stack([s[0] for s in samples], axis=0) * max_temp
    bc = np.stack([s[1] for s in samples], axis=0) * max_temp
    bc_mask = np.stack([s[2] for s in samples], axis=0)
    return x, bc, bc_mask


def gaussian(image_size):
    """utils/misc.py:80-97: separable Gaussian bump of random centre and width, peak <= 1.
    Draw order: x_mean, y_mean, s."""
    width = image_size // 2
    x_mean = np.random.randint(-width // 2, width // 2 + 1)
    y_mean = np.random.randint(-width // 2, width // 2 + 1)
    axis = np.linspace(-width, width, num=image_size, endpoint=True)
    sd = width / np.random.uniform(2, 3)
    gx = np.exp(-((axis + x_mean) ** 2) / 2 / (sd ** 2))
    gy = np.exp(-((axis + y_mean) ** 2) / 2 / (sd ** 2))
    return gx[:, np.newaxis] * gy[np.newaxis, :]
