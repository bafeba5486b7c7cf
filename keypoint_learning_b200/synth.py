"""Deterministic synthetic clouds for the benchmark configurations of BASELINE.json / SURVEY.md 8d.

No datasets can be downloaded, and the reference ships only three ~64 k-point views, so the 1 M-point
2.5D view (config 3), the 10 M-point scene (config 4) and the batch views (config 5) are generated.
All coordinates are in millimetres with the same point spacing (~0.64 mm) as the bundled views, so
radiusFeatures = 20 gives the same ~2.4-3 k neighbours per point.  numpy only.
"""
from __future__ import annotations

import numpy as np


def view_25d(width_px=1250, height_px=800, pitch=0.64, seed=1234, dtype=np.float32):
    """Unorganised list of an orthographic range image: smooth relief + bumps + sensor noise.
    Returns (xyz[n,3] float32, viewpoint).  Centred on the origin (PCL's un-centred FP32 covariance
    degrades far from it)."""
    rng = np.random.default_rng(seed)
    ix, iy = np.meshgrid(np.arange(width_px, dtype=np.float64), np.arange(height_px, dtype=np.float64), indexing="xy")
    x = (ix - (width_px - 1) / 2.0) * pitch + rng.uniform(-0.1, 0.1, ix.shape) * pitch
    y = (iy - (height_px - 1) / 2.0) * pitch + rng.uniform(-0.1, 0.1, iy.shape) * pitch
    z = np.zeros_like(x)
    ext = max(width_px, height_px) * pitch
    for _ in range(4):
        a = rng.uniform(2.0, 8.0)
        f, g = 2 * np.pi / rng.uniform(40.0, 200.0, 2)
        ph, ps = rng.uniform(0, 2 * np.pi, 2)
        z += a * np.sin(f * x + ph) * np.cos(g * y + ps)
    for _ in range(16):
        cx, cy = rng.uniform(-0.5, 0.5, 2) * np.array([width_px, height_px]) * pitch
        sg = rng.uniform(5.0, 25.0)
        hgt = rng.uniform(5.0, 20.0) * rng.choice([-1.0, 1.0])
        z += hgt * np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * sg * sg))
    z += rng.normal(0.0, 0.05, z.shape)
    xyz = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1).astype(dtype)
    del ext
    return np.ascontiguousarray(xyz), (0.0, 0.0, 1000.0)


def _sphere(rng, n, r):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v * r


def _ellipsoid(rng, n, a, b, c):
    # area-weighted rejection from the sphere parametrisation
    out = np.empty((0, 3))
    gmax = max(a * b, b * c, a * c)
    while len(out) < n:
        v = rng.normal(size=(2 * (n - len(out)) + 64, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        g = np.sqrt((b * c * v[:, 0]) ** 2 + (a * c * v[:, 1]) ** 2 + (a * b * v[:, 2]) ** 2)
        keep = rng.uniform(0, gmax, len(v)) < g
        out = np.concatenate([out, v[keep] * np.array([a, b, c])])
    return out[:n]


def _torus(rng, n, R, r):
    th = np.empty(0)
    while len(th) < n:
        t = rng.uniform(0, 2 * np.pi, 2 * (n - len(th)) + 64)
        keep = rng.uniform(0, R + r, len(t)) < (R + r * np.cos(t))
        th = np.concatenate([th, t[keep]])
    th = th[:n]
    ph = rng.uniform(0, 2 * np.pi, n)
    return np.stack([(R + r * np.cos(th)) * np.cos(ph), (R + r * np.cos(th)) * np.sin(ph), r * np.sin(th)], axis=1)


def _rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def scene_closed_surfaces(n_points=10_000_000, seed=4321, n_shapes=32, dtype=np.float32):
    """n_shapes closed surfaces (spheres, ellipsoids, tori; 100-300 mm) in a 2 x 2 x 1 m box centred on
    the origin, sampled area-uniformly; exactly n_points points, shuffled.  Returns (xyz, viewpoint).
    The surface sizes are scaled so that the mean spacing stays ~0.64 mm for any n_points."""
    rng = np.random.default_rng(seed)
    gx, gy, gz = 4, 4, 2
    assert n_shapes <= gx * gy * gz
    cells = [(i, j, k) for k in range(gz) for j in range(gy) for i in range(gx)][:n_shapes]
    shapes = []
    for (i, j, k) in cells:
        kind = rng.integers(0, 3)
        size = rng.uniform(100.0, 300.0)  # overall diameter
        if kind == 0:
            r = size / 2; area = 4 * np.pi * r * r; prm = (r,)
        elif kind == 1:
            a = size / 2; b = a * rng.uniform(0.5, 1.0); c = a * rng.uniform(0.5, 1.0)
            p = 1.6075
            area = 4 * np.pi * (((a * b) ** p + (a * c) ** p + (b * c) ** p) / 3) ** (1 / p); prm = (a, b, c)
        else:
            r = size / 2 * rng.uniform(0.2, 0.4); R = size / 2 - r; area = 4 * np.pi ** 2 * R * r; prm = (R, r)
        ctr = np.array([(i + 0.5) * 500.0 - 1000.0, (j + 0.5) * 500.0 - 1000.0, (k + 0.5) * 500.0 - 500.0]) + rng.uniform(-60, 60, 3)
        shapes.append((kind, prm, area, ctr, _rot(rng)))
    areas = np.array([s[2] for s in shapes])
    # scale all shapes so that total area = n_points * 0.64^2 * 1.05 (random sampling ~ jittered lattice density)
    scale = np.sqrt(n_points * 0.64 * 0.64 / areas.sum())
    scale = min(scale, 1.0) if n_points >= 5_000_000 else scale
    counts = np.floor(areas / areas.sum() * n_points).astype(np.int64)
    counts[0] += n_points - counts.sum()
    parts = []
    for (kind, prm, _a, ctr, Rm), cnt in zip(shapes, counts):
        if kind == 0: p = _sphere(rng, cnt, prm[0] * scale)
        elif kind == 1: p = _ellipsoid(rng, cnt, *(v * scale for v in prm))
        else: p = _torus(rng, cnt, prm[0] * scale, prm[1] * scale)
        parts.append((p @ Rm.T + ctr).astype(dtype))
    xyz = np.concatenate(parts)
    xyz = xyz[rng.permutation(len(xyz))]
    return np.ascontiguousarray(xyz), (0.0, 0.0, 0.0)


def cube_crop(xyz, m):
    """The m points nearest (Chebyshev distance) to the median point of the cloud, in their original order: a bounded
    sample with the density and geometry of the full workload (bench.py's CPU legs and parity check, tests)."""
    if m >= len(xyz):
        return np.ascontiguousarray(xyz)
    c = np.median(xyz, axis=0)
    d = np.abs(xyz - c).max(axis=1)
    sel = np.argpartition(d, m)[:m]
    return np.ascontiguousarray(xyz[np.sort(sel)])


def organized_range_image(width=320, height=240, focal=525.0, seed=7, holes=True, dtype=np.float32):
    """An ORGANIZED cloud (height, width, 3) as a depth camera delivers it: pinhole projection of a smooth relief with two
    depth steps, sensor noise, NaN holes (no measurement) -- the input of the reference's organized branch
    (impl/KeypointLearning.hpp:138-145).  Depth ~0.5-0.8 m in millimetres."""
    rng = np.random.default_rng(seed)
    v, u = np.mgrid[0:height, 0:width].astype(np.float64)
    z = 620.0 + 35.0 * np.sin(u / 23.0 + 0.3) * np.cos(v / 17.0) + 20.0 * np.cos(u / 9.0) + rng.normal(0.0, 0.15, u.shape)
    z[:, (2 * width) // 3:] += 70.0                               # a vertical depth step
    z[height // 2:, : width // 4] -= 55.0                         # and a second one
    for _ in range(6):
        cx, cy, s = rng.uniform(0, width), rng.uniform(0, height), rng.uniform(8, 25)
        z += rng.uniform(-25, 25) * np.exp(-((u - cx) ** 2 + (v - cy) ** 2) / (2 * s * s))
    x = (u - (width - 1) / 2.0) * z / focal
    y = (v - (height - 1) / 2.0) * z / focal
    xyz = np.stack([x, y, z], axis=2).astype(dtype)
    if holes:
        for _ in range(5):
            r0, c0 = int(rng.integers(0, max(1, height - 12))), int(rng.integers(0, max(1, width - 20)))
            xyz[r0:r0 + int(rng.integers(2, 12)), c0:c0 + int(rng.integers(2, 20))] = np.nan
        xyz[rng.uniform(size=(height, width)) < 0.002] = np.nan      # isolated drop-outs
    return np.ascontiguousarray(xyz), (0.0, 0.0, 0.0)


def small_patch(n_side=64, pitch=0.64, seed=0):
    """Tiny 2.5D patch for unit tests (n_side^2 points)."""
    return view_25d(n_side, n_side, pitch, seed)
