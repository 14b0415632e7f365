"""Deterministic synthetic point clouds for the BASELINE.json configs.

All clouds are float32 AoS (x, y, z, 1.0) — the 16-byte pcl::PointXYZ / PointCloud2 payload
layout (src/publisher.cpp:55).  Conventions shared by every generator:

  * point 0 is a designated interior point: the reference makes it the map origin and
    does not bin it (src/receiver.cpp:145,150);
  * points are emitted as a sensor sweeping along +y, then shuffled inside 64k-point
    windows (sensor locality without a sorted order);
  * 2 % exact (0,0,0) points are appended: every dataset tool of the reference resizes a
    W x H cloud and fills only part of it (src/test/genePcd.cpp:30-33, pose2pcd.cpp:143-146),
    so real inputs end in one pathologically heavy voxel;
  * seed = 20260000 + cfg.
"""
from dataclasses import dataclass

import numpy as np

WINDOW = 1 << 16


@dataclass
class CloudSpec:
    name: str
    n: int
    grid_len: float
    z_len: float
    slope_interval: float = 0.08
    demand: str = "slope"


CONFIGS = {
    "cfg1": CloudSpec("cfg1 ramp+step 1M", 1_000_000, 0.2, 0.1),
    "cfg2": CloudSpec("cfg2 multi-level 10M", 10_000_000, 0.2, 0.1),
    "cfg3": CloudSpec("cfg3 terrain 50M", 50_000_000, 0.1, 0.1),
    "cfg5": CloudSpec("cfg5 sweep 100M", 100_000_000, 0.2, 0.1),
}


def _rng(cfg: int, stream: int = 0):
    return np.random.Generator(np.random.Philox(key=20260000 + cfg, counter=[0, 0, 0, stream]))


def _window_shuffle(pts: np.ndarray, rng) -> np.ndarray:
    """Shuffle rows inside consecutive 64k windows, in place (row 0 stays first)."""
    n = pts.shape[0]
    for lo in range(1, n, WINDOW):
        hi = min(lo + WINDOW, n)
        pts[lo:hi] = pts[lo:hi][rng.permutation(hi - lo)]
    return pts


def _finish(parts, origin, rng, n_total, sort_by_y=True) -> np.ndarray:
    xyz = np.concatenate(parts, axis=0) if len(parts) > 1 else parts[0]
    if sort_by_y:
        xyz = xyz[np.argsort(xyz[:, 1], kind="stable")]
    n_body = xyz.shape[0]
    n_zero = max(0, n_total - 1 - n_body)
    out = np.empty((1 + n_body + n_zero, 4), np.float32)
    out[0, :3] = origin
    out[1:1 + n_body, :3] = xyz
    out[:, 3] = 1.0
    _window_shuffle(out[:1 + n_body], rng)
    out[1 + n_body:, :3] = 0.0
    return out


def _uniform(rng, n, lo, hi):
    return (lo + (hi - lo) * rng.random(n, dtype=np.float32)).astype(np.float32)


def cfg1(n=1_000_000, zero_frac=0.02) -> np.ndarray:
    """40 m x 25 m: flat | ramp of slope 0.5 (cf. genePcd.cpp:37-52) | plateau with a 0.3 m
    step; 5 mm Gaussian noise in z."""
    rng = _rng(1)
    nb = int(n * (1 - zero_frac)) - 1
    x = _uniform(rng, nb, 0.0, 40.0)
    y = np.sort(_uniform(rng, nb, 0.0, 25.0))
    z = np.where(x < 15, 0.0, np.where(x < 25, 0.5 * (x - 15.0), np.where(x < 32, 5.0, 5.3))).astype(np.float32)
    z += (0.005 * rng.standard_normal(nb, dtype=np.float32)).astype(np.float32)
    return _finish([np.stack([x, y, z], 1)], (20.013, 12.507, 0.5), rng, n, sort_by_y=False)


def cfg2(n=10_000_000, zero_frac=0.02, scale=1.0) -> np.ndarray:
    """120 m x 80 m ground + bridge deck at +3 m over an underpass + approach ramps + 8
    vertical pier walls (a scaled genePcd.cpp:54-195 scene): 2-3 surface layers per column
    under the deck, tall voxel stacks at the piers."""
    rng = _rng(2)
    nb = int(n * (1 - zero_frac)) - 1
    W, H = 120.0 * scale, 80.0 * scale
    x0, x1, y0, y1 = 50.0 * scale, 70.0 * scale, 20.0 * scale, 60.0 * scale  # deck footprint
    ramp = 6.0 * scale
    n_pier = nb // 25
    n_deck = int(nb * 0.10)
    n_ramp = int(nb * 0.06)
    n_ground = nb - n_pier - n_deck - n_ramp
    parts = []
    gx, gy = _uniform(rng, n_ground, 0, W), _uniform(rng, n_ground, 0, H)
    parts.append(np.stack([gx, gy, (0.004 * rng.standard_normal(n_ground, dtype=np.float32))], 1))
    dx, dy = _uniform(rng, n_deck, x0, x1), _uniform(rng, n_deck, y0, y1)
    parts.append(np.stack([dx, dy, 3.0 + 0.004 * rng.standard_normal(n_deck, dtype=np.float32)], 1))
    rx = _uniform(rng, n_ramp, 0, 2 * ramp)
    ry = _uniform(rng, n_ramp, y0, y1)
    up = rx < ramp
    rxx = np.where(up, x0 - ramp + rx, x1 + (rx - ramp)).astype(np.float32)
    rz = np.where(up, 3.0 * rx / ramp, 3.0 * (1.0 - (rx - ramp) / ramp)).astype(np.float32)
    parts.append(np.stack([rxx, ry, rz + 0.004 * rng.standard_normal(n_ramp, dtype=np.float32)], 1))
    k = n_pier // 8
    for i in range(8):  # 4 walls along x at y0/y1 ends, 4 along y at x0/x1
        u = _uniform(rng, k, 0.0, 4.0 * scale)
        h = _uniform(rng, k, 0.0, 3.0)
        if i < 4:
            px = (x0 if i % 2 == 0 else x1 - 4.0 * scale) + u
            py = np.full(k, y0 + 2.0 if i < 2 else y1 - 2.0, np.float32) + 0.002 * rng.standard_normal(k, dtype=np.float32)
        else:
            px = np.full(k, x0 + 1.0 if i % 2 == 0 else x1 - 1.0, np.float32) + 0.002 * rng.standard_normal(k, dtype=np.float32)
            py = (y0 + 8.0 * scale if i < 6 else y1 - 12.0 * scale) + u
        parts.append(np.stack([px.astype(np.float32), py.astype(np.float32), h], 1))
    parts = [p.astype(np.float32) for p in parts]
    return _finish(parts, (0.5 * W + 0.013, 0.5 * H + 0.007, 1.0), rng, n)


def _terrain(x, y):
    return (0.9 * np.sin(0.05 * x) + 0.7 * np.sin(0.07 * y + 1.0) + 0.4 * np.sin(0.031 * (x + y))).astype(np.float32)


def cfg3(n=50_000_000, zero_frac=0.02, extent=224.0, cfg=3) -> np.ndarray:
    """Rolling outdoor terrain (three sinusoids, ~2 m amplitude, 1 cm noise) over
    extent x extent metres; ~10 points per 0.1 m column at the full size."""
    rng = _rng(cfg)
    nb = int(n * (1 - zero_frac)) - 1
    out = np.empty((n, 4), np.float32)
    out[:, 3] = 1.0
    c = 0.5 * extent
    out[0, :3] = (c + 0.013, c + 0.007, float(_terrain(np.float32(c), np.float32(c))) + 0.5)
    chunk = 1 << 22
    for lo in range(0, nb, chunk):  # generated directly in sweep order (y grows with i)
        hi = min(lo + chunk, nb)
        m = hi - lo
        x = _uniform(rng, m, 0.0, extent)
        y = (extent * (lo + np.sort(rng.random(m, dtype=np.float32)) * m) / nb).astype(np.float32)
        z = _terrain(x, y) + 0.01 * rng.standard_normal(m, dtype=np.float32)
        blk = out[1 + lo:1 + hi]
        blk[:, 0], blk[:, 1], blk[:, 2] = x, y, z
    _window_shuffle(out[:1 + nb], rng)
    out[1 + nb:, :3] = 0.0
    return out


def cfg5(n=100_000_000, zero_frac=0.02, extent=500.0, skew=False) -> np.ndarray:
    """Cell-size sweep cloud: cfg3's terrain over 500 m x 500 m.  skew=True puts 20 % of the
    points into 0.1 % of the area (occupancy-skew / segment-imbalance stress)."""
    pts = cfg3(n, zero_frac, extent, cfg=5)
    if skew:
        rng = _rng(5, 1)
        nb = int(n * (1 - zero_frac)) - 1
        idx = 1 + rng.choice(nb, size=nb // 5, replace=False)
        side = extent * np.sqrt(1e-3)
        pts[idx, 0] = _uniform(rng, idx.size, 0.4 * extent, 0.4 * extent + side)
        pts[idx, 1] = _uniform(rng, idx.size, 0.6 * extent, 0.6 * extent + side)
        pts[idx, 2] = _terrain(pts[idx, 0], pts[idx, 1]) + 0.01 * rng.standard_normal(idx.size, dtype=np.float32)
    return pts


def scans(n_scans=10, n_per_scan=100_000, radius=20.0, cfg2_scale=1.0, seed_stream=4):
    """cfg4: lidar-like scans (20 m-radius discs around a pose moving across the cfg2
    scene), yielded one at a time; fused into a resident map by MapBuilder.update."""
    rng = _rng(4, seed_stream)
    W, H = 120.0 * cfg2_scale, 80.0 * cfg2_scale
    for s in range(n_scans):
        t = (s + 0.5) / n_scans
        cx, cy = 0.15 * W + 0.7 * W * t, 0.5 * H + 0.2 * H * np.sin(6.0 * t)
        r = radius * np.sqrt(rng.random(n_per_scan, dtype=np.float32))
        a = (2 * np.pi) * rng.random(n_per_scan, dtype=np.float32)
        x = np.clip(cx + r * np.cos(a), 0, W).astype(np.float32)
        y = np.clip(cy + r * np.sin(a), 0, H).astype(np.float32)
        deck = (x > 50 * cfg2_scale) & (x < 70 * cfg2_scale) & (y > 20 * cfg2_scale) & (y < 60 * cfg2_scale) & (rng.random(n_per_scan) < 0.5)
        z = np.where(deck, 3.0, 0.0).astype(np.float32) + 0.004 * rng.standard_normal(n_per_scan, dtype=np.float32)
        out = np.empty((n_per_scan, 4), np.float32)
        out[:, 0], out[:, 1], out[:, 2], out[:, 3] = x, y, z.astype(np.float32), 1.0
        yield out


def make(cfg: str, n=None, **kw) -> np.ndarray:
    fn = {"cfg1": cfg1, "cfg2": cfg2, "cfg3": cfg3, "cfg5": cfg5}[cfg]
    return fn(**({"n": n} if n else {}), **kw)
