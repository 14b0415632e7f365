"""GPU path against the oracle on many tiny degenerate clouds: 2..400 points, all four quadrants
around the first point, exact duplicates, points on cell edges, lines and planes (singular
scatters, the rough == 0 -> 0.01 rule), few heavy voxels, coordinates far from zero, both
demands — the same generator as the CPU-side fuzz of the oracle against the reference
(tests/test_oracle_golden.py)."""
import json

import numpy as np
import pytest

from grid_ndt_b200._abi import default_params

pytestmark = pytest.mark.gpu


def test_small_degenerate_clouds_match_oracle():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA (no fallback exists)")
    from tests import parity
    rng = np.random.default_rng(20260000)
    bad = []
    for case in range(150):
        n = int(rng.integers(2, 400))
        gl = float(rng.choice([0.1, 0.2, 0.5]))
        zl = float(rng.choice([0.05, 0.1]))
        kind = case % 5
        pts = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
        if kind == 1:
            pts = (np.round(pts / gl) * gl).astype(np.float32)
        elif kind == 2:
            pts[: n // 2, :2] = np.float32(0.3)
            pts[n // 2:, 2] = np.float32(-0.2)
        elif kind == 3:
            pts *= np.float32(0.15)
        elif kind == 4:
            pts += np.array([431.7, -209.3, 12.1], np.float32)
        cloud = np.concatenate([pts, np.ones((n, 1), np.float32)], axis=1)
        cloud[0, :3] = pts.mean(axis=0)
        demand = "true" if case % 2 else "slope"
        rep = parity.run_case(cloud, default_params(gl, zl, 0.08, demand), demand)
        if not rep["ok"]:
            bad.append((case, kind, n, {k: v for k, v in rep.items() if k not in ("stage_ms", "counts")}))
    assert not bad, json.dumps(bad[:3], default=str)[:3000]
