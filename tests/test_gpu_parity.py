"""GPU parity tests: libgndt.so (through the C ABI / TwoDmap mirror) against the CPU oracle on
the same seeded inputs, the committed golden vectors, edge cases, and size-independent
properties at the BASELINE size."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from grid_ndt_b200 import _abi, synthetic
from grid_ndt_b200._abi import default_params

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


def _gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test on a box without CUDA (no fallback exists)")
    return torch


def _run(cloud, p, demand="slope", **kw):
    from tests import parity
    rep = parity.run_case(cloud, p, demand, **kw)
    assert rep["ok"], json.dumps({k: v for k, v in rep.items() if k != "stage_ms"}, default=str)
    return rep


def _build(cloud, gl, zl, interval=0.08, demand="slope", **params):
    from grid_ndt_b200 import TwoDmap
    m = TwoDmap(gl, zl)
    m.setInterval(interval)
    for k, v in params.items():
        setattr(m.params, k, v)
    m.chatterCallback(cloud, demand)
    return m


def test_bridge_ground_golden_fixture():
    """The reference's own deterministic fixture (src/test/genePcd.cpp) with its parameter
    preset (launch/parameters.txt:53-59): counters equal the golden file made from the
    reference build; full table parity against the oracle."""
    _gpu()
    from oracle import oracle as O
    g = json.load(open(os.path.join(G, "bridge_ground.json")))
    cloud, _ = O.bridge_ground()
    rep = _run(cloud, default_params(0.1, 0.05, 0.08))
    for k in ("n_binned", "n_columns", "n_voxels", "n_fitted", "n_slopes"):
        assert rep["counts"][k] == g["counts"][k]
    assert rep["label_mismatch"] == 0


@pytest.mark.parametrize("name,n,gl,zl,kw", [
    ("cfg1", 1_000_000, 0.2, 0.1, {}),                      # BASELINE configs[0] at full size
    ("cfg2", 2_000_000, 0.2, 0.1, {"scale": 0.2 ** 0.5}),
    ("cfg3", 2_000_000, 0.1, 0.1, {"extent": 224.0 * 0.04 ** 0.5}),
    ("cfg2", 500_000, 0.5, 0.05, {"scale": 0.25}),
    ("cfg3", 400_000, 1.0, 0.1, {"extent": 60.0}),          # big cells: thousands of points per voxel
    ("cfg1", 300_000, 0.05, 0.05, {}),                      # small cells: mostly < 3 points per voxel
])
def test_synthetic_parity(name, n, gl, zl, kw):
    _gpu()
    rep = _run(synthetic.make(name, n, **kw), default_params(gl, zl, 0.08))
    assert rep["exact_count_mismatch"] == 0 and rep["exact_first_index_mismatch"] == 0


def test_demand_true_parity():
    _gpu()
    _run(synthetic.cfg2(600_000, scale=0.25), default_params(0.2, 0.1, 0.08, "true"), "true")


def test_explicit_origin_bins_every_point():
    _gpu()
    cloud = synthetic.cfg1(200_000)
    p = default_params(0.2, 0.1, 0.08, origin=(3.25, -1.5, 0.75), origin_is_first_point=0)
    rep = _run(cloud, p)
    assert rep["counts"]["n_binned"] == 200_000


def test_negative_quadrants_and_far_origin():
    _gpu()
    cloud = synthetic.cfg1(150_000)
    cloud[:, 0] -= 300.0
    cloud[:, 1] -= 700.0
    cloud[:, 2] += 55.0
    cloud[-3000:, :3] = 0.0
    _run(cloud, default_params(0.2, 0.1, 0.08))


def test_strides_and_memory_kinds_agree():
    """packed xyz (12 B), PointXYZ (16 B) and a 32-byte record; host numpy, host torch and
    device torch inputs all give byte-identical tables."""
    torch = _gpu()
    base = synthetic.cfg1(120_000)
    tables = []
    for width in (3, 4, 8):
        c = np.zeros((base.shape[0], width), np.float32)
        c[:, :3] = base[:, :3]
        for src in (c, torch.from_numpy(c), torch.from_numpy(c).cuda()):
            m = _build(src, 0.2, 0.1)
            tables.append(m.voxels.tobytes())
            m.close()
    assert all(t == tables[0] for t in tables)


def test_builds_are_deterministic():
    _gpu()
    cloud = synthetic.cfg2(400_000, scale=0.2)
    m = _build(cloud, 0.2, 0.1)
    a = (m.voxels.tobytes(), m.slopes.tobytes(), m.columns.tobytes())
    m.chatterCallback(cloud, "slope")
    b = (m.voxels.tobytes(), m.slopes.tobytes(), m.columns.tobytes())
    m.close()
    assert a == b


def test_edge_cases():
    _gpu()
    p = default_params(0.2, 0.1, 0.08)
    # only the origin point: empty map
    m = _build(np.array([[1, 2, 3, 1]], np.float32), 0.2, 0.1)
    assert m.counts()["n_voxels"] == 0 and len(m.voxels) == 0 and len(m.slopes) == 0
    m.close()
    # two points
    _run(np.array([[1, 2, 3, 1], [1.5, 2.5, 3.5, 1]], np.float32), p)
    # every point identical to the origin: one voxel at (-1,-1,-1), rough 0.01 (map2D.h:131)
    rep = _run(np.tile(np.array([[1, 2, 3, 1]], np.float32), (5000, 1)), p)
    assert rep["counts"]["n_voxels"] == 1
    # one heavy voxel spanning dozens of reduce tiles + NaN / Inf / out-of-range points
    cloud = synthetic.cfg1(100_000, zero_frac=0.6)
    cloud[5, 0] = np.nan
    cloud[6, 1] = np.inf
    cloud[7, 2] = 1e9
    cloud[8, 0] = -1e7
    rep = _run(cloud, p)
    assert rep["counts"]["n_dropped"] == 4
    # a single tall column: 3000 z levels in one x-y cell
    z = np.linspace(-100, 100, 30_000, dtype=np.float32)
    col = np.stack([np.full_like(z, 0.31), np.full_like(z, 0.47), z, np.ones_like(z)], 1)
    _run(col, p)
    # exactly planar, axis aligned (exact zero eigenvalue -> rough 0.01) and a line of points
    rng = np.random.default_rng(5)
    plane = np.stack([rng.uniform(0, 4, 40_000), rng.uniform(0, 4, 40_000), np.full(40_000, 0.5), np.ones(40_000)], 1).astype(np.float32)
    rep = _run(plane, p)
    line = np.stack([np.full(20_000, 1.25), rng.uniform(0, 9, 20_000), np.full(20_000, 0.5), np.ones(20_000)], 1).astype(np.float32)
    _run(line, p)


def test_tile_filter_matches_oracle():
    _gpu()
    cloud = synthetic.cfg2(500_000, scale=0.22)
    p = default_params(0.2, 0.1, 0.08, origin=tuple(cloud[0, :3]), origin_is_first_point=0, tile_lo=-20, tile_hi=35)
    rep = _run(cloud, p)
    assert rep["counts"]["n_outside_tile"] > 0


def test_capacity_error_is_reported():
    _gpu()
    from grid_ndt_b200 import GndtError, TwoDmap
    m = TwoDmap(0.2, 0.1)
    m.params.max_voxels = 100
    m.chatterCallback(synthetic.cfg1(50_000), "slope")
    with pytest.raises(GndtError) as e:
        m.counts()
    assert e.value.status == _abi.GNDT_ERR_CAPACITY
    m.close()


def test_bad_arguments():
    _gpu()
    from grid_ndt_b200 import GndtError, TwoDmap, lib
    m = TwoDmap(0.2, 0.1)
    with pytest.raises(GndtError):
        m.counts()  # nothing built yet
    with pytest.raises(GndtError):
        m.create2DMap("slope")  # nothing staged
    m.uniformDivision(np.zeros((10, 4), np.float32))
    with pytest.raises(GndtError):
        m.create2DMap("3d")  # the reference builds no Slopes for other demand strings
    L = lib()
    buf = np.zeros((10, 4), np.float32)
    assert L.gndt_build(m._h, buf.ctypes.data, 10, 10, 0, None) == _abi.GNDT_ERR_INVALID_ARG  # stride
    assert L.gndt_build(m._h, None, 10, 16, 0, None) == _abi.GNDT_ERR_INVALID_ARG
    assert L.gndt_build(m._h, buf.ctypes.data, 0, 16, 0, None) == _abi.GNDT_ERR_INVALID_ARG
    m.close()


def test_reference_containers_view():
    """morton_list / map_cell rebuilt on the host carry the reference's keys and order."""
    _gpu()
    from oracle import oracle as O
    cloud = synthetic.cfg1(60_000)
    m = _build(cloud, 0.2, 0.1)
    o = O.oracle_build(cloud, default_params(0.2, 0.1, 0.08))
    want = [O.oracle_morton_string(int(c["sx"]), int(c["sy"])) for c in o.columns[o.morton_list]]
    assert m.morton_list == want
    cells = m.map_cell
    assert len(cells) == o.counts["n_columns"]  # one Cell per occupied column (map2D.h:598)
    n_slopes = sum(len(c.map_slope) for c in cells.values())
    assert n_slopes == o.counts["n_slopes"]
    key, z = m.transMortonXYZ(cloud[1234, :3], origin=cloud[0, :3])
    assert key in cells
    m.close()


def test_full_size_properties_10M():
    """BASELINE configs[1] at full size: oracle comparison on the integer fields plus
    size-independent properties (sortedness, conservation of points, means inside their
    cells, PSD scatters, slope/column tables consistent, idempotence)."""
    _gpu()
    from oracle import oracle as O
    cloud = synthetic.cfg2(10_000_000)
    m = _build(cloud, 0.2, 0.1)
    v, s, c, cnt = m.voxels, m.slopes, m.columns, m.counts()
    assert int(v["count"].sum()) == cnt["n_binned"] == 9_999_999
    cz = lambda a: np.where(a > 0, a - 1, a).astype(np.int64)
    key = (cz(v["sx"]) + 32768) << 32 | (cz(v["sy"]) + 32768) << 16 | (cz(v["sz"]) + 32768)
    assert np.all(np.diff(key) > 0)  # strictly sorted, no duplicate voxel
    fit = (v["flags"] & 1) != 0
    assert np.array_equal(fit, v["count"] >= 3)
    o3 = cloud[0, :3].astype(np.float64)
    for ax, ln in ((0, 0.2), (1, 0.2), (2, 0.1)):
        idx = cz(v[("sx", "sy", "sz")[ax]])[fit]
        d = v["mean"][fit, ax].astype(np.float64) - o3[ax]
        assert np.all(d >= idx * np.float32(ln) - 1e-4) and np.all(d <= (idx + 1) * np.float32(ln) + 1e-4)
    assert np.all(v["evals"][fit, 0] >= -1e-5 * np.maximum(v["evals"][fit, 2], 1e-12))
    assert np.all(np.diff(v["evals"], axis=1) >= 0)
    assert len(s) == cnt["n_slopes"] == int(((v["flags"] & 2) != 0).sum())
    assert len(c) == cnt["n_columns"] and int(c["voxel_count"].sum()) == len(v) and int(c["slope_count"].sum()) == len(s)
    o = O.oracle_build(cloud, default_params(0.2, 0.1, 0.08))
    assert cnt["n_voxels"] == o.counts["n_voxels"] and cnt["n_slopes"] == o.counts["n_slopes"]
    for f in ("sx", "sy", "sz", "count", "first_index"):
        assert np.array_equal(v[f], o.voxels[f]), f
    assert np.array_equal(v["flags"] & 0x0F, o.voxels["flags"] & 0x0F)
    scale = np.maximum(1.0, np.abs(o.voxels["mean"]).max(axis=1))
    assert np.all(np.abs(v["mean"].astype(np.float64) - o.voxels["mean"]).max(axis=1) <= 1e-5 * scale)
    first = v.tobytes()
    m.chatterCallback(cloud, "slope")
    assert m.voxels.tobytes() == first  # idempotent / deterministic at full size
    m.close()


def test_world2_strips_nccl():
    """Two x strips on two GPUs, all-gathered; skipped on a single-GPU box."""
    torch = _gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", os.path.join(root, "tests", "multi_gpu_worker.py")], capture_output=True, timeout=600)
    assert out.returncode == 0, out.stdout.decode()[-3000:] + out.stderr.decode()[-3000:]


def test_hoisted_division_is_verified_and_equivalent():
    """The hoisted index division must have passed its exhaustive on-device check for the
    lengths in use, and a build with the plain IEEE division (GNDT_EXACT_DIV=1, separate
    process) must produce byte-identical tables."""
    _gpu()
    import subprocess, sys, hashlib
    cloud = synthetic.cfg2(300_000, scale=0.2)
    m = _build(cloud, 0.2, 0.1)
    enabled, checked = m.fast_div_status()
    assert enabled and checked > 100_000_000, (enabled, checked)
    digest = hashlib.sha256(m.voxels.tobytes() + m.slopes.tobytes() + m.columns.tobytes()).hexdigest()
    m.close()
    code = ("import hashlib; from grid_ndt_b200 import TwoDmap, synthetic; c = synthetic.cfg2(300_000, scale=0.2); "
            "m = TwoDmap(0.2, 0.1); m.setInterval(0.08); m.chatterCallback(c, 'slope'); assert not m.fast_div_status()[0]; "
            "print(hashlib.sha256(m.voxels.tobytes() + m.slopes.tobytes() + m.columns.tobytes()).hexdigest())")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, cwd=root, env=dict(os.environ, GNDT_EXACT_DIV="1"), timeout=300)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    assert out.stdout.decode().strip().splitlines()[-1] == digest
    for gl, zl in ((0.05, 0.05), (0.1, 0.05), (0.25, 0.1), (0.5, 0.1), (1.0, 0.1), (0.3, 0.07)):
        mm = _build(synthetic.cfg1(20_000), gl, zl)
        assert mm.fast_div_status()[0], (gl, zl)
        mm.close()
