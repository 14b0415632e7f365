"""The oracle (oracle/gndt_oracle.c) against the committed golden vectors that were produced
by the reference's own code (tests/golden/make_golden.py), and — where the reference build
is available — against that library directly."""
import json
import os

import numpy as np
import pytest

from grid_ndt_b200 import synthetic
from grid_ndt_b200._abi import default_params
from oracle import oracle as O
from tests.conftest import have_reference_lib
from tests.golden.make_golden import digest, table_digests

G = os.path.join(os.path.dirname(__file__), "golden")


def port_digests(m):
    """Same digests as make_golden.table_digests; the port's first_index is a cloud index,
    so the first-seen order is compared through (column first, voxel first)."""
    d = table_digests(m)
    v = m.voxels
    col_first = m.columns["first_index"][v["column"]]
    d["first_seen_rank_order"] = digest(np.lexsort((v["first_index"], col_first)).astype(np.uint32))
    return d


def test_kat_count_morton():
    kat = json.load(open(os.path.join(G, "kat_keys.json")))
    assert len(kat["count_morton"]) >= 50
    for c in kat["count_morton"]:
        assert str(O.oracle_count_morton(c["a"], c["b"])) == c["string"], c
        if c["inverse"] is not None and 0 < int(c["string"]):
            assert list(O.oracle_morton_to_xy(int(c["string"]))) == c["inverse"], c
    # the worked example in the reference's comment (Stopwatch.h:112-116) and the overflow quirks
    assert O.oracle_count_morton(5, 7) == 55
    assert O.oracle_count_morton(32768, 1) == -2147483647
    assert O.oracle_count_morton(65536, 1) == 1


def test_kat_trans_morton():
    kat = json.load(open(os.path.join(G, "kat_keys.json")))
    assert len(kat["trans"]) >= 200
    for c in kat["trans"]:
        rc, sx, sy, sz = O.oracle_trans(c["origin"], c["grid_len"], c["z_len"], c["pos"])
        assert rc == 0
        assert O.oracle_morton_string(sx, sy) == c["key"] and sz == c["z"], c


def test_survey_index_cases():
    # SURVEY.md §8(c) known answers (binary32 arithmetic, not real arithmetic)
    idx = lambda p, p0, ln: abs(O.oracle_trans((p0, p0, p0), ln, ln, (p, p0, p0))[1])
    assert idx(1.05, 1, 0.1) == 1 and idx(1.2, 1, 0.1) == 3 and idx(1.3, 1, 0.1) == 3
    assert idx(1.4, 1, 0.2) == 2 and idx(0.8, 1, 0.1) == 2 and idx(3.0, 1, 0.1) == 20
    assert idx(1.6, 1, 0.2) == 3 and idx(1.1, 1, 0.05) == 3
    rc, sx, sy, sz = O.oracle_trans((1, 1, 1), 0.1, 0.1, (1.0, 1.0, 1.0))
    assert (sx, sy, sz) == (-1, -1, -1)  # d == 0 lands in negative-side cell 1 (quadrant D)


def test_bridge_ground_fixture_and_golden():
    g = json.load(open(os.path.join(G, "bridge_ground.json")))
    pts, assigned = O.bridge_ground()
    assert assigned == g["assigned_points"] == 295841
    assert digest(pts) == g["cloud_sha"]
    m = O.oracle_build(pts, default_params(0.1, 0.05, 0.08, "slope"))
    assert m.counts == g["counts"]
    assert port_digests(m) == g["digests"]


@pytest.mark.parametrize("demand", ["slope", "true"])
def test_cfg1_golden(demand):
    g = json.load(open(os.path.join(G, "cfg1_50k.json")))
    cloud = synthetic.cfg1(50_000)
    assert digest(cloud) == g["cloud_sha"]
    m = O.oracle_build(cloud, default_params(0.2, 0.1, 0.08, demand))
    assert m.counts == g[demand]["counts"]
    assert port_digests(m) == g[demand]["digests"]


@pytest.mark.skipif(not have_reference_lib(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name,n,gl,zl,kw", [
    ("cfg1", 120_000, 0.2, 0.1, {}),
    ("cfg2", 300_000, 0.2, 0.1, {"scale": 0.18}),
    ("cfg3", 200_000, 0.1, 0.1, {"extent": 14.0}),
    ("cfg2", 150_000, 0.5, 0.05, {"scale": 0.15}),
])
@pytest.mark.parametrize("demand", ["slope", "true"])
def test_port_equals_reference(name, n, gl, zl, kw, demand):
    """Bit-exact agreement of the restatement with the reference's own control flow on
    every observable: keys, counts, first-seen orders, centroids, scatters, eigen outputs,
    Slope set, down flags, reachability."""
    cloud = synthetic.make(name, n, **kw)
    p = default_params(gl, zl, 0.08, demand)
    a, b = O.oracle_build(cloud, p), O.ref_build(cloud, p)
    assert a.counts == b.counts
    for f in ("sx", "sy", "sz", "count", "mean", "scatter", "evals", "normal", "rough", "column", "slope"):
        assert np.array_equal(a.voxels[f], b.voxels[f]), f
    assert np.array_equal(a.voxels["flags"] & 0x1F3, b.voxels["flags"] & 0x1F3)
    sl = (a.voxels["flags"] & 2) != 0
    assert np.array_equal((a.voxels["flags"] & 0x0C)[sl], (b.voxels["flags"] & 0x0C)[sl])
    assert np.array_equal(a.morton_list, b.morton_list)
    col_first = a.columns["first_index"][a.voxels["column"]]
    order_a = np.lexsort((a.voxels["first_index"], col_first))
    order_b = np.argsort(b.voxels["first_index"], kind="stable")
    assert np.array_equal(order_a, order_b)
    for f in ("sx", "sy", "voxel_begin", "voxel_count", "slope_begin", "slope_count"):
        assert np.array_equal(a.columns[f], b.columns[f]), f


@pytest.mark.skipif(not have_reference_lib(), reason="oracle/_ref not built")
def test_key_arithmetic_matches_reference_randomised():
    rng = np.random.default_rng(7)
    for _ in range(3000):
        o = rng.uniform(-100, 100, 3).astype(np.float32)
        gl = float(rng.choice([0.05, 0.1, 0.2, 0.25, 0.5, 1.0]))
        zl = float(rng.choice([0.05, 0.1]))
        k = rng.integers(-300, 300, 3)
        pos = (o + np.array([k[0] * gl, k[1] * gl, k[2] * zl], np.float32)).astype(np.float32)
        if rng.random() < 0.5:
            pos = np.nextafter(pos, rng.choice([-np.inf, np.inf]), dtype=np.float32)
        rc, sx, sy, sz = O.oracle_trans(o, gl, zl, pos)
        key, z = O.ref_trans(o, gl, zl, pos)
        assert rc == 0 and O.oracle_morton_string(sx, sy) == key and sz == z


def test_truth64_close_to_faithful32():
    cloud = synthetic.cfg1(80_000)
    p = default_params(0.2, 0.1, 0.08)
    a, t = O.oracle_build(cloud, p, "faithful32"), O.oracle_build(cloud, p, "truth64")
    assert np.array_equal(a.voxels["count"], t.voxels["count"])
    fit = (a.voxels["flags"] & 1) != 0
    s = np.abs(t.voxels["scatter"][fit]).max(axis=1)
    assert (np.abs(a.voxels["scatter"][fit] - t.voxels["scatter"][fit]).max(axis=1) <= 1e-4 * s + 1e-12).all()


def test_oracle_edge_cases():
    p = default_params(0.2, 0.1, 0.08)
    one = np.array([[1, 2, 3, 1]], np.float32)
    m = O.oracle_build(one, p)  # only the origin point: nothing binned
    assert m.counts["n_voxels"] == 0 and m.counts["n_binned"] == 0
    same = np.tile(np.array([[1, 2, 3, 1]], np.float32), (1000, 1))
    m = O.oracle_build(same, p)  # d == 0 on all axes -> one voxel at (-1,-1,-1)
    v = m.voxels[0]
    assert m.counts["n_voxels"] == 1 and (v["sx"], v["sy"], v["sz"], v["count"]) == (-1, -1, -1, 999)
    assert v["rough"] == np.float32(0.01) and np.all(v["scatter"] == 0)
    bad = np.array([[0, 0, 0, 1], [np.nan, 0, 0, 1], [np.inf, 1, 1, 1], [1e9, 0, 0, 1], [0.5, 0.5, 0.5, 1]], np.float32)
    m = O.oracle_build(bad, p)
    assert m.counts["n_dropped"] == 3 and m.counts["n_binned"] == 1


@pytest.mark.skipif(not have_reference_lib(), reason="oracle/_ref not built")
def test_port_equals_reference_on_small_degenerate_clouds():
    """Fuzz: the restatement against the reference's own code on 150 tiny clouds that stress
    what the synthetic scenes do not — 2..400 points, all four quadrants around the first point,
    exact duplicates, points on cell edges, lines and planes (singular scatters, the
    rough == 0 -> 0.01 rule), 1- and 2-point voxels next to fitted ones, both demands."""
    rng = np.random.default_rng(20260000)
    for case in range(150):
        n = int(rng.integers(2, 400))
        gl = float(rng.choice([0.1, 0.2, 0.5]))
        zl = float(rng.choice([0.05, 0.1]))
        kind = case % 5
        pts = rng.uniform(-1.5, 1.5, (n, 3)).astype(np.float32)
        if kind == 1:    # on a lattice of cell edges / exact duplicates
            pts = (np.round(pts / gl) * gl).astype(np.float32)
        elif kind == 2:  # a vertical line and a horizontal plane
            pts[: n // 2, :2] = np.float32(0.3)
            pts[n // 2:, 2] = np.float32(-0.2)
        elif kind == 3:  # few voxels, many points each
            pts *= np.float32(0.15)
        elif kind == 4:  # far from the origin of coordinates
            pts += np.array([431.7, -209.3, 12.1], np.float32)
        cloud = np.concatenate([pts, np.ones((n, 1), np.float32)], axis=1)
        cloud[0, :3] = pts.mean(axis=0)  # the map origin sits inside the cloud: four quadrants
        p = default_params(gl, zl, 0.08, "true" if case % 2 else "slope")
        a, b = O.oracle_build(cloud, p), O.ref_build(cloud, p)
        assert a.counts == b.counts, case
        for f in ("sx", "sy", "sz", "count", "mean", "scatter", "evals", "normal", "rough", "column", "slope"):
            assert np.array_equal(a.voxels[f], b.voxels[f]), (case, f)
        assert np.array_equal(a.voxels["flags"] & 0x1F3, b.voxels["flags"] & 0x1F3), case
        sl = (a.voxels["flags"] & 2) != 0
        assert np.array_equal((a.voxels["flags"] & 0x0C)[sl], (b.voxels["flags"] & 0x0C)[sl]), case
        assert np.array_equal(a.morton_list, b.morton_list), case
