"""adapter/gndt_twodmap_adapter.h: the tables of a build, pushed through the header-only C++
adapter, must recreate exactly the containers the reference's own code builds (morton_list
order, map_cell / map_slope contents, map_xy node order).  Runs on CPU: the tables come from
the oracle port, the comparison runs inside oracle/_ref against the reference's global map."""
import ctypes as C

import numpy as np
import pytest

from grid_ndt_b200 import synthetic
from grid_ndt_b200._abi import SLOPE_DTYPE, Params, default_params
from oracle import oracle as O
from tests.conftest import have_reference_lib

pytestmark = pytest.mark.skipif(not have_reference_lib(), reason="oracle/_ref not built")


def slopes_of(v):
    idx = np.nonzero(v["flags"] & 2)[0]
    s = np.zeros(len(idx), SLOPE_DTYPE)
    for f in ("sx", "sy", "sz", "mean", "normal", "rough", "flags"):
        s[f] = v[f][idx]
    s["voxel"] = idx
    return s


@pytest.mark.parametrize("demand", ["slope", "true"])
@pytest.mark.parametrize("name,n,gl,kw", [("cfg1", 60_000, 0.2, {}), ("cfg2", 150_000, 0.2, {"scale": 0.15})])
def test_adapter_rebuilds_reference_containers(name, n, gl, kw, demand):
    cloud = synthetic.make(name, n, **kw)
    p = default_params(gl, 0.1, 0.08, demand)
    O.ref_build(cloud, p)  # leaves the reference's global map2D populated
    o = O.oracle_build(cloud, p)
    vox, cols, sl = o.voxels, o.columns, slopes_of(o.voxels)
    lib = O._lib("ref")
    fn = lib.gndt_ref_adapter_check
    fn.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    fn.restype = C.c_int
    origin = np.ascontiguousarray(cloud[0, :3], np.float32)
    bad = fn(origin.ctypes.data, C.byref(p), vox.ctypes.data, len(vox), sl.ctypes.data, len(sl), cols.ctypes.data, len(cols), 1)
    assert bad == 0, f"adapter mismatch code {bad}"


@pytest.mark.parametrize("name,n,gl,kw", [("cfg1", 60_000, 0.2, {}), ("cfg2", 150_000, 0.2, {"scale": 0.15})])
def test_integer_lookups_equal_reference_string_lookups(name, n, gl, kw):
    """adapter CellIndex / include/gndt_lookup.h against map_cell.find(key) and the neighbour
    keys of the reference's own (private) TwoDmap::countLRFB, for every cell, a ring of empty
    cells around each, and all four directions."""
    cloud = synthetic.make(name, n, **kw)
    p = default_params(gl, 0.1, 0.08, "slope")
    o = O.oracle_build(cloud, p)
    vox, cols, sl = o.voxels, o.columns, slopes_of(o.voxels)
    # the port's slope table is compacted like the library's: slope_begin/count index into it
    lib = O._lib("ref")
    fn = lib.gndt_ref_lookup_check
    fn.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                   C.POINTER(C.c_double)]
    fn.restype = C.c_int
    origin = np.ascontiguousarray(cloud[0, :3], np.float32)
    rates = (C.c_double * 2)()
    bad = fn(origin.ctypes.data, C.byref(p), vox.ctypes.data, len(vox), sl.ctypes.data, len(sl), cols.ctypes.data, len(cols), rates)
    assert bad == 0, f"{bad} lookups differ between the integer tables and the reference's string maps"
    print(f"\n{name}: {len(cols)} cells; neighbour lookups/s: reference strings {rates[0]:.3g}, integer tables {rates[1]:.3g} "
          f"({rates[1] / rates[0]:.0f}x)")
    assert rates[1] > rates[0]


@pytest.mark.parametrize("name,n,gl,kw", [("cfg1", 250_000, 0.2, {}), ("cfg2", 400_000, 0.2, {"scale": 0.2})])
def test_graph_equals_reference_accessible_neighbors_and_cost_map(name, n, gl, kw):
    """the CSR traversability graph (here from the oracle port; the GPU's is
    compared with it in tests/test_gpu_graph.py) served through adapter SlopeGraph must return, for
    EVERY Slope, exactly the list the reference's own TwoDmap::AccessibleNeighbors returns (same
    Slope objects, same order), and computeCostFast must leave the same h on every Slope as the
    reference's own TwoDmap::computeCost (map2D.h:1285-1397) from the same goal."""
    cloud = synthetic.make(name, n, **kw)
    p = default_params(gl, 0.1, 0.08, "slope")
    o = O.oracle_build(cloud, p)
    vox, cols, sl = o.voxels, o.columns, slopes_of(o.voxels)
    off, tgt = O.oracle_edges(o, p)
    assert len(off) == len(sl) + 1 and off[-1] == len(tgt)
    # goal: a Slope in the middle of the map with neighbours in all four directions
    deg = np.diff(off.astype(np.int64))
    g = int(np.argsort(-deg, kind="stable")[len(deg) // 50])
    goal = np.ascontiguousarray(sl["mean"][g], np.float32)
    lib = O._lib("ref")
    fn = lib.gndt_ref_graph_check
    fn.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
    fn.restype = C.c_int
    origin = np.ascontiguousarray(cloud[0, :3], np.float32)
    out = (C.c_double * 6)()
    bad = fn(origin.ctypes.data, C.byref(p), vox.ctypes.data, len(vox), sl.ctypes.data, len(sl), cols.ctypes.data, len(cols),
             off.ctypes.data, tgt.ctypes.data, goal.ctypes.data, out)
    assert bad == 0, f"{bad} Slopes differ (neighbour lists or cost-map h)"
    print(f"\n{name}: {len(sl)} slopes, {len(tgt)} edges; AccessibleNeighbors/s: reference {out[0]:.3g}, graph {out[1]:.3g} ({out[1] / out[0]:.0f}x); "
          f"computeCost: reference {out[2] * 1e3:.1f} ms, computeCostFast {out[3] * 1e3:.2f} ms ({out[2] / max(out[3], 1e-9):.0f}x); "
          f"{int(out[4])} Slopes reached, {int(out[5])} traversable")
    assert out[4] > 100, "the goal reached almost nothing: pick a better goal"
    assert out[1] > out[0] and out[3] < out[2]
