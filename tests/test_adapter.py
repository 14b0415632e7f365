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
