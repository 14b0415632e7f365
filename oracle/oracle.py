"""ctypes wrappers around the CPU checkers (TEST INFRASTRUCTURE ONLY).

  * libgndt_oracle.so      — oracle/gndt_oracle.c, the plain-C restatement ("port")
  * _ref/libgndt_ref.so    — the reference's own sources built against inert shims
                             ("reference"); absent only if it was never built

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package (grid_ndt_b200) must never do so.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from grid_ndt_b200._abi import COLUMN_DTYPE, VOXEL_DTYPE, Params

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libgndt_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libgndt_ref.so")
REFERENCE_TREE = "/root/reference"


class _Result(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_input", "n_binned", "n_dropped", "n_outside_tile",
        "n_columns", "n_voxels", "n_fitted", "n_slopes")] + [
        ("voxels", C.c_void_p), ("columns", C.c_void_p), ("morton_list", C.c_void_p),
        ("division_s", C.c_double), ("calculate_s", C.c_double), ("edges_s", C.c_double)]


@dataclass
class OracleMap:
    counts: dict
    voxels: np.ndarray       # VOXEL_DTYPE, canonical (cx,cy,cz) order
    columns: np.ndarray      # COLUMN_DTYPE, same order
    morton_list: np.ndarray  # column ids in first-seen order (morton_list of the reference)
    division_s: float
    calculate_s: float
    edges_s: float


def build(force=False):
    """Compile the checkers (make -C oracle).  _ref is rebuilt only where the reference
    tree exists (this container); elsewhere the prebuilt file is used as shipped."""
    if force or not os.path.exists(ORACLE_SO) or (os.path.isdir(REFERENCE_TREE) and not os.path.exists(REF_SO)):
        subprocess.check_call(["make", "-C", HERE, "all"], stdout=subprocess.DEVNULL)


_libs = {}


def _lib(kind):
    if kind not in _libs:
        path = ORACLE_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            build()
        lib = C.CDLL(path)
        pre = "gndt_oracle" if kind == "port" else "gndt_ref"
        fn = getattr(lib, pre + "_build")
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.POINTER(Params), C.c_int, C.POINTER(C.POINTER(_Result))]
        fn.restype = C.c_int
        getattr(lib, pre + "_free").argtypes = [C.POINTER(_Result)]
        getattr(lib, pre + "_free").restype = None
        _libs[kind] = lib
    return _libs[kind]


def have_ref():
    return os.path.exists(REF_SO) or os.path.isdir(REFERENCE_TREE)


def _run(kind, xyzw: np.ndarray, params: Params, mode: int) -> OracleMap:
    lib = _lib(kind)
    pts = np.ascontiguousarray(xyzw, dtype=np.float32)
    assert pts.ndim == 2 and pts.shape[1] >= 3
    res = C.POINTER(_Result)()
    pre = "gndt_oracle" if kind == "port" else "gndt_ref"
    rc = getattr(lib, pre + "_build")(pts.ctypes.data, pts.shape[0], pts.shape[1], C.byref(params), mode, C.byref(res))
    if rc != 0:
        raise RuntimeError(f"{pre}_build failed: {rc}")
    r = res.contents
    nv, nc = int(r.n_voxels), int(r.n_columns)
    vox = np.frombuffer((C.c_char * (nv * VOXEL_DTYPE.itemsize)).from_address(r.voxels), dtype=VOXEL_DTYPE).copy() if nv else np.zeros(0, VOXEL_DTYPE)
    cols = np.frombuffer((C.c_char * (nc * COLUMN_DTYPE.itemsize)).from_address(r.columns), dtype=COLUMN_DTYPE).copy() if nc else np.zeros(0, COLUMN_DTYPE)
    ml = np.frombuffer((C.c_char * (nc * 4)).from_address(r.morton_list), dtype=np.uint32).copy() if nc else np.zeros(0, np.uint32)
    counts = {n: int(getattr(r, n)) for n in ("n_input", "n_binned", "n_dropped", "n_outside_tile", "n_columns", "n_voxels", "n_fitted", "n_slopes")}
    out = OracleMap(counts, vox, cols, ml, float(r.division_s), float(r.calculate_s), float(r.edges_s))
    getattr(lib, pre + "_free")(res)
    return out


def oracle_build(xyzw, params, mode="faithful32") -> OracleMap:
    """The plain-C restatement.  mode: 'faithful32' (the reference's arithmetic) or 'truth64'."""
    return _run("port", xyzw, params, {"faithful32": 0, "truth64": 1}[mode])


def ref_build(xyzw, params) -> OracleMap:
    """The reference's own receiver.cpp/map2D.h (compiled against shims).  In its output
    `first_index` fields are first-seen RANKS, not cloud indices (the reference does not
    keep point indices)."""
    return _run("ref", xyzw, params, 0)


def oracle_edges(o: OracleMap, params: Params):
    """(offsets[n_slopes + 1], targets): the AccessibleNeighbors lists of every Slope of an oracle
    build (map2D.h:530-548), slope index = rank of the SLOPE voxel in table order."""
    lib = _lib("port")
    fn = lib.gndt_oracle_edges
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(Params), C.POINTER(C.POINTER(C.c_uint32)),
                   C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    fn.restype = C.c_int
    off, tgt, ns, nt = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.c_size_t(), C.c_size_t()
    vox, cols = np.ascontiguousarray(o.voxels), np.ascontiguousarray(o.columns)
    rc = fn(vox.ctypes.data, len(vox), cols.ctypes.data, len(cols), C.byref(params), C.byref(off), C.byref(tgt), C.byref(ns), C.byref(nt))
    assert rc == 0, rc
    offsets = np.ctypeslib.as_array(off, (ns.value + 1,)).copy()
    targets = np.ctypeslib.as_array(tgt, (max(nt.value, 1),)).copy()[: nt.value]
    lib.gndt_oracle_free_edges.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.gndt_oracle_free_edges(off, tgt)
    return offsets, targets


# ---- key helpers -------------------------------------------------------------------------

def oracle_count_morton(a, b):
    lib = _lib("port")
    lib.gndt_oracle_count_morton.restype = C.c_int32
    return int(lib.gndt_oracle_count_morton(int(a), int(b)))


def oracle_morton_to_xy(m):
    lib = _lib("port")
    a, b = C.c_int(), C.c_int()
    lib.gndt_oracle_morton_to_xy(int(m), C.byref(a), C.byref(b))
    return a.value, b.value


def oracle_trans(origin, grid_len, z_len, pos):
    lib = _lib("port")
    o = (C.c_float * 3)(*[np.float32(v) for v in origin])
    p = (C.c_float * 3)(*[np.float32(v) for v in pos])
    sx, sy, sz = C.c_int32(), C.c_int32(), C.c_int32()
    lib.gndt_oracle_trans_morton_xyz.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib.gndt_oracle_trans_morton_xyz(o, np.float32(grid_len), np.float32(z_len), p, C.byref(sx), C.byref(sy), C.byref(sz))
    return rc, sx.value, sy.value, sz.value


def oracle_morton_string(sx, sy):
    lib = _lib("port")
    buf = C.create_string_buffer(32)
    lib.gndt_oracle_morton_string(int(sx), int(sy), buf)
    return buf.value.decode()


def ref_count_morton(a, b):
    lib = _lib("ref")
    buf = C.create_string_buffer(64)
    lib.gndt_ref_count_morton(int(a), int(b), buf)
    return buf.value.decode()


def ref_morton_to_xy(m):
    lib = _lib("ref")
    a, b = C.c_int(), C.c_int()
    lib.gndt_ref_morton_to_xy(int(m), C.byref(a), C.byref(b))
    return a.value, b.value


def ref_trans(origin, grid_len, z_len, pos):
    lib = _lib("ref")
    o = (C.c_float * 3)(*[np.float32(v) for v in origin])
    p = (C.c_float * 3)(*[np.float32(v) for v in pos])
    key = C.create_string_buffer(64)
    sz = C.c_int()
    lib.gndt_ref_trans_morton_xyz.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.gndt_ref_trans_morton_xyz(o, np.float32(grid_len), np.float32(z_len), p, key, C.byref(sz))
    return key.value.decode(), sz.value


def bridge_ground():
    """The genePcd.cpp fixture cloud: 360000 x (x,y,z,0) float32, zero tail included."""
    lib = _lib("port")
    pts = np.zeros((360000, 4), np.float32)
    lib.gndt_oracle_bridge_ground.argtypes = [C.c_void_p, C.c_size_t]
    lib.gndt_oracle_bridge_ground.restype = C.c_size_t
    n = lib.gndt_oracle_bridge_ground(pts.ctypes.data, pts.shape[0])
    return pts, int(n)
