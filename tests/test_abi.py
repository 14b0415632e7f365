"""The C-ABI shared library: loads, exports every symbol include/gndt.h declares, record
layouts match the header, host helpers agree with the oracle, and — without a GPU — the
product path fails loudly instead of falling back to anything."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import grid_ndt_b200 as g
from grid_ndt_b200 import _abi, _lib
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gndt.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gndt_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 20
    L = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/gndt.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "python binding table out of sync with the header"
    assert g.lib().gndt_version().startswith(b"gndt")


def test_record_layouts_match_header():
    hdr = open(os.path.join(ROOT, "include", "gndt.h")).read()
    src = "#include <stdio.h>\n#include <stddef.h>\n#include \"gndt.h\"\nint main(){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(gndt_voxel), sizeof(gndt_slope), sizeof(gndt_column), sizeof(gndt_params), offsetof(gndt_voxel, flags), offsetof(gndt_voxel, scatter), offsetof(gndt_params, max_voxels), sizeof(gndt_counts_t)); return 0;}\n"
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I" + os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        out = subprocess.check_output([os.path.join(d, "t")]).split()
    sizes = [int(x) for x in out]
    assert sizes[0] == _abi.VOXEL_DTYPE.itemsize == 96
    assert sizes[1] == _abi.SLOPE_DTYPE.itemsize == 48
    assert sizes[2] == _abi.COLUMN_DTYPE.itemsize == 32
    assert sizes[3] == C.sizeof(_abi.Params)
    assert sizes[4] == _abi.VOXEL_DTYPE.fields["flags"][1] == 84
    assert sizes[5] == _abi.VOXEL_DTYPE.fields["scatter"][1]
    assert sizes[6] == _abi.Params.max_voxels.offset
    assert sizes[7] == C.sizeof(_abi.Counts)
    assert "GNDT_ABI_VERSION 2" in hdr


def test_host_key_helpers_match_oracle():
    rng = np.random.default_rng(3)
    L = g.lib()
    for _ in range(2000):
        a, b = (int(x) for x in rng.integers(1, 32768, 2))
        assert L.gndt_count_morton(a, b) == O.oracle_count_morton(a, b)
        x, y = C.c_uint32(), C.c_uint32()
        L.gndt_morton_to_xy(L.gndt_count_morton(a, b), C.byref(x), C.byref(y))
        assert (x.value, y.value) == (a, b)
        sx, sy = (int(v) for v in rng.choice([-1, 1], 2) * (a, b))
        assert g.morton_string(sx, sy) == O.oracle_morton_string(sx, sy)
    sx = rng.integers(-500, 500, 1000); sx[sx == 0] = 1
    sy = rng.integers(-500, 500, 1000); sy[sy == 0] = -1
    assert list(g.morton_strings(sx, sy)) == [O.oracle_morton_string(a, b) for a, b in zip(sx, sy)]
    for _ in range(2000):
        o = rng.uniform(-50, 50, 3).astype(np.float32)
        pos = (o + rng.uniform(-30, 30, 3)).astype(np.float32)
        oo, pp = (C.c_float * 3)(*o), (C.c_float * 3)(*pos)
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        assert L.gndt_trans_morton_xyz(oo, np.float32(0.1), np.float32(0.05), pp, C.byref(a), C.byref(b), C.byref(c)) == 0
        assert (0, a.value, b.value, c.value) == O.oracle_trans(o, 0.1, 0.05, pos)


def test_default_params_match_reference_defaults():
    p = _abi.Params()
    g.lib().gndt_default_params(C.byref(p))
    q = _abi.default_params()
    for f, _ in _abi.Params._fields_:
        if f != "origin":
            assert getattr(p, f) == getattr(q, f), f
    assert (p.min_points, p.rough_max, p.angle_max_deg) == (3, 100.0, 30.0) and abs(p.reach_height - 0.15) < 1e-7


def _no_cuda():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_cuda(), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly_no_fallback():
    with pytest.raises(g.GndtError) as e:
        g.TwoDmap(0.2, 0.1)
    assert e.value.status == _abi.GNDT_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "grid_ndt_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"libgndt_oracle", r"libgndt_ref", r"gndt_oracle_", r"gndt_ref_", r"oracle[/\\.]"):
                    assert not re.search(pat, txt, flags=re.M), f"product file {f} references the oracle ({pat})"


def test_create_rejects_bad_params():
    L = g.lib()
    h = C.c_void_p()
    p = _abi.default_params(0.0, 0.1)
    assert L.gndt_create(C.byref(p), 0, C.byref(h)) == _abi.GNDT_ERR_INVALID_ARG
    assert L.gndt_create(None, 0, C.byref(h)) == _abi.GNDT_ERR_INVALID_ARG
    assert L.gndt_build(None, None, 0, 16, 0, None) == _abi.GNDT_ERR_INVALID_ARG
    assert L.gndt_counts(None, None) == _abi.GNDT_ERR_INVALID_ARG
    # the strip exchange (both transports) refuses a missing handle before touching the device
    for name in ("gndt_xchg_run", "gndt_xchg_stage", "gndt_xchg_send"):
        assert getattr(L, name)(None, None) == _abi.GNDT_ERR_INVALID_ARG, name
    assert L.gndt_xchg_counts_ready(None) == _abi.GNDT_ERR_INVALID_ARG
    assert L.gndt_xchg_view_get(None, None) == _abi.GNDT_ERR_INVALID_ARG
