"""One process driving several GPUs through gndt_multi_* (include/gndt.h): the shape of the
reference's receiver, a single C++ process (src/receiver.cpp:283), with N B200s behind it.
Everything happens inside libgndt.so — strip planning, the cloud's fan-out, the strip builds
on one stream per device, the peer-memory exchange; this file only marshals the call."""
import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _abi
from ._abi import COLUMN_DTYPE, SLOPE_DTYPE, VOXEL_DTYPE, XchgView, default_params
from ._lib import GndtError, lib

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class MultiTwoDmap:
    def __init__(self, res: float, zres: float, interval: float, devices: Sequence[int], capacity: int = 8_000_000,
                 halo_records: int = 65536, gather=("voxels", "slopes", "columns"), demand: str = "slope"):
        self._p = default_params(res, zres, interval, demand)
        self.devices = list(devices)
        self.what = sum({"voxels": _abi.X_VOXELS, "slopes": _abi.X_SLOPES, "columns": _abi.X_COLUMNS}[g] for g in gather)
        self._m = C.c_void_p()
        devs = (C.c_int * len(self.devices))(*self.devices)
        rc = lib().gndt_multi_create(C.byref(self._p), devs, len(self.devices), int(capacity), int(halo_records), self.what, C.byref(self._m))
        if rc != 0:
            raise GndtError(rc, lib().gndt_multi_last_error(None).decode())
        self._keep = None

    def _check(self, rc):
        if rc != 0:
            raise GndtError(rc, lib().gndt_multi_last_error(self._m).decode())

    def _marshal(self, cloud):
        if torch is not None and isinstance(cloud, torch.Tensor):
            if cloud.is_cuda and cloud.device.index != self.devices[0]:
                raise GndtError(-1, "a device cloud must live on devices[0]")
            return cloud.data_ptr(), cloud.shape[0], cloud.shape[1] * 4, _abi.GNDT_MEM_DEVICE if cloud.is_cuda else _abi.GNDT_MEM_HOST, cloud
        arr = np.ascontiguousarray(cloud, dtype=np.float32)
        return arr.ctypes.data, arr.shape[0], arr.shape[1] * 4, _abi.GNDT_MEM_HOST, arr

    def chatterCallback(self, cloud):
        """One cloud -> the whole map on every GPU (asynchronous; view() waits)."""
        ptr, n, stride, mem, keep = self._marshal(cloud)
        if mem == _abi.GNDT_MEM_DEVICE:
            torch.cuda.current_stream(cloud.device).synchronize()  # the library works on its own streams
        self._check(lib().gndt_multi_build(self._m, ptr, n, stride, mem))
        self._keep = keep

    def change2DMap(self, scan):
        ptr, n, stride, mem, keep = self._marshal(scan)
        if mem == _abi.GNDT_MEM_DEVICE:
            torch.cuda.current_stream(scan.device).synchronize()
        self._check(lib().gndt_multi_update(self._m, ptr, n, stride, mem))
        self._keep = keep

    def view(self, index: int = 0) -> XchgView:
        v = XchgView()
        self._check(lib().gndt_multi_view(self._m, index, C.byref(v)))
        return v

    def tables(self, index: int = 0) -> dict:
        """Host copies of the gathered tables as they lie on GPU `index`."""
        v = self.view(index)
        out = {"strip_voxels": list(v.strip_voxels[: v.world])}
        with torch.cuda.device(self.devices[index]):
            for name, ptr, n, dt in (("voxels", v.voxels, v.n_voxels, VOXEL_DTYPE), ("slopes", v.slopes, v.n_slopes, SLOPE_DTYPE),
                                     ("columns", v.columns, v.n_columns, COLUMN_DTYPE)):
                if not ptr:
                    continue
                if n:
                    from .tiles import _as_tensor
                    dev = torch.device("cuda", self.devices[index])
                    out[name] = _as_tensor(ptr, int(n) * dt.itemsize, dev).cpu().numpy().view(dt)
                else:
                    out[name] = np.zeros(0, dt)
        return out

    def cuts(self) -> np.ndarray:
        c = (C.c_int32 * (len(self.devices) + 1))()
        self._check(lib().gndt_multi_cuts(self._m, c, len(c)))
        return np.array(c[:], np.int32)

    def close(self):
        if self._m:
            lib().gndt_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
