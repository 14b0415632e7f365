"""Multi-GPU map build: one x strip of the map per GPU of a single box, one process per GPU.

Everything up to and including the surface labels depends only on the points of one x-y
column, so strips need no exchange while they are built.  The only cross-
strip dependence is the forward/back reachability of the first and last x row of a strip, so
those two rows (a thin halo, fixed-size buffers, no host round trip) are swapped with the two
neighbour strips BEFORE the gather and each rank finishes its own boundary rows.  Then the
strip sizes are all-gathered (the only host synchronisation of a build) and every rank sends
its final records straight out of the builder's table into the other ranks' copy of the whole
map (batched point-to-point = an all-gather with uneven sizes, no padding, no staging copies;
measured faster than padded all_gather / broadcasts, tools/gather_bench.py).  Last, strip-local
column / slope indices are made global by adding per-strip offsets.  torch.distributed is only
the plumbing (process group, streams, device buffers).
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from ._abi import VOXEL_DTYPE
from ._lib import lib
from .builder import TwoDmap, _check

REC = VOXEL_DTYPE.itemsize  # 96


class TiledTwoDmap:
    """Strip r of `world` strips.  `cuts` are contiguous signed x column indices; strip r
    keeps columns with cuts[r] <= cx < cuts[r+1] (gndt_params.tile_lo/hi)."""

    def __init__(self, res, zres, interval, rank: int, world: int, device: Optional[int] = None, group=None,
                 halo_records: int = 32768):
        self.rank, self.world, self.group = rank, world, group
        self.map = TwoDmap(res, zres, device=device)
        self.map.setInterval(interval)
        self.device = torch.device("cuda", self.map._device)
        self._counts = torch.zeros(world * 4, dtype=torch.int32, device=self.device)
        self._table = None
        self._halo = None
        self.halo_records = halo_records  # capacity of one halo row (records); overflow is reported, not truncated
        self.offsets = None

    def plan(self, cloud, origin=None) -> np.ndarray:
        """Balanced cuts from a histogram over x columns.  Every rank that sees the same
        cloud computes the same cuts; no communication."""
        if origin is not None:
            self.map.setCloudFirst(origin)
        return self.map.plan_tiles(cloud, self.world)

    def build(self, cloud, demand="slope", origin=None, cuts=None, filter_points=True):
        """Build this rank's strip and assemble the whole map on every rank.  `cloud` may be
        the whole cloud (filter_points=True: points of other strips are dropped on the
        device) or only this strip's share.  Returns (table tensor [V_total, 96] uint8 on
        this rank's GPU, offsets[world+1]).

        Order of work (everything stream-ordered except the one size exchange):
          1. local build of the strip (no communication: labels are per column)
          2. thin halo: first / last x row swapped with the two neighbour strips, the strip's
             own boundary rows get their cross-strip forward/back reachability bits
          3. strip sizes all-gathered (the only host synchronisation)
          4. the now-final records go straight from the builder's table into every rank's
             copy of the map (batched point-to-point = all-gather with uneven sizes)
          5. strip-local column / slope indices become global (add per-strip offsets)"""
        m, L = self.map, lib()
        if origin is not None:
            m.setCloudFirst(origin)
        if cuts is not None and filter_points:
            m.setTile(int(cuts[self.rank]), int(cuts[self.rank + 1]))
        else:
            m.setTile(0, 0)
        m.uniformDivision(cloud)
        m.create2DMap(demand)  # (1) asynchronous
        st = torch.cuda.current_stream(self.device).cuda_stream
        # (2) halo rows: fixed-size buffers so that no size has to be known on the host
        slot = (self.halo_records + 1) * REC
        if self._halo is None:
            self._halo = torch.zeros(4 * slot, dtype=torch.uint8, device=self.device)
        send_first, send_last, recv_prev, recv_next = (self._halo[i * slot:(i + 1) * slot] for i in range(4))
        _check(m._h, L.gndt_halo_pack(m._h, send_first.data_ptr(), send_last.data_ptr(), self.halo_records, st))
        ops = []
        if self.rank > 0:
            ops += [dist.P2POp(dist.isend, send_first, self.rank - 1, group=self.group),
                    dist.P2POp(dist.irecv, recv_prev, self.rank - 1, group=self.group)]
        if self.rank + 1 < self.world:
            ops += [dist.P2POp(dist.isend, send_last, self.rank + 1, group=self.group),
                    dist.P2POp(dist.irecv, recv_next, self.rank + 1, group=self.group)]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        _check(m._h, L.gndt_halo_edges(m._h, recv_prev.data_ptr() if self.rank > 0 else None,
                                       recv_next.data_ptr() if self.rank + 1 < self.world else None, st))
        # (3) strip sizes: {n_voxels, n_columns, n_slopes, n_fitted} of every strip
        p_cnt, p_tab, cap = C.c_void_p(), C.c_void_p(), C.c_size_t()
        _check(m._h, L.gndt_device_count_ptr(m._h, C.byref(p_cnt)))
        _check(m._h, L.gndt_device_table_ptr(m._h, C.byref(p_tab), C.byref(cap)))
        mine = _as_tensor(p_cnt.value, 16, self.device).view(torch.int32)
        dist.all_gather_into_tensor(self._counts, mine, group=self.group)
        counts4 = self._counts.cpu().numpy().astype(np.int64).reshape(self.world, 4)
        counts = counts4[:, 0]
        self.offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
        col_off = np.concatenate([[0], np.cumsum(counts4[:, 1])])[:-1].astype(np.uint32)
        slope_off = np.concatenate([[0], np.cumsum(counts4[:, 2])])[:-1].astype(np.uint32)
        total, n = int(self.offsets[-1]), int(counts[self.rank])
        # (4) records: straight from the builder's table into every rank's map
        if self._table is None or self._table.numel() < max(total, 1) * REC:
            self._table = torch.empty(int(max(total, 1) * 1.25) * REC, dtype=torch.uint8, device=self.device)
        table = self._table
        local = _as_tensor(p_tab.value, max(n, 1) * REC, self.device)[: n * REC]
        lo = int(self.offsets[self.rank]) * REC
        ops = []
        for r in range(self.world):
            if r == self.rank or counts[r] == 0:
                continue
            ops.append(dist.P2POp(dist.irecv, table[int(self.offsets[r]) * REC: int(self.offsets[r + 1]) * REC], r, group=self.group))
        if n:
            for r in range(self.world):
                if r != self.rank:
                    ops.append(dist.P2POp(dist.isend, local, r, group=self.group))
            table[lo: lo + n * REC].copy_(local, non_blocking=True)
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        # (5) global column / slope indices
        off = (C.c_uint64 * (self.world + 1))(*[int(x) for x in self.offsets])
        co = (C.c_uint32 * self.world)(*[int(x) for x in col_off])
        so = (C.c_uint32 * self.world)(*[int(x) for x in slope_off])
        _check(m._h, L.gndt_apply_strip_offsets(m._h, table.data_ptr(), off, co, so, self.world, st))
        self.strip_counts = counts4
        return table[: total * REC].view(-1, REC), self.offsets

    def gathered_numpy(self) -> np.ndarray:
        total = int(self.offsets[-1])
        return self._table[: total * REC].cpu().numpy().view(VOXEL_DTYPE).reshape(-1)

    def close(self):
        self.map.close()


def allgather_strips(local: torch.Tensor, world: int, group=None):
    """All-gather variable-length strip tables (flat uint8, 96 B records) into one compact
    table in rank order with collectives only (sizes, then padded records).  Device-agnostic:
    NCCL on CUDA tensors, gloo on CPU tensors — the CPU form is what the world_size-2 unit test
    exercises; TiledTwoDmap.build uses the unpadded point-to-point form of the same exchange.
    Returns (table, offsets)."""
    dev = local.device
    n = local.numel() // REC
    cnt = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(cnt, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
    counts = cnt.cpu().numpy()
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    total, vmax = int(offsets[-1]), max(int(counts.max()), 1)
    send = torch.zeros(vmax * REC, dtype=torch.uint8, device=dev)
    send[: n * REC] = local
    recv = torch.empty(world * vmax * REC, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    table = torch.empty(max(total, 1) * REC, dtype=torch.uint8, device=dev)
    for r in range(world):
        c = int(counts[r])
        table[offsets[r] * REC: (offsets[r] + c) * REC] = recv[r * vmax * REC: r * vmax * REC + c * REC]
    return table, offsets


def _as_tensor(ptr: int, nbytes: int, device) -> torch.Tensor:
    """Zero-copy uint8 view of library-owned device memory (valid until the next build)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    return torch.as_tensor(h, device=device)
