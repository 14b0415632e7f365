"""Multi-GPU map build: one x strip of the map per GPU of a single box, one process per GPU.

Everything up to and including the surface labels depends only on the points of one x-y
column, so strips need no exchange while they are built.  NCCL is used
once per build, to all-gather the finished strip tables so that every rank (and the host
planner behind it) holds the whole map; the neighbour-reachability bits of the columns on
strip boundaries are then recomputed against the gathered table (their 1-cell halo).
torch.distributed is only the plumbing (process group, streams, device buffers).
"""
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from ._abi import VOXEL_DTYPE
from .builder import TwoDmap

REC = VOXEL_DTYPE.itemsize  # 96


class TiledTwoDmap:
    """Strip r of `world` strips.  `cuts` are contiguous signed x column indices; strip r
    keeps columns with cuts[r] <= cx < cuts[r+1] (gndt_params.tile_lo/hi)."""

    def __init__(self, res, zres, interval, rank: int, world: int, device: Optional[int] = None, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.map = TwoDmap(res, zres, device=device)
        self.map.setInterval(interval)
        self.device = torch.device("cuda", self.map._device)
        self._gathered = None
        self.offsets = None

    def plan(self, cloud, origin=None) -> np.ndarray:
        """Balanced cuts from a histogram over x columns.  Every rank that sees the same
        cloud computes the same cuts; no communication."""
        if origin is not None:
            self.map.setCloudFirst(origin)
        return self.map.plan_tiles(cloud, self.world)

    def build(self, cloud, demand="slope", origin=None, cuts=None, filter_points=True):
        """Build this rank's strip and all-gather the strips.  `cloud` may be the whole
        cloud (filter_points=True: points of other strips are dropped on the device) or
        only this strip's share.  Returns (gathered table tensor [V_total, 96] uint8,
        offsets[world+1])."""
        m = self.map
        if origin is not None:
            m.setCloudFirst(origin)
        if cuts is not None and filter_points:
            m.setTile(int(cuts[self.rank]), int(cuts[self.rank + 1]))
        else:
            m.setTile(0, 0)
        m.uniformDivision(cloud)
        m.create2DMap(demand)
        ptr, n = m.device_voxels()
        local = _as_tensor(ptr, max(n, 1) * REC, self.device)[: n * REC]
        # (1)+(2) strip sizes, then a padded all-gather of the finished records
        table, self.offsets = allgather_strips(local, self.world, self.group)
        counts = np.diff(self.offsets)
        total, vmax = int(self.offsets[-1]), int(counts.max())
        # (3) halo: reachability bits of this strip against the whole map
        st = torch.cuda.current_stream(self.device).cuda_stream
        m.label_edges(table.data_ptr(), total, int(self.offsets[self.rank]), n, st)
        # (4) publish the refreshed flag words of this strip (4 B per voxel)
        tv = table.view(torch.int32).view(-1, REC // 4)[:total]
        my_flags = torch.zeros(max(vmax, 1), dtype=torch.int32, device=self.device)
        my_flags[:n] = tv[self.offsets[self.rank]: self.offsets[self.rank] + n, 21]
        all_flags = torch.empty(self.world * max(vmax, 1), dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(all_flags, my_flags, group=self.group)
        for r in range(self.world):
            c = int(counts[r])
            tv[self.offsets[r]: self.offsets[r] + c, 21] = all_flags[r * vmax: r * vmax + c]
        self._gathered = table
        return table[: total * REC].view(-1, REC), self.offsets

    def gathered_numpy(self) -> np.ndarray:
        total = int(self.offsets[-1])
        return self._gathered[: total * REC].cpu().numpy().view(VOXEL_DTYPE).reshape(-1)

    def close(self):
        self.map.close()


def allgather_strips(local: torch.Tensor, world: int, group=None):
    """All-gather variable-length strip tables (flat uint8, 96 B records) into one compact
    table in rank order.  Device-agnostic: NCCL on CUDA tensors, gloo on CPU tensors (the
    CPU form is what the world_size-2 unit test exercises).  Returns (table, offsets)."""
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
