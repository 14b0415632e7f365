"""Micro-benchmark: ways to all-gather ~53 MB strips (uneven sizes) across the GPUs of one box."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
REC = 96
counts = [549226 + 137 * r for r in range(world)]
offs = [0]
for c in counts: offs.append(offs[-1] + c)
local_t = torch.full((counts[rank] * REC,), rank, dtype=torch.uint8, device=dev)
table = torch.empty(offs[-1] * REC, dtype=torch.uint8, device=dev)
vmax = max(counts)
padded_send = torch.zeros(vmax * REC, dtype=torch.uint8, device=dev)
padded_recv = torch.empty(world * vmax * REC, dtype=torch.uint8, device=dev)

def p2p():
    ops = []
    for r in range(world):
        if r != rank: ops.append(dist.P2POp(dist.irecv, table[offs[r] * REC: offs[r + 1] * REC], r))
    for r in range(world):
        if r != rank: ops.append(dist.P2POp(dist.isend, local_t, r))
    table[offs[rank] * REC: offs[rank + 1] * REC].copy_(local_t, non_blocking=True)
    for q in dist.batch_isend_irecv(ops): q.wait()

def padded():
    padded_send[: counts[rank] * REC].copy_(local_t, non_blocking=True)
    dist.all_gather_into_tensor(padded_recv, padded_send)
    for r in range(world):
        table[offs[r] * REC: offs[r + 1] * REC].copy_(padded_recv[r * vmax * REC: r * vmax * REC + counts[r] * REC], non_blocking=True)

def bcast():
    table[offs[rank] * REC: offs[rank + 1] * REC].copy_(local_t, non_blocking=True)
    hs = [dist.broadcast(table[offs[r] * REC: offs[r + 1] * REC], src=r, async_op=True) for r in range(world)]
    for h in hs: h.wait()

def ag_list():
    outs = [table[offs[r] * REC: offs[r + 1] * REC] for r in range(world)]
    dist.all_gather(outs, local_t)

for name, fn in (("p2p", p2p), ("padded_allgather+compact", padded), ("broadcasts", bcast), ("all_gather(list, uneven)", ag_list)):
    try:
        for _ in range(3): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ok = all(int(table[offs[r] * REC]) == r and int(table[offs[r + 1] * REC - 1]) == r for r in range(world))
        if rank == 0: print(f"{name:28s} {e0.elapsed_time(e1) / 10:.3f} ms/iter  ok={ok}  (recv {(offs[-1] - counts[rank]) * REC / 1e6:.0f} MB)", flush=True)
    except Exception as ex:
        if rank == 0: print(name, "failed:", str(ex)[:120], flush=True)
dist.destroy_process_group()
