"""Host-to-device bandwidth per GPU when all ranks upload at once (the e2e arm's limiter at N = 8)."""
import json, os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
host = torch.empty(160_000_000, dtype=torch.uint8).pin_memory(); d = torch.empty_like(host, device=dev)
back = torch.empty(33_000_000, dtype=torch.uint8).pin_memory()
for mode in ("h2d", "h2d+d2h"):
    for _ in range(3): d.copy_(host, non_blocking=True)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    s2 = torch.cuda.Stream()
    for _ in range(10):
        d.copy_(host, non_blocking=True)
        if mode == "h2d+d2h":
            with torch.cuda.stream(s2): back.copy_(d[:33_000_000], non_blocking=True)
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / 10
    t = torch.tensor([dt], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(json.dumps({"world": world, "mode": mode, "h2d_gbs_per_gpu": 0.16 / float(t.item()), "aggregate_gbs": 0.16 * world / float(t.item())}), flush=True)
dist.destroy_process_group()
