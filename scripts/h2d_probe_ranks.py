"""Concurrent host->device ceiling: every rank uploads 178 MB clips from pinned memory at the same time (torchrun).
Prints one line per run: per-rank GB/s (min / mean over ranks) and the aggregate -- the e2e ceiling of bench.py at N ranks."""
import json, os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 178257920
pin = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
dev = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
back = torch.empty(2 * 1024 * 1024, dtype=torch.uint8).pin_memory()
small = torch.empty(2 * 1024 * 1024, dtype=torch.uint8, device="cuda")
streams = [torch.cuda.Stream() for _ in range(2)]
def run(reps):
    for i in range(reps):
        with torch.cuda.stream(streams[i % 2]):
            dev[i % 2].copy_(pin[i % 2], non_blocking=True)
            back.copy_(small, non_blocking=True)
    torch.cuda.synchronize()
run(4)
if world > 1: dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter(); run(20); dt = time.perf_counter() - t0
gbs = torch.tensor([n * 20 / dt / 1e9], device="cuda")
if world > 1:
    allg = [torch.zeros_like(gbs) for _ in range(world)]; dist.all_gather(allg, gbs); vals = [float(x) for x in allg]
else:
    vals = [float(gbs)]
if rank == 0:
    print(json.dumps(dict(ranks=world, h2d_gbs_per_rank_min=min(vals), h2d_gbs_per_rank_mean=sum(vals) / len(vals), h2d_gbs_aggregate=sum(vals),
                          frames_per_s_ceiling=sum(vals) * 1e9 / n * 2, bytes_per_clip=n)))
if world > 1: dist.destroy_process_group()
