"""Host <-> device copy bandwidth per rank with all ranks copying at once: H2D alone, D2H alone, both directions together."""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
n = 1 << 29                                  # 2 GiB of int32
hin, hout = torch.empty(n, dtype=torch.int32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()
din, dout = torch.empty(n, dtype=torch.int32, device="cuda"), torch.zeros(n, dtype=torch.int32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    if h2d:
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier()
    return time.perf_counter() - t0
for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
    run(a, b)
    t = min(run(a, b) for _ in range(3))
    gb = (a + b) * n * 4 / 1e9
    if rank == 0:
        print(f"{world} ranks {name}: {t*1e3:.1f} ms per 2 GiB per direction -> {gb/t:.1f} GB/s per GPU, {gb*world/t:.1f} GB/s aggregate", flush=True)
dist.destroy_process_group()
