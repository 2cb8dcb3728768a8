"""Where the CSR all-gather's time goes (torchrun, N ranks): NCCL all-gather variants on 66 M int32 per rank."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 65_700_000 + 1000 * rank            # nearly even, like the CSR pieces of a balanced range partition
arr = torch.arange(n, dtype=torch.int32, device="cuda")
sizes = [65_700_000 + 1000 * r for r in range(world)]
mx = max(sizes)

def timed(name, fn, reps=4):
    ts = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    if rank == 0:
        print(f"{name:44s} best {min(ts)*1e3:7.2f} ms  last {ts[-1]*1e3:7.2f} ms  ({sum(sizes)*4/min(ts)/1e9:6.1f} GB/s gathered)", flush=True)

def even():
    out = torch.empty(world * mx, dtype=torch.int32, device="cuda")
    src = torch.empty(mx, dtype=torch.int32, device="cuda"); src[:n].copy_(arr)
    dist.all_gather_into_tensor(out, src)
def even_prealloc(out=torch.empty(world * mx, dtype=torch.int32, device="cuda"), src=torch.empty(mx, dtype=torch.int32, device="cuda")):
    dist.all_gather_into_tensor(out, src)
def inplace_padded():
    pad = torch.empty(world * mx, dtype=torch.int32, device="cuda")
    src = pad[rank * mx:(rank + 1) * mx]; src[:n].copy_(arr)
    dist.all_gather_into_tensor(pad, src)
    out = torch.empty(sum(sizes), dtype=torch.int32, device="cuda")
    torch.cat([pad[j * mx:j * mx + sizes[j]] for j in range(world)], out=out)
def uneven_list():
    out = torch.empty(sum(sizes), dtype=torch.int32, device="cuda")
    offs = [0]
    for s in sizes: offs.append(offs[-1] + s)
    dist.all_gather([out[offs[j]:offs[j + 1]] for j in range(world)], arr)
def cat_only(pad=torch.empty(world * mx, dtype=torch.int32, device="cuda")):
    out = torch.empty(sum(sizes), dtype=torch.int32, device="cuda")
    torch.cat([pad[j * mx:j * mx + sizes[j]] for j in range(world)], out=out)
def alloc_only():
    a = torch.empty(world * mx, dtype=torch.int32, device="cuda"); b = torch.empty(sum(sizes), dtype=torch.int32, device="cuda"); return a, b
def small_ints():
    t = torch.tensor([n], dtype=torch.int64).cuda(); o = torch.empty((world, 1), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(o, t); return o.cpu()
for name, fn in (("all_gather_into_tensor, preallocated", even_prealloc), ("all_gather_into_tensor + alloc + copy", even),
                 ("in-place padded + cat (distributed.py)", inplace_padded), ("all_gather(list of uneven views)", uneven_list),
                 ("cat only", cat_only), ("alloc only", alloc_only), ("small int all-gather + .cpu()", small_ints)):
    timed(name, fn)
dist.destroy_process_group()
