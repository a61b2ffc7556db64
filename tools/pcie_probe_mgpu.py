"""torchrun --nproc-per-node N tools/pcie_probe_mgpu.py: pinned<->device copy rate of every rank while all ranks copy at
the same time (16 MB transfers, the size of a cfg2 stream), and the multi-threaded host memcpy rate beside them."""
import os, time, torch, numpy as np
import torch.distributed as dist
lr = int(os.environ.get("LOCAL_RANK", 0))
dev = torch.device("cuda", lr)
torch.cuda.set_device(dev)
dist.init_process_group("gloo")
MB = 1 << 20
n = 16 * MB
src = torch.empty(n, dtype=torch.uint8).pin_memory(); src.fill_(1)
back = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device=dev)
page = np.ones(n, dtype=np.uint8); page2 = np.empty(n, dtype=np.uint8)
def rate(fn, secs=1.5):
    torch.cuda.synchronize(); dist.barrier()
    t = time.perf_counter(); k = 0
    while time.perf_counter() - t < secs:
        fn(); k += 1
    torch.cuda.synchronize()
    return k * n / (time.perf_counter() - t) / 1e9
def h2d(): dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
def d2h(): back.copy_(dst, non_blocking=True); torch.cuda.synchronize()
def both():
    dst.copy_(src, non_blocking=True)
    with torch.cuda.stream(side): back.copy_(dst2, non_blocking=True)
    torch.cuda.synchronize()
def hcopy(): page2[:] = page
side = torch.cuda.Stream(); dst2 = torch.empty_like(dst)
res = {"h2d": rate(h2d), "d2h": rate(d2h), "h2d+d2h": 2 * rate(both), "host_memcpy_1thread": rate(hcopy)}
out = [None] * dist.get_world_size()
dist.all_gather_object(out, {k: round(v, 1) for k, v in res.items()})
if dist.get_rank() == 0:
    print("world", dist.get_world_size(), "GB/s per rank:")
    for r, o in enumerate(out): print(" rank", r, o)
