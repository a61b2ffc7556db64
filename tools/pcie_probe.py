"""PCIe probe: pinned <-> device copy rates for whole and chunked transfers, with and without a concurrent host memcpy."""
import torch, time, threading, numpy as np
dev = torch.device("cuda", 0)
MB = 1 << 20
n = 16 * MB
src = torch.empty(n, dtype=torch.uint8).pin_memory(); src.fill_(3)
dst = torch.empty(n, dtype=torch.uint8, device=dev)
back = torch.empty(n, dtype=torch.uint8).pin_memory()
side = torch.cuda.Stream()
def t_copy(fn, reps=5):
    res = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1))
    return min(res), sorted(res)[len(res)//2]
def h2d(chunk):
    def f():
        for k in range(0, n, chunk): dst[k:k+chunk].copy_(src[k:k+chunk], non_blocking=True)
    return f
def d2h(chunk):
    def f():
        for k in range(0, n, chunk): back[k:k+chunk].copy_(dst[k:k+chunk], non_blocking=True)
    return f
for chunk in (16*MB, 4*MB, MB, 256*1024):
    print("H2D chunk", chunk // 1024, "KB: ms min/med", t_copy(h2d(chunk)), " D2H", t_copy(d2h(chunk)))
# with a host thread writing the pinned source just before each chunk goes out (the staging pattern)
page = np.frombuffer(bytes(n), dtype=np.uint8)
srcn = src.numpy()
def staged(chunk):
    def f():
        for k in range(0, n, chunk):
            srcn[k:k+chunk] = page[k:k+chunk]
            dst[k:k+chunk].copy_(src[k:k+chunk], non_blocking=True)
    return f
for chunk in (4*MB, MB):
    t = time.perf_counter(); r = t_copy(staged(chunk)); 
    print("staged H2D chunk", chunk // 1024, "KB: ms min/med", r)
t = time.perf_counter()
for _ in range(5): srcn[:] = page
print("host memcpy 16 MB single thread ms", (time.perf_counter() - t) / 5 * 1e3)
