"""BASIC_TRACE=1 python tools/trace_step.py: host / device timeline of the y-path calls of a cfg2 step (stderr)."""
import os, sys, time, torch
os.environ.setdefault("BASIC_TRACE", "1")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
dev = torch.device("cuda", 0)
y, prior, w = bench.make_inputs("cfg2", 0)
coder = bench.build_coder("cfg2", w, 0, dev)
yd, pd = y.to(dev), prior.to(dev)
yp, pp = y.pin_memory(), prior.pin_memory()
for i in range(4):
    sys.stderr.write(f"--- resident step {i}\n")
    torch.cuda.synchronize(); t = time.perf_counter()
    bs = coder.encode(yd, prior=pd)
    t1 = time.perf_counter()
    out = coder.decode(bs, prior=pd)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    sys.stderr.write(f"encode() {1e3*(t1-t):.3f} ms  decode() {1e3*(t2-t1):.3f} ms\n")
for i in range(3):
    sys.stderr.write(f"--- host-input step {i}\n")
    torch.cuda.synchronize(); t = time.perf_counter()
    bs = coder.encode(yp, prior=pp)
    t1 = time.perf_counter()
    out = coder.decode(bs, prior=pp)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    sys.stderr.write(f"encode() {1e3*(t1-t):.3f} ms  decode() {1e3*(t2-t1):.3f} ms\n")
