"""Where the host time of an encode() call goes: the C call (queues everything, returns with the D2H copies in flight) against
the delivery into a bytes object (basic_coder_take_output) and into nothing (basic_coder_last_output: wait only)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from cbench_basic_b200 import _native as N
dev = torch.device("cuda", 0)
y, prior, w = bench.make_inputs("cfg2", 0)
coder = bench.build_coder("cfg2", w, 0, dev)
yd, pd = y.to(dev), prior.to(dev)
B, Cc, H, W = y.shape
coder._set_map(coder._get_pgm(y.shape, None))
h = coder.ans_encoder.handle
for mode in ("bytes", "view", "bytes", "view"):
    ts = []
    for _ in range(8):
        torch.cuda.synchronize()
        out_len = C.c_int64(0)
        t0 = time.perf_counter()
        N.check(N.lib().basic_ypath_encode(h, coder._ctx, yd.data_ptr(), pd.data_ptr(), B, Cc, H, W, 0, None, 0, C.byref(out_len), None, 0))
        t1 = time.perf_counter()
        if mode == "bytes":
            bs = N.last_output(h)
        else:
            bs = N.last_output_view(h)
        t2 = time.perf_counter()
        del bs
        t3 = time.perf_counter()
        ts.append((t1 - t0, t2 - t1, t3 - t2))
    ts = ts[2:]
    print(mode, "C call %.3f ms  delivery %.3f ms  free %.3f ms" % tuple(1e3 * sum(t[i] for t in ts) / len(ts) for i in range(3)))
