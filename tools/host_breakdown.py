"""Where a resident-input step goes on the host side: C call vs byte-string plumbing (cfg2)."""
import os, sys, time, ctypes as C, numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from cbench_basic_b200 import _native as N
dev = torch.device("cuda", 0)
y, prior, w = bench.make_inputs("cfg2", 0)
coder = bench.build_coder("cfg2", w, 0, dev)
yd, pd = y.to(dev), prior.to(dev)
for _ in range(3):
    bs = coder.encode(yd, prior=pd); out = coder.decode(bs, prior=pd)
torch.cuda.synchronize()
def T(f, n=5):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, r
t_enc, bs = T(lambda: coder.encode(yd, prior=pd))
t_dec, _ = T(lambda: coder.decode(bs, prior=pd))
B, Cc, H, W = yd.shape
h = coder.ans_encoder.handle
cap = int(N.lib().basic_ypath_encode_bound(h, B, Cc, H, W, 0))
outb = np.empty(cap, dtype=np.uint8); ol = C.c_int64(0)
coder._set_map(coder._get_pgm(yd.shape, None))
t_c, _ = T(lambda: N.check(N.lib().basic_ypath_encode(h, coder._ctx, yd.data_ptr(), pd.data_ptr(), B, Cc, H, W, 0, outb.ctypes.data, cap, C.byref(ol), None, 0)))
t_tb, _ = T(lambda: outb[:ol.value].tobytes())
enc = np.frombuffer(bs, dtype=np.uint8)
yh = torch.empty_like(yd)
t_d, _ = T(lambda: N.check(N.lib().basic_ypath_decode(coder.ans_decoder.handle, coder._ctx, enc.ctypes.data, enc.size, pd.data_ptr(), B, Cc, H, W, 0, yh.data_ptr(), 0)))
print(f"encode() {t_enc:.2f} ms = C call {t_c:.2f} + tobytes {t_tb:.2f} (+ np.empty etc); decode() {t_dec:.2f} ms = C call {t_d:.2f}; stream {len(bs)/1e6:.1f} MB")
