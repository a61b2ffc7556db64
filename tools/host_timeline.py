"""Host-side timeline of one resident-input step (cfg2): where the wall time of encode()/decode() goes."""
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
def T(f, n=10):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = f()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, r
print("encode()", T(lambda: coder.encode(yd, prior=pd))[0])
print("decode()", T(lambda: coder.decode(bs, prior=pd))[0])
tg = coder._get_pgm(yd.shape, None)
print("_get_pgm", T(lambda: coder._get_pgm(yd.shape, None))[0])
print("_set_map", T(lambda: coder._set_map(tg))[0])
B, Cc, H, W = yd.shape
h = coder.ans_encoder.handle
ol = C.c_int64(0)
st = torch.cuda.current_stream(dev).cuda_stream
print("C encode (pinned out)", T(lambda: N.check(N.lib().basic_ypath_encode(h, coder._ctx, yd.data_ptr(), pd.data_ptr(), B, Cc, H, W, 0, None, 0, C.byref(ol), None, st)))[0])
print("last_output (bytes copy)", T(lambda: N.last_output(h))[0], ol.value)
enc = np.frombuffer(bs, dtype=np.uint8)
yh = torch.empty_like(yd)
print("C decode", T(lambda: N.check(N.lib().basic_ypath_decode(coder.ans_decoder.handle, coder._ctx, enc.ctypes.data, enc.size, pd.data_ptr(), B, Cc, H, W, 0, yh.data_ptr(), st)))[0])
N.profile(True); N.profile_read()
N.check(N.lib().basic_ypath_encode(h, coder._ctx, yd.data_ptr(), pd.data_ptr(), B, Cc, H, W, 0, None, 0, C.byref(ol), None, st))
print("enc phases", N.profile_read())
N.check(N.lib().basic_ypath_decode(coder.ans_decoder.handle, coder._ctx, enc.ctypes.data, enc.size, pd.data_ptr(), B, Cc, H, W, 0, yh.data_ptr(), st))
print("dec phases", N.profile_read())
ts = []
for _ in range(8):
    torch.cuda.synchronize(); t = time.perf_counter(); o = coder.decode(bs, prior=pd); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
print("decode() individually", [round(x, 2) for x in ts])
ts = []
for _ in range(8):
    torch.cuda.synchronize(); t = time.perf_counter(); b2 = coder.encode(yd, prior=pd); torch.cuda.synchronize(); ts.append((time.perf_counter() - t) * 1e3)
print("encode() individually", [round(x, 2) for x in ts])
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): o = coder.decode(bs, prior=pd)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
