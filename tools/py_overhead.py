"""Python-side overhead of encode()/decode() around the C calls (cfg2)."""
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
L = N.lib()
orig_enc, orig_dec, orig_take = L.basic_ypath_encode, L.basic_ypath_decode, L.basic_coder_take_output
acc = {"enc_c": 0.0, "dec_c": 0.0, "take_c": 0.0}
def wrap(fn, key):
    def f(*a):
        t = time.perf_counter(); r = fn(*a); acc[key] += time.perf_counter() - t; return r
    return f
L.basic_ypath_encode = wrap(orig_enc, "enc_c"); L.basic_ypath_decode = wrap(orig_dec, "dec_c"); L.basic_coder_take_output = wrap(orig_take, "take_c")
n = 20
N.profile(True); N.profile_read()
torch.cuda.synchronize(); t0 = time.perf_counter(); te = 0.0; td = 0.0
for _ in range(n):
    t = time.perf_counter(); bs = coder.encode(yd, prior=pd); te += time.perf_counter() - t
    t = time.perf_counter(); out = coder.decode(bs, prior=pd); td += time.perf_counter() - t
torch.cuda.synchronize(); tot = time.perf_counter() - t0
ph = N.profile_read(); N.profile(False)
print({k: round(v[0] / n, 3) for k, v in ph.items()})
print(f"step {tot / n * 1e3:.3f} ms: encode() {te / n * 1e3:.3f} = C {acc['enc_c'] / n * 1e3:.3f} + take {acc['take_c'] / n * 1e3:.3f} + python {(te - acc['enc_c'] - acc['take_c']) / n * 1e3:.3f};"
      f" decode() {td / n * 1e3:.3f} = C {acc['dec_c'] / n * 1e3:.3f} + python {(td - acc['dec_c']) / n * 1e3:.3f}")
