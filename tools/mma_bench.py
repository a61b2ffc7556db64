"""Cycles per tcgen05.mma (M = 128) on one SM, issued back to back by one warp: shapes the context kernel uses or could use."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cbench_basic_b200 import _native as N
import torch
torch.cuda.init()
iters = 512
for mode, name in ((1, "f16 K=16"), (0, "tf32 K=8")):
    for ts in (1, 0):
        for n in (64, 128, 256):
            for same in (1, 0):
                if not same and 2 * n > 448:
                    continue
                cyc = (C.c_longlong * 2)()
                N.check(N.lib().basic_debug_mma_bench(mode, ts, n, iters, same, cyc))
                print(f"{name:9s} A-in-{'TMEM' if ts else 'smem'} N={n:3d} {'one accumulator ' if same else 'two accumulators'}: "
                      f"issue {cyc[0] / iters:6.1f}  total {cyc[1] / iters:6.1f} cycles/MMA")
