"""Wall time of the z-node coder on the cfg2 geometry (24 images, 192 channels, 8x12 hyper-latents)."""
import os, sys, time, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from cbench_basic_b200 import z_coder
from oracle import z_oracle as Z
C_, B, H, W = 192, 24, 8, 12
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
coder = z_coder.CompressAIEntropyBottleneckPriorCoder(entropy_bottleneck_channels=C_, lanes=lanes)
eb = coder.entropy_bottleneck
eb.load_state_dict(dict(Z.init_params(C_, seed=3), target=eb.target.clone(), _offset=torch.IntTensor(),
                        _quantized_cdf=torch.IntTensor(), _cdf_length=torch.IntTensor()))
t = time.perf_counter(); coder.update_state(); print("update_state ms", (time.perf_counter() - t) * 1e3)
x = (4 * torch.randn(B, C_, H, W)).cuda()
for _ in range(3):
    bs = coder.encode(x); y = coder.decode(bs)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10): bs = coder.encode(x)
torch.cuda.synchronize(); te = (time.perf_counter() - t) / 10
t = time.perf_counter()
for _ in range(10): y = coder.decode(bs)
torch.cuda.synchronize(); td = (time.perf_counter() - t) / 10
print(f"lanes {lanes}: encode {te*1e3:.3f} ms, decode {td*1e3:.3f} ms, {len(bs)} bytes, {B*C_*H*W} symbols, max err {float((y - x).abs().max()):.3f}")
