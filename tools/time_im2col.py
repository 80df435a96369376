"""Event-timed im2col of the BiSeNet stem (B=16, 1024^2)."""
import sys
sys.path.insert(0, ".")
import torch
from hairfastgan_b200 import _lib, nn16
x = torch.rand(16, 3, 1024, 1024, device="cuda")
cols = torch.empty(16, 512, 512, 160, device="cuda", dtype=nn16.torch_dtype())
lib = _lib.lib()
def run():
    _lib.check(lib.hf_im2col7x7s2_nhwc16(x.data_ptr(), cols.data_ptr(), 16, 1024, 1024, nn16.default_dtype(), _lib.stream_ptr()), "im2col")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
print("im2col B=16 1024^2: %.3f ms, %.0f GB/s written" % (ms, cols.numel() * 2 / ms / 1e6))
