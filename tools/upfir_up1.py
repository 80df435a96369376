"""Standalone rate of the blur (upfirdn2d up 1, 4x4, pad (1,1)) on the north_star shape [256,1025,1025] -> 1024^2."""
import sys
import torch
sys.path.insert(0, ".")
import hairfastgan_b200.op as OP
dev = "cuda"
kern = torch.tensor([1., 3., 3., 1.], device=dev)
k2 = kern[None, :] * kern[:, None]
k2 = k2 / k2.sum()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
xu = torch.randn(1, 256, 1025, 1025, device=dev)
for _ in range(3):
    OP.upfirdn2d(xu, k2 * 4, pad=(1, 1))
torch.cuda.synchronize()
ts = []
for _ in range(8):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    OP.upfirdn2d(xu, k2 * 4, pad=(1, 1))
    e1.record()
    e1.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print("up1 k4 [256,1025,1025]: median %.3f ms  %.0f GB/s  (min %.3f)" % (ms, (xu.numel() + 256 * 1024 * 1024) * 4 / ms / 1e6, min(ts)))
