"""CUDA-event time of the two fused stem kernels at the bench's batch sizes (and their HBM rate)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hairfastgan_b200.nn16 as N
torch.set_grad_enabled(False)
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def avg(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot = 0
    for _ in range(n):
        flush.fill_(1); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n
bn = torch.nn.BatchNorm2d(64).eval().to(dev)
for B in (48, 96):
    x = torch.rand(B, 3, 256, 256, device=dev)
    s3 = N.PackedStem3x3(torch.randn(64, 3, 3, 3, device=dev), bn, torch.rand(64, device=dev))
    aff = (torch.rand(64, device=dev), torch.rand(64, device=dev))
    ms = avg(lambda: s3(x, y16b_affine=aff))
    byt = x.numel() * 4 + 2 * B * 256 * 256 * 64 * 2
    print(f"stem3x3 B={B} 256^2: {ms*1e3:.0f} us, {byt/ms/1e6:.0f} GB/s")
for B, R in ((48, 512), (16, 1024)):
    x = torch.rand(B, 3, R, R, device=dev)
    s7 = N.PackedStem7x7(torch.randn(64, 3, 7, 7, device=dev), bn)
    ms = avg(lambda: s7(x))
    byt = x.numel() * 4 + B * (R // 2) ** 2 * 64 * 2
    print(f"stem7x7 B={B} {R}^2: {ms*1e3:.0f} us, {byt/ms/1e6:.0f} GB/s")
