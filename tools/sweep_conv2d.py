"""Time the encoder-style convolution (hf_conv2d_forward: folded BN + PReLU epilogue, 16-bit NHWC in/out) over the
shapes of the e4e / FSE / PostProcess / BiSeNet trunks.  Usage: python tools/sweep_conv2d.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hairfastgan_b200 import nn16

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
dev = torch.device("cuda", 0)
SHAPES = [  # cin, cout, r_in, k, stride, groups
    (64, 64, 128, 3, 1, 1), (64, 64, 256, 3, 2, 1), (64, 128, 128, 3, 1, 1), (128, 128, 128, 3, 2, 1),
    (128, 128, 64, 3, 1, 1), (128, 256, 64, 3, 1, 1), (256, 256, 64, 3, 2, 1), (256, 256, 32, 3, 1, 1),
    (256, 512, 32, 3, 1, 1), (512, 512, 32, 3, 2, 1), (512, 512, 16, 3, 1, 1), (64, 128, 128, 1, 2, 1),
    (256, 512, 32, 1, 1, 1), (512 * 11, 512 * 11, 64, 3, 2, 11), (1024, 1024, 64, 3, 1, 1), (768, 768, 64, 3, 1, 1),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for cin, cout, r, k, s, g in SHAPES:
    b = B if cin < 1024 * 4 else max(1, B // 4)
    w = torch.randn(cout, cin // g, k, k, device=dev) / (k * (cin // g) ** 0.5)
    pc = nn16.PackedConv2d(w, torch.ones(cout, device=dev), stride=s, groups=g)
    x = torch.randn(b, r, r, cin, device=dev).to(nn16.torch_dtype())
    shift, slope = torch.zeros(cout, device=dev), torch.full((cout,), 0.25, device=dev)
    for _ in range(3):
        pc(x, shift=shift, act=1, slope=slope)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(5):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pc(x, shift=shift, act=1, slope=slope); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 5
    ro = r // s
    gf = 2.0 * (cin // g) * cout * k * k * ro * ro * b / 1e9
    print(f"{cin:5d}->{cout:5d} k{k} s{s} g{g:2d} r_in={r:3d} B={b:3d}  {ms * 1e3:8.1f} us  {gf / ms:7.1f} TFLOP/s", flush=True)
