"""Two launches of each standalone upfirdn2d configuration of bench.py's `ops_hbm` leg (for an ncu capture)."""
import sys
import torch
sys.path.insert(0, ".")
import hairfastgan_b200.op as OP
dev = "cuda"
kern = torch.tensor([1., 3., 3., 1.], device=dev)
k2 = kern[None, :] * kern[:, None]
k2 = k2 / k2.sum()
xu = torch.randn(1, 256, 1025, 1025, device=dev)
xd = torch.randn(1, 192, 1024, 1024, device=dev)
xs = torch.randn(48, 3, 512, 512, device=dev)
for _ in range(2):
    OP.upfirdn2d(xu, k2 * 4, pad=(1, 1))
    OP.upfirdn2d(xd, k2, down=2, pad=(1, 1))
    OP.upfirdn2d(xs, k2 * 4, up=2, pad=(2, 1))
torch.cuda.synchronize()
