"""Where does an encoder's output start to depend on the batch size?  Runs the e4e trunk on the same two images inside
batches of 2 and 32 and prints, per block, max |diff| / rms of the raw block output (16-bit tensors)."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import hairfastgan_b200.encoders as E
from hairfastgan_b200 import nn16
from oracle import encoders_oracle as EO

torch.set_grad_enabled(False)
x = torch.rand(32, 3, 256, 256, generator=torch.Generator().manual_seed(70)).cuda() * 2 - 1
net = E.Encoder4Editing(50, "ir_se", types.SimpleNamespace(stylegan_size=1024)).eval()
net.load_state_dict(EO.synth_params_like(net, seed=11), strict=True)
net = net.cuda()
pk = net._pack()
blocks = pk["blocks"]


def trunk(xb):
    outs = []
    x16 = nn16.to_nhwc16(xb, c_pad=32)
    raw, bn, _ = pk["stem"](x16, shift=pk["stem_shift"], act=1, slope=pk["stem_slope"], y16b_affine=blocks[0].pre)
    outs.append(raw)
    for i, blk in enumerate(blocks):
        nxt = blocks[i + 1].pre if i + 1 < len(blocks) else None
        raw, bn = blk(raw, bn, nxt, want_raw=True)
        outs.append(raw)
    return outs


a = trunk(x[:2])
b = trunk(x)
for i, (u, v) in enumerate(zip(a, b)):
    u, v = u.float(), v[:2].float()
    d = float((u - v).abs().max()) / float(u.pow(2).mean().sqrt())
    n = int((u != v).sum())
    print(f"{'stem' if i == 0 else 'block %2d' % (i - 1)}  shape {tuple(u.shape)}  max|diff|/rms {d:.3e}  differing elements {n}")
w2, w32 = net(x[:2]), net(x)[:2]
print("w: max|diff|/rms", float((w2 - w32).abs().max()) / float(w2.pow(2).mean().sqrt()))
