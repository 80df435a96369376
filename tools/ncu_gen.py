"""One full 1024^2 generator forward (B = argv[1], default 4) between cudaProfilerStart/Stop, after two warm
forwards.  For
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_ \
        -o gpurun_out/gen_layers python tools/ncu_gen.py 4
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import hairfastgan_b200.model as M

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gen = M.Generator(1024, 512, 8).to(dev).eval()
lat = torch.randn(B, 18, 512, device=dev)
for _ in range(2):
    gen([lat], input_is_latent=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
gen([lat], input_is_latent=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
