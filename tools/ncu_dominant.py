"""Launch the dominant kernel (512->512 3x3 @64^2, bf16, batch = argv[1], default 4) a few times so that
    ncu --set full --clock-control none --import-source on -k regex:conv_halo --launch-skip 5 -c 1 \
        -o gpurun_out/dominant python tools/ncu_dominant.py 48
captures one warm launch.  Prints the event-timed figure of the same launches (not a bench value under ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import hairfastgan_b200.model as M

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
gen = M.Generator(1024, 512, 8).to(dev).eval()
print(bench.time_dominant_kernel(gen, dev, int(sys.argv[1]) if len(sys.argv) > 1 else 4))
