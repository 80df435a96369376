"""One forward of a drop-in network between cudaProfilerStart/Stop (after two warm forwards), for
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches_<net>.csv python tools/prof_net.py <net> [batch] [size]
net: bisenet | e4e | fse | pp_enc | pp_res"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

torch.set_grad_enabled(False)
net_name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
size = int(sys.argv[3]) if len(sys.argv) > 3 else (512 if net_name == "bisenet" else 256)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if net_name == "bisenet":
    import hairfastgan_b200.bisenet as M
    net, x = M.BiSeNet(19), torch.rand(B, 3, size, size, device=dev) * 2 - 1
elif net_name == "e4e":
    import hairfastgan_b200.encoders as M
    net = M.Encoder4Editing(50, "ir_se", types.SimpleNamespace(stylegan_size=1024))
    x = torch.rand(B, 3, size, size, device=dev) * 2 - 1
elif net_name == "fse":
    import hairfastgan_b200.encoders as M
    net, x = M.fs_encoder_v2(n_styles=18, opts=None, stride=(2, 2)), torch.rand(B, 3, size, size, device=dev) * 2 - 1
elif net_name == "pp_enc":
    import hairfastgan_b200.postprocess as M
    net, x = M.FeatureEncoderMult(fs_layers=[9], opts=None), torch.rand(B, 3, size, size, device=dev) * 2 - 1
else:
    import hairfastgan_b200.postprocess as M
    net, x = M.FeatureiResnet([[1024, 2], [768, 2], [512, 2]]), torch.randn(B, 1024, 64, 64, device=dev)
net = net.to(dev).eval()
for _ in range(2):
    net(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
net(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
