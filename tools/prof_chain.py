"""Per-launch timing of one full 1024^2 generator forward (HF_GEN_PROFILE=1 makes the C library bracket
every conv / rgb_combine launch with CUDA events and print them).  Usage: python tools/prof_chain.py [B]"""
import os
import sys

os.environ.setdefault("HF_GEN_PROFILE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

torch.set_grad_enabled(False)
import hairfastgan_b200.model as M  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
gen = M.Generator(1024, 512, 8).cuda().eval()
lat = torch.randn(B, 18, 512, device="cuda")
for i in range(3):
    print(f"--- forward {i} (B={B}, HF_CONV_V1={os.environ.get('HF_CONV_V1', '0')})", file=sys.stderr)
    gen([lat], input_is_latent=True)
torch.cuda.synchronize()
os.environ["HF_GEN_PROFILE"] = "0"
