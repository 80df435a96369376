"""DEBUG: run the stock reference's HairFast.swap() on the CPU of the build container (no GPU here) to validate the
stand-in modules and the synthetic checkpoints before GPU minutes are spent.  Hard-coded 'cuda' spots are redirected
by baseline/refenv.cpu_dryrun_patches().  Not a parity or bench tool."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import refenv, synth_checkpoints  # noqa: E402

work = "/tmp/hf_work"
overlay = "--overlay" in sys.argv          # construction only: this package's modules have no CPU forward
refenv.activate(overlay=overlay, workdir=work)
import torch  # noqa: E402
refenv.cpu_dryrun_patches()
synth_checkpoints.write_all(work, 0)
from hair_swap import HairFast, get_parser  # noqa: E402

args = get_parser().parse_args(["--device", "cpu"])
t0 = time.time()
hf = HairFast(args)
print("init %.1fs" % (time.time() - t0), type(hf.net.generator).__module__, flush=True)
if overlay:
    sys.exit(0)
imgs = [torch.rand(3, 1024, 1024, generator=torch.Generator().manual_seed(s)) for s in range(3)]
t0 = time.time()
out = hf.swap(*imgs)
print("swap %.1fs" % (time.time() - t0), out.shape, out.dtype, float(out.min()), float(out.max()), float(out.mean()),
      bool(torch.isfinite(out).all()))
