"""Prebuild the reference's two JIT CUDA extensions (models/stylegan2/op/fused_act.py:10-16, upfirdn2d.py:10-16) for
sm_100 into baseline/_ref/ext, from the staged, unmodified sources, with torch's own JIT builder.  The cache then
travels to the GPU box with the snapshot; `baseline/refenv.py` points TORCH_EXTENSIONS_DIR at it, so importing the
reference there is a ninja no-op (or, if the timestamps did not survive the copy, the same 2.5 min rebuild)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refenv  # noqa: E402

if __name__ == "__main__":
    refenv.activate(overlay=False, chdir=False)
    import models.stylegan2.op as op          # triggers both torch.utils.cpp_extension.load calls
    print("built:", sorted(os.listdir(os.environ["TORCH_EXTENSIONS_DIR"])), op.__file__)
