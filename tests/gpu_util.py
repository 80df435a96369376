"""Shared helpers for the -m gpu parity tests (oracle = checker only)."""
import json
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# Stated tolerances (DESIGN.md section "Parity"): max-abs error relative to the RMS of the reference
# output.  16-bit operands, fp32 accumulate: per-product relative error <= 2^-8 (bf16) / 2^-11 (fp16).
TOL_SINGLE = {"bf16": 2.5e-2, "fp16": 4e-3}     # one conv layer
TOL_RGB = {"bf16": 6e-2, "fp16": 1e-2}          # generator RGB after up to 17 chained convs
TOL_FP32 = 2e-5                                  # pure fp32 SIMT ops (upfirdn2d, bias_act, ToRGB)


def dtype_name():
    return "fp16" if os.environ.get("HAIRFAST_DTYPE", "bf16").lower() in ("fp16", "f16", "half") else "bf16"


def enc_dtype_name():
    """dtype of the encoder family (hairfastgan_b200/nn16.default_dtype): fp16 unless an env var says bf16."""
    v = os.environ.get("HAIRFAST_ENC_DTYPE") or os.environ.get("HAIRFAST_DTYPE") or "fp16"
    return "bf16" if v.lower() in ("bf16", "bfloat16") else "fp16"


def rel_err(y: torch.Tensor, ref: torch.Tensor):
    y = y.detach().float().cpu()
    ref = ref.detach().float().cpu()
    rms = float(ref.pow(2).mean().sqrt()) + 1e-12
    return float((y - ref).abs().max()) / rms, rms


def record(name: str, **kv):
    """Append a measured parity number to gpurun_out/parity.jsonl (travels back from the GPU box)."""
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity.jsonl"), "a") as f:
        f.write(json.dumps({"name": name, "dtype": dtype_name(), **kv}) + "\n")
