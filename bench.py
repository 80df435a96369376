#!/usr/bin/env python
"""bench.py -- hair-swap hot-path throughput on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--triples T]

A *step* = the generator hot path of T HairFast.swap() triples: the eight Generator.forward calls one
triple makes (full x3 [FSE recon], 3->3 x3, 0->3 x3, full x1, 0->3 x2, full x1,
4->8 x1, 5->8 x1 = 1057.6 GFLOP per triple at 1024^2), with the T triples' calls batched together
(independent triples, BASELINE config 5).  Encoders and the out-of-scope nets (BiSeNet/SEAN/CLIP) are not
part of the step.  `value` times the step with everything resident in HBM; `e2e` times the same step
through the public Generator.forward API with HOST (pinned) latents / layer_in features copied in and the
T final images copied out inside the timed region.

One JSON line on stdout (rank 0).  Multi-GPU: one process per GPU under torchrun, weights broadcast from
rank 0 over NCCL at init, no collective in the step; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_FULL, GFLOP_0_3, GFLOP_3_3, GFLOP_4_8, GFLOP_5_8 = 148.52, 8.007, 6.043, 140.51, 116.34
GFLOP_PER_TRIPLE = 3 * GFLOP_FULL + 3 * GFLOP_3_3 + 3 * GFLOP_0_3 + GFLOP_FULL + 2 * GFLOP_0_3 + GFLOP_FULL \
    + GFLOP_4_8 + GFLOP_5_8                               # = 1057.6
ROOFLINE_US_PER_IMG = 142.5                               # SURVEY Appendix A, sum of per-layer maxima


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        smax = int(float(self.rows[0][1])) if self.rows else None
        busy = [v for v in sm if smax and v > 0.5 * smax] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def census(T: int):
    """(start_layer, end_layer, batch, layer_in resolution) of the 8 generator calls of T batched triples."""
    return [(0, 8, 3 * T, None), (3, 3, 3 * T, 16), (0, 3, 3 * T, None), (0, 8, T, None), (0, 3, 2 * T, None),
            (0, 8, T, None), (4, 8, T, 32), (5, 8, T, 64)]


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from hairfastgan_b200 import _lib
    import hairfastgan_b200.model as M
    torch.set_grad_enabled(False)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = _lib.lib()                                  # raises if the CUDA library is missing: no fallback
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T = args.triples

    # ---- synthetic generator: seeded random weights of the reference architecture (no checkpoints exist)
    torch.manual_seed(0)
    gen = M.Generator(1024, 512, 8).to(dev).eval()
    for name, prm in gen.named_parameters():
        if name.endswith("noise.weight") or name.endswith("activate.bias") or name == "to_rgb1.bias" \
                or (name.startswith("to_rgbs.") and name.endswith(".bias") and name.count(".") == 2):
            prm.data.normal_(0, 0.1)
    if world > 1:                                     # weights replicated: one NCCL broadcast at init
        for t in list(gen.parameters()) + list(gen.buffers()):
            dist.broadcast(t.data, src=0)
    calls = census(T)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    host_lat = [torch.randn(b, 18, 512, generator=g).pin_memory() for (_, _, b, _) in calls]
    host_lin = [None if r is None else torch.randn(b, 512, r, r, generator=g).pin_memory() for (_, _, b, r) in calls]
    dev_lat = [t.to(dev) for t in host_lat]
    dev_lin = [None if t is None else t.to(dev) for t in host_lin]
    host_out = torch.empty(T, 3, 1024, 1024).pin_memory()
    launches = [0]

    def step(e2e: bool):
        final = None
        for i, (s, e, b, r) in enumerate(calls):
            lat = host_lat[i].to(dev, non_blocking=True) if e2e else dev_lat[i]
            lin = None if r is None else (host_lin[i].to(dev, non_blocking=True) if e2e else dev_lin[i])
            out = gen([lat], input_is_latent=True, start_layer=s, end_layer=e, layer_in=lin)   # random noise, as swap()
            launches[0] += lib.hf_last_launch_count()
            if i == len(calls) - 1:
                final = out[0]
        if e2e:
            host_out.copy_(final, non_blocking=True)
        return final

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def timed(e2e: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(e2e)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches[0] = 0
        total_ms = 0.0
        wall0 = time.perf_counter()
        for _ in range(steps):
            flush.fill_(1)                                             # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(e2e)
            e1.record()
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - wall0
        t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall, launches[0]

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, wall, n_launch = timed(False, args.steps, args.warmup)
    ms_e2e, _, _ = timed(True, args.steps, max(3, args.warmup // 2))
    clocks = sampler.stop() if rank == 0 else None

    # ---- configs[1]: full 1024^2 generator forward, B=4, and the dominant kernel alone (configs[0] shape x4)
    extra = {}
    if rank == 0 and world == 1:
        lat4 = torch.randn(4, 18, 512, device=dev)
        for _ in range(3):
            gen([lat4], input_is_latent=True)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(5):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gen([lat4], input_is_latent=True); e1.record(); e1.synchronize()
            tot += e0.elapsed_time(e1)
        us_img = tot / 5 / 4 * 1e3
        extra["generator_b4"] = {"img_per_s": round(1e6 / us_img, 1), "us_per_img": round(us_img, 1),
                                 "frac_of_roofline_142.5us": round(ROOFLINE_US_PER_IMG / us_img, 4),
                                 "tflops_algorithmic": round(GFLOP_FULL / us_img * 1e3, 1)}
        extra["roofline"] = time_dominant_kernel(gen, dev)
    if world > 1:
        dist.destroy_process_group()
    return ms, ms_e2e, n_launch, clocks, extra, host_lat, host_lin


def time_dominant_kernel(gen, dev):
    """conv_igemm_kernel on the 512->512 3x3 @64^2 layer (BASELINE configs[0] shape), B=4: the layer class that
    carries most of the tensor-core time.  Algorithmic FLOPs = 2*512*512*9*64^2 per sample."""
    import ctypes as C
    import torch
    from hairfastgan_b200 import _lib
    import hairfastgan_b200.model as M
    lib = _lib.lib()
    conv = gen.convs[7].conv                       # convs.7 = 512->512 @64^2
    B = 4
    desc, blob = conv._packed.get(conv, M.default_dtype())
    x = torch.randn(B, 512, 64, 64, device=dev); st = torch.randn(B, 512, device=dev)
    y = torch.empty(B, 512, 64, 64, device=dev)
    ws = torch.empty(lib.hf_conv_workspace_bytes(C.byref(desc), B, 64, 64), dtype=torch.uint8, device=dev)
    io = _lib.hf_conv_io()
    io.batch, io.height, io.width = B, 64, 64
    io.x, io.style, io.style_dim, io.style_stride = x.data_ptr(), st.data_ptr(), 512, 512
    mw, mb = conv.modulation.weight.data, conv.modulation.bias.data
    io.mod_weight, io.mod_bias, io.demodulate = mw.data_ptr(), mb.data_ptr(), 1
    io.y, io.workspace = y.data_ptr(), ws.data_ptr()
    ms = C.c_float(0)
    _lib.check(lib.hf_conv_time_kernel(C.byref(desc), blob.data_ptr(), C.byref(io), 20, C.byref(ms),
                                       torch.cuda.current_stream().cuda_stream), "hf_conv_time_kernel")
    flops = 2.0 * 512 * 512 * 9 * 64 * 64 * B
    achieved = flops / (ms.value * 1e-3) / 1e12
    pk = peaks()
    return {"kernel": "conv_halo_kernel<64,bf16> 512->512 3x3 @64^2 B=4 (+fp32 NCHW store)", "bound": "tensor",
            "achieved": round(achieved, 1), "peak": pk["tf_burst"], "unit": "TFLOP/s",
            "frac": round(achieved / pk["tf_burst"], 4), "peak_source": pk["src"] + " burst (kernel timed alone)",
            "launch_ms": round(ms.value, 4), "traffic": None}


def cpu_oracle_sample(threads=None):
    """The reference algorithm on host cores (oracle port; the reference itself cannot travel to the GPU box):
    one full 1024^2 generator forward, B=1 (148.52 of the 1057.6 GFLOP of a triple), scaled to triples/s."""
    import torch
    from oracle import stylegan2_oracle as O
    torch.set_grad_enabled(False)
    if threads:
        torch.set_num_threads(threads)
    p = O.synth_generator_params(size=1024, seed=0)
    lat = torch.randn(1, 18, 512, generator=torch.Generator().manual_seed(0))
    noise = O.synth_noise(1024, batch=1, seed=1)
    t0 = time.perf_counter()
    O.generator_ref(p, lat, noise)
    dt = time.perf_counter() - t0
    return dt, (GFLOP_FULL / GFLOP_PER_TRIPLE) / dt, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--triples", type=int, default=4, help="independent triples batched per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = ("generator hot path of HairFast.swap(): the 8 Generator.forward calls per triple "
                "(SURVEY App. B, 1057.6 GFLOP/triple, 1024^2, randomize_noise=True), synthetic weights; "
                "encoders and out-of-scope nets excluded")

    if args.impl == "reference":
        if rank != 0:
            return
        # the reference's CPU implementation of the path (oracle port, all host threads), bounded sample
        times = []
        for i in range(args.warmup + args.steps):
            dt, tps, thr = cpu_oracle_sample()
            if i >= args.warmup:
                times.append(dt)
        dt = sum(times) / len(times)
        val = (GFLOP_FULL / GFLOP_PER_TRIPLE) / dt
        sample = "one full 1024^2 generator forward B=1 = 148.52 of 1057.6 GFLOP per triple, scaled"
        print(json.dumps({
            "impl": "reference", "metric": "hair_swap_triples_per_sec", "value": val, "unit": "triples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": val, "unit": "triples/s", "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    ms, ms_e2e, n_launch, clocks, extra, host_lat, host_lin = run_ours(args, rank, world, local_rank)
    if rank != 0:
        return
    T = args.triples
    step_ms = ms / args.steps
    value = world * T / (step_ms * 1e-3)
    e2e_value = world * T / (ms_e2e / args.steps * 1e-3)
    h2d = sum(t.numel() * 4 for t in host_lat) + sum(t.numel() * 4 for t in host_lin if t is not None)
    out = {
        "metric": "hair_swap_triples_per_sec", "value": round(value, 3), "unit": "triples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_ms, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": os.environ.get("HAIRFAST_DTYPE", "bf16") + " operands, f32 accumulate", "data": "synthetic",
        "config": {"workload": workload, "triples_per_step_per_gpu": T, "size": 1024,
                   "parallelism": f"dp{world} (independent triples per rank, no step collective)",
                   "l2": "256 MiB flush write between timed steps; per-step CUDA events summed"},
        "tflops_algorithmic": round(GFLOP_PER_TRIPLE * value / 1e3, 1),
        "e2e": {"value": round(e2e_value, 3), "unit": "triples/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": T * 3 * 1024 * 1024 * 4},
        "gpu_launches": n_launch, "clocks": clocks,
    }
    out.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        dt, tps, thr = cpu_oracle_sample()
        out["cpu_baseline"] = {"value": round(tps, 5), "unit": "triples/s", "cores": thr, "kind": "port",
                               "sample": f"oracle generator_ref, one full 1024^2 forward B=1 ({dt:.1f} s), "
                                         "scaled by 148.52/1057.6 GFLOP"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
