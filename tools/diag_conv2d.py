"""GPU bring-up of hf_conv2d_forward (plain conv path of the encoders) against torch CPU fp32 convs.
Test infrastructure.  Writes gpurun_out/diag_conv2d.txt."""
import os
import sys
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
torch.set_grad_enabled(False)
import hairfastgan_b200.nn16 as N  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
log = open(os.path.join(ROOT, "gpurun_out", "diag_conv2d.txt"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s); log.write(s + "\n"); log.flush()


def case(name, cin, cout, r, k, stride, groups=1, cin_pad=None, act=0, residual=False, dual=False, batch=2):
    try:
        g = torch.Generator().manual_seed(hash(name) % 1000)
        x = torch.randn(batch, cin, r, r, generator=g)
        w = torch.randn(cout, cin // groups, k, k, generator=g) / (k * (cin // groups) ** 0.5)
        osc = torch.rand(cout, generator=g) + 0.5
        shift = torch.randn(cout, generator=g) * 0.3
        slope = torch.rand(cout, generator=g) * 0.5
        ref = F.conv2d(x, w, None, stride, k // 2, 1, groups) * osc.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
        if act == 1:
            ref = torch.where(ref > 0, ref, ref * slope.view(1, -1, 1, 1))
        elif act == 2:
            ref = F.leaky_relu(ref, 0.01)
        elif act == 3:
            ref = F.relu(ref)
        res = torch.randn_like(ref) if residual else None
        if residual:
            ref = ref + res
        s2, b2 = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
        pc = N.PackedConv2d(w.cuda(), osc.cuda(), stride=stride, groups=groups, cin_pad=cin_pad)
        x16 = N.to_nhwc16(x.cuda(), c_pad=cin_pad)
        r16 = N.to_nhwc16(res.cuda()) if residual else None
        y16, y16b, y32 = pc(x16, shift=shift.cuda(), act=act, slope=slope.cuda() if act == 1 else None, slope0=0.01,
                            residual16=r16, want_y16=True, y16b_affine=(s2.cuda(), b2.cuda()) if dual else None,
                            want_y32=True)
        torch.cuda.synchronize()
        rms = float(ref.pow(2).mean().sqrt())
        e32 = float((y32.cpu() - ref).abs().max()) / rms
        e16 = float((N.to_nchw32(y16).cpu() - ref).abs().max()) / rms
        msg = f"[{name}] {cin}->{cout} r={r} k={k} s={stride} g={groups} act={act} res={int(residual)}: y32 err/rms={e32:.4g} y16 err/rms={e16:.4g}"
        if dual:
            refb = ref * s2.view(1, -1, 1, 1) + b2.view(1, -1, 1, 1)
            msg += f" y16b err/rms={float((N.to_nchw32(y16b).cpu() - refb).abs().max()) / float(refb.pow(2).mean().sqrt()):.4g}"
        P(msg)
        if e32 > 0.05:
            err = (y32.cpu() - ref).abs()
            P("   by out row:", [round(float(v), 2) for v in (err.amax(dim=(0, 1, 3)) / rms)[:16]])
            P("   by out col:", [round(float(v), 2) for v in (err.amax(dim=(0, 1, 2)) / rms)[:16]])
            P("   by 32-ch  :", [round(float(v), 2) for v in (err.amax(dim=(0, 2, 3)).view(-1, 32).amax(1) / rms)[:16]])
    except Exception:
        P(f"[{name}] EXCEPTION\n" + traceback.format_exc())


P("device:", torch.cuda.get_device_name(0))
case("s1_k3_halo", 64, 64, 32, 3, 1, act=1)
case("s1_k3_small", 64, 128, 8, 3, 1, act=2)
case("s1_k1", 128, 512, 32, 1, 1)
case("s2_k3", 64, 64, 32, 3, 2, act=0)
case("s2_k3_128", 128, 256, 64, 3, 2, act=1, dual=True)
case("s2_k1", 64, 128, 32, 1, 2)
case("s2_k3_tiny", 512, 512, 4, 3, 2, act=2)
case("s2_k3_2to1", 512, 512, 2, 3, 2, act=2)
case("stem_pad", 3, 64, 64, 3, 1, cin_pad=32, act=1)
case("residual", 64, 64, 32, 3, 1, residual=True, dual=True)
case("grouped_s2", 4 * 64, 4 * 64, 16, 3, 2, groups=4, act=2)
case("grouped_s1", 3 * 128, 3 * 64, 16, 3, 1, groups=3)
P("done")
