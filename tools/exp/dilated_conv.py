"""Experiment: library routes for the dilated 3x3 convolutions of the inference plan (layer4 d = 2/4/8 on 512 ch,
ASPP d = 6/12/18 on 2048 -> 256 ch, 32x32 maps, 45 tiles, channels-last f32/TF32).  Prints the time and the
kernels each route launches, so the plan can pick a route that stays on sm100 kernels."""
import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pylc_b200.models.fused import _Conv

torch.backends.cudnn.benchmark = True
B = 45
cl = torch.channels_last


def t(fn, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    e.synchronize()
    return s.elapsed_time(e) / n * 1e3


def kernels(fn):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    out = []
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            out.append("%s %.0fus" % (ev.name[:60], (ev.time_range.end - ev.time_range.start)))
    return out


def tap_split(x, w, b, d):
    """3x3 dilated conv (pad = d) as nine 1x1 convolutions on the windows that do not fall into the padding."""
    Bn, C, H, W = x.shape
    out = F.conv2d(x, w[:, :, 1:2, 1:2].contiguous(memory_format=cl), b)
    for ty in (-1, 0, 1):
        for tx in (-1, 0, 1):
            if ty == 0 and tx == 0:
                continue
            iy0, iy1 = max(0, d * ty), H + min(0, d * ty)
            ix0, ix1 = max(0, d * tx), W + min(0, d * tx)
            if iy1 <= iy0 or ix1 <= ix0:
                continue
            win = x[:, :, iy0:iy1, ix0:ix1].contiguous(memory_format=cl)
            y = F.conv2d(win, w[:, :, ty + 1:ty + 2, tx + 1:tx + 2].contiguous(memory_format=cl))
            out[:, :, iy0 - d * ty:iy1 - d * ty, ix0 - d * tx:ix1 - d * tx] += y
    return out.relu_()


for name, cin, cout, d in [("layer4 d8", 512, 512, 8), ("aspp d6", 2048, 256, 6), ("aspp d12", 2048, 256, 12), ("aspp d18", 2048, 256, 18)]:
    x = torch.randn(B, cin, 32, 32, device="cuda").contiguous(memory_format=cl)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") * 0.01).contiguous(memory_format=cl)
    b = torch.randn(cout, device="cuda")
    xn, wn = x.contiguous(), w.contiguous()
    conv_m = torch.nn.Conv2d(cin, cout, 3, padding=d, dilation=d, bias=True).cuda()
    with torch.no_grad():
        conv_m.weight.copy_(w)
        conv_m.bias.copy_(b)
    plan_conv = _Conv(conv_m, None, True, True)
    xp = F.pad(x, (d, d, d, d)).contiguous(memory_format=cl)
    routes = {
        "fused cudnn_convolution_relu (NHWC)": lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (d, d), (d, d), 1),
        "F.conv2d + relu_ (NHWC, benchmark)": lambda: F.conv2d(x, w, b, 1, d, d).relu_(),
        "F.conv2d + relu_ (NCHW, benchmark)": lambda: F.conv2d(xn, wn, b, 1, d, d).relu_(),
        "explicit zero pad + conv pad 0 (NHWC)": lambda: F.conv2d(F.pad(x, (d, d, d, d)), w, b, 1, 0, d).relu_(),
        "conv pad 0 on a pre-padded input": lambda: F.conv2d(xp, w, b, 1, 0, d).relu_(),
        "tap split (9 x 1x1 on valid windows)": lambda: tap_split(x, w, b, d),
        "plan route 2: GEMM tap split": lambda: plan_conv._tap_split(x),
    }
    ref = routes["F.conv2d + relu_ (NCHW, benchmark)"]()
    print("== %s  %d->%d" % (name, cin, cout))
    for rn, fn in routes.items():
        try:
            us = t(fn)
            err = float((fn() - ref).abs().max())
            ks = kernels(fn)
            print("  %-40s %7.0f us  maxdiff %.2e   %s" % (rn, us, err, " | ".join(ks[:6])))
        except Exception as ex:  # noqa: BLE001
            print("  %-40s failed: %s" % (rn, str(ex)[:80]))
