"""Network-only timing of DeepLabv3+/ResNet-101 inference on one batch of 512x512 tiles, for the
execution variants the pipeline can use.  Prints one JSON line per variant (ms per tile, TFLOP/s at
~180 GFLOP per tile, max |diff| against the eager NCHW fp32 network)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pylc_b200.models.deeplab import DeepLab  # noqa: E402
from pylc_b200.models.fused import FusedDeepLab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=45)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = DeepLab(9).cuda().eval()
    x = torch.randn(args.batch, 3, 512, 512, device="cuda")
    with torch.no_grad():
        ref = net(x)

    def timeit(name, fn):
        with torch.no_grad():
            for _ in range(2):
                y = fn()
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(args.iters):
                y = fn()
            e.record()
            torch.cuda.synchronize()
        ms = s.elapsed_time(e) / args.iters
        print(json.dumps({"variant": name, "batch": args.batch, "ms_per_batch": round(ms, 2),
                          "ms_per_tile": round(ms / args.batch, 3),
                          "tflops_at_180gf_per_tile": round(0.18 * args.batch / (ms * 1e-3), 1),
                          "max_abs_diff": float((y.float() - ref).abs().max()), "ref_abs_max": float(ref.abs().max())}),
              flush=True)

    timeit("eager nchw fp32(tf32 conv)", lambda: net(x))
    net_cl = DeepLab(9).cuda().eval()
    net_cl.load_state_dict(net.state_dict())
    net_cl = net_cl.to(memory_format=torch.channels_last)
    xcl = x.contiguous(memory_format=torch.channels_last)
    timeit("eager channels_last fp32(tf32 conv)", lambda: net_cl(xcl))
    fused = FusedDeepLab(net, channels_last=True)
    timeit("fused channels_last fp32(tf32 conv)", lambda: fused(x))
    fused_nchw = FusedDeepLab(net, channels_last=False)
    timeit("fused nchw fp32(tf32 conv)", lambda: fused_nchw(x))

    def autocast():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return net_cl(xcl)
    timeit("eager channels_last bf16 autocast", autocast)
    fused16 = FusedDeepLab(net, channels_last=True, dtype=torch.bfloat16)
    timeit("fused channels_last bf16", lambda: fused16(x))
    torch.backends.cudnn.allow_tf32 = False
    timeit("eager channels_last strict fp32 (no tf32)", lambda: net_cl(xcl))


if __name__ == "__main__":
    main()
