"""GPU-idle fraction and kernel shares of one timed bench step, from a CUPTI kernel trace.

nsys is not in this image; torch.profiler records the same CUPTI activity records (start, duration,
stream of every kernel and memcpy, whoever launched it -- the ctypes launches of libpylc_b200.so
included).  The script runs bench.py's resident step (configs[1] images already in HBM) and its
host-buffer step (TiledSegmenter.run_host) on `--images` images, and reports for each:

    span_ms      first kernel start -> last kernel end
    busy_ms      union of all kernel / memcpy intervals (any stream)
    idle_frac    1 - busy / span: time in which NO engine of the GPU ran anything
    ours_frac    share of the summed kernel time spent in pylc:: kernels
    top          the ten kernels with the largest total time

    python tools/gpu_idle.py [--images 8] [--out profiles/gpu_idle_r2.json]

A number printed under the profiler is not a bench value; only the fractions are meant to be read.
"""
import argparse
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def trace(fn):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
    iv, per = [], defaultdict(float)
    for ev in prof.events():
        if ev.device_type != torch.autograd.DeviceType.CUDA:
            continue
        t0 = ev.time_range.start
        t1 = ev.time_range.end
        if t1 <= t0:
            continue
        iv.append((t0, t1))
        per[ev.name] += t1 - t0
    iv.sort()
    busy, cur0, cur1 = 0.0, None, None
    for t0, t1 in iv:
        if cur1 is None or t0 > cur1:
            if cur1 is not None:
                busy += cur1 - cur0
            cur0, cur1 = t0, t1
        else:
            cur1 = max(cur1, t1)
    if cur1 is not None:
        busy += cur1 - cur0
    span = iv[-1][1] - iv[0][0] if iv else 0.0
    span = max(t1 for _, t1 in iv) - iv[0][0] if iv else 0.0
    total = sum(per.values())
    ours = sum(v for k, v in per.items() if "pylc::" in k)
    top = sorted(per.items(), key=lambda kv: -kv[1])[:10]
    return {"records": len(iv), "span_ms": span / 1e3, "busy_ms": busy / 1e3, "idle_frac": 1.0 - busy / span if span else None,
            "kernel_time_sum_ms": total / 1e3, "ours_frac": ours / total if total else None,
            "top": [{"kernel": k[:100], "ms": v / 1e3, "share": v / total} for k, v in top]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from pylc_b200 import _lib
    from pylc_b200.pipeline import TiledSegmenter
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    _lib.load()
    torch.backends.cudnn.benchmark = True
    model = bench.build_model(device)
    seg = TiledSegmenter(model, batch_tiles=45, channels_last=True, host_workers=6, fuse_network=True, device_fit=True, fuse_upsample=True)
    gidx = list(range(args.images))
    imgs, masks = bench.make_inputs(gidx, model.meta.palette_rgb)
    resident = [seg.stage_device(imgs[i], masks[i], index=gidx[i]) for i in range(len(gidx))]
    for _ in range(3):
        seg.reset()
        seg.run_resident(resident)
        seg.reset()
        seg.run_host(imgs, masks, global_indices=gidx)
    torch.cuda.synchronize()

    def resident_step():
        seg.reset()
        seg.run_resident(resident)

    def host_step():
        seg.reset()
        seg.run_host(imgs, masks, global_indices=gidx)

    out = {"images": args.images, "workload": "bench.py configs[1] images (3000x2000 colour + RGB mask), DeepLabv3+/ResNet-101 inference plan",
           "method": "torch.profiler CUPTI kernel+memcpy records; idle = 1 - union(busy intervals) / span",
           "resident_step": trace(resident_step), "host_step": trace(host_step)}
    line = json.dumps(out)
    print(line)
    if args.out:
        with open(args.out, "w") as f:
            f.write(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
