"""N-GPU rows for the BASELINE.json configurations that are not the headline metric (bench.py runs configs[1]
and configs[4]): run under torchrun, one process per GPU, NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/nbench.py --config {2,3} [--steps K] [--out gpurun_out/nbench.jsonl]

  --config 2   training-step multi-loss (weighted CE + Dice + Focal), 512x512 tiles, batch 64 PER GPU, data
               parallel: pylc_multiloss_reduce -> all-reduce of the 2C+3 f64 partials -> pylc_multiloss_grad
               (reference models/modules/loss.py:23-215 + autograd; models/model.py:317-325).  Reported per
               step as the max over ranks; the all-reduce's cost is the difference to the same two launches
               without it.
  --config 3   extraction / profile sweep: 10 000 synthetic 2000x1500 grayscale image + RGB mask pairs,
               schema_b (11 classes), dealt round-robin over the ranks: image tile gather + moments, mask
               gather + encode + per-tile histograms, ONE all-reduce of the dataset histogram and moment sums
               at the end (reference utils/extract.py:106-231, utils/profile.py:21-150).  The pairs cycle
               through a pool of distinct device-resident images larger than L2.
One JSON line per row (rank 0); timing = CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pylc_b200 import dist as pdist   # noqa: E402
from pylc_b200 import ops, synth     # noqa: E402


def peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    pdist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    pdist.barrier()
    return pdist.max_over_ranks(e0.elapsed_time(e1)) / steps


def config2(args, rank, world, rows):
    T, B = 512, 64
    peak, kind = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for C in (9, 11):
        g = torch.Generator(device="cuda").manual_seed(100 + rank)
        z = torch.randn((B, C, T, T), generator=g, device="cuda") * 3      # 604 MB (C=9) / 738 MB (C=11): > L2
        t = torch.randint(0, C, (B, T, T), generator=g, device="cuda")
        w = torch.linspace(0.3, 1.0, C, device="cuda")
        cfg = ops.loss_cfg()
        grad = torch.empty_like(z)
        npx = B * T * T

        def step(reduce_across):
            part = torch.zeros((2 * C + 3,), dtype=torch.float64, device="cuda")
            ops.multiloss_reduce(z, t, cfg, class_w=w, partials=part)
            n = npx
            if reduce_across and world > 1:
                pdist.all_reduce_(part)
                n = npx * world
            ops.multiloss_finalize(part, C, n, cfg)
            ops.multiloss_grad(z, t, cfg, part, n, class_w=w, out=grad)
        ms_dp = timed(lambda: step(True), args.steps)
        ms_local = timed(lambda: step(False), args.steps)
        # the same step as ONE cooperative launch with the exchange inside the kernel (peer memory), and the
        # single-GPU single launch it is to be compared with
        ex = pdist.loss_exchange() if world > 1 else None
        ms_fused_dp = None
        if ex is not None:
            def fused_dp():
                part = torch.zeros((2 * C + 3,), dtype=torch.float64, device="cuda")
                ops.multiloss_fwd_bwd_dp(z, t, cfg, ex["ptrs_dev"], ex["rank"], ex["world"], ex["next_epoch"](), w, out=grad, partials=part)
            ms_fused_dp = timed(fused_dp, args.steps)

        def fused_local():
            part = torch.zeros((2 * C + 3,), dtype=torch.float64, device="cuda")
            ops.multiloss_fwd_bwd(z, t, cfg, w, out=grad, partials=part)
        ms_fused_local = timed(fused_local, args.steps)
        alg = npx * (8 * C + 8)
        if rank == 0:
            rows.append({"config": "configs[2]: multi-loss fwd+bwd, 512x512 tiles, batch 64/GPU, C=%d, weighted CE" % C,
                         "n_gpus": world, "ms_per_step": round(ms_dp, 4), "ms_per_step_without_allreduce": round(ms_local, 4),
                         "allreduce_share": round(max(0.0, ms_dp - ms_local) / ms_dp, 4),
                         "ms_per_step_fused_exchange_in_kernel": None if ms_fused_dp is None else round(ms_fused_dp, 4),
                         "ms_per_step_fused_single_gpu_launch": round(ms_fused_local, 4),
                         "in_kernel_exchange_cost_ms": None if ms_fused_dp is None else round(ms_fused_dp - ms_fused_local, 4),
                         "tiles_per_s_all_gpus": round(B * world / (ms_dp * 1e-3), 1),
                         "alg_bytes_per_gpu": alg, "achieved_gbs_per_gpu": round(alg / (ms_dp * 1e-3) / 1e9, 1),
                         "frac_of_peak_per_gpu": round(alg / (ms_dp * 1e-3) / 1e9 / peak, 4), "peak_gbs": peak, "peak_kind": kind,
                         "exchange": "ncclAllReduce(sum) of %d f64 partials between the two passes" % (2 * C + 3)})
        del z, t, grad
    del flush


def config3(args, rank, world, rows):
    from pylc_b200.config import Parameters
    T, S, W, H, n_pairs, pool_n = 512, 512, 2000, 1500, args.pairs, 24
    meta = Parameters({"schema": "./schemas/schema_b.json"})
    pal, C = meta.palette_rgb, meta.n_classes
    assert C == 11
    # 24 x (3 MB + 9 MB) = 288 MB of distinct inputs (> L2), resident as two stacks [24, H, pitch]
    d_imgs, ip, _ = ops.upload_stack([synth.image(rank * pool_n + i, W, H, 1) for i in range(pool_n)])
    d_masks, mp, _ = ops.upload_stack([synth.mask(rank * pool_n + i, W, H, pal, skew=True) for i in range(pool_n)])
    mine = pdist.shard_indices(n_pairs, rank, world)
    nH, nW = ops.tile_grid(H, W, T, S)
    per = nH * nW
    # every tile of the rank's share is kept, as the extractor does (utils/extract.py:182,214)
    tiles = torch.empty((len(mine) * per, 1, T, T), dtype=torch.uint8, device="cuda")
    masks = torch.empty((len(mine) * per, T, T), dtype=torch.uint8, device="cuda")
    px_dist = torch.zeros((len(mine) * per, C), dtype=torch.int64, device="cuda")
    stat = torch.zeros((len(mine) * per, 1, 2), dtype=torch.int64, device="cuda")

    def finish():
        hist = px_dist.sum(0)
        mom = stat.sum((0, 1)).view(-1)
        if world > 1:
            pdist.all_reduce_(hist)
            pdist.all_reduce_(mom)
        return hist, mom

    def sweep_pairs():           # the reference's file loop: two launches per pair
        px_dist.zero_()
        for k in range(len(mine)):
            j = k % pool_n                  # the rank's k-th pair: the same pool image the stacked sweep reads
            ops.tile_gather_u8_stack(d_imgs[j:j + 1], H, W, 1, ip, T, S, stats=True, out=tiles[k * per:(k + 1) * per],
                                     stat_out=stat[k * per:(k + 1) * per])
            ops.mask_gather_encode_hist(d_masks[j], H, W, mp, T, S, pal, out=masks[k * per:(k + 1) * per],
                                        px_dist=px_dist[k * per:(k + 1) * per])
        return finish()

    def sweep_stacks():          # equally sized pairs as stacks of 24: two launches per 24 pairs
        px_dist.zero_()
        stat.zero_()
        lib, pal_c = ops._lib.load(), ops._lib.palette_array(pal)[0]
        st = ops._stream()
        for k0 in range(0, len(mine), pool_n):
            n = min(pool_n, len(mine) - k0)
            lo, hi = k0 * per, (k0 + n) * per
            ops.check(lib.pylc_tile_gather_u8_stack(ops._p(d_imgs), n, H * ip, H, W, 1, ip, T, S, ops._p(tiles[lo:hi]),
                                                    ops._p(stat[lo:hi]), st), "pylc_tile_gather_u8_stack")
            ops.check(lib.pylc_mask_gather_encode_hist_stack(ops._p(d_masks), n, H * mp, H, W, mp, T, S, pal_c, C,
                                                             ops._p(masks[lo:hi]), ops._p(px_dist[lo:hi]), st),
                      "pylc_mask_gather_encode_hist_stack")
        return finish()

    total_px = n_pairs * per * T * T
    alg = n_pairs * per * T * T * (2 + 4)          # gray gather 1+1 B, mask gather 3+1 B per tile pixel
    peak, kind = peak_gbs()
    results = {}
    for name, fn, launches in (("stacks", sweep_stacks, 2.0 / pool_n), ("pairs", sweep_pairs, 2.0)):
        ms = timed(fn, max(1, args.steps // 4), warmup=1)
        hist, mom = fn()
        results[name] = (ms, hist.clone(), mom.clone(), tiles[:per * pool_n].clone(), masks[:per * pool_n].clone())
        if rank == 0:
            rows.append({"config": "configs[3]: extraction/profile sweep, %d synthetic 2000x1500 gray image + mask pairs, schema_b (11 classes)" % n_pairs,
                         "form": {"stacks": "stacks of %d equally sized pairs per call (pylc_*_stack)" % pool_n,
                                  "pairs": "one call per pair (the reference's file loop)"}[name],
                         "n_gpus": world, "ms_per_sweep": round(ms, 2), "pairs_per_s_all_gpus": round(n_pairs / (ms * 1e-3), 1),
                         "mpx_per_s_all_gpus": round(n_pairs * W * H / 1e6 / (ms * 1e-3), 1),
                         "launches_per_pair": round(launches, 4), "us_per_pair_per_gpu": round(ms * 1e3 / len(mine), 2),
                         "alg_bytes_all_gpus": alg, "achieved_gbs_all_gpus": round(alg / (ms * 1e-3) / 1e9, 1),
                         "frac_of_peak_per_gpu": round(alg / world / (ms * 1e-3) / 1e9 / peak, 4), "peak_gbs": peak, "peak_kind": kind,
                         "histogram_total_equals_pixel_count": int(hist.sum()) == total_px,
                         "exchange": "one ncclAllReduce(sum) of the [11] i64 dataset histogram and one of the moment sums per sweep"})
    a, b = results["stacks"], results["pairs"]
    same = all(torch.equal(x, y) for x, y in zip(a[1:], b[1:]))
    if rank == 0:
        rows[-2]["identical_to_per_pair_calls"] = same
        rows[-2]["speedup_over_per_pair_calls"] = round(b[0] / a[0], 2)
    if not same:
        raise SystemExit("configs[3]: stacked and per-pair sweeps disagree")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[2, 3])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--pairs", type=int, default=10000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, world, local = pdist.init_from_env()
    torch.cuda.set_device(local)
    ops._lib.load()
    rows = []
    (config2 if args.config == 2 else config3)(args, rank, world, rows)
    if rank == 0:
        for r in rows:
            print(json.dumps(r), flush=True)
        if args.out:
            os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
            with open(args.out, "a") as f:
                for r in rows:
                    f.write(json.dumps(r) + "\n")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
