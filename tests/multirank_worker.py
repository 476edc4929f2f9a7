"""
Worker of tests/test_gpu_multirank.py: run under torch.distributed.run with one process per GPU (NCCL).

Every rank computes BOTH the sharded result (its shard + the all-reduce) and, independently, the
single-rank result over the whole set, and compares them on the spot; rank 0 prints one JSON line.
The loop being sharded is the reference's test.py:52-115 (images), utils/profile.py:98-111 (tiles) and
one training step's loss (models/model.py:317-325); integer results must be bit-identical for any GPU
count, the loss within 1e-4 relative (north_star).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pylc_b200 import dist as pdist     # noqa: E402
from pylc_b200 import ops, synth       # noqa: E402


def tiny_model(device):
    from pylc_b200.config import defaults
    from pylc_b200.models.model import Model
    torch.manual_seed(0)
    model = Model()
    model.track = False
    model.device = device
    model.update_meta({"ch": 3, "arch": "deeplab", "backbone": "resnet", "pretrained": False,
                       "px_mean": [130.0, 140.0, 150.0], "px_std": [25.0, 22.0, 19.0], "weights": [1.0] * 9,
                       "normalize_default": False, "weighted": False, "schema": defaults.schema})
    model.build()
    model.net.eval()
    return model


def main():
    rank, world, local = pdist.init_from_env()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    ops._lib.load()
    out = {"world": world}

    # ---- (1) tiled inference + confusion: 6 images dealt round-robin vs all 6 on one rank ----------
    from pylc_b200.pipeline import TiledSegmenter
    model = tiny_model(device)
    pal = model.meta.palette_rgb
    n_img, W, H = 6, 1100, 1060
    imgs = [synth.image(i, W, H, 3) for i in range(n_img)]
    gts = [synth.mask(i, W, H, pal) for i in range(n_img)]
    mine = pdist.shard_indices(n_img)
    seg = TiledSegmenter(model, batch_tiles=8, fuse_network=True)
    conf_sharded, _ = seg.run_host([imgs[i] for i in mine], [gts[i] for i in mine], distributed=True,
                                   global_indices=mine)
    seg.reset()
    conf_single, _ = seg.run_host(imgs, gts, distributed=False)
    out["conf_equal"] = bool(np.array_equal(conf_sharded, conf_single))
    out["conf_sum"] = int(conf_sharded.sum())
    out["conf_expected_sum"] = n_img * W * H
    # default partition of distributed=True: equal contiguous shards, offset rank * len(images)
    seg.reset()
    per = n_img // world
    lo = rank * per
    conf_contig, _ = seg.run_host(imgs[lo:lo + per], gts[lo:lo + per], distributed=True)
    seg.reset()
    conf_single_sub, _ = seg.run_host(imgs[:per * world], gts[:per * world], distributed=False)
    out["conf_contig_equal"] = bool(np.array_equal(conf_contig, conf_single_sub))

    # ---- (2) get_profile(distributed=True): tiles sharded vs all tiles on one rank -----------------
    from pylc_b200.config import Parameters
    from pylc_b200.db.dataset import MLPDataset
    from pylc_b200.utils.profile import get_profile
    T, n_tiles = 512, 10
    rng = np.random.default_rng(7)
    t_imgs = rng.integers(0, 256, (n_tiles, 3, T, T), dtype=np.uint8)
    t_masks = synth.labels(3, T, T * n_tiles, 9).reshape(n_tiles, T, T)

    def profile_of(idx, distributed):
        meta = Parameters()
        meta.update({"ch": 3})
        dset = MLPDataset(input_data={"img": t_imgs[idx], "mask": t_masks[idx], "meta": meta})
        return get_profile(dset, distributed=distributed)

    tmine = pdist.shard_indices(n_tiles)
    m_sh = profile_of(tmine, True)
    m_one = profile_of(list(range(n_tiles)), False)
    out["hist_equal"] = m_sh.dset_px_dist == m_one.dset_px_dist and m_sh.n_samples == m_one.n_samples
    out["probs_equal"] = m_sh.probs == m_one.probs and m_sh.weights == m_one.weights
    out["mean_std_close"] = bool(np.allclose(m_sh.px_mean, m_one.px_mean, rtol=1e-6) and
                                 np.allclose(m_sh.px_std, m_one.px_std, rtol=1e-6))

    # ---- (3) MultiLoss(distributed=True): batch shards vs the concatenated batch --------------------
    from pylc_b200.models.modules.loss import MultiLoss
    C, B, hw = 9, 4, 96
    g = torch.Generator().manual_seed(11)
    z_all = (torch.randn(B * world, C, hw, hw, generator=g) * 3).to(device)
    t_all = torch.randint(0, C, (B * world, hw, hw), generator=g).to(device)
    lw = {"weighted": True, "weights": list(np.linspace(0.3, 1.0, C)), "ce": 0.5, "dice": 0.5, "focal": 0.5}
    schema = {"n_classes": C, "class_codes": ["c%d" % i for i in range(C)], "class_labels": ["l%d" % i for i in range(C)]}
    z_loc = z_all[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    crit_d = MultiLoss(lw, schema, distributed=True)
    crit_d.weights = crit_d.weights.to(device)
    loss_d = crit_d(z_loc, t_all[rank * B:(rank + 1) * B])
    loss_d.backward()
    z_one = z_all.clone().requires_grad_(True)
    crit_1 = MultiLoss(lw, schema, distributed=False)
    crit_1.weights = crit_1.weights.to(device)
    loss_1 = crit_1(z_one, t_all)
    loss_1.backward()
    g_one = z_one.grad[rank * B:(rank + 1) * B]
    out["loss_rel"] = abs(float(loss_d) - float(loss_1)) / abs(float(loss_1))
    den = g_one.abs().max().item()
    out["grad_rel"] = (z_loc.grad - g_one).abs().max().item() / den
    # ddp_average: the logit gradient carries the world-size factor DDP's mean removes again
    z_loc2 = z_all[rank * B:(rank + 1) * B].clone().requires_grad_(True)
    crit_a = MultiLoss(lw, schema, distributed=True, ddp_average=True)
    crit_a.weights = crit_a.weights.to(device)
    crit_a(z_loc2, t_all[rank * B:(rank + 1) * B]).backward()
    out["grad_ddp_rel"] = (z_loc2.grad / world - g_one).abs().max().item() / den

    # ---- (3b) the two data-parallel routes agree: in-kernel exchange over peer memory vs NCCL between launches ----
    ex = pdist.loss_exchange()               # the MultiLoss calls above already created it (or found none)
    out["dp_fused_available"] = ex is not None
    out["dp_partials_rel"] = out["dp_grad_rel"] = out["dp_loss_rel"] = 0.0
    out["dp_partials_same_on_all_ranks"] = True
    if ex is not None:
        cfg = crit_d._cfg(0.5, 0.5, 0.5)
        zl, tl = z_all[rank * B:(rank + 1) * B].contiguous(), t_all[rank * B:(rank + 1) * B].contiguous()
        n_total = tl.numel() * world
        part = ops.multiloss_reduce(zl, tl, cfg, crit_d.weights)
        pdist.all_reduce_(part)
        vals_n = ops.multiloss_finalize(part, C, n_total, cfg)
        grad_n = ops.multiloss_grad(zl, tl, cfg, part, n_total, crit_d.weights)
        for _ in range(3):                   # consecutive calls: both slots of the workspace, epochs in step
            vals_f, grad_f, part_f = ops.multiloss_fwd_bwd_dp(zl, tl, cfg, ex["ptrs_dev"], ex["rank"], ex["world"], ex["next_epoch"](),
                                                              crit_d.weights)
        out["dp_partials_rel"] = float(((part_f - part).abs() / part.abs().clamp_min(1e-300)).max())
        out["dp_loss_rel"] = float((vals_f - vals_n).abs().max() / vals_n.abs().max())
        out["dp_grad_rel"] = float((grad_f - grad_n).abs().max() / grad_n.abs().max())
        gathered = [torch.empty_like(part_f) for _ in range(world)]
        torch.distributed.all_gather(gathered, part_f)
        out["dp_partials_same_on_all_ranks"] = all(torch.equal(gathered[0], q) for q in gathered)

    # ---- (4) one data-parallel training step: DDP + MultiLoss(distributed, ddp_average) ---------------
    # parameter gradients after DDP's averaging == gradients of the single large batch on one rank
    # (BatchNorm in eval mode so that batch statistics do not depend on the shard; fp32 convolutions)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from pylc_b200.models.deeplab import DeepLab
    torch.manual_seed(1)
    net = DeepLab(n_classes=C).to(device).eval()
    import copy
    net_ref = copy.deepcopy(net)            # the single-rank reference runs on an unwrapped copy
    gx = torch.Generator().manual_seed(21)
    x_all = torch.randn(2 * world, 3, 96, 96, generator=gx).to(device)
    y_all = torch.randint(0, C, (2 * world, 96, 96), generator=gx).to(device)
    ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local])
    crit_t = MultiLoss(lw, schema, distributed=True, ddp_average=True)
    crit_t.weights = crit_t.weights.to(device)
    ddp.zero_grad()
    crit_t(ddp(x_all[rank * 2:(rank + 1) * 2]), y_all[rank * 2:(rank + 1) * 2]).backward()
    g_ddp = [p.grad.detach().clone() for p in net.parameters() if p.grad is not None]
    crit_s = MultiLoss(lw, schema, distributed=False)
    crit_s.weights = crit_s.weights.to(device)
    crit_s(net_ref(x_all), y_all).backward()
    g_one = [p.grad.detach().clone() for p in net_ref.parameters() if p.grad is not None]
    num = torch.sqrt(sum(((a - b) ** 2).sum() for a, b in zip(g_ddp, g_one)))
    den2 = torch.sqrt(sum((b ** 2).sum() for b in g_one))
    out["param_grad_rel"] = float(num / den2)
    out["n_param_grads"] = len(g_one)

    # every rank must agree: reduce the boolean verdicts (min) and the errors (max)
    flags = torch.tensor([float(out[k]) for k in ("conf_equal", "conf_contig_equal", "hist_equal", "probs_equal",
                                                  "mean_std_close")], device=device)
    torch.distributed.all_reduce(flags, op=torch.distributed.ReduceOp.MIN)
    errs = torch.tensor([out["loss_rel"], out["grad_rel"], out["grad_ddp_rel"], out["param_grad_rel"], out["dp_partials_rel"],
                         out["dp_loss_rel"], out["dp_grad_rel"]], device=device, dtype=torch.float64)
    torch.distributed.all_reduce(errs, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        for k, v in zip(("conf_equal", "conf_contig_equal", "hist_equal", "probs_equal", "mean_std_close"), flags.tolist()):
            out[k] = bool(v)
        (out["loss_rel"], out["grad_rel"], out["grad_ddp_rel"], out["param_grad_rel"], out["dp_partials_rel"], out["dp_loss_rel"],
         out["dp_grad_rel"]) = errs.tolist()
        out["launches"] = int(ops._lib.launch_count())
        print(json.dumps(out), flush=True)
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
