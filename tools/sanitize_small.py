"""One small launch of every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck are 10-100x
slower than a plain run, so the inputs are small; results are still checked against the oracle).

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pylc_oracle as orc  # noqa: E402  (checker)
from pylc_b200 import ops  # noqa: E402
from pylc_b200.config import Parameters  # noqa: E402

pal = Parameters().palette_rgb
C, T = len(pal), 512
W, H = 1300, 1100                     # 2 x 2 tiles at S = 512, ragged right / bottom edges

mask = orc.synth_mask(1, W, H, pal, skew=True, off_palette=0.01)
img1 = orc.synth_image(1, W, H, 1)
img3 = orc.synth_image(2, W, H, 3)
d_mask, mp = ops.upload_image(mask)
d1, p1 = ops.upload_image(img1)
d3, p3 = ops.upload_image(img3)

# TMA mask gather + encode + histogram, per-thread twin
for no_tma in ("0", "1"):
    os.environ["PYLC_NO_TMA"] = no_tma
    tiles, hist = ops.mask_gather_encode_hist(d_mask, H, W, mp, T, 512, pal)
    ref = orc.class_encode(orc.split_tiles(mask, T, 512), pal)
    assert np.array_equal(tiles.cpu().numpy(), ref) and np.array_equal(hist.cpu().numpy(), orc.tile_histograms(ref, C))
    enc = ops.class_encode_hwc(d_mask, H, W, mp, pal, hist=True)[0]
    assert np.array_equal(enc.cpu().numpy().reshape(H, W), orc.class_encode_hwc(mask, pal))
os.environ["PYLC_NO_TMA"] = "0"

# image gathers
for d, p, img, ch in ((d1, p1, img1, 1), (d3, p3, img3, 3)):
    g = ops.tile_gather_u8(d, H, W, ch, p, T, 256)
    assert np.array_equal(g.cpu().numpy(), orc.split_tiles(img, T, 256))

# stacks of equally sized sources: one launch (3-D tensor map over the stack; grid.y = image)
masks = [mask, orc.synth_mask(2, W, H, pal, off_palette=0.01), orc.synth_mask(3, W, H, pal, skew=True)]
imgs = [img1, orc.synth_image(3, W, H, 1), orc.synth_image(4, W, H, 1)]
d_ms, mps, _ = ops.upload_stack(masks)
d_is, ips, _ = ops.upload_stack(imgs)
ref_m = np.concatenate([orc.class_encode(orc.split_tiles(m, T, 256), pal) for m in masks])
for no_tma in ("0", "1"):
    os.environ["PYLC_NO_TMA"] = no_tma
    tiles, hist = ops.mask_gather_encode_hist_stack(d_ms, H, W, mps, T, 256, pal)
    assert np.array_equal(tiles.cpu().numpy(), ref_m) and np.array_equal(hist.cpu().numpy(), orc.tile_histograms(ref_m, C))
os.environ["PYLC_NO_TMA"] = "0"
g, st = ops.tile_gather_u8_stack(d_is, H, W, 1, ips, T, 256, stats=True)
ref_i = np.concatenate([orc.split_tiles(im, T, 256) for im in imgs])
assert np.array_equal(g.cpu().numpy(), ref_i)
assert np.array_equal(st.cpu().numpy()[:, 0, 0], ref_i.astype(np.int64).reshape(len(ref_i), -1).sum(-1))

# fit resize (TMA + f32x2) against cv2 via the oracle restatement
import cv2  # noqa: E402
for d, p, img, ch in ((d1, p1, img1, 1), (d3, p3, img3, 3)):
    w, h = orc.fit_dims(W, H, T)
    out, po = ops.fit_resize_area(d, H, W, ch, p, h, w)
    want = cv2.resize(img, (w, h), interpolation=cv2.INTER_AREA)
    assert np.array_equal(out.cpu().numpy()[:, :w * ch].reshape(want.shape), want)

# stitch + argmax; resample + encode + confusion (TMA and per-thread)
w, h = orc.fit_dims(W, H, T)
nr, nc = h // 256 - 1, w // 256 - 1
gen = torch.Generator().manual_seed(0)
logits = torch.randn(nr * nc, C, T, T, generator=gen) * 3
labels, _, stitched = ops.stitch_argmax_colour(logits.cuda(), nr, nc, T, 256, want_stitched=True)
ref_map = orc.stitch_map(logits.numpy(), nr, nc, T, 256)
assert np.allclose(stitched.cpu().numpy(), ref_map, rtol=1e-5, atol=1e-7)
lab = torch.from_numpy(orc.synth_labels(3, w, h, C, block=37)).cuda()
yt, yp = orc.inject_coverage(orc.class_encode_hwc(mask, pal), orc.resample_labels(lab.cpu().numpy(), W, H), C)
want = orc.confusion_counts(yt, yp, C)
for no_tma in ("0", "1"):
    os.environ["PYLC_NO_TMA"] = no_tma
    res = ops.resample_encode_confusion(lab, W, H, gt_rgb=d_mask, gt_pitch=mp, palette=pal, n_inject=C)
    assert np.array_equal(res["conf"].cpu().numpy(), want)
os.environ["PYLC_NO_TMA"] = "0"

# multi-loss: two-pass and single cooperative launch
z = torch.randn(2, C, 64, 64, generator=gen) * 3
t = torch.randint(0, C, (2, 64, 64), generator=gen)
cfg = ops.loss_cfg()
part = ops.multiloss_reduce(z.cuda(), t.cuda(), cfg)
grad = ops.multiloss_grad(z.cuda(), t.cuda(), cfg, part, t.numel())
out4, grad2, _ = ops.multiloss_fwd_bwd(z.cuda(), t.cuda(), cfg)
ref = orc.multiloss(z.numpy(), t.numpy(), C)
assert np.allclose(grad.cpu().numpy(), ref[4], rtol=2e-3, atol=1e-9) and np.allclose(grad2.cpu().numpy(), ref[4], rtol=2e-3, atol=1e-9)
torch.cuda.synchronize()
print("sanitize_small: all kernel families ran and matched the oracle")
