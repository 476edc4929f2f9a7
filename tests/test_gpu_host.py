"""GPU parity of the host-side mirror (Extractor, tools.*, Evaluator, MultiLoss, Model, the tiled
pipeline) against the CPU oracle's *_port functions -- the reference's own call sequences.
These tests read like PyLC's call sites: same object chains, same argument meaning."""
import numpy as np
import pytest
import torch

import pylc_oracle as orc

pytestmark = pytest.mark.gpu

T = 512


@pytest.fixture(scope="module")
def ops():
    from pylc_b200 import ops as _ops
    _ops._lib.load()
    return _ops


def _params(**kw):
    from pylc_b200.config import Parameters
    p = Parameters()
    p.update(kw)
    return p


# ---------------------------------------------------------------------------------------------
# Extractor: extract -> profile  (preprocess.py:37-38)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("ch,schema", [(1, "a"), (3, "b")])
def test_extractor_extract_profile(ops, palettes, ch, schema):
    from pylc_b200.config import Parameters
    from pylc_b200.utils.extract import Extractor
    pal = palettes[schema]
    meta = Parameters({"schema": "./schemas/schema_%s.json" % schema})
    meta.update({"ch": ch})
    imgs = [orc.synth_image(i, 1300 + 40 * i, 1100, ch) for i in range(3)]
    masks = [orc.synth_mask(i, 1300 + 40 * i, 1100, pal, skew=(i == 1)) for i in range(3)]
    ex = Extractor(meta)
    ex.verbose = False
    ex.load_arrays(imgs, masks).extract().profile()
    ref_i = np.concatenate([orc.split_tiles(im, T, T) for im in imgs])
    ref_m = np.concatenate([orc.class_encode_port(orc.split_tiles(m, T, T), pal) for m in masks])
    got_i, got_m = ex.host()
    assert np.array_equal(got_i, ref_i)                      # bit-exact tiles
    assert np.array_equal(got_m, ref_m)                      # bit-exact encoded masks
    want = orc.profile_port(ref_i, ref_m, len(pal), T)
    m = ex.get_meta()
    assert m.n_samples == len(ref_i) == 12
    assert np.array_equal(np.array(m.px_dist), want["px_dist"])          # bit-exact histograms
    assert np.array_equal(np.array(m.dset_px_dist), want["dset_px_dist"])
    assert m.dset_px_count == want["dset_px_count"]
    np.testing.assert_allclose(m.px_mean, want["px_mean"], rtol=1e-5)
    np.testing.assert_allclose(m.px_std, want["px_std"], rtol=1e-5)
    np.testing.assert_allclose(m.probs, want["probs"], rtol=0, atol=0)
    np.testing.assert_allclose(m.weights, want["weights"], rtol=1e-15)
    assert m.m2 == pytest.approx(want["m2"], rel=1e-15) and m.jsd == pytest.approx(want["jsd"], rel=1e-15)
    # get_profile over the dataset object (the `profile --db` route) gives the same metadata
    from pylc_b200.utils.profile import get_profile
    m2 = get_profile(ex.get_data())
    assert np.array_equal(np.array(m2.px_dist), want["px_dist"])
    np.testing.assert_allclose(m2.px_std, want["px_std"], rtol=1e-5)


@pytest.mark.parametrize("stack_max", [32, 2])
def test_extractor_stacks_equal_sized_files(ops, palettes, stack_max):
    """Runs of equally sized files go through the *_stack entry points (one launch per kernel and stack);
    file order, tiles, masks and the profile are what the per-file loop of extract.py:132-222 gives."""
    from pylc_b200.config import Parameters
    from pylc_b200.utils.extract import Extractor
    pal = palettes["b"]
    meta = Parameters({"schema": "./schemas/schema_b.json"})
    meta.update({"ch": 1})
    sizes = [(1300, 1100)] * 3 + [(1100, 1300)] + [(1300, 1100)] * 2       # two runs around an odd one out
    imgs = [orc.synth_image(50 + i, w, h, 1) for i, (w, h) in enumerate(sizes)]
    masks = [orc.synth_mask(50 + i, w, h, pal, skew=bool(i & 1)) for i, (w, h) in enumerate(sizes)]
    ex = Extractor(meta)
    ex.verbose = False
    ex.STACK_MAX = stack_max
    n0 = ops._lib.launch_count()
    ex.load_arrays(imgs, masks).extract().profile()
    launches = ops._lib.launch_count() - n0
    assert launches == (6 if stack_max == 32 else 8)         # 2 kernels x stacks (3 | 1 | 2  or  2, 1 | 1 | 2)
    ref_i = np.concatenate([orc.split_tiles(im, T, T) for im in imgs])
    ref_m = np.concatenate([orc.class_encode_port(orc.split_tiles(m, T, T), pal) for m in masks])
    got_i, got_m = ex.host()
    assert np.array_equal(got_i, ref_i) and np.array_equal(got_m, ref_m)
    want = orc.profile_port(ref_i, ref_m, len(pal), T)
    m = ex.get_meta()
    assert np.array_equal(np.array(m.px_dist), want["px_dist"])
    assert np.array_equal(np.array(m.dset_px_dist), want["dset_px_dist"])
    np.testing.assert_allclose(m.px_mean, want["px_mean"], rtol=1e-5)
    np.testing.assert_allclose(m.px_std, want["px_std"], rtol=1e-5)
    np.testing.assert_allclose(m.weights, want["weights"], rtol=1e-15)


def test_extractor_coshuffle_is_one_permutation(ops, palettes):
    from pylc_b200.utils.extract import Extractor
    meta = _params(ch=1)
    img = orc.synth_image(7, 1600, 1100, 1)
    mask = orc.synth_mask(7, 1600, 1100, palettes["a"])
    ex = Extractor(meta)
    ex.verbose = False
    ex.load_arrays([img], [mask]).extract()
    i0, m0 = ex.host()
    ex.coshuffle()
    i1, m1 = ex.host()
    perm = ex._perm
    assert sorted(perm.tolist()) == list(range(len(i0)))
    assert np.array_equal(i1, i0[perm]) and np.array_equal(m1, m0[perm])
    ex.profile()                                             # per-tile rows follow the permutation
    assert np.array_equal(np.array(ex.get_meta().px_dist), orc.tile_histograms(m1, 9))


def test_extractor_fit_matches_reference_geometry(ops):
    from pylc_b200.utils.extract import Extractor
    img = orc.synth_image(2, 2000, 1500, 1)
    ex = Extractor(_params(ch=1))
    ex.verbose = False
    ex.load_arrays([img]).extract(fit=True, stride=256)
    e = ex.get_meta().extract
    assert (e["w_fitted"], e["h_fitted"], e["offset"], e["n"]) == (1536, 1024, 0, 15)
    import cv2
    fitted = cv2.resize(img, (1536, 1024), interpolation=cv2.INTER_AREA)
    assert np.array_equal(ex.host()[0], orc.split_tiles(fitted, T, 256))


# ---------------------------------------------------------------------------------------------
# tools.class_encode / colourize / reconstruct
# ---------------------------------------------------------------------------------------------

def test_tools_class_encode_and_colourize(ops, palettes):
    from pylc_b200.utils import tools
    pal = palettes["a"]
    mask = orc.synth_mask(3, 640, 512, pal, off_palette=0.01)
    nchw = torch.from_numpy(np.moveaxis(mask, 2, 0)[None].copy())
    enc = tools.class_encode(nchw, pal)
    assert enc.dtype == torch.uint8 and not enc.is_cuda
    assert np.array_equal(enc.numpy(), orc.class_encode_port(nchw.numpy(), pal))
    rgb = tools.colourize(enc.numpy().astype(np.int64), 9, palette=pal)
    assert rgb.dtype == np.int64
    assert np.array_equal(rgb, orc.colourize_port(enc.numpy().astype(np.int64), 9, pal))
    with pytest.raises(AssertionError):
        tools.class_encode(torch.zeros((1, 4, 8, 8), dtype=torch.uint8), pal)


@pytest.mark.parametrize("w_full,h_full", [(2000, 1500), (1700, 1300)])
def test_tools_reconstruct_matches_port(ops, palettes, w_full, h_full):
    """tools.reconstruct(list of batches of 8, meta) == the reference's loops (tools.py:209-319)."""
    from pylc_b200.utils import tools
    meta = _params(ch=1, stride=256)
    w_fit, h_fit = orc.fit_dims(w_full, h_full, T)
    nr, nc = h_fit // 256 - 1, w_fit // 256 - 1
    meta.extract = {"fid": "x", "n": nr * nc, "w_full": w_full, "h_full": h_full, "w_scaled": w_full,
                    "h_scaled": h_full, "w_fitted": w_fit, "h_fitted": h_fit, "offset": 0}
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(nr * nc, 9, T, T, generator=g) * 3
    batches = [logits[i:i + 8].cuda() for i in range(0, nr * nc, 8)]
    got = tools.reconstruct(batches, meta)
    want = orc.reconstruct_port([b.cpu().numpy() for b in batches], h_fit, w_fit, w_full, h_full, T, 256,
                                palettes["a"], 9)
    assert got.dtype == np.float32 and got.shape == want.shape == (h_full, w_full, 3)
    bad = np.any(got != want, axis=2)
    if bad.any():   # only where the reference's own top-1 is a near tie
        margin = orc.resample_labels(orc.top2_margin(orc.stitch_map(logits.numpy(), nr, nc, T, 256)), w_full, h_full)
        assert not (bad & (margin > 1e-6)).any()
    assert bad.mean() < 1e-5


# ---------------------------------------------------------------------------------------------
# Evaluator
# ---------------------------------------------------------------------------------------------

def _eval_case(palettes, seed, w_full=1000, h_full=760):
    pal = palettes["a"]
    gt = orc.synth_mask(seed, w_full, h_full, pal, skew=True)
    pred_lab = orc.synth_labels(seed + 50, w_full, h_full, 9, skew=False, block=37)
    pred_rgb = np.asarray(pal, dtype=np.uint8)[pred_lab].astype(np.float32)
    meta = _params(ch=1)
    meta.extract = {"fid": "img_%d" % seed, "n": 0, "w_full": w_full, "h_full": h_full, "w_scaled": w_full,
                    "h_scaled": h_full, "w_fitted": 0, "h_fitted": 0, "offset": 0}
    return gt, pred_lab, pred_rgb, meta


def test_evaluator_matches_sklearn_port(ops, palettes, tmp_path, monkeypatch):
    from pylc_b200.utils.evaluate import Evaluator
    monkeypatch.chdir(tmp_path)
    gt, pred_lab, pred_rgb, meta = _eval_case(palettes, 1)
    ev = Evaluator(meta)
    ev.load(pred_rgb, meta, mask_true=gt).evaluate()
    yt, yp = orc.inject_coverage(orc.class_encode_hwc(gt, palettes["a"]), pred_lab, 9)
    want = orc.metrics_port(yt, yp, meta.class_codes)
    assert np.array_equal(ev.conf.cpu().numpy(), orc.confusion_counts(yt, yp, 9))    # bit-exact
    r = ev.metrics.results
    assert r["f1"] == want["f1"] and r["iou"] == want["iou"] and r["mcc"] == want["mcc"]
    assert np.array_equal(ev.metrics.cmatrix, want["cmatrix"])
    for k, v in want["report"].items():
        assert r["report"][k] == pytest.approx(v, rel=0, abs=0) if not isinstance(v, dict) else r["report"][k] == v
    files = ev.save_metrics()
    assert all(f.endswith(s) for f, s in zip(files, ("_eval.json", "_cmap.pdf", "_cmap.npy")))
    assert np.array_equal(np.load(files[2]), want["cmatrix"])
    assert ev.save_image().endswith("img_1.png")


def test_evaluator_aggregate_equals_concatenation(ops, palettes, tmp_path, monkeypatch):
    from pylc_b200.utils.evaluate import Evaluator
    monkeypatch.chdir(tmp_path)
    ev = None
    yts, yps = [], []
    for seed in (1, 2, 3):
        gt, pred_lab, pred_rgb, meta = _eval_case(palettes, seed, 900 + 16 * seed, 700)
        ev = ev or Evaluator(meta)
        ev.load(pred_rgb, meta, mask_true=gt)
        ev.reset()
        yts.append(orc.class_encode_hwc(gt, palettes["a"]).ravel())
        yps.append(pred_lab.ravel())
    ev.evaluate(aggregate=True)
    yt, yp = orc.inject_coverage(np.concatenate(yts), np.concatenate(yps), 9)       # evaluate.py:158-174
    assert np.array_equal(ev.metrics.counts, orc.confusion_counts(yt, yp, 9))
    want = orc.metrics_port(yt, yp, ev.labels)
    assert ev.metrics.results["iou"] == want["iou"] and ev.fid == "aggregate_metrics"


# ---------------------------------------------------------------------------------------------
# MultiLoss (autograd Function) and Model
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("weighted", [False, True])
def test_multiloss_module_forward_backward(ops, weighted):
    from pylc_b200.models.modules.loss import MultiLoss
    from pylc_b200.config import defaults
    C = 9
    g = torch.Generator().manual_seed(11)
    z = (torch.randn(3, C, 96, 80, generator=g) * 2)
    t = torch.randint(0, C, (3, 96, 80), generator=g)
    w = (torch.rand(C, generator=g) + 0.2).tolist()
    crit = MultiLoss({"weighted": weighted, "weights": w, "ce": 0.5, "dice": 0.5, "focal": 0.5},
                     {"n_classes": C, "class_codes": defaults.class_codes, "class_labels": defaults.class_labels})
    zc = z.cuda().requires_grad_(True)
    loss = crit.forward(zc, t.cuda())
    (loss * 2.0).backward()
    ref = orc.multiloss_port(z.numpy(), t.numpy(), C, weights=w, weighted=weighted)
    assert loss.item() == pytest.approx(ref[0], rel=1e-4)
    assert crit.ce.item() == pytest.approx(ref[1], rel=1e-4)
    assert crit.dsc.item() == pytest.approx(ref[2], rel=1e-4)
    assert crit.fl.item() == pytest.approx(ref[3], rel=1e-4)
    np.testing.assert_allclose(zc.grad.cpu().numpy(), 2.0 * ref[4], rtol=2e-3, atol=1e-9)
    # component calls used by Model.eval (model.py:360-362)
    assert crit.ce_loss(zc.detach(), t.cuda()).item() == pytest.approx(ref[1], rel=1e-4)
    assert crit.dice_loss(zc.detach(), t.cuda()).item() == pytest.approx(ref[2], rel=1e-4)
    assert crit.focal_loss(zc.detach(), t.cuda()).item() == pytest.approx(ref[3], rel=1e-4)
    with pytest.raises(AssertionError):
        crit.forward(zc, t[:, :50].cuda())


def _tiny_model(ch):
    from pylc_b200.config import defaults
    from pylc_b200.models.model import Model
    torch.manual_seed(0)
    model = Model()
    model.track = False
    stats = ([120.0], [40.0]) if ch == 1 else ([130.0, 140.0, 150.0], [25.0, 22.0, 19.0])
    model.update_meta({"ch": ch, "arch": "deeplab", "backbone": "resnet", "pretrained": False, "px_mean": stats[0],
                       "px_std": stats[1], "weights": [1.0] * 9, "normalize_default": False, "weighted": False,
                       "schema": defaults.schema})
    model.build()
    model.net.eval()
    return model


@pytest.mark.parametrize("ch", [1, 3])
def test_model_test_equals_fused_tile_path(ops, ch):
    """Model.test(u8 tiles) (reference route: float -> normalize -> cat x3 -> net) and
    Model.test_tiles(pylc_tile_gather_norm_f32 output) see bit-identical network inputs."""
    model = _tiny_model(ch)
    img = orc.synth_image(1, 1024, 1024, ch)
    d, pitch = ops.upload_image(img)
    mean, std, post, out_ch = model.norm_params()
    fused = ops.tile_gather_norm_f32(d, 1024, 1024, ch, pitch, T, 256, mean, std, post, out_ch)
    tiles = torch.from_numpy(orc.split_tiles(img, T, 256)).float()
    x_ref = model._prepare(tiles)
    assert torch.equal(fused, x_ref)
    assert torch.equal(model.test_tiles(fused[:2])[0], model.test(tiles[:2])[0])


def test_model_train_step_runs_and_reports(ops):
    model = _tiny_model(3)
    model.net.train()
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (2, 3, 128, 128), generator=g).float()
    y = torch.randint(0, 9, (2, 128, 128), generator=g)
    before = [p.detach().clone() for p in list(model.net.parameters())[:3]]
    loss = model.train(x, y)
    assert torch.isfinite(loss) and model.iter == 1
    assert any(not torch.equal(a, b) for a, b in zip(before, list(model.net.parameters())[:3]))
    model.eval(x, y)


# ---------------------------------------------------------------------------------------------
# Tiled pipeline (test.py's loop, GPU-resident)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("fuse", [False, True])
def test_pipeline_matches_reference_sequence(ops, palettes, fuse):
    """TiledSegmenter on a 1600x1200 colour image == reference sequence evaluated by the oracle on
    the very logits the network produced (network is stock torch, so it is not under test)."""
    import cv2
    from pylc_b200.pipeline import TiledSegmenter
    model = _tiny_model(3)
    pal = palettes["a"]
    W, H = 1600, 1200
    img = orc.synth_image(9, W, H, 3)
    gt = orc.synth_mask(9, W, H, pal, skew=True)
    seg = TiledSegmenter(model, batch_tiles=8, channels_last=fuse, keep_masks=True, fuse_network=fuse)
    conf, results = seg.run_host([img], [gt])
    res = results[0]
    # oracle on the same network outputs
    w_fit, h_fit = orc.fit_dims(W, H, T)
    fitted = cv2.resize(img, (w_fit, h_fit), interpolation=cv2.INTER_AREA)
    tiles = torch.from_numpy(orc.split_tiles(fitted, T, 256))
    if fuse:    # BN-folded cuDNN-fused plan: same batches through the same plan (space-to-depth stem layout)
        assert seg.s2d
        d_fit, fp = ops.upload_image(fitted)
        xs = ops.tile_gather_norm_s2d(d_fit, h_fit, w_fit, 3, fp, T, 256, seg.mean, seg.std, seg.post_div)
        outs = seg.forward_tiles(xs, s2d=True)
        eager = torch.cat([model.test(tiles[i:i + 8])[0] for i in range(0, len(tiles), 8)])
        assert (torch.cat(outs) - eager).abs().max() <= 5e-3 * eager.abs().max()   # TF32 conv noise level
    else:
        outs = [model.test(tiles[i:i + 8])[0] for i in range(0, len(tiles), 8)]
    logits = torch.cat(outs).cpu().numpy()
    nr, nc = h_fit // 256 - 1, w_fit // 256 - 1
    ref_map = orc.stitch_map(logits, nr, nc, T, 256)
    ref_lab = orc.stitch_labels(ref_map)
    near_tie = orc.top2_margin(ref_map) <= 1e-6
    got_lab = res["labels"].cpu().numpy()
    assert not ((got_lab != ref_lab) & ~near_tie).any()
    pred_full = orc.resample_labels(got_lab, W, H)
    assert np.array_equal(res["pred_full"].cpu().numpy(), pred_full)
    assert np.array_equal(res["pred_rgb"].cpu().numpy(), np.asarray(pal, dtype=np.uint8)[pred_full])
    yt, yp = orc.inject_coverage(orc.class_encode_hwc(gt, pal), pred_full, 9)
    assert np.array_equal(conf, orc.confusion_counts(yt, yp, 9))
    # resident route gives the same matrix
    seg.reset()
    seg.run_resident([seg.stage(img, gt, index=0)])
    assert np.array_equal(seg.conf.cpu().numpy(), conf)
    s = seg.scores(conf)
    assert s["iou"] == orc.metrics_port(yt, yp, model.meta.class_codes)["iou"]


def test_pipeline_fused_upsample_flag_changes_nothing(ops, palettes):
    """TiledSegmenter(fuse_upsample=True) (the default with the inference plan: the stitch reads the decoder
    output) and fuse_upsample=False (up-sample kernel + stitch kernel) give the same label
    map, full-size prediction and confusion matrix, bit for bit."""
    from pylc_b200.pipeline import TiledSegmenter
    model = _tiny_model(3)
    pal = palettes["a"]
    W, H = 1600, 1200
    img = orc.synth_image(31, W, H, 3)
    gt = orc.synth_mask(31, W, H, pal, skew=True)
    seg_a = TiledSegmenter(model, batch_tiles=8, keep_masks=True, fuse_network=True, fuse_upsample=True)
    seg_b = TiledSegmenter(model, batch_tiles=8, keep_masks=True, fuse_network=True, fuse_upsample=False)
    assert seg_a.fuse_upsample and not seg_b.fuse_upsample
    seg_b.fused = seg_a.fused            # one plan (one set of per-layer route choices) for both
    conf_a, res_a = seg_a.run_host([img], [gt])
    conf_b, res_b = seg_b.run_host([img], [gt])
    assert torch.equal(res_a[0]["labels"], res_b[0]["labels"])
    assert torch.equal(res_a[0]["pred_full"], res_b[0]["pred_full"])
    assert np.array_equal(conf_a, conf_b)


@pytest.mark.parametrize("fuse", [False, True])
def test_pipeline_grayscale_image(ops, palettes, fuse):
    """ch = 1 (historic photographs, BASELINE configs[0] / configs[4]): the gather replicates the gray
    plane x3 and normalises with the single profiled mean / std (reference model.py:372-377,433-435);
    everything downstream equals the reference sequence evaluated by the oracle on the same logits."""
    import cv2
    from pylc_b200.pipeline import TiledSegmenter
    model = _tiny_model(1)
    pal = palettes["a"]
    W, H = 2000, 1500                                          # configs[0] geometry: fitted 1536 x 1024, 15 tiles
    img = orc.synth_image(21, W, H, 1)
    gt = orc.synth_mask(21, W, H, pal, skew=True)
    seg = TiledSegmenter(model, batch_tiles=8, channels_last=fuse, keep_masks=True, fuse_network=fuse)
    conf, results = seg.run_host([img], [gt])
    res = results[0]
    w_fit, h_fit = orc.fit_dims(W, H, T)
    assert (w_fit, h_fit) == (1536, 1024)
    fitted = cv2.resize(img, (w_fit, h_fit), interpolation=cv2.INTER_AREA)
    tiles = torch.from_numpy(orc.split_tiles(fitted, T, 256))
    assert tiles.shape == (15, 1, T, T)
    if fuse:
        d_fit, fp = ops.upload_image(fitted)
        if seg.s2d:
            xs = ops.tile_gather_norm_s2d(d_fit, h_fit, w_fit, 1, fp, T, 256, seg.mean, seg.std, seg.post_div)
        else:
            xs = ops.tile_gather_norm_f32(d_fit, h_fit, w_fit, 1, fp, T, 256, seg.mean, seg.std, seg.post_div, seg.out_ch)
        outs = seg.forward_tiles(xs, s2d=seg.s2d)
        eager = torch.cat([model.test(tiles[i:i + 8])[0] for i in range(0, len(tiles), 8)])
        assert (torch.cat(outs) - eager).abs().max() <= 5e-3 * eager.abs().max()   # TF32 conv noise level
    else:
        outs = [model.test(tiles[i:i + 8])[0] for i in range(0, len(tiles), 8)]
    logits = torch.cat(outs).cpu().numpy()
    nr, nc = h_fit // 256 - 1, w_fit // 256 - 1
    ref_map = orc.stitch_map(logits, nr, nc, T, 256)
    near_tie = orc.top2_margin(ref_map) <= 1e-6
    got_lab = res["labels"].cpu().numpy()
    assert not ((got_lab != orc.stitch_labels(ref_map)) & ~near_tie).any()
    pred_full = orc.resample_labels(got_lab, W, H)
    assert np.array_equal(res["pred_full"].cpu().numpy(), pred_full)
    yt, yp = orc.inject_coverage(orc.class_encode_hwc(gt, pal), pred_full, 9)
    assert np.array_equal(conf, orc.confusion_counts(yt, yp, 9))
    assert int(conf.sum()) == W * H


# ---------------------------------------------------------------------------------------------
# Augmentor.optimize (sample-rate grid search on the device)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", ["a", "b"])
def test_augmentor_optimize_golden(ops, golden, name):
    """Augmentor.optimize: every grid point and the winner equal the reference's own run and the oracle,
    to the last bit (int64 histograms on the device, float64 tail on the host)."""
    import types
    from pylc_b200.utils.augment import Augmentor
    g = golden("augment")
    N, C, px_count = (int(v) for v in g["aug_%s_meta" % name])
    meta = types.SimpleNamespace(px_dist=g["aug_%s_px_dist" % name], tile_px_count=px_count,
                                 probs=g["aug_%s_probs" % name], n_classes=C)
    aug = Augmentor().load_profile(meta, n_tiles=N).optimize()
    assert np.array_equal(aug.rates, g["aug_%s_rates" % name])
    om, ref = aug.optim_meta, g["aug_%s_optim" % name]
    assert [om["threshold"], om["rate_coef"], om["jsd"], om["m2"], om["n_samples"], om["aug_n_samples"]] == list(ref)
    assert np.array_equal(om["probs"], g["aug_%s_optim_probs" % name])
    best, data = orc.augment_optimize_port(meta.px_dist, px_count, meta.probs, C, N)
    assert len(data) == len(aug.profile_data)
    for a, b in zip(aug.profile_data, data):
        assert a["jsd"] == b["jsd"] and a["m2"] == b["m2"] and a["aug_n_samples"] == b["aug_n_samples"]
        assert np.array_equal(a["probs"], b["probs"])


def test_augmentor_oversample(ops, golden, palettes):
    """Augmentor.oversample (reference augment.py:184-239): every input tile once, plus rates[i] warped copies
    whose bytes are the reference's (golden warp vectors, RandomState(j)); the augmented set is profiled on the
    device and its histogram is the bincount of the output masks."""
    from pylc_b200.config import Parameters
    from pylc_b200.db.dataset import MLPDataset
    from pylc_b200.utils.augment import Augmentor
    g = golden("warp")
    img, mask = orc.augment_fixture_tile(3)
    rng = np.random.default_rng(2)
    other_img = rng.integers(0, 256, size=img.shape, dtype=np.uint8)
    other_mask = orc.synth_labels(6, 512, 512, 9)[None]
    meta = Parameters()
    meta.update({"ch": 3})
    dset = MLPDataset(input_data={"img": np.concatenate([img, other_img]), "mask": np.concatenate([mask, other_mask]),
                                  "meta": meta})
    aug = Augmentor().load(dset)
    aug.rates = np.array([2, 0])                      # two warped copies of tile 0, none of tile 1
    aug.oversample(shuffle=False)
    assert aug.output_imgs.shape == (4, 3, 512, 512) and aug.output_masks.shape == (4, 512, 512)
    assert np.array_equal(aug.output_imgs[0], img[0]) and np.array_equal(aug.output_masks[0], mask[0])
    for j in (0, 1):
        assert np.array_equal(aug.output_imgs[1 + j], g["warp_ch3_s%d_img" % j])
        assert np.array_equal(aug.output_masks[1 + j], g["warp_ch3_s%d_mask" % j])
    assert np.array_equal(aug.output_imgs[3], other_img[0]) and np.array_equal(aug.output_masks[3], other_mask[0])
    m = aug.output_meta
    assert m.n_samples == 4 and m.id.startswith("_aug")
    assert m.dset_px_dist == np.bincount(aug.output_masks.ravel(), minlength=9).tolist()
    shuffled = Augmentor().load(dset)
    shuffled.rates = np.array([1, 1])
    shuffled.oversample()
    assert shuffled.output_imgs.shape[0] == 4


def test_augmentor_oversample_device_equals_host_path(ops):
    """Augmentor.oversample(device=True) -- copies made by pylc_augment_tiles_u8 -- gives the arrays of the default
    path through the reference's OpenCV calls (reference utils/augment.py:184-239), and the same profile."""
    from pylc_b200.config import Parameters
    from pylc_b200.db.dataset import MLPDataset
    from pylc_b200.utils.augment import Augmentor
    imgs = np.stack([orc.synth_image(80 + i, 512, 512, 3).transpose(2, 0, 1) for i in range(3)])
    masks = np.stack([orc.synth_labels(80 + i, 512, 512, 9, skew=True) for i in range(3)]).astype(np.uint8)
    meta = Parameters()
    meta.update({"ch": 3})
    got = []
    for device in (False, True):
        aug = Augmentor().load(MLPDataset(input_data={"img": imgs, "mask": masks, "meta": meta}))
        aug.rates = np.array([1, 0, 2])
        aug.oversample(shuffle=False, device=device)
        got.append(aug)
    assert np.array_equal(got[0].output_imgs, got[1].output_imgs) and np.array_equal(got[0].output_masks, got[1].output_masks)
    assert got[0].output_meta.dset_px_dist == got[1].output_meta.dset_px_dist and got[1].output_meta.n_samples == 6


def test_sample_rate_grid_wide_classes(ops):
    rng = np.random.default_rng(3)
    N, C = 1234, 20
    px = rng.integers(0, 5000, size=(N, C)).astype(np.int64)
    scores = rng.random(N) * 4
    coefs, thr = np.arange(1, 6, 1.), np.arange(0, 2, 0.25)
    sr, full = ops.sample_rate_grid(torch.from_numpy(scores).cuda(), torch.from_numpy(px).cuda(),
                                    torch.from_numpy(coefs).cuda(), torch.from_numpy(thr).cuda(), 0, 4)
    for i, rc in enumerate(coefs):
        for j, t in enumerate(thr):
            rates = np.clip(np.multiply(scores > t, rc * scores).astype(int), 0, 4)
            gi = i * len(thr) + j
            assert int(sr[gi]) == rates.sum()
            assert np.array_equal(full[gi].cpu().numpy(), (px + rates[:, None] * px).sum(axis=0))
