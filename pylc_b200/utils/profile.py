"""
Dataset profiling -- host-side mirror of PyLC's utils/profile.py (reference profile.py:21-207).

get_profile(dset) keeps the reference's contract (fills meta.n_samples, px_mean, px_std, px_dist,
dset_px_dist, dset_px_count, probs, weights, m2, jsd, tile_px_count) but replaces the per-tile
Python loop (DataLoader batch 1 -> torch.mean/std -> one_hot -> np.sum; profile.py:98-111) by one
launch of pylc_profile_tiles per chunk of tiles: per-tile class histograms [n,C] i64 and exact
integer per-tile/channel sum(x), sum(x^2), from which the host derives the reference's
"mean of per-tile mean / unbiased std" in float64.

Data-parallel: every rank profiles its own tiles; get_profile(distributed=True) all-reduces the
dataset histogram and the moment sums so that all ranks derive identical metadata.
"""
import numpy as np
import torch

from .. import dist as pdist
from .. import ops
from .metrics import jsd, m2

_CHUNK = 4096  # tiles per launch (1 GiB of u8 tiles at ch=1)


def tile_moments_to_stats(stat, tile_px):
    """[n,ch,2] integer (sum x, sum x^2) -> per-tile mean [n,ch] and unbiased std [n,ch] (f64):
    torch.mean / torch.std of profile.py:101-106 (std over all pixels of the tile per channel)."""
    s1 = stat[..., 0].astype(np.float64)
    s2 = stat[..., 1].astype(np.float64)
    mean = s1 / tile_px
    var = (s2 - s1 * s1 / tile_px) / (tile_px - 1)
    return mean, np.sqrt(np.maximum(var, 0.0))


def profile_arrays(imgs, masks, n_classes):
    """Run the profiling kernels over tile arrays (host ndarray or CUDA tensor).
    Returns (px_dist [n,C] i64 ndarray, stat [n,ch,2] i64 ndarray)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    n = len(imgs)
    dists, stats = [], []
    for lo in range(0, n, _CHUNK):
        hi = min(lo + _CHUNK, n)
        im, mk = imgs[lo:hi], masks[lo:hi]
        if not torch.is_tensor(im):
            im = torch.from_numpy(np.ascontiguousarray(im)).pin_memory().to(dev, non_blocking=True)
        if not torch.is_tensor(mk):
            mk = torch.from_numpy(np.ascontiguousarray(mk)).pin_memory().to(dev, non_blocking=True)
        stat, px_dist = ops.profile_tiles(im, mk, n_classes)
        dists.append(px_dist)
        stats.append(stat)
    if not dists:
        ch = imgs.shape[1] if len(imgs.shape) > 1 else 1
        return np.zeros((0, n_classes), np.int64), np.zeros((0, ch, 2), np.int64)
    return torch.cat(dists).cpu().numpy(), torch.cat(stats).cpu().numpy()


def finish_profile(meta, dset_px_dist, mean_sum, std_sum, n_samples):
    """The [C]-vector tail of get_profile (profile.py:114-150), float64 on the host."""
    meta.n_samples = int(n_samples)
    px_mean = (mean_sum / n_samples).astype(np.float32)
    px_std = (std_sum / n_samples).astype(np.float32)
    dset_px_dist = np.asarray(dset_px_dist, dtype=np.int64)
    dset_px_count = np.sum(dset_px_dist)
    probs = dset_px_dist / dset_px_count
    meta.tile_px_count = meta.tile_size * meta.tile_size
    assert dset_px_count / meta.tile_px_count == meta.n_samples, "Pixel distribution does not match tile count."
    weights = 1 / (np.log(1.02 + probs))
    weights = weights / np.max(weights)
    balanced = np.full(meta.n_classes, 1 / meta.n_classes)
    meta.m2 = m2(probs, meta.n_classes)
    meta.jsd = jsd(probs, balanced)
    meta.px_mean = px_mean.tolist()
    meta.px_std = px_std.tolist()
    meta.probs = probs.tolist()
    meta.weights = weights.tolist()
    meta.dset_px_count = int(dset_px_count)
    meta.dset_px_dist = dset_px_dist.tolist()
    return meta


def get_profile(dset, distributed=False):
    """Statistical profile of a tile dataset (reference profile.py:21-150).  With
    `distributed=True` (one process per GPU, each holding a shard of the tiles) the dataset-level
    sums are all-reduced; meta.px_dist then holds this rank's rows only."""
    meta = dset.get_meta()
    db = dset.db
    imgs = db.data['img'][db.start:db.end]
    masks = db.data['mask'][db.start:db.end]
    px_dist, stat = profile_arrays(imgs, masks, meta.n_classes)
    mean, std = tile_moments_to_stats(stat, meta.tile_size * meta.tile_size)
    dset_hist = px_dist.sum(axis=0)
    moments = np.concatenate([mean.sum(axis=0), std.sum(axis=0), [float(len(px_dist))]])
    if distributed and pdist.world_size() > 1:
        dset_hist = pdist.all_reduce_i64(dset_hist)     # exact for any GPU count
        moments = pdist.all_reduce_f64(moments)
    ch = mean.shape[1]
    meta.px_dist = px_dist.tolist()
    return finish_profile(meta, dset_hist, moments[:ch], moments[ch:2 * ch], int(round(moments[-1])))


def print_meta(meta):
    """Console read-out of profile metadata (reference profile.py:153-207)."""
    hline = '-' * 50
    print('\nProfile Metadata')
    print(hline)
    print('{:30s} {}'.format('ID', meta.id))
    print('{:30s} {} ({})'.format('Channels', meta.ch, 'Grayscale' if meta.ch == 1 else 'Colour'))
    print('{:30s} {}'.format('Classes', meta.n_classes))
    print('{:30s} {}'.format('Samples', meta.n_samples))
    print('{:30s} {}px x {}px'.format('Tile size (WxH)', meta.tile_size, meta.tile_size))
    print('{:30s} {}'.format('Pixel mean', meta.px_mean))
    print('{:30s} {}'.format('Pixel std-dev', meta.px_std))
    print('{:30s} {}'.format('M2', meta.m2))
    print('{:30s} {}'.format('JSD', meta.jsd))
    print('\n{:8s}{:25s}{:>14s}{:>10s}'.format('Code', 'Name', 'Probs', 'Weights'))
    print(hline)
    for i in range(meta.n_classes):
        print('{:8s}{:25s}{:14.5f}{:10.5f}'.format(
            meta.class_codes[i], meta.class_labels[i], meta.probs[i], meta.weights[i]))
    print('{:30s} {}'.format('Tile pixel count', meta.tile_px_count))
    print('{:30s} {}'.format('Dataset pixel count', meta.dset_px_count))
    print()
