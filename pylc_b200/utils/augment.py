"""
Augmentor -- host-side mirror of PyLC's utils/augment.py for the sample-rate optimisation
(reference augment.py:92-187), the consumer of the per-tile class histograms `px_dist` that the
extraction / profiling kernels produce.

`optimize()` is the reference's grid search over (rate coefficient, threshold): the per-tile scores
are a handful of NumPy operations on [N, C] (kept verbatim, so thresholds decisions see the same
float64 values), the grid itself -- |coefs| x |thresholds| passes over [N, C], each building the
over-sampling rates and the class histogram of the over-sampled dataset -- is ONE kernel launch
(pylc_sample_rate_grid, exact int64), and the O(grid x C) tail (probabilities, M2, JSD, argmin)
is the reference's float64 NumPy again.

`oversample()` (reference augment.py:184-239) copies every tile and appends `rates[i]` warped copies of
tile i.  The warps (tools.augment_transform: perspective jitter + brightness shift) are the reference's own
OpenCV calls on host arrays -- same library, same arguments, same bytes (pinned by a golden vector from the
reference) -- and the profile of the augmented set runs on the device kernels (utils/profile.py).
"""
import numpy as np
import torch

from .. import ops
from ..config import defaults
from . import metrics


class Augmentor(object):
    def __init__(self):
        self.input_path = None
        self.input_dset = None
        self.input_meta = None
        self.input_size = 0
        self.output_meta = None
        self.optim_meta = None
        self.profile_data = []
        self.rates = []

    def load(self, db_path):
        """`db_path`: a database path (reference augment.py:45-76) or, as an extension, an MLPDataset."""
        from ..db.dataset import MLPDataset
        self.input_dset = db_path if isinstance(db_path, MLPDataset) else MLPDataset(db_path)
        self.input_path = db_path if isinstance(db_path, str) else None
        self.input_meta = self.input_dset.get_meta()
        self.input_size = self.input_dset.size
        return self

    def load_profile(self, meta, n_tiles=None):
        """Use profile metadata directly (px_dist [N,C], tile_px_count, probs [C], n_classes)."""
        self.input_meta = meta
        self.input_size = int(n_tiles if n_tiles is not None else len(meta.px_dist))
        return self

    def scores(self):
        """Per-tile over-sampling scores, reference augment.py:104-116 verbatim."""
        eps = 1e-8
        px_dist = np.array(self.input_meta.px_dist, dtype='long')
        px_count = self.input_meta.tile_px_count
        dset_probs = np.array(self.input_meta.probs, dtype='float32') + eps
        oversample_filter = np.clip(1 / self.input_meta.n_classes - dset_probs, a_min=0, a_max=1.)
        probs = px_dist / px_count
        probs_weighted = np.multiply(np.multiply(probs, 1 / dset_probs), oversample_filter)
        return np.sqrt(np.sum(probs_weighted, axis=1)), px_dist, px_count

    def optimize(self, device=None):
        """Grid search for the sample rates that minimise the JSD to a balanced distribution
        (reference augment.py:92-187).  Sets .optim_meta, .rates, .profile_data."""
        device = device or torch.device("cuda")
        n_classes = self.input_meta.n_classes
        scores, px_dist, px_count = self.scores()
        rate_coefs = np.arange(min(defaults.aug_rate_coef_range), max(defaults.aug_rate_coef_range), 1.)
        thresholds = np.arange(min(defaults.aug_threshold_range), max(defaults.aug_threshold_range), 0.05)
        assert (rate_coefs >= 1).all(), 'Rate coefficient must be >= 1.'
        lo, hi = defaults.aug_oversample_rate_range
        d_scores = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float64)).to(device)
        d_dist = torch.from_numpy(np.ascontiguousarray(px_dist, dtype=np.int64)).to(device)
        sum_rates, full = ops.sample_rate_grid(d_scores, d_dist, torch.from_numpy(rate_coefs).to(device),
                                               torch.from_numpy(thresholds).to(device), lo, hi)
        sum_rates, full = sum_rates.cpu().numpy(), full.cpu().numpy()       # [G], [G, C]: a few KB

        balanced_px_prob = np.empty(n_classes)
        balanced_px_prob.fill(1 / n_classes)
        limit = int(defaults.aug_n_samples_ratio * self.input_size)
        profile_data, jsd = [], []
        for i, rate_coef in enumerate(rate_coefs):
            for j, threshold in enumerate(thresholds):
                g = i * len(thresholds) + j
                if sum_rates[g] < limit:
                    full_px_dist_sum = full[g]
                    full_px_probs = full_px_dist_sum / np.sum(full_px_dist_sum)
                    m2_sample = metrics.m2(full_px_probs, n_classes)
                    jsd_sample = metrics.jsd(full_px_probs, balanced_px_prob)
                    jsd += [jsd_sample]
                    profile_data += [{
                        'probs': full_px_probs, 'threshold': threshold, 'rate_coef': rate_coef,
                        'n_samples': int(np.sum(full_px_dist_sum) / px_count), 'aug_n_samples': int(sum_rates[g]),
                        'n_samples_max': defaults.aug_n_samples_ratio, 'jsd': jsd_sample, 'm2': m2_sample}]
        assert len(jsd) > 0, 'No augmentation optimization found.'
        self.profile_data = profile_data
        self.optim_meta = profile_data[int(np.argmin(np.asarray(jsd)))]
        # the per-tile rates of the winning grid point only (the reference keeps them for every point)
        over = scores > self.optim_meta['threshold']
        rates = np.multiply(over, self.optim_meta['rate_coef'] * scores).astype(int)
        self.optim_meta['rates'] = np.clip(rates, lo, hi)
        self.rates = self.optim_meta['rates']
        return self

    def oversample(self, shuffle=True, device=None):
        """Originals + `rates[i]` augmented copies of tile i (reference augment.py:184-239): copy j of a tile is
        tools.augment_transform with RandomState(j), exactly as the reference seeds it.  Sets .output_imgs,
        .output_masks, .output_meta (profiled on the device).
        device=True (or PYLC_AUG_DEVICE=1): the copies are made by pylc_augment_tiles_u8 -- the same bytes as the
        OpenCV chain (csrc/augment_math.cuh), one launch for all copies; default: the reference's own OpenCV calls
        on the host."""
        import os
        from ..config import Parameters
        from ..db.dataset import MLPDataset
        from .profile import get_profile
        from .tools import augment_params, augment_transform, coshuffle
        assert self.input_dset is not None and self.input_dset.size > 0, "Loaded input dataset is empty."
        assert len(self.rates) == self.input_dset.size, "Run optimize() first: one rate per input tile."
        if device is None:
            device = os.environ.get("PYLC_AUG_DEVICE", "0") == "1"
        db = self.input_dset.db
        src_imgs, src_masks = db.data['img'][db.start:db.end], db.data['mask'][db.start:db.end]
        if torch.is_tensor(src_imgs):
            src_imgs, src_masks = src_imgs.cpu().numpy(), src_masks.cpu().numpy()
        n_out = int(self.input_dset.size + np.sum(self.rates))
        imgs = np.empty((n_out,) + tuple(src_imgs.shape[1:]), dtype=np.uint8)
        masks = np.empty((n_out,) + tuple(src_masks.shape[1:]), dtype=np.uint8)
        if device:
            # job table on the host (17 numbers per copy, the reference's RandomState(j) draws), pixels on the device
            rates = np.asarray(self.rates, dtype=np.int64)
            T = int(src_masks.shape[-1])
            job_src = np.repeat(np.arange(len(rates)), rates)
            params = [augment_params(np.random.RandomState(j), T) for i in range(len(rates)) for j in range(int(rates[i]))]
            first = np.arange(len(rates)) + np.concatenate([[0], np.cumsum(rates)[:-1]])      # output slot of original i
            copy_slots = np.setdiff1d(np.arange(n_out), first)                                 # copies follow their original
            imgs[first], masks[first] = np.asarray(src_imgs), np.asarray(src_masks)
            if len(job_src):
                from .tools import _device
                dev = _device()
                out_i, out_m = ops.augment_tiles(torch.from_numpy(np.ascontiguousarray(src_imgs, dtype=np.uint8)).to(dev),
                                                 torch.from_numpy(np.ascontiguousarray(src_masks, dtype=np.uint8)).to(dev),
                                                 job_src, np.stack([p[0] for p in params]), np.array([p[1] for p in params]))
                imgs[copy_slots], masks[copy_slots] = out_i.cpu().numpy(), out_m.cpu().numpy()
            return self._finish_oversample(imgs, masks, shuffle)
        idx = 0
        for i in range(self.input_dset.size):
            img, mask = np.asarray(src_imgs[i:i + 1]), np.asarray(src_masks[i:i + 1])
            imgs[idx], masks[idx] = img[0], mask[0]
            idx += 1
            # the loader hands the reference float32 images and int64 masks (db/dataset.py:62-63)
            img_f, mask_l = img.astype(np.float32), mask.astype(np.int64)
            for j in range(int(self.rates[i])):
                inp, tgt = augment_transform(img_f, mask_l, np.random.RandomState(j))
                imgs[idx] = np.asarray(torch.as_tensor(inp, dtype=torch.uint8).numpy()).reshape(imgs.shape[1:])
                masks[idx] = torch.as_tensor(tgt, dtype=torch.uint8).numpy()
                idx += 1
        assert idx == n_out
        return self._finish_oversample(imgs, masks, shuffle)

    def _finish_oversample(self, imgs, masks, shuffle):
        from ..config import Parameters
        from .profile import get_profile
        from .tools import coshuffle
        if shuffle:
            imgs, masks = coshuffle(imgs, masks)
        self.output_imgs, self.output_masks = imgs, masks
        self.output_meta = Parameters().update(vars(self.input_meta))
        self.output_meta = get_profile(self.get_data())
        self.output_meta.id = '_aug' + str(self.input_meta.id)
        return self

    def print_settings(self):
        """Augmentation read-out: augmented against input sample counts, M2, JSD and the winning grid point
        (reference augment.py:348-378)."""
        rule = '_' * 70
        row = ' {:30s}{:20s}{:20s}'.format
        lines = ['', '', 'Augmentation Results', rule, row(' ', 'Augmented', 'Input'), rule,
                 row('Total Samples', str(self.output_meta.n_samples), str(self.input_meta.n_samples)),
                 row('Total Generated', str(len(self.output_imgs) - self.input_size), '-'),
                 row('M2', str(round(self.output_meta.m2, 4)), str(round(self.input_meta.m2, 4))),
                 row('JSD', str(round(self.output_meta.jsd, 4)), str(round(self.input_meta.jsd, 4))), rule,
                 row('Threshold [Range]', str(self.optim_meta['threshold']), str(defaults.aug_threshold_range)),
                 row('Rate Coefficient [Range]', str(self.optim_meta['rate_coef']), str(defaults.aug_rate_coef_range)),
                 ' {:30s}{:20s}'.format('Max Rate', str(defaults.aug_oversample_rate_range[1])), rule, '']
        print('\n'.join(lines))
        return self

    def get_data(self):
        from ..db.dataset import MLPDataset
        return MLPDataset(input_data={'img': self.output_imgs, 'mask': self.output_masks, 'meta': self.output_meta})
