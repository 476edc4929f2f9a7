"""
Tile extractor -- host-side mirror of PyLC's utils/extract.py (reference extract.py:25-385).

Same chainable interface: Extractor(params).load(img, mask).extract(fit, stride, scale)
.coshuffle().profile().get_data(); same `meta.extract` dict, capacity check and messages.  The
per-pixel work is done by the sm_100a kernels:

    image tiles  Extractor.__split + np.copyto (extract.py:279-310,182)      -> pylc_tile_gather_u8[_stack]
    mask tiles   __split + tools.class_encode + np.copyto (extract.py:195-214) -> pylc_mask_gather_encode_hist[_stack]
                 (the per-tile class histograms profile() needs fall out of the same pass)

Decode and the optional fit-resize stay on the host (OpenCV, bit-identical by construction), on a few
threads that run ahead of the loop in file order (tools.ordered_prefetch).  Consecutive files of one
size form a STACK: one pinned staging copy, one upload and ONE launch per kernel for up to STACK_MAX
files (pylc_tile_gather_u8_stack / pylc_mask_gather_encode_hist_stack); tiles are written straight
into device-resident buffers sized exactly (the reference pre-allocates n_files*700 host tiles).
`.imgs` / `.masks` are CUDA u8 tensors [N,ch,T,T] / [N,T,T]; `.host()` returns NumPy copies.
"""
import os
import time

import cv2
import numpy as np
import torch

from .. import ops
from ..config import Parameters, defaults
from ..db.dataset import MLPDataset
from . import tools as utils
from .profile import finish_profile, get_profile, print_meta, tile_moments_to_stats


class Extractor(object):
    STACK_MAX = 32             # images per extraction launch
    STACK_BYTES = 1 << 30      # bound on one stack's device + pinned staging footprint (image + RGB mask)

    def __init__(self, params=None):
        self.meta = Parameters(params) if params is not None else defaults
        self.verbose = True
        self.reset()

    # ---- reference API -----------------------------------------------------------------------
    def load(self, img_path, mask_path=None):
        """Collate image / mask paths (reference extract.py:58-104)."""
        self.reset()
        self.img_path = img_path
        self.mask_path = mask_path
        self.files = utils.collate(img_path, mask_path)
        self.n_files = len(self.files)
        if self.n_files == 0:
            print('File list is empty. Extraction stopped.')
            exit(1)
        self.imgs_capacity = self.masks_capacity = self.n_files * self.meta.tiles_per_image
        return self

    def load_arrays(self, images, masks=None, names=None):
        """Extension: extract from already decoded arrays (u8 [H,W] / [H,W,3]; masks RGB [H,W,3])
        instead of files -- used by benchmarks and tests, which start from decoded pixels."""
        self.reset()
        self.files = [{'img_data': im, 'mask_data': (masks[i] if masks is not None else None),
                       'name': names[i] if names else 'array_%d' % i} for i, im in enumerate(images)]
        self.n_files = len(self.files)
        self.mask_path = 'arrays' if masks is not None else None
        self.img_path = 'arrays'
        self.imgs_capacity = self.masks_capacity = self.n_files * self.meta.tiles_per_image
        return self

    def extract(self, fit=False, stride=None, scale=None):
        """Tile every loaded image (and mask) into [N,ch,T,T] / [N,T,T] u8 (extract.py:106-231)."""
        if stride:
            self.meta.stride = stride
        if scale:
            self.meta.scales = [scale]
            self.meta.tiles_per_image = int(self.meta.tiling_factor * scale)
        self.fit = fit
        if self.verbose:
            self.print_settings()
        T, S, ch = self.meta.tile_size, self.meta.stride, self.meta.ch
        dev = utils._device()
        img_parts, mask_parts, dist_parts, stat_parts = [], [], [], []
        n_img = n_mask = 0
        # Consecutive files of one size are extracted as a STACK: one upload and one launch per kernel for up to
        # STACK_MAX images (pylc_tile_gather_u8_stack / pylc_mask_gather_encode_hist_stack), tiles in file order.
        pending = []

        def shape_of(a):
            return None if a is None else a.shape

        def flush():
            if not pending:
                return
            imgs, masks = [p[0] for p in pending], [p[1] for p in pending]
            H, W = imgs[0].shape[:2]
            d_imgs, pitch, self._staging[0] = ops.upload_stack(imgs, dev, self._staging[0])
            tiles, stat = ops.tile_gather_u8_stack(d_imgs, H, W, ch, pitch, T, S, stats=True)
            img_parts.append(tiles)
            stat_parts.append(stat)
            if masks[0] is not None:
                mh, mw = masks[0].shape[:2]
                d_masks, mpitch, self._staging[1] = ops.upload_stack(masks, dev, self._staging[1])
                m_tiles, px_dist = ops.mask_gather_encode_hist_stack(d_masks, mh, mw, mpitch, T, S, self.meta.palette_rgb)
                mask_parts.append(m_tiles)
                dist_parts.append(px_dist)
            pending.clear()

        for scale in self.meta.scales:
            if self.verbose:
                print('\nExtraction --- Scaling Factor: {}'.format(scale))
            def decode_fit(fpair, scale=scale):
                img, mask, img_name, mask_name, dims = self._decode(fpair, scale)
                fitted = utils.adjust_to_tile(img, T, S, ch) if self.fit else (img, dims[2], dims[3], 0)
                return fitted, mask, img_name, mask_name, dims

            # files are decoded (and fitted) on host threads ahead of the loop, in file order
            for fitted, mask, img_name, mask_name, dims in utils.ordered_prefetch(decode_fit, self.files):
                w_full, h_full, w_scaled, h_scaled = dims
                img, w_fitted, h_fitted, offset = fitted
                H, W = img.shape[:2]
                nH, nW = ops.tile_grid(H, W, T, S)
                n_tiles = nH * nW
                self.meta.extract = {
                    'fid': os.path.basename(img_name.replace('.', '_')) + '_scale_' + str(scale),
                    'n': n_tiles, 'w_full': w_full, 'h_full': h_full, 'w_scaled': w_scaled, 'h_scaled': h_scaled,
                    'w_fitted': w_fitted, 'h_fitted': h_fitted, 'offset': offset}
                if self.verbose:
                    self.print_result("Image", img_name, self.meta.extract)
                if n_tiles > self.imgs_capacity:
                    print('Data array reached capacity. Increase the number of tiles per image.')
                    exit(1)
                n_img += n_tiles
                if mask is not None:
                    assert mask.shape[1] == w_scaled and mask.shape[0] == h_scaled, \
                        "Dimensions do not match: \n\tImage {}\n\tMask {}.".format(img_name, mask_name)
                    # the mask is tiled at its own (scaled, unfitted) size, as in the reference (extract.py:189-195)
                    mH, mW = ops.tile_grid(mask.shape[0], mask.shape[1], T, S)
                    if self.verbose:
                        md = dict(self.meta.extract, n=mH * mW,
                                  fid=os.path.basename(mask_name.replace('.', '_')) + '_scale_' + str(scale))
                        self.print_result("Mask", mask_name, md)
                    n_mask += mH * mW
                if pending and (pending[0][0].shape != img.shape or shape_of(pending[0][1]) != shape_of(mask)
                                or len(pending) >= self.STACK_MAX
                                or (len(pending) + 1) * img.size * 4 > self.STACK_BYTES):
                    flush()
                pending.append((img, mask))
        flush()
        self.imgs = torch.cat(img_parts) if len(img_parts) > 1 else img_parts[0]
        self.img_idx = n_img
        self._stat = torch.cat(stat_parts) if len(stat_parts) > 1 else stat_parts[0]
        if mask_parts:
            self.masks = torch.cat(mask_parts) if len(mask_parts) > 1 else mask_parts[0]
            self._px_dist = torch.cat(dist_parts) if len(dist_parts) > 1 else dist_parts[0]
        else:
            # the reference leaves an uninitialised mask buffer in place (extract.py:96-102,221-222)
            self.masks = torch.zeros((n_img, T, T), dtype=torch.uint8, device=dev)
            self._px_dist = None
        self.mask_idx = n_mask
        self.meta.n_tiles = len(self.imgs)
        if self.verbose:
            print()
            print('{:30s}{}'.format('Total image tiles generated:', self.meta.n_tiles))
            if self.mask_path:
                print('{:30s}{}'.format('Total mask tiles generated:', len(self.masks)))
            print()
        return self

    def reset(self):
        self.img_path = None
        self.mask_path = None
        self.files = None
        self.n_files = 0
        self.img_idx = 0
        self.imgs = None
        self.imgs_capacity = 0
        self.mask_idx = 0
        self.masks = None
        self.masks_capacity = 0
        self._px_dist = None
        self._stat = None
        self._perm = None
        self._staging = [None, None]
        self.fit = False
        self.meta.id = '_db_pylc_' + self.meta.ch_label + '_' + str(int(time.time()))
        return self

    def profile(self, distributed=False):
        """Profile metadata of the extracted tiles (reference extract.py:262-269).  The per-tile
        histograms and pixel moments were accumulated by the extraction kernels, so no second
        pass over the tiles is made unless the data was replaced after extract()."""
        if self._px_dist is None or self._stat is None or len(self._px_dist) != len(self.imgs):
            self.meta = get_profile(self.get_data(), distributed=distributed)
        else:
            from .. import dist as pdist
            px_dist = self._px_dist.cpu().numpy()
            mean, std = tile_moments_to_stats(self._stat.cpu().numpy(), self.meta.tile_size * self.meta.tile_size)
            hist = px_dist.sum(axis=0)
            moments = np.concatenate([mean.sum(axis=0), std.sum(axis=0), [float(len(px_dist))]])
            if distributed and pdist.world_size() > 1:
                hist = pdist.all_reduce_i64(hist)
                moments = pdist.all_reduce_f64(moments)
            ch = mean.shape[1]
            self.meta.px_dist = px_dist.tolist()
            self.meta = finish_profile(self.meta, hist, moments[:ch], moments[ch:2 * ch], int(round(moments[-1])))
        if self.verbose:
            print_meta(self.meta)
        return self

    def coshuffle(self, permutation=None):
        """One permutation applied to images, masks (and their per-tile statistics)."""
        n = len(self.imgs)
        perm = np.random.permutation(n) if permutation is None else np.asarray(permutation)
        idx = torch.as_tensor(perm, device=self.imgs.device)
        self.imgs, self.masks = self.imgs[idx], self.masks[idx]
        if self._px_dist is not None:
            self._px_dist = self._px_dist[idx]
        if self._stat is not None:
            self._stat = self._stat[idx]
        self._perm = perm
        return self

    def get_meta(self):
        return self.meta

    def get_data(self):
        return MLPDataset(input_data={'img': self.imgs, 'mask': self.masks, 'meta': self.meta})

    def host(self):
        """(imgs, masks) as NumPy arrays -- what the reference's `.imgs` / `.masks` hold."""
        return self.imgs.cpu().numpy(), self.masks.cpu().numpy()

    # ---- helpers -----------------------------------------------------------------------------
    def _decode(self, fpair, scale):
        if isinstance(fpair, dict) and 'img_data' in fpair:
            img = np.ascontiguousarray(fpair['img_data'], dtype=np.uint8)
            h, w = img.shape[:2]
            mask = fpair['mask_data']
            return img, (None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)), \
                fpair['name'], fpair['name'] + '_mask', (w, h, w, h)
        if isinstance(fpair, dict) and 'img' in fpair and 'mask' in fpair:
            img_path, mask_path = fpair.get('img'), fpair.get('mask')
        else:
            img_path, mask_path = fpair, None
        img, w_full, h_full, w_scaled, h_scaled = utils.get_image(
            img_path, self.meta.ch, scale=scale, interpolate=cv2.INTER_AREA)
        mask = None
        if self.mask_path and mask_path:
            mask, _, _, w_m, h_m = utils.get_image(mask_path, 3, scale=scale, interpolate=cv2.INTER_NEAREST)
            assert w_m == w_scaled and h_m == h_scaled, \
                "Dimensions do not match: \n\tImage {}\n\tMask {}.".format(img_path, mask_path)
        return img, mask, img_path, mask_path, (w_full, h_full, w_scaled, h_scaled)

    def print_settings(self):
        hline = '-' * 40
        print('\nExtraction Configuration')
        print(hline)
        print('{:30s} {}'.format('ID', self.meta.id))
        print('{:30s} {}'.format('Image(s) path', self.img_path))
        print('{:30s} {}'.format('Masks(s) path', self.mask_path))
        print('{:30s} {}'.format('Output path', self.meta.output_dir))
        print('{:30s} {}'.format('Number of files', self.n_files))
        print('{:30s} {}'.format('Scaling', self.meta.scales))
        print('{:30s} {} ({})'.format('Channels', self.meta.ch, 'Grayscale' if self.meta.ch == 1 else 'Colour'))
        print('{:30s} {}px'.format('Stride', self.meta.stride))
        print('{:30s} {}px x {}px'.format('Tile size (WxH)', self.meta.tile_size, self.meta.tile_size))
        print('{:30s} {}'.format('Maximum tiles/image', self.meta.tiles_per_image))
        print(hline)

    def print_result(self, img_type, img_path, md):
        print()
        print('{:30s} {}'.format('{} File'.format(img_type), os.path.basename(img_path)))
        print('- {:28s} {}px x {}px'.format('W x H Original', md['w_full'], md['h_full']))
        if md['w_scaled'] != md['w_full'] or md['h_scaled'] != md['h_full']:
            print('- {:28s} {}px x {}px'.format(
                'W x H Scaled ({})'.format(round(md['w_scaled'] / md['w_full'], 2)), md['w_scaled'], md['h_scaled']))
        if md['w_fitted'] != md['w_scaled'] or md['h_fitted'] != md['h_scaled']:
            print('- {:28s} {}px x {}px'.format('W x H Fitted for Tiling', md['w_fitted'], md['h_fitted']))
        if md['offset']:
            print('- {:28s} {}px'.format('Crop (offset)', md['offset']))
        print('- {:28s} {}'.format('Number of Tiles', md['n']))
