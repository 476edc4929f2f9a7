"""
Parameters / defaults -- host-side mirror of PyLC's config.py (reference config.py:22-329).

Same attribute names, defaults and `update()` / `get_schema()` behaviour, so metadata written by
either implementation (HDF5 `meta` attribute, model-file `meta`) loads in the other.
Intentional deviations, all documented in DESIGN.md:
  * schema files resolve against the caller's path first and then against the schemas packaged
    with pylc_b200 (the reference only looks in ./schemas relative to the cwd, config.py:108);
  * a dict argument may carry `schema` (the reference only honours it on attribute-style
    namespaces, config.py:107);
  * import does not reseed the global NumPy / random generators (config.py:163-165).
"""
import json
import os
import random
import sys

import numpy as np
import torch

_PKG_SCHEMAS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "schemas")


def resolve_schema_path(path):
    if path and os.path.isfile(path):
        return path
    if path:
        cand = os.path.join(_PKG_SCHEMAS, os.path.basename(path))
        if os.path.isfile(cand):
            return cand
    return path


class Schema(object):
    pass


class Parameters:
    """Flat attribute bag of package parameters (general, schema, extraction, profile, network)."""

    def __init__(self, args=None):
        # general
        self.id = None
        self.ch = 3
        self.ch_options = [1, 3]
        self.ch_label = 'grayscale' if self.ch == 1 else 'colour'

        # device
        self.device = "cuda:0" if torch.cuda.is_available() else "cpu"
        self.n_workers = 0

        # run modes
        self.TRAIN, self.VALID, self.TEST = 'train', 'valid', 'test'
        self.EXTRACT, self.AUGMENT, self.PROFILE = 'extract', 'augment', 'profile'
        self.GRAYSCALE, self.MERGE = 'grayscale', 'merge'

        # categorisation schema (default: LCC.A)
        schema_arg = None
        if args is not None:
            schema_arg = args.get('schema') if isinstance(args, dict) else getattr(args, 'schema', None)
        self.schema = schema_arg if schema_arg else './schemas/schema_a.json'
        self.schema_name = str(os.path.splitext(os.path.basename(self.schema))[0])
        schema = self.get_schema(self.schema)
        self.class_labels = schema.class_labels
        self.class_codes = schema.class_codes
        self.palette_hex = schema.palette_hex
        self.palette_rgb = schema.palette_rgb
        self.n_classes = schema.n_classes
        self.class_labels_hex = {schema.palette_hex[i]: schema.class_labels[i] for i in range(len(self.palette_hex))}

        # default paths
        self.root = './data/'
        self.img_dir = './data/raw/images/'
        self.mask_dir = './data/raw/masks/'
        self.db_dir = './data/db/'
        self.output_dir = './data/outputs/'
        self.save_dir = './data/save/'
        self.model_dir = './data/models/'
        self.meta_grayscale_path = './data/metadata/meta_ch1_schema_a.npy'
        self.meta_colour_path = './data/metadata/meta_ch3_schema_a.npy'

        # extraction
        self.n_samples = 0
        self.tile_size = 512
        self.stride = 512
        self.scale = 1.
        self.scales = [1.]
        self.tiling_factor = 700
        self.tiles_per_image = self._tiles_per_image()
        self.tile_px_count = self.tile_size * self.tile_size

        # augmentation
        self.aug_n_samples_ratio = 0.36
        self.aug_oversample_rate_range = (0, 4)
        self.aug_rate_coef = 0.
        self.aug_rate_coef_range = (1, 21)
        self.aug_threshold = 0.
        self.aug_threshold_range = (0, 3.)
        self.alpha = 0.19

        # database
        self.buffer_size = 1000
        self.partition = 0.2
        self.clip = 1.
        self.clip_overfit = 0.003

        self.seed = random.randrange(sys.maxsize)

        # normalisation defaults (normally computed by profiling)
        self.normalize_default = False
        self.gs_mean = 0.456
        self.gs_std = 0.225
        self.px_rgb_mean = [132.47, 144.47, 149.45]
        self.px_rgb_std = [24.85, 22.04, 18.77]
        self.px_grayscale_mean = 142.01
        self.px_grayscale_std = 23.66

        # profile metadata
        self.px_mean = None
        self.px_std = None
        self.px_dist = None
        self.dset_px_dist = None
        self.dset_px_count = 0
        self.probs = None
        self.weights = None
        self.m2 = 0.
        self.jsd = 1.

        # network
        self.pretrained = './data/models/resnet101-5d3b4d8f.pth'
        self.n_epochs = 20
        self.batch_size = 8
        self.dropout = 0.5
        self.crop_target = False
        self.lr = 0.0001
        self.lr_min = 1e-6
        self.lr_max = 0.1
        self.gamma = 0.9
        self.l2_reg = 1e-4
        self.in_channels = 3
        self.momentum = 0.9
        self.weighted = False
        self.dice_weight = 0.5
        self.ce_weight = 0.5
        self.focal_weight = 0.5
        self.dice_smooth = 1.
        self.weight_decay = 5e-5
        self.fl_gamma = 2
        self.fl_alpha = 0.25
        self.fl_reduction = 'mean'
        self.grad_steps = 16
        self.test_intv = 70
        self.optim_options = ['adam', 'sgd']
        self.optim_type = self.optim_options[0]
        self.sched_options = ['step_lr', 'cyclic_lr', 'anneal']
        self.sched_type = self.sched_options[0]
        self.arch_options = ['deeplab', 'unet', 'resunet']
        self.arch = self.arch_options[0]
        self.backbone_options = ['resnet', 'xception']
        self.backbone = self.backbone_options[0]
        self.norm_options = ['batch', 'instance', 'layer', 'synbatch']
        self.norm_type = self.norm_options[0]
        self.activ_options = ['relu', 'lrelu', 'selu', 'synbatch']
        self.activ_type = self.activ_options[0]

        # U-Net geometry (kept for metadata compatibility)
        self.output_size = 324
        self.input_size = 512
        self.pad_size = (self.input_size - self.output_size) // 2
        self.up_mode = 'upsample'
        self.up_mode_options = ['upconv', 'upsample']
        self.crop_left = self.pad_size
        self.crop_right = self.pad_size + self.output_size
        self.crop_up = self.pad_size
        self.crop_down = self.pad_size + self.output_size

        # loss tracking / metrics
        self.resume_checkpoint = False
        self.report = 20
        self.save_logits = False
        self.aggregate_metrics = False

        if args:
            self.update(args)

    def _tiles_per_image(self):
        # reference: int(sum(tiling_factor * scales)) -- list repetition, 700 per scale of 1.0
        return int(sum(self.tiling_factor * list(self.scales)))

    def update(self, args):
        """Copy every same-named key of a dict / namespace / Parameters onto this object."""
        params = args if isinstance(args, dict) else vars(args)
        for key, value in params.items():
            if hasattr(self, key):
                if isinstance(value, (np.ndarray, torch.Tensor)):
                    value = value.tolist()
                setattr(self, key, value)
        self.ch_label = 'grayscale' if self.ch == 1 else 'colour'
        self.tiles_per_image = self._tiles_per_image()
        return self

    def get_schema(self, schema_path):
        schema_path = resolve_schema_path(schema_path if schema_path else self.schema)
        if not schema_path or not os.path.isfile(schema_path):
            print('Schema file not found:\n\t{}'.format(schema_path))
            exit(1)
        schema = Schema()
        with open(schema_path) as f:
            classes = json.load(f)['classes']
        schema.class_labels = [c['label'] for c in classes]
        schema.class_codes = [c['code'] for c in classes]
        schema.palette_hex = [c['colour']['hex'] for c in classes]
        schema.palette_rgb = [c['colour']['rgb'] for c in classes]
        schema.n_classes = len(classes)
        return schema

    def print(self):
        readout = '\nGlobal Parameters\n------\n'
        for key, value in vars(self).items():
            readout += '\n{:20s}{:20s}'.format(str(key), str(value))
        readout += '\n------\n'
        print(readout)


defaults: Parameters = Parameters()
