"""
pylc_b200 -- B200-native implementation of PyLC's data-parallel tiled-segmentation hot path.

The per-pixel work (tile gather, palette encode, class histograms, stitching with softmax /
argmax / colourise, confusion matrix, the CE+Dice+Focal multi-loss) runs in hand-written sm_100a
CUDA kernels behind the C ABI in include/pylc_b200.h; this package is the Python host layer that
keeps PyLC's own interfaces (Extractor, Evaluator, MultiLoss, Model, MLPDataset, Parameters).
"""
__version__ = "0.1.0"
