"""
Deterministic synthetic inputs for benchmarks and examples: there is no network,
so the DST.A-shaped workloads of BASELINE.json are generated, not downloaded.

    image   u8, 64-px blocky field + per-pixel noise; three different fields for colour so the
            image is not grey (PyLC's is_grayscale check, reference tools.py:27-43)
    mask    RGB u8 drawn from the schema palette over 50-px label blocks, uniform or with the
            skewed class distribution of the real DST.A profile (pylc_gpu.ipynb cell 9), plus
            0.1 % off-palette pixels (exercises class_encode's "unmatched -> class 1" rule)
"""
import numpy as np

MLP_SKEW = (0.5495, 0.0, 0.2215, 0.0804, 0.1015, 0.0007, 0.0010, 0.0321, 0.0132)


def _blocks(rng_values, H, W, block):
    return np.repeat(np.repeat(rng_values, block, axis=0), block, axis=1)[:H, :W]


def image(index, W, H, ch):
    rng = np.random.default_rng(1000 + index)
    planes = []
    for _ in range(ch):
        coarse = rng.integers(40, 216, size=((H + 63) // 64, (W + 63) // 64), dtype=np.int16)
        field = _blocks(coarse, H, W, 64) + rng.integers(-32, 33, size=(H, W), dtype=np.int16)
        planes.append(np.clip(field, 0, 255).astype(np.uint8))
    return planes[0] if ch == 1 else np.stack(planes, axis=2)


def labels(index, W, H, n_classes, skew=True, block=50):
    rng = np.random.default_rng(5000 + index)
    if skew:
        p = np.zeros(n_classes)
        k = min(n_classes, len(MLP_SKEW))
        p[:k] = MLP_SKEW[:k]
        p /= p.sum()
    else:
        p = np.full(n_classes, 1.0 / n_classes)
    coarse = rng.choice(n_classes, size=((H + block - 1) // block, (W + block - 1) // block), p=p).astype(np.uint8)
    return _blocks(coarse, H, W, block)


def mask(index, W, H, palette, skew=True, off_palette=0.001):
    lab = labels(index, W, H, len(palette), skew=skew)
    rgb = np.asarray(palette, dtype=np.uint8)[lab]
    n_off = int(W * H * off_palette)
    if n_off:
        rng = np.random.default_rng(9000 + index)
        rgb[rng.integers(0, H, n_off), rng.integers(0, W, n_off)] = rng.integers(0, 256, size=(n_off, 3), dtype=np.uint8)
    return rgb
