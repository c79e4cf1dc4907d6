"""Seeded synthetic inputs for the MC-CNN stereo hot path.

No data and no checkpoint ship with the reference (its eval/, test_data/ dirs hold
empty readme.txt files), so every parity vector and every benchmark input is made
here: textured u8 stereo pairs with a known piecewise-constant disparity field, pure
noise pairs (adversarial for the unsigned-wrap penalty test), Glorot-uniform conv
weights in the reference's ``{'conv{i}/weights:0': HWIO, 'conv{i}/biases:0'}`` dict
layout (mc_cnn_brunch.py:61-66, 76-77), and unit-norm feature maps.

NumPy only (scipy for the blur) so that tests, bench and the oracle can all use it.
"""
from __future__ import annotations

import numpy as np

# (W, H, D) of BASELINE.json's five configs, in order.
CONFIGS = {
    "c1": (463, 370, 80),
    "c2": (695, 555, 128),
    "c3": (1440, 994, 400),
    "c4": (2880, 1988, 800),
    "c5": (1242, 375, 228),
}


def _blur(a: np.ndarray, sigma: float) -> np.ndarray:
    from scipy.ndimage import gaussian_filter

    return gaussian_filter(a, sigma, mode="wrap")


def textured_pair(H: int, W: int, D: int, seed: int = 0):
    """Left/right u8 images + integer ground-truth left disparity.

    Left: three octaves of blurred Gaussian noise rescaled to 0..255, so neighbour
    differences fall on both sides of the SGM threshold 30. Disparity: random
    rectangles of constant value in [D/8, 7D/8] (clipped to < W). Right = left
    shifted by the disparity, occlusion holes filled with fresh noise.
    """
    rng = np.random.default_rng(seed)
    tex = np.zeros((H, W), np.float64)
    for sigma, amp in ((0.7, 1.0), (2.5, 1.5), (8.0, 2.0)):
        tex += amp * _blur(rng.standard_normal((H, W)), sigma) * sigma
    tex = (tex - tex.min()) / max(float(np.ptp(tex)), 1e-12)
    left = np.clip(np.rint(tex * 255.0), 0, 255).astype(np.uint8)

    dmax = max(1, min(D - 1, W - 1))
    lo, hi = max(0, dmax // 8), max(1, (7 * dmax) // 8)
    disp = np.full((H, W), int(rng.integers(lo, hi + 1)), np.int32)
    for _ in range(12):
        y0, x0 = int(rng.integers(0, H)), int(rng.integers(0, W))
        h, w = int(rng.integers(max(1, H // 8), max(2, H // 2))), int(rng.integers(max(1, W // 8), max(2, W // 2)))
        disp[y0:y0 + h, x0:x0 + w] = int(rng.integers(lo, hi + 1))

    right = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
    ys, xs = np.mgrid[0:H, 0:W]
    xr = xs - disp
    ok = xr >= 0
    # nearer (larger disparity) surfaces win where two left pixels land on one right pixel
    order = np.argsort(disp[ok], kind="stable")
    yy, xx, dd = ys[ok][order], xr[ok][order], disp[ok][order]
    right[yy, xx] = left[yy, xx + dd]
    return left, right, disp.astype(np.float32)


def noise_pair(H: int, W: int, seed: int = 0):
    """Independent uniform u8 noise for both images (uint-wrap quirk stress)."""
    rng = np.random.default_rng(seed)
    return (rng.integers(0, 256, size=(H, W), dtype=np.uint8),
            rng.integers(0, 256, size=(H, W), dtype=np.uint8))


def glorot_weights(num_layers: int = 5, num_features: int = 64, ksize: int = 3, seed: int = 7) -> dict:
    """Random-init weights in the reference's .npy dict layout.

    tf.get_variable's default initializer is Glorot-uniform for every variable,
    biases included (mc_cnn_brunch.py:76-77): limit = sqrt(6 / (fan_in + fan_out)).
    """
    rng = np.random.default_rng(seed)
    out = {}
    cin = 1
    for i in range(1, num_layers + 1):
        fan_in, fan_out = ksize * ksize * cin, ksize * ksize * num_features
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        out[f"conv{i}/weights:0"] = rng.uniform(-lim, lim, (ksize, ksize, cin, num_features)).astype(np.float32)
        # a rank-1 [64] variable has fan_in = fan_out = 64 under TF's fan rule
        blim = np.sqrt(6.0 / (num_features + num_features))
        out[f"conv{i}/biases:0"] = rng.uniform(-blim, blim, (num_features,)).astype(np.float32)
        cin = num_features
    return out


def glorot_fc_weights(num_features: int = 64, units: int = 384, seed: int = 11, gain: float = 1.0) -> dict:
    """Random-init weights of the MC-CNN-accurate head in the reference's variable naming (fc(), mc_cnn_brunch.py:95-106):
    fc1 [2*features, units], fc2 / fc3 [units, units], fc4 [units, 1], Glorot-uniform like every tf.get_variable default.
    `gain` scales the weight matrices (tests use > 1 so that the sigmoid leaves its linear range)."""
    rng = np.random.default_rng(seed)
    out = {}
    shapes = [(2 * num_features, units), (units, units), (units, units), (units, 1)]
    for i, (fi, fo) in enumerate(shapes, 1):
        lim = np.sqrt(6.0 / (fi + fo)) * gain
        out[f"fc{i}/weights:0"] = rng.uniform(-lim, lim, (fi, fo)).astype(np.float32)
        blim = np.sqrt(6.0 / (fo + fo))
        out[f"fc{i}/biases:0"] = rng.uniform(-blim, blim, (fo,)).astype(np.float32)
    return out


def standardise(image_u8: np.ndarray) -> np.ndarray:
    """(I - mean) / std with population std, as match_single.py:34-43; returns [H,W,1] f32."""
    img = image_u8.astype(np.float32)
    img = (img - np.mean(img, axis=(0, 1))) / np.std(img, axis=(0, 1))
    return np.expand_dims(img, axis=2).astype(np.float32)


def unit_features(H: int, W: int, F: int = 64, seed: int = 0):
    """Two random unit-norm feature maps [H,W,F] f32 (no conv tower needed)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(2):
        f = rng.standard_normal((H, W, F)).astype(np.float32)
        f /= np.sqrt(np.sum(f.astype(np.float64) ** 2, axis=-1, keepdims=True)).astype(np.float32)
        out.append(f.astype(np.float32))
    return out[0], out[1]


def correlated_features(H: int, W: int, D: int, F: int = 64, seed: int = 0, noise: float = 0.35):
    """Unit-norm features whose best match follows textured_pair's disparity field.

    fr is a smooth random field; fl[y,x] = fr[y, x-d(y,x)] + noise. Gives a cost
    volume with a real minimum structure without running the conv tower.
    """
    rng = np.random.default_rng(seed)
    _, _, disp = textured_pair(H, W, D, seed)
    disp = disp.astype(np.int64)
    base = rng.standard_normal((H, W, F))
    base = np.stack([_blur(base[..., i], 1.2) for i in range(F)], axis=-1)
    fr = base + 0.05 * rng.standard_normal((H, W, F))
    ys, xs = np.mgrid[0:H, 0:W]
    xr = np.clip(xs - disp, 0, W - 1)
    fl = base[ys, xr] + noise * base.std() * rng.standard_normal((H, W, F))

    def nrm(f):
        f = f / np.sqrt(np.sum(f * f, axis=-1, keepdims=True))
        return f.astype(np.float32)

    return nrm(fl), nrm(fr), disp.astype(np.float32)
