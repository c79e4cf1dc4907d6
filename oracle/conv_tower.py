"""ORACLE (test infrastructure, never the product path): CPU restatement of the
reference's siamese conv tower.

Follows /root/reference/mc_cnn_brunch.py:31-48 (Net.construct: conv1 1->64 ReLU,
conv2..conv{nl-1} 64->64 ReLU, conv{nl} linear, all 3x3 VALID stride 1, then
tf.nn.l2_normalize over channels), mc_cnn_brunch.py:70-92 (conv = conv2d + bias_add
[+ relu]) and /root/reference/process_functional.py:13-19 (zero-pad the standardised
image ONCE by (patch-1)//2 on each side; not per-layer SAME padding).

The arithmetic is owned by TensorFlow 1.x (unpinned, not installed here), so this
part of the oracle is "parity unpinned": it is pinned only by definition
(l2_normalize(x) = x * rsqrt(max(sum(x^2), 1e-12))) and by an fp64 twin.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.
"""
from __future__ import annotations

import numpy as np
import torch


def pad_image(image_hw1: np.ndarray, patch_height: int = 11, patch_width: int = 11) -> np.ndarray:
    """process_functional.py:13-19 -> [1, H+ph-1, W+pw-1, 1] f32."""
    height, width = image_hw1.shape[0:2]
    out = np.zeros([1, height + patch_height - 1, width + patch_width - 1, 1], dtype=np.float32)
    r0, c0 = (patch_height - 1) // 2, (patch_width - 1) // 2
    out[0, r0:height + r0, c0:width + c0] = image_hw1.reshape(height, width, 1)
    return out


def conv_tower(padded_nhwc: np.ndarray, weights: dict, num_layers: int = 5,
               dtype=torch.float32, num_threads: int | None = None) -> np.ndarray:
    """[1,Hp,Wp,1] -> [H,W,F] features, same dtype as requested (f32 oracle / f64 twin)."""
    if num_threads:
        torch.set_num_threads(num_threads)
    x = torch.from_numpy(np.ascontiguousarray(padded_nhwc)).to(dtype).permute(0, 3, 1, 2)
    with torch.no_grad():
        for i in range(1, num_layers + 1):
            w = torch.from_numpy(weights[f"conv{i}/weights:0"]).to(dtype).permute(3, 2, 0, 1).contiguous()  # HWIO->OIHW
            b = torch.from_numpy(weights[f"conv{i}/biases:0"]).to(dtype)
            x = torch.nn.functional.conv2d(x, w, b, stride=1, padding=0)
            if i < num_layers:
                x = torch.relu(x)
        ss = torch.sum(x * x, dim=1, keepdim=True)
        x = x * torch.rsqrt(torch.clamp(ss, min=1e-12))
    return x[0].permute(1, 2, 0).contiguous().numpy()


def compute_feature(left_image, right_image, patch_height, patch_width, num_of_feature_maps, weights,
                    dtype=torch.float32):
    """Oracle twin of process_functional.compute_feature (:11-45) taking the weight dict."""
    nl = patch_height // 2
    fl = conv_tower(pad_image(left_image, patch_height, patch_width), weights, nl, dtype)
    fr = conv_tower(pad_image(right_image, patch_height, patch_width), weights, nl, dtype)
    assert fl.shape[-1] == num_of_feature_maps
    return fl, fr
