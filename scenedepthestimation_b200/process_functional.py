"""Drop-in for the reference's process_functional.py (same names, arguments and returns).

NumPy in / NumPy out, like the reference; everything between runs on the B200 through the
C ABI (engine.py -> libmccnn_b200.so). Differences a caller can see:
  * `checkpoint` is the reference's .npy weight dict (Net.save_weights_dict, mc_cnn_brunch.py:61-66), such a dict, or
    the prefix of a tf.train.Saver checkpoint, read without TensorFlow by tf_checkpoint.py;
  * the disparity count is a parameter (`ndisp`, default 128 = the reference's hard-coded range,
    process_functional.py:125) instead of a constant;
  * the device is the current torch CUDA device, not the reference's cuda.select_device(1) (:1095);
  * the returned right map is the raw right WTA map (the reference returns the median of an
    uninitialised buffer there, App. A6);
  * no per-disparity prints;
  * `params` (optional, a _lib.SgmParams from `sgm_params(...)`) switches on the stages the reference names but
    does not run: cross-based aggregation (cbca_iters), sub-pixel refinement, bilateral filter. Default = reference.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine as _e

NDISP = 128  # process_functional.py:125


def sgm_params(**overrides):
    """The reference's constants (process_functional.py:1141-1144) with optional overrides, e.g.
    sgm_params(cbca_iters=2, subpixel=1)."""
    from . import _lib

    p = _lib.default_sgm_params()
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown parameter {k!r}")
        setattr(p, k, v)
    return p

_weights_cache: dict = {}   # key -> packed device blob; keyed by CONTENT for dicts (never by id(): addresses are reused)
_CACHE_SLOTS = 4


def _content_key(arrays) -> str:
    """Digest of the values, shapes and dtypes of a sequence of (name, array): ~1 ms for the tower's 0.6 MB."""
    import hashlib

    h = hashlib.blake2b(digest_size=16)
    for name, a in arrays:
        a = np.ascontiguousarray(a)
        h.update(f"{name}|{a.dtype.str}|{a.shape}|".encode())
        h.update(a.view(np.uint8).reshape(-1).data)
    return h.hexdigest()


def _cache_put(cache: dict, key, value):
    while len(cache) >= _CACHE_SLOTS:
        cache.pop(next(iter(cache)))   # oldest entry (dicts keep insertion order)
    cache[key] = value
    return value


def _load_weights(checkpoint, num_layers):
    if isinstance(checkpoint, dict):
        weights = checkpoint
        names = [f"conv{i}/{k}:0" for i in range(1, num_layers + 1) for k in ("weights", "biases")]
        missing = [n for n in names if n not in weights]
        if missing:
            raise KeyError(f"weights dict lacks {missing[:4]} ...")
        key = ("dict", _content_key((n, weights[n]) for n in names), num_layers)
    else:
        path = os.fspath(checkpoint)
        from . import tf_checkpoint as _tfc

        if path.endswith(".npy"):
            key = ("file", path, os.path.getmtime(path), num_layers)
            weights = None
        elif _tfc.is_checkpoint_prefix(path):
            # a tf.train.Saver checkpoint prefix (what the reference passes, process_functional.py:11,33): read by this repo's own
            # reader of the tensor-bundle format (tf_checkpoint.py; TensorFlow is not needed)
            key = ("ckpt", path, os.path.getmtime(path + ".index"), num_layers)
            weights = _tfc.checkpoint_to_weights(path) if key not in _weights_cache else None
        else:
            raise RuntimeError(
                f"checkpoint {path!r}: neither a .npy weight dict (Net.save_weights_dict, mc_cnn_brunch.py:61-66) nor the prefix of "
                "a tf.train.Saver checkpoint (<prefix>.index / <prefix>.data-00000-of-00001)")
    if key not in _weights_cache:
        if weights is None:
            weights = np.load(checkpoint, encoding="bytes", allow_pickle=True).item()
            weights = {(k.decode() if isinstance(k, bytes) else k): v for k, v in weights.items()}
        _cache_put(_weights_cache, key, _e.pack_weights(weights, num_layers))
    return _weights_cache[key]


def compute_feature(left_image, right_image, patch_height, patch_width, num_of_feature_maps, checkpoint):
    """process_functional.py:11-45: [H,W,1] standardised f32 images -> two [H,W,64] f32 feature maps."""
    _e._require_cuda()
    if patch_height != patch_width or patch_height % 2 == 0:
        raise ValueError("square odd patches only (the tower has patch//2 3x3 layers)")
    if num_of_feature_maps != _e.FEATURES:
        raise ValueError(f"num_of_feature_maps must be {_e.FEATURES}")
    nl = patch_height // 2
    packed = _load_weights(checkpoint, nl)
    height, width = left_image.shape[0:2]
    out = []
    for img in (left_image, right_image):
        d = _e._dev(np.asarray(img, dtype=np.float32).reshape(height, width), torch.float32)
        out.append(_e.conv_tower(_e.pad_f32(d, (patch_height - 1) // 2), packed, nl).cpu().numpy())
    return out[0], out[1]


def compute_cost_volume(featuresl, featuresr, ndisp):
    """process_functional.py:48-73 (the reference's CPU path): -> f32 [ndisp,H,W], invalid entries 0 (negated)."""
    _e._require_cuda()
    fl, fr = _e._dev(featuresl, torch.float32), _e._dev(featuresr, torch.float32)
    cl, _ = _e.cost_volume(fl, fr, int(ndisp), fill=-0.0, right=False)
    return _e.volume_to_dhw(cl, int(ndisp)).cpu().numpy()


def WTA(left_cost_volume):
    """process_functional.py:76-93: argmin over the last axis of [H,W,D]."""
    _e._require_cuda()
    vol = np.asarray(left_cost_volume, dtype=np.float32)
    H, W, D = vol.shape
    Dp = _e.disp_pitch(D)
    dev = torch.zeros((H, W, Dp), dtype=torch.float32, device="cuda")
    dev[:, :, :D] = _e._dev(vol, torch.float32)
    return _e.wta(dev, D).cpu().numpy()


def WTA1(left_cost_volume):
    """process_functional.py:96-113: argmin over the first axis of [D,H,W]."""
    _e._require_cuda()
    return _e.wta_dhw(_e._dev(left_cost_volume, torch.float32)).cpu().numpy()


def disparity_compute_by_gpu(imagel, imager, featuresl, featuresr, detail_time, ndisp=None, params=None, mode="exact"):
    """process_functional.py:1093-1267: u8 images + features -> (left disparity, right disparity, detail_time)."""
    _e._require_cuda()
    assert imagel.shape == imager.shape
    D = int(NDISP if ndisp is None else ndisp)
    il, ir = _e._dev(imagel, torch.uint8), _e._dev(imager, torch.uint8)
    fl, fr = _e._dev(featuresl, torch.float32), _e._dev(featuresr, torch.float32)
    stage = np.zeros(7, np.float32)
    dl, dr = _e.disparity_pipeline(il, ir, fl, fr, D, params=params, stage_ms=stage, mode=mode)
    if detail_time is not None:
        detail_time += (stage / 1000.0).astype(detail_time.dtype)  # the reference accumulates seconds
    return dl.cpu().numpy(), dr.cpu().numpy(), detail_time


_head_cache = {}


def _load_head(head):
    """MC-CNN-accurate head weights: a dict / .npy path holding fc1..fc4 (the reference's fc() variable names,
    mc_cnn_brunch.py:95-106) -> device copy in the kernel's layout, cached per object / path."""
    if head is None or isinstance(head, _e.FcHeadWeights):
        return head
    if isinstance(head, (str, os.PathLike)):
        path = os.fspath(head)
        key = ("file", path, os.path.getmtime(path))
        w = None
    else:
        w = {(k.decode() if isinstance(k, bytes) else k): v for k, v in head.items()}
        key = ("dict", _content_key(sorted((k, np.asarray(v)) for k, v in w.items() if k.startswith("fc"))))
    if key not in _head_cache:
        if w is None:
            w = np.load(head, encoding='bytes', allow_pickle=True).item()
            w = {(k.decode() if isinstance(k, bytes) else k): v for k, v in w.items()}
        _cache_put(_head_cache, key, _e.FcHeadWeights(w))
    return _head_cache[key]


def match_pair(left_u8, right_u8, checkpoint, ndisp=None, patch=11, detail_time=None, params=None, head=None, mode="exact"):
    """Fused match_single.py:34-55: u8 pair -> (left disparity f32, right raw WTA f32), one C call. `head` (weights dict
    or .npy path with fc1..fc4) switches the matching cost to the MC-CNN-accurate decision head.
    Weight dicts are recognised by content (a digest of the arrays), so a new or an updated dict is always repacked.
    mode="fused" selects the opt-in throughput mode (engine.FUSED: 1e-4 contract instead of the reference's bits)."""
    _e._require_cuda()
    D = int(NDISP if ndisp is None else ndisp)
    nl = patch // 2
    packed = _load_weights(checkpoint, nl)
    il, ir = _e._dev(left_u8, torch.uint8), _e._dev(right_u8, torch.uint8)
    stage = np.zeros(7, np.float32) if detail_time is not None else None
    dl, dr = _e.match_pair(il, ir, packed, D, nl, params=params, stage_ms=stage, head=_load_head(head), mode=mode)
    if detail_time is not None:
        detail_time += (stage / 1000.0).astype(detail_time.dtype)
    return dl.cpu().numpy(), dr.cpu().numpy()
