"""Drop-in for the reference's error_calculate.py: PFM loader + bad-pixel rate.

load_pfm mirrors error_calculate.py:6-45 (bottom-up rows flipped, scale sign = endianness). The
reference's script body (:49-88, hard-coded /home/rjt1 paths, a Python double loop per image) becomes
evaluate(), whose counting runs on the GPU (mccnn_bad_pixels): a pixel is bad iff the ground truth is
finite and non-zero and |disp - gt/2| > 1; the rate divides by ALL H*W pixels (:83).
"""
from __future__ import annotations

import re

import numpy as np
import torch

from . import engine as _e


def load_pfm(fname):
    with open(fname, 'rb') as file:
        header = file.readline().decode().rstrip()
        if header == 'PF':
            channels = 3
        elif header == 'Pf':
            channels = 1
        else:
            raise Exception('Not a PFM file.')
        dim_match = re.match(r'^(\d+)\s(\d+)\s$', file.readline().decode('utf-8'))
        if not dim_match:
            raise Exception('Malformed PFM header')
        width, height = map(int, dim_match.groups())
        scale = float(file.readline().decode().rstrip())
        endian = '<f' if scale < 0 else '>f'
        scale = abs(scale)
        data = np.fromfile(file, endian)
    disparity = np.flipud(np.reshape(data, (height, width, channels)))
    return disparity, scale


def save_pfm(fname, image, scale=1.0):
    image = np.asarray(image, dtype=np.float32)
    if image.ndim == 3 and image.shape[2] == 1:
        image = image[:, :, 0]
    with open(fname, 'wb') as f:
        f.write(b'Pf\n' if image.ndim == 2 else b'PF\n')
        f.write(f'{image.shape[1]} {image.shape[0]}\n'.encode())
        f.write(f'{-abs(scale)}\n'.encode())
        np.flipud(image).astype('<f4').tofile(f)


def error_rate(disp_u8, true_disp):
    """error_calculate.py:63-83 for one image; true_disp is the full-resolution ground truth. The map may be the reference's
    uint8 or the uint16 written for disparity ranges uint8 cannot hold (match_single.output_dtype)."""
    import cv2

    _e._require_cuda()
    disp_u8 = np.asarray(disp_u8)
    if disp_u8.dtype not in (np.uint8, np.uint16):
        raise TypeError(f"disparity map must be uint8 or uint16, got {disp_u8.dtype}")
    height, width = disp_u8.shape[0:2]
    gt = np.asarray(true_disp, dtype=np.float32)
    if gt.ndim == 3:
        gt = gt[:, :, 0]
    gt = cv2.resize(gt, (width, height)) / 2
    dev = _e._dev(disp_u8, torch.uint8) if disp_u8.dtype == np.uint8 else _e._dev(disp_u8.view(np.int16), torch.int16)
    bad, _ = _e.bad_pixels(dev, _e._dev(gt.astype(np.float32), torch.float32))
    return bad / (height * width)


def evaluate(result_paths, true_paths):
    import cv2

    total = 0.0
    for i, (r, t) in enumerate(zip(result_paths, true_paths)):
        disp = cv2.imread(r, cv2.IMREAD_GRAYSCALE | cv2.IMREAD_ANYDEPTH)   # keeps 16-bit maps 16-bit
        true_disp, _ = load_pfm(t)
        rate = error_rate(disp, true_disp)
        total += rate
        print('error rate of disp{}: {}'.format(i, rate))
    print('mean error rate: {}'.format(total / len(result_paths)))
    return total / len(result_paths)
