"""Drop-in for the reference's match_single.py / match_single_ui.py (same CLI: -g, -i, -f).

Flow of match_single.py:20-55: read ./eval/left_{id}.png / right_{id}.png as greyscale, standardise,
conv tower, disparity pipeline, write ./result/{file}/ld{id}.png as uint8. The whole of it between
imread and imwrite is one C call (mccnn_match_pair). Extra flags (not in the reference): --weights,
--ndisp, --image-dir, --scale (2 reproduces match_single_ui.py:55).
"""
from __future__ import annotations

import argparse
import os

import numpy as np

parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                                 description="stereo matching based on trained model and post-processing")
parser.add_argument("-g", "--gpu", type=str, default=None, help="gpu id to use, multiple ids should be separated "
                    "by commons(e.g. 0,1,2,3); the first visible device is used")
parser.add_argument("-i", "--id", type=int, default=0, help="image_id")
parser.add_argument("-f", "--file", type=str, default="11_11", help="file to save result")
parser.add_argument("--weights", type=str, default="./check_points_11_11/model_epoch14.npy",
                    help=".npy weight dict (Net.save_weights_dict layout); 'random' = seeded Glorot init")
parser.add_argument("--ndisp", type=int, default=128, help="disparity range (the reference hard-codes 128)")
parser.add_argument("--image-dir", type=str, default="./eval/")
parser.add_argument("--scale", type=int, default=1, help="multiply the uint8 map (match_single_ui.py uses 2)")
parser.add_argument("--mode", type=str, default="exact", choices=["exact", "fused"], help="exact: the reference's arithmetic bit for bit; "
                    "fused: the opt-in throughput mode (1e-4 contract, about twice as fast on large pairs)")
parser.add_argument("--pfm", action="store_true", help="also write the fp32 map as ./result/{file}/ld{id}.pfm")
parser.add_argument("--head-weights", type=str, default=None, help="MC-CNN-accurate: .npy dict with fc1..fc4 (fc() naming of "
                    "mc_cnn_brunch.py:95-106), or 'random'; the matching cost is then the fully-connected decision head")


def output_dtype(ndisp: int, scale: int = 1):
    """The reference writes `astype('uint8')` (match_single.py:55; `* 2` in match.py:90 / match_single_ui.py:55), which wraps as
    soon as a scaled disparity exceeds 255; its range is hard-coded to 128 disparities, so it never does there. With ndisp a
    parameter the map is written as a 16-bit PNG whenever (ndisp - 1) * scale > 255: same integer values (truncation toward
    zero, then the scale), no wrap. error_calculate reads either depth."""
    return np.uint8 if (int(ndisp) - 1) * int(scale) <= 255 else np.uint16


def encode_disparity(disparity_f32: np.ndarray, ndisp: int, scale: int = 1) -> np.ndarray:
    dt = output_dtype(ndisp, scale)
    return (disparity_f32.astype(dt) * dt(scale)).astype(dt)


def match_images(left_u8: np.ndarray, right_u8: np.ndarray, weights, ndisp: int = 128, scale: int = 1, head=None, mode="exact") -> np.ndarray:
    """match_single.py:34-55 for in-memory images -> the integer map the reference would write (uint8; uint16 for ranges
    the reference's uint8 cannot hold, see output_dtype)."""
    from . import process_functional as pf

    left_disparity, _ = pf.match_pair(left_u8, right_u8, weights, ndisp=ndisp, head=head, mode=mode)
    return encode_disparity(left_disparity, ndisp, scale)


def main(argv=None):
    args = parser.parse_args(argv)
    if args.gpu is not None:
        os.environ['CUDA_VISIBLE_DEVICES'] = args.gpu
    import cv2

    from . import synthetic

    left_image_path = os.path.join(args.image_dir, 'left_{}.png'.format(args.id))
    right_image_path = os.path.join(args.image_dir, 'right_{}.png'.format(args.id))
    left = cv2.imread(left_image_path, cv2.IMREAD_GRAYSCALE)
    right = cv2.imread(right_image_path, cv2.IMREAD_GRAYSCALE)
    if left is None or right is None:
        raise FileNotFoundError(f"{left_image_path} / {right_image_path}")
    weights = synthetic.glorot_weights() if args.weights == 'random' else args.weights
    head = synthetic.glorot_fc_weights() if args.head_weights == 'random' else args.head_weights
    from . import process_functional as pf

    left_disparity, _ = pf.match_pair(left, right, weights, ndisp=args.ndisp, head=head, mode=args.mode)
    out_dir = './result/{}'.format(args.file)
    os.makedirs(out_dir, exist_ok=True)
    # 8-bit PNG as in the reference; 16-bit where uint8 would wrap (output_dtype)
    cv2.imwrite(os.path.join(out_dir, 'ld{}.png'.format(args.id)), encode_disparity(left_disparity, args.ndisp, args.scale))
    if args.pfm:
        from .error_calculate import save_pfm

        save_pfm(os.path.join(out_dir, 'ld{}.pfm'.format(args.id)), left_disparity)


if __name__ == "__main__":
    main()
