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
parser.add_argument("--head-weights", type=str, default=None, help="MC-CNN-accurate: .npy dict with fc1..fc4 (fc() naming of "
                    "mc_cnn_brunch.py:95-106), or 'random'; the matching cost is then the fully-connected decision head")


def match_images(left_u8: np.ndarray, right_u8: np.ndarray, weights, ndisp: int = 128, scale: int = 1, head=None) -> np.ndarray:
    """match_single.py:34-55 for in-memory images -> the uint8 map the reference would write."""
    from . import process_functional as pf

    left_disparity, _ = pf.match_pair(left_u8, right_u8, weights, ndisp=ndisp, head=head)
    return (left_disparity.astype('uint8') * scale).astype('uint8')


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
    out = match_images(left, right, weights, args.ndisp, args.scale, head)
    out_dir = './result/{}'.format(args.file)
    os.makedirs(out_dir, exist_ok=True)
    cv2.imwrite(os.path.join(out_dir, 'ld{}.png'.format(args.id)), out)


if __name__ == "__main__":
    main()
