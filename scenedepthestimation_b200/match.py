"""Drop-in for the reference's match.py (same CLI: -g): the batch loop over ./test/left_{i}.jpg, i = 1..18.

match.py:46-90 processes the pairs strictly one after another (imread -> standardise -> pad ->
2x sess.run -> disparity_compute_by_gpu -> imwrite(uint8*2)). Here the weights are packed once, the
workspace is allocated once, and each pair is one mccnn_match_pair call; with several visible GPUs
and torch.distributed initialised, pairs are dealt round-robin to the ranks (one pair per rank at a
time, no data-path collective).
"""
from __future__ import annotations

import argparse
import os

import numpy as np

parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                                 description="stereo matching based on trained model and post-processing")
parser.add_argument("-g", "--gpu", type=str, default=None, help="gpu id to use, multiple ids should be separated "
                    "by commons(e.g. 0,1,2,3)")
parser.add_argument("--weights", type=str, default="./check_points_11_11/model_epoch14.npy")
parser.add_argument("--ndisp", type=int, default=128)
parser.add_argument("--image-dir", type=str, default="./test/")
parser.add_argument("--out-dir", type=str, default="./disparity/")
parser.add_argument("--first", type=int, default=1)
parser.add_argument("--last", type=int, default=18)


def shard(ids, rank: int, world: int):
    """Pairs are independent units: rank r takes ids[r::world]."""
    return list(ids)[rank::world]


def match_batch(pairs, weights, ndisp=128, scale=2, detail_time=None):
    """pairs: iterable of (left_u8, right_u8) -> list of uint8 maps (match.py:90 writes uint8*2)."""
    from . import process_functional as pf

    out = []
    for left, right in pairs:
        dl, _ = pf.match_pair(left, right, weights, ndisp=ndisp, detail_time=detail_time)
        out.append((dl.astype('uint8') * scale).astype('uint8'))
    return out


def main(argv=None):
    args = parser.parse_args(argv)
    if args.gpu is not None:
        os.environ['CUDA_VISIBLE_DEVICES'] = args.gpu
    import cv2
    import torch

    from . import synthetic

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)) % max(1, torch.cuda.device_count()))
    weights = synthetic.glorot_weights() if args.weights == 'random' else args.weights
    detail_time = np.zeros(shape=[7], dtype=np.float32)
    os.makedirs(args.out_dir, exist_ok=True)
    for i in shard(range(args.first, args.last + 1), rank, world):
        left = cv2.imread(os.path.join(args.image_dir, 'left_{}.jpg'.format(i)), cv2.IMREAD_GRAYSCALE)
        right = cv2.imread(os.path.join(args.image_dir, 'right_{}.jpg'.format(i)), cv2.IMREAD_GRAYSCALE)
        if left is None or right is None:
            raise FileNotFoundError(f"pair {i} under {args.image_dir}")
        out = match_batch([(left, right)], weights, args.ndisp, 2, detail_time)[0]
        cv2.imwrite(os.path.join(args.out_dir, 'ld{}.png'.format(i)), out)
    names = ["features", "cost volume", '"*" cost aggregation', "SGM", "WTA & Subpixel refinement", "LR Check", "Filtering"]
    for n, t in zip(names, detail_time):
        print('time of {}: {}s'.format(n, t))


if __name__ == "__main__":
    main()
