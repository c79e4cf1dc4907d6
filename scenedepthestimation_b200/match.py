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
parser.add_argument("--mode", type=str, default="exact", choices=["exact", "fused"], help="exact: the reference's arithmetic bit for bit; "
                    "fused: the opt-in throughput mode (1e-4 contract)")
parser.add_argument("--depth", type=int, default=3, help="pairs in flight on separate CUDA streams (1 = the reference's "
                    "strictly sequential loop, with the per-stage times of match.py:95-103 printed at the end)")


def shard(ids, rank: int, world: int):
    """Pairs are independent units: rank r takes ids[r::world]."""
    return list(ids)[rank::world]


def match_batch(pairs, weights, ndisp=128, scale=2, detail_time=None, mode="exact"):
    """pairs: iterable of (left_u8, right_u8) -> list of integer maps (match.py:90 writes uint8*2; uint16 where that would
    wrap, match_single.output_dtype)."""
    from . import process_functional as pf
    from .match_single import encode_disparity

    out = []
    for left, right in pairs:
        dl, _ = pf.match_pair(left, right, weights, ndisp=ndisp, detail_time=detail_time, mode=mode)
        out.append(encode_disparity(dl, ndisp, scale))
    return out


class StreamedMatcher:
    """match.py's loop (:46-90) as a pipeline: `depth` pairs in flight on their own CUDA streams, each with pinned
    host buffers and its own workspace, so the H2D copy, the kernels and the D2H copy of consecutive pairs overlap
    (and the latency-bound small-D kernels of two pairs share the SMs). Results come back in submission order."""

    def __init__(self, H, W, weights, ndisp=128, scale=2, depth=2, num_layers=5, mode="exact"):
        import torch

        from . import engine as eng
        from . import process_functional as pf
        from .match_single import output_dtype

        eng._require_cuda()
        self.torch, self.eng = torch, eng
        self.H, self.W, self.D, self.scale, self.nl = H, W, int(ndisp), int(scale), num_layers
        self.mode = eng._mode(mode)   # "exact" (the reference's bits, default) or "fused" (opt-in throughput mode)
        self.wide = output_dtype(ndisp, scale) is np.uint16   # 16-bit maps where the reference's uint8 would wrap
        odt = torch.int16 if self.wide else torch.uint8       # int16 storage, reinterpreted as uint16 on the host
        self.packed = pf._load_weights(weights, num_layers)
        nws = eng.match_workspace_bytes(H, W, self.D, num_layers)
        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(
                stream=torch.cuda.Stream(), done=torch.cuda.Event(), busy=False, tag=None,
                h_in=torch.empty((2, H, W), dtype=torch.uint8).pin_memory(),
                h_out=torch.empty((H, W), dtype=odt).pin_memory(),
                d_in=torch.empty((2, H, W), dtype=torch.uint8, device="cuda"),
                d_out=torch.empty((H, W), dtype=odt, device="cuda"),
                disp=(torch.empty((H, W), dtype=torch.float32, device="cuda"), torch.empty((H, W), dtype=torch.float32, device="cuda")),
                ws=torch.empty(nws, dtype=torch.uint8, device="cuda")))
        self.next = 0

    def _collect(self, slot):
        slot["done"].synchronize()
        slot["busy"] = False
        img = slot["h_out"].numpy().copy()
        return slot["tag"], (img.view(np.uint16) if self.wide else img)

    def submit(self, left_u8, right_u8, tag=None):
        """Enqueue one pair; returns the (tag, integer map) of the pair that previously used the slot, or None."""
        torch, eng = self.torch, self.eng
        slot = self.slots[self.next]
        self.next = (self.next + 1) % len(self.slots)
        out = self._collect(slot) if slot["busy"] else None
        slot["h_in"][0].copy_(torch.from_numpy(np.ascontiguousarray(left_u8)))
        slot["h_in"][1].copy_(torch.from_numpy(np.ascontiguousarray(right_u8)))
        with torch.cuda.stream(slot["stream"]):
            slot["d_in"].copy_(slot["h_in"], non_blocking=True)
            eng.match_pair(slot["d_in"][0], slot["d_in"][1], self.packed, self.D, self.nl, out=slot["disp"], workspace=slot["ws"],
                           mode=self.mode)
            lib = eng._lib.load()
            if self.wide:   # trunc(d) * scale in 16 bits: mccnn_encode_u16 takes the scale as a power of two or the map is scaled after
                eng._lib.check(lib.mccnn_encode_u16(slot["disp"][0].data_ptr(), slot["d_out"].data_ptr(), self.H, self.W, 0,
                                                    slot["stream"].cuda_stream), "mccnn_encode_u16")
                if self.scale != 1:
                    slot["d_out"].mul_(self.scale)
            else:
                eng._lib.check(lib.mccnn_encode_u8(slot["disp"][0].data_ptr(), slot["d_out"].data_ptr(), self.H, self.W, self.scale,
                                                   slot["stream"].cuda_stream), "mccnn_encode_u8")
            slot["h_out"].copy_(slot["d_out"], non_blocking=True)
            slot["done"].record(slot["stream"])
        slot["busy"], slot["tag"] = True, tag
        return out

    def drain(self):
        """Results still in flight, oldest first."""
        out = []
        for k in range(len(self.slots)):
            slot = self.slots[(self.next + k) % len(self.slots)]
            if slot["busy"]:
                out.append(self._collect(slot))
        return out


def match_stream(pairs, weights, ndisp=128, scale=2, depth=2, mode="exact"):
    """Generator over (left_u8, right_u8) pairs of one shape -> uint8 maps in order, `depth` pairs in flight."""
    m = None
    for i, (left, right) in enumerate(pairs):
        if m is None:
            m = StreamedMatcher(left.shape[0], left.shape[1], weights, ndisp, scale, depth, mode=mode)
        r = m.submit(left, right, i)
        if r is not None:
            yield r[1]
    if m is not None:
        for _, img in m.drain():
            yield img


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
    ids = shard(range(args.first, args.last + 1), rank, world)

    def read(i):
        left = cv2.imread(os.path.join(args.image_dir, 'left_{}.jpg'.format(i)), cv2.IMREAD_GRAYSCALE)
        right = cv2.imread(os.path.join(args.image_dir, 'right_{}.jpg'.format(i)), cv2.IMREAD_GRAYSCALE)
        if left is None or right is None:
            raise FileNotFoundError(f"pair {i} under {args.image_dir}")
        return left, right

    def write(i, out):
        cv2.imwrite(os.path.join(args.out_dir, 'ld{}.png'.format(i)), out)

    if args.depth <= 1:
        for i in ids:
            write(i, match_batch([read(i)], weights, args.ndisp, 2, detail_time, mode=args.mode)[0])
        names = ["features", "cost volume", '"*" cost aggregation', "SGM", "WTA & Subpixel refinement", "LR Check", "Filtering"]
        for n, t in zip(names, detail_time):
            print('time of {}: {}s'.format(n, t))
        return
    # streamed: imread of pair k+1, the copies and the kernels of the pairs in flight, and imwrite of pair k-1 overlap;
    # a new matcher (streams, pinned buffers, workspaces) whenever the image shape changes
    m, shape = None, None
    for i in ids:
        left, right = read(i)
        if left.shape != shape:
            if m is not None:
                for tag, img in m.drain():
                    write(tag, img)
            m, shape = StreamedMatcher(left.shape[0], left.shape[1], weights, args.ndisp, 2, args.depth, mode=args.mode), left.shape
        done = m.submit(left, right, i)
        if done is not None:
            write(*done)
    if m is not None:
        for tag, img in m.drain():
            write(tag, img)


if __name__ == "__main__":
    main()
