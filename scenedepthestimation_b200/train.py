"""Drop-in for the compute of the reference's train.py: one training step of the siamese tower on the GPU.

train.py:71-99 builds three weight-sharing Net branches over [B,p,p,1] patch placeholders, the cosine hinge loss and a
tf.train.MomentumOptimizer, and runs them once per batch (:143-150). Here the whole step (forward, loss, backward, momentum
update) is one C call, mccnn_train_step; `Trainer` holds the flat parameter / accumulator vectors on the device and converts
from and to the reference's weight dict ({'conv{i}/weights:0': HWIO, 'conv{i}/biases:0'}, Net.save_weights_dict). The data
generator, TensorBoard summaries and checkpointing of train.py are outside the hot path and are not rebuilt; `main` runs
the loop on seeded synthetic patches.
"""
from __future__ import annotations

import argparse

import numpy as np
import torch

from . import _lib
from . import engine as _e
from . import synthetic as _syn


class Trainer:
    def __init__(self, weights: dict | None = None, num_layers: int = 5, margin: float = 0.3, learning_rate: float = 0.001,
                 beta: float = 0.9):
        _e._require_cuda()
        self.lib = _lib.load()
        self.num_layers, self.patch = num_layers, 2 * num_layers + 1
        self.margin, self.lr, self.beta = float(margin), float(learning_rate), float(beta)
        self.n = self.lib.mccnn_train_param_count(num_layers)
        self.params = torch.zeros(self.n, dtype=torch.float32, device="cuda")
        self.velocity = torch.zeros_like(self.params)
        self.grads = torch.zeros_like(self.params)
        self.loss = torch.zeros(1, dtype=torch.float32, device="cuda")
        self._ws = None
        self.set_weights(weights if weights is not None else _syn.glorot_weights(num_layers))

    # ---- flat vector <-> the reference's dict
    def _layout(self):
        off, cin = 0, 1
        for i in range(1, self.num_layers + 1):
            nw = 9 * cin * _e.FEATURES
            yield f"conv{i}/weights:0", off, (3, 3, cin, _e.FEATURES)
            yield f"conv{i}/biases:0", off + nw, (_e.FEATURES,)
            off += nw + _e.FEATURES
            cin = _e.FEATURES

    def set_weights(self, weights: dict):
        flat = np.zeros(self.n, np.float32)
        for name, off, shape in self._layout():
            key = name if name in weights else name[:-2]
            flat[off:off + int(np.prod(shape))] = np.asarray(weights[key], np.float32).reshape(-1)
        self.params.copy_(torch.from_numpy(flat))

    def _to_dict(self, vec: torch.Tensor) -> dict:
        flat = vec.cpu().numpy()
        return {name: flat[off:off + int(np.prod(shape))].reshape(shape).copy() for name, off, shape in self._layout()}

    def weights_dict(self) -> dict:
        return self._to_dict(self.params)

    def grads_dict(self) -> dict:
        return self._to_dict(self.grads)

    def save_weights_dict(self, file_name='pretrain.npy'):
        np.save(file_name, self.weights_dict())

    # ---- one step (train.py:143-150: sess.run([train, loss], feed_dict={leftx, rightx_pos, rightx_neg, factor}))
    def step(self, left, right_pos, right_neg, factor: float = 1.0, update: bool = True) -> float:
        dev = lambda a: _e._dev(np.asarray(a, np.float32).reshape(-1, self.patch, self.patch), torch.float32)
        l, p, n = dev(left), dev(right_pos), dev(right_neg)
        B = l.shape[0]
        if p.shape != l.shape or n.shape != l.shape:
            raise ValueError("the three patch batches must have one shape [B, p, p]")
        nws = self.lib.mccnn_train_workspace_bytes(B, self.patch, self.num_layers)
        if self._ws is None or self._ws.numel() < nws:
            self._ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
        _lib.check(self.lib.mccnn_train_step(l.data_ptr(), p.data_ptr(), n.data_ptr(), self.params.data_ptr(), self.velocity.data_ptr(),
                                             self.grads.data_ptr(), self.loss.data_ptr(), self._ws.data_ptr(), self._ws.numel(), B,
                                             self.patch, self.num_layers, self.margin, self.lr / float(factor), self.beta,
                                             1 if update else 0, torch.cuda.current_stream().cuda_stream), "mccnn_train_step")
        return float(self.loss.item())


def synthetic_patches(batch: int, patch: int = 11, seed: int = 0):
    """Seeded stand-in for the reference's patch generator: positives are noisy copies of the left patch, negatives are
    unrelated patches (standardised like match_single.py:40-43)."""
    rng = np.random.default_rng(seed)
    left = rng.standard_normal((batch, patch, patch)).astype(np.float32)
    pos = (left + 0.2 * rng.standard_normal(left.shape)).astype(np.float32)
    neg = rng.standard_normal(left.shape).astype(np.float32)
    return left, pos, neg


def main(argv=None):
    ap = argparse.ArgumentParser(description="training of the siamese tower (compute of the reference's train.py) on synthetic patches")
    ap.add_argument("-ps", "--patch_size", type=int, default=11)
    ap.add_argument("-bs", "--batch_size", type=int, default=128)
    ap.add_argument("-mr", "--margin", type=float, default=0.3)
    ap.add_argument("-lr", "--learning_rate", type=float, default=0.001)
    ap.add_argument("-bt", "--beta", type=float, default=0.9)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--save", type=str, default=None, help=".npy weight dict to write (Net.save_weights_dict layout)")
    a = ap.parse_args(argv)
    tr = Trainer(None, a.patch_size // 2, a.margin, a.learning_rate, a.beta)
    for it in range(a.steps):
        loss = tr.step(*synthetic_patches(a.batch_size, a.patch_size, it))
        if it % 20 == 0 or it == a.steps - 1:
            print(f"step {it}: hinge loss {loss:.5f}")
    if a.save:
        tr.save_weights_dict(a.save)


if __name__ == "__main__":
    main()
