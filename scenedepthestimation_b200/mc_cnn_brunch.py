"""Drop-in for the reference's mc_cnn_brunch.py: the siamese branch ("brunch") network.

Same constructor signature and attribute names as Net (mc_cnn_brunch.py:4-48). The reference builds
a TF1 graph over a placeholder and evaluates `.features` with sess.run; here `inputs` is an NHWC
f32 array and `.features` is computed on first access by the CUDA conv tower. Weights use the
reference's dict layout {'conv{i}/weights:0': [3,3,Cin,64], 'conv{i}/biases:0': [64]} (:61-66, :76-77).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import engine as _e
from . import synthetic as _syn


# The reference's variables live in the TF graph under "conv{i}/weights:0" / "conv{i}/biases:0"; a Net built with
# is_branch=True re-uses them (scope.reuse_variables(), mc_cnn_brunch.py:73-75). Here the "graph" is this module-level
# store: the first Net of a configuration creates the dict, branches alias the SAME dict object, so load_initial_weights /
# save_weights_dict / training updates on one branch are seen by its siamese twins.
_VARIABLE_STORE: dict = {}


def reset_default_graph():
    """tf.reset_default_graph(): forget the shared variables (a new first Net draws a fresh Glorot init)."""
    _VARIABLE_STORE.clear()


class Net:
    def __init__(self, inputs, weights_path='DEFAULT', input_patch_size=11, num_of_conv_layers=4,
                 num_of_conv_feature_maps=64, conv_kernel_size=3, batch_size=128, is_branch=False):
        self.input = inputs
        self.input_patch_size = input_patch_size
        self.num_of_conv_layers = num_of_conv_layers
        self.num_of_conv_feature_maps = num_of_conv_feature_maps
        self.conv_kernel_size = conv_kernel_size
        self.batch_size = batch_size
        self.is_branch = is_branch
        self.weights_path = 'pretrain.npy' if weights_path == 'DEFAULT' else weights_path
        if conv_kernel_size != 3 or num_of_conv_feature_maps != _e.FEATURES:
            raise ValueError("only 3x3 kernels and 64 feature maps are supported (the reference's configuration)")
        self.construct()

    def construct(self):
        """Variables are created at graph construction in the reference (tf.get_variable, Glorot-uniform)."""
        key = (self.num_of_conv_layers, self.num_of_conv_feature_maps, self.conv_kernel_size)
        if self.is_branch:
            if key not in _VARIABLE_STORE:
                raise ValueError("is_branch=True re-uses the variables of an earlier Net (scope.reuse_variables(), "
                                 "mc_cnn_brunch.py:73-75), but none has been built with this configuration")
        else:
            _VARIABLE_STORE[key] = _syn.glorot_weights(*key)
        self.weights = _VARIABLE_STORE[key]   # the same dict object for every branch
        self._features = None

    @property
    def features(self):
        from .process_functional import _content_key

        digest = _content_key(sorted(self.weights.items()))   # shared weights may have been replaced through a twin
        if self._features is None or self._features[0] != digest:
            self._features = (digest, self.run(self.input))
        return self._features[1]

    def run(self, inputs):
        """sess.run(features, feed_dict={x: inputs}): [N,h,w,1] -> [N,h-2nl,w-2nl,64] (VALID convolutions)."""
        _e._require_cuda()
        x = np.asarray(inputs, dtype=np.float32)
        if x.ndim != 4 or x.shape[-1] != 1:
            raise ValueError("inputs must be NHWC with one channel")
        nl = self.num_of_conv_layers
        from .process_functional import _load_weights

        packed = _load_weights(self.weights, nl)   # content-keyed cache: repacked only when the shared weights change
        N, h, w = x.shape[0:3]
        if N > 1 and h * w <= 64 * 64:
            # a batch of small patches (train.py feeds 128 x 11 x 11): ONE launch over the patches laid side by side. VALID
            # convolutions: a patch's own output windows lie inside the patch; an output column whose window straddles two
            # patches belongs to neither and is dropped, so no gutter is needed.
            strip = np.ascontiguousarray(x[:, :, :, 0].transpose(1, 0, 2).reshape(h, N * w))
            f = _e.conv_tower(_e._dev(strip, torch.float32), packed, nl).cpu().numpy()
            ow = w - 2 * nl
            return np.stack([f[:, n * w:n * w + ow] for n in range(N)], axis=0)
        outs = [_e.conv_tower(_e._dev(x[n, :, :, 0], torch.float32), packed, nl).cpu().numpy() for n in range(N)]
        return np.stack(outs, axis=0)

    def load_initial_weights(self, session=None):
        weights_dict = np.load(self.weights_path, encoding='bytes', allow_pickle=True).item()
        for name in weights_dict:
            key = name.decode() if isinstance(name, bytes) else name
            if key not in self.weights:
                raise KeyError(key)
            self.weights[key] = np.asarray(weights_dict[name], dtype=np.float32)   # in place: every branch sees it
        self._features = None

    def save_weights_dict(self, session=None, file_name='pretrain.npy'):
        np.save(file_name, dict(self.weights))
        print('weights saved in file {}'.format(file_name))
