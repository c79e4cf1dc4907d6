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
        self.weights = _syn.glorot_weights(self.num_of_conv_layers, self.num_of_conv_feature_maps, self.conv_kernel_size)
        self._features = None

    @property
    def features(self):
        if self._features is None:
            self._features = self.run(self.input)
        return self._features

    def run(self, inputs):
        """sess.run(features, feed_dict={x: inputs}): [N,h,w,1] -> [N,h-2nl,w-2nl,64] (VALID convolutions)."""
        _e._require_cuda()
        x = np.asarray(inputs, dtype=np.float32)
        if x.ndim != 4 or x.shape[-1] != 1:
            raise ValueError("inputs must be NHWC with one channel")
        nl = self.num_of_conv_layers
        packed = _e.pack_weights(self.weights, nl)
        outs = [_e.conv_tower(_e._dev(x[n, :, :, 0], torch.float32), packed, nl).cpu().numpy() for n in range(x.shape[0])]
        return np.stack(outs, axis=0)

    def load_initial_weights(self, session=None):
        weights_dict = np.load(self.weights_path, encoding='bytes', allow_pickle=True).item()
        for name in weights_dict:
            key = name.decode() if isinstance(name, bytes) else name
            if key not in self.weights:
                raise KeyError(key)
            self.weights[key] = np.asarray(weights_dict[name], dtype=np.float32)
        self._features = None

    def save_weights_dict(self, session=None, file_name='pretrain.npy'):
        np.save(file_name, dict(self.weights))
        print('weights saved in file {}'.format(file_name))
