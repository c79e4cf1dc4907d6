"""A minimal stand-in for the TensorFlow-1.x API surface the reference's mc_cnn_brunch.py / process_functional.py touch
(TEST INFRASTRUCTURE ONLY; TensorFlow is not installed in this image).

Purpose: let tests import the REFERENCE files themselves (from /root/reference, never copied) so that the tower's
wiring -- layer count, kernel shapes, ReLU placement, variable names, the normalisation axis, the padding done by
compute_feature -- comes from the reference's code and not from a reading of it. The arithmetic of each op is plain
torch CPU (fp32 by default, fp64 on request): TensorFlow's own kernels stay "unpinned".

Covered: placeholder, variable_scope(+reuse_variables), get_variable (Glorot-uniform, like tf's default initializer),
trainable_variables, nn.conv2d / bias_add / relu / l2_normalize / xw_plus_b, identity, reshape, split, concat,
Session / ConfigProto / GPUOptions, train.Saver.restore (reads the .npy dict of Net.save_weights_dict: the only
checkpoint format this repo reads too). Graph mode is imitated with lazily evaluated nodes.
"""
from __future__ import annotations

import contextlib
import sys
import types

import numpy as np
import torch

float32 = np.float32
_DTYPE = torch.float32


def set_dtype(dt):
    global _DTYPE
    _DTYPE = dt


class Node:
    def __init__(self, fn, inputs, shape, name=None):
        self.fn, self.inputs, self.shape, self.name = fn, inputs, list(shape), name

    def get_shape(self):
        shape = self.shape
        return types.SimpleNamespace(as_list=lambda: list(shape))

    def eval(self, feed, cache):
        if id(self) in cache:
            return cache[id(self)]
        if self in feed:
            v = torch.as_tensor(np.asarray(feed[self])).to(_DTYPE)
        else:
            v = self.fn(*[i.eval(feed, cache) for i in self.inputs])
        cache[id(self)] = v
        return v

    __hash__ = object.__hash__


class Variable(Node):
    def __init__(self, name, shape, rng):
        super().__init__(None, [], shape, name)
        # tf.get_variable's default initializer: glorot_uniform, fans by TF's rule (receptive field x channels)
        if len(shape) == 1:
            fan_in = fan_out = shape[0]
        elif len(shape) == 2:
            fan_in, fan_out = shape
        else:
            rf = int(np.prod(shape[:-2]))
            fan_in, fan_out = rf * shape[-2], rf * shape[-1]
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        self.value = rng.uniform(-lim, lim, shape).astype(np.float32)

    def eval(self, feed, cache):
        return torch.from_numpy(self.value).to(_DTYPE)

    def assign(self, value):
        def run():
            value_ = np.asarray(value, dtype=np.float32)
            assert list(value_.shape) == self.shape, (self.name, value_.shape, self.shape)
            self.value = value_
        return _Op(run)


class _Op:
    def __init__(self, fn):
        self.fn = fn


class _Graph:
    def __init__(self):
        self.reset()

    def reset(self, seed=0):
        self.vars = {}
        self.scope = []
        self.reuse = False
        self.rng = np.random.default_rng(seed)


_G = _Graph()


def reset_default_graph(seed=0):
    _G.reset(seed)


def placeholder(dtype, shape, name=None):
    return Node(None, [], shape, name)


class _Scope:
    def __init__(self, name):
        self.name = name

    def reuse_variables(self):
        _G.reuse = True


@contextlib.contextmanager
def variable_scope(name):
    _G.scope.append(name)
    old = _G.reuse
    try:
        yield _Scope("/".join(_G.scope))
    finally:
        _G.scope.pop()
        _G.reuse = old


def get_variable(name, shape=None, trainable=True):
    full = "/".join(_G.scope + [name]) + ":0"
    if full in _G.vars:
        if not _G.reuse:
            raise ValueError(f"Variable {full} already exists, disallowed (reuse not set)")
        return _G.vars[full]
    if _G.reuse:
        raise ValueError(f"Variable {full} does not exist (reuse set)")
    v = Variable(full, list(shape), _G.rng)
    _G.vars[full] = v
    return v


def trainable_variables():
    return list(_G.vars.values())


def identity(x, name=None):
    return Node(lambda a: a, [x], x.shape, name)


def reshape(x, shape, name=None):
    return Node(lambda a: a.reshape(shape), [x], shape, name)


def split(axis, num_or_size_splits, value):
    n = num_or_size_splits
    shp = list(value.shape)
    shp[axis] //= n
    return [Node(lambda a, k=k: torch.chunk(a, n, dim=axis)[k], [value], shp) for k in range(n)]


def concat(axis, values):
    shp = list(values[0].shape)
    shp[axis] = sum(v.shape[axis] for v in values)
    return Node(lambda *a: torch.cat(a, dim=axis), list(values), shp)


def _conv2d(i, k, strides, padding):
    assert padding in ("VALID", "SAME") and strides[0] == strides[3] == 1
    n, h, w, _ = i.shape
    kh, kw, _, co = k.shape
    if padding == "VALID":
        oh, ow = (h - kh) // strides[1] + 1, (w - kw) // strides[2] + 1
        pad = 0
    else:
        assert strides[1] == strides[2] == 1 and kh % 2 == 1 and kw % 2 == 1
        oh, ow = h, w
        pad = (kh // 2, kw // 2)

    def run(a, wt):  # NHWC x HWIO
        y = torch.nn.functional.conv2d(a.permute(0, 3, 1, 2), wt.permute(3, 2, 0, 1).contiguous(), None,
                                       stride=(strides[1], strides[2]), padding=pad)
        return y.permute(0, 2, 3, 1)

    return Node(run, [i, k], [n, oh, ow, co])


def _l2_normalize(x, dim=-1, axis=None, epsilon=1e-12, name=None):
    ax = dim if axis is None else axis
    return Node(lambda a: a * torch.rsqrt(torch.clamp((a * a).sum(dim=ax, keepdim=True), min=epsilon)), [x], x.shape, name)


nn = types.SimpleNamespace(
    conv2d=_conv2d,
    bias_add=lambda x, b: Node(lambda a, bb: a + bb, [x, b], x.shape),
    relu=lambda x, name=None: Node(torch.relu, [x], x.shape, name),
    l2_normalize=_l2_normalize,
    xw_plus_b=lambda x, w, b, name=None: Node(lambda a, ww, bb: a @ ww + bb, [x, w, b], [x.shape[0], w.shape[1]]),
)


class GPUOptions:
    def __init__(self, **kw):
        pass


class ConfigProto:
    def __init__(self, **kw):
        pass


class Session:
    def __init__(self, config=None):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        if isinstance(fetches, _Op):
            return fetches.fn()
        if isinstance(fetches, (list, tuple)):
            return [self.run(f, feed_dict) for f in fetches]
        with torch.no_grad():
            return fetches.eval(feed_dict or {}, {}).to(torch.float32).numpy()


class _Saver:
    """restore() takes the .npy dict written by Net.save_weights_dict (mc_cnn_brunch.py:61-66)."""

    def __init__(self, max_to_keep=None):
        pass

    def restore(self, sess, checkpoint):
        d = np.load(checkpoint, encoding="bytes", allow_pickle=True).item()
        for name, value in d.items():
            key = name.decode() if isinstance(name, bytes) else name
            sess.run(_G.vars[key].assign(value))


train = types.SimpleNamespace(Saver=_Saver)


def install():
    """Register this module as `tensorflow` (only if the real one is absent)."""
    sys.modules["tensorflow"] = sys.modules[__name__]
    return sys.modules[__name__]
