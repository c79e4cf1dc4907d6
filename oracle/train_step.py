"""TEST INFRASTRUCTURE ONLY (tests/ may import this; the product must not).

CPU restatement of one training step of the reference (train.py:71-99): three weight-sharing branches of Net (mc_cnn_brunch.py:
31-48: 3x3 VALID conv + bias, ReLU on all but the last layer, tf.nn.l2_normalize with epsilon 1e-12), cosine similarities,
hinge loss mean(max(0, margin - cos_pos + cos_neg)) (:83-89), tf.train.MomentumOptimizer (accum = beta * accum + grad;
var -= lr * accum, :97-99). torch autograd in fp64 stands in for TensorFlow's graph (not installed; its arithmetic is not
pinned by any test of the reference: parity is a tolerance, stated in the tests).
"""
import numpy as np
import torch


def _branch(x, ws, bs):
    for i, (w, b) in enumerate(zip(ws, bs)):
        x = torch.nn.functional.conv2d(x, w.permute(3, 2, 0, 1)) + b.view(1, -1, 1, 1)  # HWIO -> OIHW, VALID
        if i + 1 < len(ws):
            x = torch.relu(x)
    x = x[:, :, 0, 0]  # tf.squeeze(features, [1, 2])
    return x * torch.rsqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=1e-12))


def loss_and_grads(weights: dict, left, right_pos, right_neg, margin=0.3, num_layers=5, dtype=torch.float64):
    """-> (loss, {name: grad}) for [B,p,p] patches and the reference's weight dict."""
    ws = [torch.tensor(np.asarray(weights[f"conv{i}/weights:0"]), dtype=dtype, requires_grad=True) for i in range(1, num_layers + 1)]
    bs = [torch.tensor(np.asarray(weights[f"conv{i}/biases:0"]), dtype=dtype, requires_grad=True) for i in range(1, num_layers + 1)]
    t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)[:, None]
    fl, fp, fn = (_branch(t(a), ws, bs) for a in (left, right_pos, right_neg))
    loss = torch.clamp(margin - (fl * fp).sum(-1) + (fl * fn).sum(-1), min=0).mean()
    loss.backward()
    grads = {}
    for i in range(num_layers):
        grads[f"conv{i + 1}/weights:0"] = ws[i].grad.numpy()
        grads[f"conv{i + 1}/biases:0"] = bs[i].grad.numpy()
    return float(loss.detach()), grads


def momentum_update(weights: dict, velocity: dict, grads: dict, lr: float, beta: float):
    """tf.train.MomentumOptimizer (use_nesterov=False)."""
    new_w, new_v = {}, {}
    for k in weights:
        v = beta * np.asarray(velocity[k], np.float64) + grads[k]
        new_v[k] = v
        new_w[k] = np.asarray(weights[k], np.float64) - lr * v
    return new_w, new_v
