"""The reference's OWN files (imported from /root/reference, never copied) against the oracle, on the CPU:

  * mc_cnn_brunch.Net (:4-48, conv() :70-92) built over tests/tf_shim.py (a torch-CPU stand-in for the TF-1.x calls the file
    makes): layer count, kernel shapes, variable names, ReLU placement, is_branch weight sharing and the normalisation axis come
    from the reference's code; the result must equal oracle.conv_tower on the same weights;
  * process_functional.compute_feature (:11-45) end to end through the same shim (padding, squeeze, the two sess.run calls);
  * the reference's NumPy CPU path compute_cost_volume / WTA / WTA1 (:48-113) against the oracle's restatements.
TensorFlow's own arithmetic stays unpinned (it is not installed); what is pinned here is the wiring.
Skipped where /root/reference does not exist (the GPU box)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "mc_cnn_brunch.py")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import tf_shim

    tf = tf_shim.install()
    import numba.cuda as nbcuda

    nbcuda.select_device = lambda idx: None
    sys.path.insert(0, REF)
    try:
        for m in ("mc_cnn_brunch", "process_functional"):
            sys.modules.pop(m, None)
        import mc_cnn_brunch as ref_net
        import process_functional as ref_pf
    finally:
        sys.path.remove(REF)
    assert os.path.realpath(ref_net.__file__).startswith(REF) and os.path.realpath(ref_pf.__file__).startswith(REF)
    return tf, ref_net, ref_pf


def test_reference_net_wiring_equals_oracle_tower(ref):
    from oracle import conv_tower as ct
    from scenedepthestimation_b200 import synthetic as syn

    tf, ref_net, _ = ref
    tf.reset_default_graph()
    rng = np.random.default_rng(5)
    data = rng.standard_normal((2, 15, 19, 1)).astype(np.float32)
    x = tf.placeholder(tf.float32, [2, 15, 19, 1])
    net = ref_net.Net(x, input_patch_size=11, num_of_conv_layers=5, batch_size=2)
    names = sorted(v.name for v in tf.trainable_variables())
    weights = syn.glorot_weights(5)
    assert names == sorted(weights), "the reference's variable names are the keys of the .npy dict layout"
    for v in tf.trainable_variables():
        assert list(weights[v.name].shape) == v.shape, v.name
    # a second branch shares the first one's variables (scope.reuse_variables(), :73-75)
    twin = ref_net.Net(tf.placeholder(tf.float32, [2, 15, 19, 1]), num_of_conv_layers=5, batch_size=2, is_branch=True)
    assert len(tf.trainable_variables()) == 10 and twin.features is not net.features
    with tf.Session() as sess:
        for v in tf.trainable_variables():
            sess.run(v.assign(weights[v.name]))
        got = sess.run(net.features, feed_dict={x: data})
    assert got.shape == (2, 5, 9, 64)
    for n in range(2):
        exp = ct.conv_tower(data[n:n + 1], weights, 5)
        np.testing.assert_allclose(got[n], exp, rtol=0, atol=1e-6)
    # ReLU placement: layers 1..nl-1 are rectified, the last is linear (:37-46)
    with tf.Session() as sess:
        c4, c5 = sess.run([net.conv4, net.conv5], feed_dict={x: data})
    assert c4.min() >= 0 and c5.min() < 0


def test_reference_default_net_has_four_layers(ref):
    """Net's own default is num_of_conv_layers=4 (9x9 receptive field); compute_feature passes patch // 2 = 5 (:23)."""
    tf, ref_net, _ = ref
    tf.reset_default_graph()
    net = ref_net.Net(tf.placeholder(tf.float32, [1, 11, 11, 1]), batch_size=1)
    assert len(tf.trainable_variables()) == 8 and net.features.shape == [1, 3, 3, 64]


def test_reference_compute_feature_equals_oracle(ref, tmp_path):
    from oracle import conv_tower as ct
    from scenedepthestimation_b200 import synthetic as syn

    tf, _, ref_pf = ref
    tf.reset_default_graph()
    il, ir, _ = syn.textured_pair(21, 33, 16, 9)
    left, right = syn.standardise(il), syn.standardise(ir)
    weights = syn.glorot_weights(5)
    ckpt = str(tmp_path / "w.npy")
    np.save(ckpt, weights)
    with contextlib.redirect_stdout(io.StringIO()):
        fl, fr = ref_pf.compute_feature(left, right, 11, 11, 64, ckpt)
    exp_l, exp_r = ct.compute_feature(left, right, 11, 11, 64, weights)
    assert fl.shape == (21, 33, 64)
    np.testing.assert_allclose(fl, exp_l, rtol=0, atol=1e-6)
    np.testing.assert_allclose(fr, exp_r, rtol=0, atol=1e-6)
    import torch

    twin = ct.compute_feature(left, right, 11, 11, 64, weights, dtype=torch.float64)[0]
    assert np.abs(fl - twin).max() < 2e-6


def test_reference_cpu_path_equals_oracle(ref):
    from oracle import stereo as st
    from scenedepthestimation_b200 import synthetic as syn

    _, _, ref_pf = ref
    fl, fr = syn.unit_features(7, 41, 64, 3)
    with contextlib.redirect_stdout(io.StringIO()):
        vol = ref_pf.compute_cost_volume(fl, fr, 24)
        d1 = ref_pf.WTA1(vol)
        d2 = ref_pf.WTA(np.ascontiguousarray(np.transpose(vol, (1, 2, 0))))
    assert np.array_equal(vol, st.cost_volume_cpu_reference(fl, fr, 24))
    assert np.array_equal(d1, st.wta_dhw(vol)) and np.array_equal(d2, d1)
