"""CPU-side checks: the C-ABI library loads and exports every symbol include/mccnn_b200.h declares,
argument validation answers without touching a GPU, and the host-side mirror behaves like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mccnn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mccnn_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from scenedepthestimation_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from scenedepthestimation_b200 import _lib

    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mccnn_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names


def test_constants_and_sizes(lib):
    from scenedepthestimation_b200 import _lib

    assert lib.mccnn_abi_version() == 4
    assert [lib.mccnn_disp_pitch(d) for d in (1, 4, 80, 81, 228, 800)] == [4, 4, 80, 84, 228, 800]
    p = _lib.default_sgm_params()
    # process_functional.py:1141-1144, stored as fp32 (:149), reduced pair computed in fp64 then rounded (:141-142)
    assert p.P1 == np.float32(2.3) and p.P2 == np.float32(55.9)
    assert p.P1_red == np.float32(2.3 / 4) and p.P2_red == np.float32(55.9 / 4) and p.threshold == 30
    assert (p.subpixel, p.bilateral, p.cbca_iters) == (0, 0, 0)  # stages the reference does not run stay off
    assert (p.cbca_L1, p.cbca_tau) == (14, 6)
    fp32_bytes = 4 * (9 * 64 + 64 + 4 * (9 * 64 * 64 + 64))
    # fp32 section (padded to 1 KB) + per 64->64 layer the fp16 hi/lo tensor-core tiles (2 x 9 taps x 8 KB)
    assert lib.mccnn_conv_packed_weight_bytes(5) == ((fp32_bytes + 1023) // 1024) * 1024 + 4 * 2 * 9 * 8192
    w1 = lib.mccnn_pipeline_workspace_bytes(370, 463, 80)
    w4 = lib.mccnn_pipeline_workspace_bytes(1988, 2880, 800)
    assert 4 * 370 * 463 * 80 * 4 <= w1 < w4 < 80e9  # four volumes + maps; c4 fits one 180 GB B200
    assert lib.mccnn_match_workspace_bytes(1988, 2880, 800, 5) > w4


def test_argument_errors_without_gpu(lib):
    from scenedepthestimation_b200 import _lib

    p = _lib.default_sgm_params()
    assert lib.mccnn_cost_volume(None, None, None, None, 4, 4, 4, 1.0, None) == -1
    assert b"null" in lib.mccnn_last_error()
    assert lib.mccnn_sgm(1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 4096,
                         2, 2, 8, C.byref(p), 0, 0, None) == -1
    assert b"too small" in lib.mccnn_last_error()
    big = 1 << 40   # (one workspace size serves the exact and the fused mode: tens of MB)
    assert lib.mccnn_sgm_workspace_bytes(8, 8, 16) > 4096 and lib.mccnn_sgm_workspace_bytes(0, 8, 16) == 0
    assert lib.mccnn_sgm(1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, big,
                         8, 8, 2000, C.byref(p), 0, 0, None) == -1
    assert lib.mccnn_sgm(1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, big,
                         8, 8, 16, C.byref(p), 7, 0, None) == -1
    assert b"mode" in lib.mccnn_last_error()
    neg = _lib.default_sgm_params()
    neg.P2_red = -1.0   # the fused mode starts a path from the zero state, which needs penalties >= 0
    assert lib.mccnn_sgm(1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, big,
                         8, 8, 16, C.byref(neg), 1, 0, None) == -1
    assert b"penalties" in lib.mccnn_last_error()
    assert lib.mccnn_sgm(1 << 20, (1 << 20) + 4, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, big,
                         8, 8, 16, C.byref(p), 0, 0, None) == -2
    assert lib.mccnn_disparity_pipeline(1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 1 << 20, 16,
                                        8, 8, 16, C.byref(p), 0, None, None) == -3


def test_pack_weights_layout(lib):
    from scenedepthestimation_b200 import synthetic as syn

    w = syn.glorot_weights()
    ws = [np.ascontiguousarray(w[f"conv{i}/weights:0"]) for i in range(1, 6)]
    bs = [np.ascontiguousarray(w[f"conv{i}/biases:0"]) for i in range(1, 6)]
    out = np.zeros(lib.mccnn_conv_packed_weight_bytes(5) // 4, np.float32)
    blob = out.view(np.uint8)
    wp = (C.c_void_p * 5)(*[a.ctypes.data for a in ws])
    bp = (C.c_void_p * 5)(*[a.ctypes.data for a in bs])
    assert lib.mccnn_pack_weights_host(wp, bp, 5, out.ctypes.data) == 0
    assert np.array_equal(out[:576], ws[0].ravel()) and np.array_equal(out[576:640], bs[0])
    assert np.array_equal(out[640:640 + 36864], ws[1].ravel())
    # tensor-core section: w = hi + lo / 2048 in fp16, [cout][cin] K-major 128-byte rows, 16-byte chunks XOR (row % 8)
    fp32_bytes = 4 * (9 * 64 + 64 + 4 * (9 * 64 * 64 + 64))
    tc0 = ((fp32_bytes + 1023) // 1024) * 1024
    tap, n, k = 5, 13, 42
    off = tc0 + tap * 8192 + (n // 8) * 1024 + (n % 8) * 128 + (((k // 8) ^ (n % 8)) * 16) + (k % 8) * 2
    w = ws[1][tap // 3, tap % 3, k, n]
    hi = blob[off:off + 2].view(np.float16)[0]
    lo = blob[off + 9 * 8192:off + 9 * 8192 + 2].view(np.float16)[0]
    assert hi == np.float16(w) and abs(float(hi) + float(lo) / 2048 - float(w)) <= abs(float(w)) * 2.0 ** -21


def test_no_cpu_fallback(monkeypatch):
    """Without a CUDA device the reference-shaped API raises instead of computing on the CPU."""
    import torch

    from scenedepthestimation_b200 import process_functional as pf

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pf.compute_cost_volume(np.zeros((4, 4, 64), np.float32), np.zeros((4, 4, 64), np.float32), 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pf.disparity_compute_by_gpu(np.zeros((4, 4), np.uint8), np.zeros((4, 4), np.uint8),
                                    np.zeros((4, 4, 64), np.float32), np.zeros((4, 4, 64), np.float32), np.zeros(7, np.float32))


def test_unknown_checkpoint_is_refused_loudly():
    from scenedepthestimation_b200 import process_functional as pf

    with pytest.raises(RuntimeError, match="neither a .npy weight dict"):
        pf._load_weights("./check_points_11_11/model_epoch14.ckpt", 5)


def test_pfm_roundtrip(tmp_path):
    from scenedepthestimation_b200 import error_calculate as ec

    a = np.random.default_rng(0).random((5, 7)).astype(np.float32)
    ec.save_pfm(tmp_path / "d.pfm", a)
    b, scale = ec.load_pfm(tmp_path / "d.pfm")
    assert scale == 1.0 and b.shape == (5, 7, 1) and np.array_equal(b[:, :, 0], a)


def test_cli_signatures_match_reference():
    """-g / -i / -f as in match_single.py:12-18 and -g as in match.py:12-16."""
    from scenedepthestimation_b200 import match, match_single

    a = match_single.parser.parse_args(["-g", "0", "-i", "3", "-f", "x"])
    assert (a.gpu, a.id, a.file) == ("0", 3, "x")
    assert match_single.parser.parse_args([]).file == "11_11"
    assert match.parser.parse_args(["-g", "0,1"]).gpu == "0,1"
    assert match.shard(range(1, 19), 1, 8) == [2, 10, 18]


def test_synthetic_inputs_are_seeded():
    from scenedepthestimation_b200 import synthetic as syn

    a = syn.textured_pair(20, 30, 16, 3)
    b = syn.textured_pair(20, 30, 16, 3)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    w = syn.glorot_weights()
    assert w["conv1/weights:0"].shape == (3, 3, 1, 64) and w["conv5/biases:0"].shape == (64,)
    assert abs(float(np.abs(w["conv2/weights:0"]).max()) - np.sqrt(6 / (576 + 576))) < 1e-3


def test_weight_cache_key_is_content():
    """Host logic of the weight caches (no GPU): equal arrays give equal keys whatever the dict object, any change gives a new key."""
    from scenedepthestimation_b200 import process_functional as pf, synthetic as syn

    a, b = syn.glorot_weights(seed=3), syn.glorot_weights(seed=3)
    key = lambda w: pf._content_key(sorted(w.items()))
    assert a is not b and key(a) == key(b)
    b["conv2/biases:0"] = b["conv2/biases:0"].copy()
    b["conv2/biases:0"][5] += np.float32(1e-6)
    assert key(a) != key(b)
    assert key(a) != key({k: v.astype(np.float64) for k, v in a.items()})
    cache = {}
    for i in range(10):
        pf._cache_put(cache, i, i)
    assert list(cache) == [6, 7, 8, 9]


def test_net_variable_store_without_gpu(tmp_path):
    """mc_cnn_brunch drop-in: branches alias one weight dict (scope.reuse_variables(), mc_cnn_brunch.py:73-75)."""
    from scenedepthestimation_b200 import mc_cnn_brunch as mb, synthetic as syn

    mb.reset_default_graph()
    x = np.zeros((2, 11, 11, 1), np.float32)
    with pytest.raises(ValueError):
        mb.Net(x, num_of_conv_layers=5, is_branch=True)
    first = mb.Net(x, num_of_conv_layers=5)
    twin = mb.Net(x, num_of_conv_layers=5, is_branch=True)
    other = mb.Net(x, num_of_conv_layers=4)              # another configuration has its own variables
    assert twin.weights is first.weights and other.weights is not first.weights
    w = syn.glorot_weights(seed=42)
    np.save(tmp_path / "w.npy", w)
    twin.weights_path = str(tmp_path / "w.npy")
    twin.load_initial_weights()
    assert all(np.array_equal(first.weights[k], w[k]) for k in w)


def test_output_dtype_contract():
    from scenedepthestimation_b200 import match_single as ms

    assert ms.output_dtype(128, 1) is np.uint8 and ms.output_dtype(128, 2) is np.uint8 and ms.output_dtype(256, 1) is np.uint8
    assert ms.output_dtype(257, 1) is np.uint16 and ms.output_dtype(129, 2) is np.uint16 and ms.output_dtype(800, 1) is np.uint16
    d = np.array([[0.0, 3.75, 254.0, 399.5]], np.float32)
    assert np.array_equal(ms.encode_disparity(d, 128, 2), np.array([[0, 6, 252, 30]], np.uint8))   # the reference's wrap, kept
    assert np.array_equal(ms.encode_disparity(d, 400, 2), np.array([[0, 6, 508, 798]], np.uint16))


def test_fused_sharded_argument_errors_without_gpu(lib):
    """mccnn_sgm_fused_sharded rejects bad bands before any CUDA call."""
    import ctypes as C
    from scenedepthestimation_b200 import _lib

    assert lib.mccnn_sgm_fused_shard_exchange_bytes(0, 16) == 0 and lib.mccnn_sgm_fused_shard_exchange_bytes(64, 2000) == 0
    n64, n1000 = lib.mccnn_sgm_fused_shard_exchange_bytes(100, 64), lib.mccnn_sgm_fused_shard_exchange_bytes(100, 1000)
    assert 0 < n64 < n1000 and n64 % 256 == 0
    prm = _lib.default_sgm_params()
    p, big = 1 << 20, 1 << 40

    def call(shard):
        return lib.mccnn_sgm_fused_sharded(p, p, p, p, p, p, p, p, p, big, 64, 16, C.byref(prm), 0, shard, 15, None)

    assert call(None) == -1 and b"null shard" in lib.mccnn_last_error()
    bad_order = _lib.Shard(1, 2, 32, 0, 16, p, p, None, 1, None, 0)           # rank 1 must not start at row 0
    assert call(C.byref(bad_order)) == -1 and b"tile the image" in lib.mccnn_last_error()
    no_epoch = _lib.Shard(0, 2, 32, 0, 16, p, None, p, 0, None, 0)
    assert call(C.byref(no_epoch)) == -1 and b"epoch" in lib.mccnn_last_error()
    no_peer = _lib.Shard(0, 2, 32, 0, 16, p, None, None, 1, None, 0)
    assert call(C.byref(no_peer)) == -1 and b"peer" in lib.mccnn_last_error()
