"""Device-side plumbing above the C ABI: torch tensors for memory and streams, nothing else.

Every function here takes/returns CUDA torch tensors and forwards raw pointers to
libmccnn_b200.so. The reference-shaped NumPy API lives in process_functional.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

FEATURES = 64
FC_UNITS = 384  # MCCNN_FC_UNITS: hidden width of the MC-CNN-accurate head
EXACT = 0  # MCCNN_SGM_EXACT: the reference's arithmetic, bit for bit (default everywhere)
FUSED = 1  # MCCNN_SGM_FUSED: opt-in throughput mode (fp32 SGM state, 4 sweeps, fp32-accumulated cost volume; 1e-4 contract)


def _mode(mode) -> int:
    if mode in (EXACT, "exact", None):
        return EXACT
    if mode in (FUSED, "fused"):
        return FUSED
    raise ValueError(f"unknown SGM mode {mode!r} (use 'exact' or 'fused')")


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("scenedepthestimation_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t) -> int:
    return 0 if t is None else t.data_ptr()


def disp_pitch(D: int) -> int:
    return _lib.load().mccnn_disp_pitch(int(D))


def _dev(x, dtype) -> torch.Tensor:
    """numpy / torch -> contiguous CUDA tensor of dtype."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.to(device="cuda", dtype=dtype, non_blocking=True).contiguous()


class Workspace:
    """Caller-owned scratch memory (the C ABI never allocates), cached per (device, CUDA stream): two streams, or two host
    threads on their own streams, never share the volumes of a call in flight. A buffer that has to grow is handed back to
    torch's caching allocator, which keeps it alive for the work already queued on its stream (the tensor was allocated
    and only ever used on that stream)."""

    def __init__(self):
        self._bufs = {}

    def get(self, nbytes: int) -> torch.Tensor:
        key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            self._bufs.pop(key, None)
            buf = None
            with torch.cuda.stream(torch.cuda.current_stream()):
                buf = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
            self._bufs[key] = buf
        return buf

    def clear(self):
        self._bufs.clear()


_ws = Workspace()


# ------------------------------------------------------------------------------ weights
def pack_weights(weights: dict, num_layers: int = 5) -> torch.Tensor:
    """Reference dict {'conv{i}/weights:0': HWIO, 'conv{i}/biases:0'} (mc_cnn_brunch.py:61-66) -> device blob."""
    _require_cuda()
    lib = _lib.load()
    ws, bs = [], []
    for i in range(1, num_layers + 1):
        w = np.ascontiguousarray(weights[f"conv{i}/weights:0"], dtype=np.float32)
        b = np.ascontiguousarray(weights[f"conv{i}/biases:0"], dtype=np.float32)
        cin = 1 if i == 1 else FEATURES
        if w.shape != (3, 3, cin, FEATURES) or b.shape != (FEATURES,):
            raise ValueError(f"conv{i}: expected weights (3,3,{cin},{FEATURES}) and biases ({FEATURES},), got {w.shape} {b.shape}")
        ws.append(w)
        bs.append(b)
    nbytes = lib.mccnn_conv_packed_weight_bytes(num_layers)
    host = np.zeros(nbytes, np.uint8)
    wp = (C.c_void_p * num_layers)(*[w.ctypes.data for w in ws])
    bp = (C.c_void_p * num_layers)(*[b.ctypes.data for b in bs])
    _lib.check(lib.mccnn_pack_weights_host(wp, bp, num_layers, host.ctypes.data), "mccnn_pack_weights_host")
    return torch.from_numpy(host).cuda()


# ------------------------------------------------------------------------------ stages
def standardize_pad(image_u8: torch.Tensor, pad: int) -> torch.Tensor:
    H, W = image_u8.shape
    out = torch.empty((H + 2 * pad, W + 2 * pad), dtype=torch.float32, device="cuda")
    scratch = torch.empty(4, dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().mccnn_standardize_pad(_p(image_u8), _p(out), _p(scratch), H, W, pad, _stream()), "mccnn_standardize_pad")
    return out


def pad_f32(image: torch.Tensor, pad: int) -> torch.Tensor:
    H, W = image.shape
    out = torch.empty((H + 2 * pad, W + 2 * pad), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().mccnn_pad_f32(_p(image), _p(out), H, W, pad, _stream()), "mccnn_pad_f32")
    return out


def conv_tower(padded: torch.Tensor, packed: torch.Tensor, num_layers: int = 5, fp32: bool = False) -> torch.Tensor:
    """Tensor-core tower (tcgen05, fp16 hi/lo split); fp32=True runs the CUDA-core fp32 twin."""
    lib = _lib.load()
    Hp, Wp = padded.shape
    H, W = Hp - 2 * num_layers, Wp - 2 * num_layers
    feat = torch.empty((H, W, FEATURES), dtype=torch.float32, device="cuda")
    nws = lib.mccnn_conv_workspace_bytes(H, W, num_layers)
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    fn = lib.mccnn_conv_tower_fp32 if fp32 else lib.mccnn_conv_tower
    _lib.check(fn(_p(padded), _p(packed), _p(feat), _p(ws), nws, H, W, num_layers, _stream()), "mccnn_conv_tower")
    return feat


def cost_volume(fl: torch.Tensor, fr: torch.Tensor, D: int, fill: float = 1.0, right: bool = True):
    H, W, F = fl.shape
    assert F == FEATURES and fr.shape == fl.shape
    Dp = disp_pitch(D)
    CL = torch.empty((H, W, Dp), dtype=torch.float32, device="cuda")
    CR = torch.empty((H, W, Dp), dtype=torch.float32, device="cuda") if right else None
    _lib.check(_lib.load().mccnn_cost_volume(_p(fl), _p(fr), _p(CL), _p(CR), H, W, D, float(fill), _stream()), "mccnn_cost_volume")
    return CL, CR


def cost_volume_fast(fl: torch.Tensor, fr: torch.Tensor, D: int, fill: float = 1.0, right: bool = True, tensor_cores: bool = True):
    """The fused mode's cost volume (not the reference's bits; |difference| <= 4e-6): on the tensor cores (fp16 hi/lo split,
    fp32 accumulation in TMEM), or with tensor_cores=False the CUDA-core band GEMM with fp32 FMA accumulation."""
    lib = _lib.load()
    H, W, F = fl.shape
    assert F == FEATURES and fr.shape == fl.shape
    Dp = disp_pitch(D)
    CL = torch.empty((H, W, Dp), dtype=torch.float32, device="cuda")
    CR = torch.empty((H, W, Dp), dtype=torch.float32, device="cuda") if right else None
    if tensor_cores:
        nws = lib.mccnn_cost_volume_fast_tc_workspace_bytes(H, W)
        ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
        _lib.check(lib.mccnn_cost_volume_fast_tc(_p(fl), _p(fr), _p(CL), _p(CR), _p(ws), nws, H, W, D, float(fill), _stream()),
                   "mccnn_cost_volume_fast_tc")
    else:
        _lib.check(lib.mccnn_cost_volume_fast(_p(fl), _p(fr), _p(CL), _p(CR), H, W, D, float(fill), _stream()), "mccnn_cost_volume_fast")
    return CL, CR


def cost_volume_tc(fl: torch.Tensor, fr: torch.Tensor, D: int, fill: float = 1.0, right: bool = True):
    """Tensor-core variant of cost_volume: same contract, same bits."""
    lib = _lib.load()
    H, W, F = fl.shape
    assert F == FEATURES and fr.shape == fl.shape
    Dp = disp_pitch(D)
    CL = torch.empty((H, W, Dp), dtype=torch.float32, device="cuda")
    CR = torch.empty((H, W, Dp), dtype=torch.float32, device="cuda") if right else None
    nws = lib.mccnn_cost_volume_tc_workspace_bytes(H, W)
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    _lib.check(lib.mccnn_cost_volume_tc(_p(fl), _p(fr), _p(CL), _p(CR), _p(ws), nws, H, W, D, float(fill), _stream()),
               "mccnn_cost_volume_tc")
    return CL, CR


class FcHeadWeights:
    """Device copy of the MC-CNN-accurate head's weights in the layout mccnn_cost_volume_accurate wants
    (fc1 split per image, fc2 / fc3 as the pre-swizzled fp16 blocks of mccnn_pack_fc_matrix_host); keeps the tensors alive
    for the ctypes struct."""

    def __init__(self, weights: dict):
        _require_cuda()

        def get(name):
            for k in (name, name + ":0"):
                if k in weights:
                    return np.asarray(weights[k], dtype=np.float32)
            raise KeyError(f"{name} missing from the weights dict (keys: {sorted(weights)[:8]} ...)")

        w1, w2, w3, w4 = (get(f"fc{i}/weights") for i in (1, 2, 3, 4))
        if w1.shape != (2 * FEATURES, FC_UNITS) or w2.shape != (FC_UNITS, FC_UNITS) or w3.shape != (FC_UNITS, FC_UNITS) \
                or w4.reshape(-1).shape != (FC_UNITS,):
            raise ValueError(f"head must be fc1 [{2 * FEATURES},{FC_UNITS}], fc2/fc3 [{FC_UNITS},{FC_UNITS}], fc4 [{FC_UNITS},1]")
        dev = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt).contiguous()
        lib = _lib.load()

        def blocks(w):  # [384 in][384 out] fp32 -> the kernel's pre-swizzled fp16 blocks (one definition of the layout: the C side)
            src = np.ascontiguousarray(w, dtype=np.float32)
            out = np.empty(lib.mccnn_fc_matrix_blocks_bytes(), np.uint8)
            _lib.check(lib.mccnn_pack_fc_matrix_host(src.ctypes.data, out.ctypes.data), "mccnn_pack_fc_matrix_host")
            return torch.from_numpy(out).cuda()

        self.t = dict(w1_left=dev(w1[:FEATURES]), w1_right=dev(w1[FEATURES:]), b1=dev(get("fc1/biases")),
                      w2b=blocks(w2), b2=dev(get("fc2/biases")), w3b=blocks(w3),
                      b3=dev(get("fc3/biases")), w4=dev(w4.reshape(-1)))
        t = self.t
        self.c = _lib.FcWeights(t["w1_left"].data_ptr(), t["w1_right"].data_ptr(), t["b1"].data_ptr(), t["w2b"].data_ptr(),
                                t["b2"].data_ptr(), t["w3b"].data_ptr(), t["b3"].data_ptr(), t["w4"].data_ptr(),
                                float(get("fc4/biases").reshape(-1)[0]))


def cost_volume_accurate(fl: torch.Tensor, fr: torch.Tensor, head: FcHeadWeights, D: int, fill: float = 1.0, right: bool = True):
    """MC-CNN-accurate matching cost (fully-connected head on tcgen05) -> (CL, CR) like cost_volume."""
    lib = _lib.load()
    H, W, F = fl.shape
    assert F == FEATURES and fr.shape == fl.shape
    Dp = disp_pitch(D)
    CL = torch.empty((H, W, Dp), dtype=torch.float32, device="cuda")
    CR = torch.empty((H, W, Dp), dtype=torch.float32, device="cuda") if right else None
    nws = lib.mccnn_fc_head_workspace_bytes(H, W)
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    _lib.check(lib.mccnn_cost_volume_accurate(_p(fl), _p(fr), C.byref(head.c), _p(CL), _p(CR), _p(ws), nws, H, W, D, float(fill),
                                              _stream()), "mccnn_cost_volume_accurate")
    return CL, CR


def volume_to_dhw(vol: torch.Tensor, D: int) -> torch.Tensor:
    H, W, _ = vol.shape
    out = torch.empty((D, H, W), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().mccnn_volume_to_dhw(_p(vol), _p(out), H, W, D, _stream()), "mccnn_volume_to_dhw")
    return out


def cross_arms(image_u8: torch.Tensor, L1: int = 14, tau: int = 6) -> torch.Tensor:
    """u8 [H,W] -> u8 [H,W,4] arm lengths (left, right, up, down) for the cross-based aggregation."""
    H, W = image_u8.shape
    arms = torch.empty((H, W, 4), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().mccnn_cross_arms(_p(image_u8), _p(arms), H, W, int(L1), int(tau), _stream()), "mccnn_cross_arms")
    return arms


def cbca(CL, CR, imageL, imageR, D: int, iters: int = 2, L1: int = 14, tau: int = 6):
    """`iters` cross-based aggregation passes of both volumes (north_star stage 3; absent from the reference)."""
    lib = _lib.load()
    H, W, _ = CL.shape
    al, ar = cross_arms(imageL, L1, tau), cross_arms(imageR, L1, tau)
    tmp = torch.empty_like(CL)
    for _ in range(int(iters)):
        nl, nr = torch.empty_like(CL), torch.empty_like(CR)
        _lib.check(lib.mccnn_cbca(_p(CL), _p(nl), _p(tmp), _p(al), _p(ar), H, W, D, -1, int(L1), _stream()), "mccnn_cbca")
        _lib.check(lib.mccnn_cbca(_p(CR), _p(nr), _p(tmp), _p(ar), _p(al), H, W, D, 1, int(L1), _stream()), "mccnn_cbca")
        CL, CR = nl, nr
    return CL, CR


def sgm(CL, CR, imageL, imageR, D: int, params=None, keep_volumes: bool = True, mode=EXACT):
    """8-path SGM + fused WTA. Returns (SL, SR, dispL, dispR). mode: EXACT (reference bits) or FUSED (4 sweeps, fp32 state)."""
    lib = _lib.load()
    H, W, Dp = CL.shape
    params = params or _lib.default_sgm_params()
    SL, SR = torch.empty_like(CL), torch.empty_like(CR)
    dl = torch.empty((H, W), dtype=torch.float32, device="cuda")
    dr = torch.empty((H, W), dtype=torch.float32, device="cuda")
    nws = lib.mccnn_sgm_workspace_bytes(H, W, D)
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    _lib.check(lib.mccnn_sgm(_p(CL), _p(CR), _p(imageL), _p(imageR), _p(SL), _p(SR), _p(dl), _p(dr), _p(ws), nws,
                             H, W, D, C.byref(params), _mode(mode), 1 if keep_volumes else 0, _stream()), "mccnn_sgm")
    return SL, SR, dl, dr


def sgm_single_path(Cv, image, S, D: int, path: int, params=None):
    lib = _lib.load()
    H, W, _ = Cv.shape
    params = params or _lib.default_sgm_params()
    ws = torch.empty(256, dtype=torch.uint8, device="cuda")
    _lib.check(lib.mccnn_sgm_single_path(_p(Cv), _p(image), _p(S), _p(ws), 256, H, W, D, C.byref(params), path, _stream()),
               "mccnn_sgm_single_path")
    return S


def wta(S: torch.Tensor, D: int) -> torch.Tensor:
    H, W, _ = S.shape
    out = torch.empty((H, W), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().mccnn_wta(_p(S), _p(out), H, W, D, _stream()), "mccnn_wta")
    return out


def wta_subpixel(S: torch.Tensor, D: int) -> torch.Tensor:
    H, W, _ = S.shape
    out = torch.empty((H, W), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().mccnn_wta_subpixel(_p(S), _p(out), H, W, D, _stream()), "mccnn_wta_subpixel")
    return out


def encode_u16(disp, frac_bits: int = 0):
    H, W = disp.shape
    out = torch.empty((H, W), dtype=torch.int16, device="cuda")
    _lib.check(_lib.load().mccnn_encode_u16(_p(disp), _p(out), H, W, int(frac_bits), _stream()), "mccnn_encode_u16")
    return out.view(torch.uint16) if hasattr(torch, "uint16") else out


def wta_dhw(vol: torch.Tensor) -> torch.Tensor:
    D, H, W = vol.shape
    out = torch.empty((H, W), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().mccnn_wta_dhw(_p(vol), _p(out), H, W, D, _stream()), "mccnn_wta_dhw")
    return out


def lr_flags(dl, dr, right: bool = True):
    H, W = dl.shape
    fl = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    fr = torch.empty((H, W), dtype=torch.uint8, device="cuda") if right else None
    _lib.check(_lib.load().mccnn_lr_flags(_p(dl), _p(dr), _p(fl), _p(fr), H, W, _stream()), "mccnn_lr_flags")
    return fl, fr


def lrc_fill(dl, flag_l):
    H, W = dl.shape
    out = torch.empty_like(dl)
    lib = _lib.load()
    nws = lib.mccnn_lrc_fill_workspace_bytes(H, W)
    ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    _lib.check(lib.mccnn_lrc_fill(_p(dl), _p(flag_l), _p(out), _p(ws), nws, H, W, _stream()), "mccnn_lrc_fill")
    return out


def median5(filled, wta_map):
    H, W = filled.shape
    out = torch.empty_like(filled)
    _lib.check(_lib.load().mccnn_median5(_p(filled), _p(wta_map), _p(out), H, W, _stream()), "mccnn_median5")
    return out


def bilateral9(image_u8, disp):
    H, W = disp.shape
    out = torch.empty_like(disp)
    _lib.check(_lib.load().mccnn_bilateral9(_p(image_u8), _p(disp), _p(out), H, W, _stream()), "mccnn_bilateral9")
    return out


def encode_u8(disp, scale: int = 1):
    H, W = disp.shape
    out = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load().mccnn_encode_u8(_p(disp), _p(out), H, W, int(scale), _stream()), "mccnn_encode_u8")
    return out


def bad_pixels(disp_int, gt_half):
    """disp_int: u8 map, or a 16-bit map (torch.int16 / uint16 storage, read as unsigned)."""
    H, W = disp_int.shape
    counts = torch.empty(2, dtype=torch.int64, device="cuda")
    fn = _lib.load().mccnn_bad_pixels if disp_int.element_size() == 1 else _lib.load().mccnn_bad_pixels_u16
    _lib.check(fn(_p(disp_int), _p(gt_half), _p(counts), H, W, _stream()), "mccnn_bad_pixels")
    bad, valid = counts.tolist()
    return bad, valid


# ------------------------------------------------------------------------------ whole path
def disparity_pipeline(imageL, imageR, fl, fr, D: int, params=None, stage_ms: np.ndarray | None = None,
                       out=None, workspace: torch.Tensor | None = None, mode=EXACT):
    """mccnn_disparity_pipeline on device tensors -> (dispL filtered, dispR raw WTA)."""
    lib = _lib.load()
    H, W = imageL.shape
    params = params or _lib.default_sgm_params()
    nws = lib.mccnn_pipeline_workspace_bytes(H, W, D)
    ws = workspace if workspace is not None else _ws.get(nws)
    dl, dr = out if out is not None else (torch.empty((H, W), dtype=torch.float32, device="cuda"),
                                          torch.empty((H, W), dtype=torch.float32, device="cuda"))
    sm = stage_ms.ctypes.data if stage_ms is not None else None
    _lib.check(lib.mccnn_disparity_pipeline(_p(imageL), _p(imageR), _p(fl), _p(fr), _p(dl), _p(dr), _p(ws), ws.numel(),
                                            H, W, D, C.byref(params), _mode(mode), sm, _stream()), "mccnn_disparity_pipeline")
    return dl, dr


def match_workspace_bytes(H: int, W: int, D: int, num_layers: int = 5) -> int:
    return _lib.load().mccnn_match_workspace_bytes(H, W, D, num_layers)


def match_accurate_workspace_bytes(H: int, W: int, D: int, num_layers: int = 5) -> int:
    return _lib.load().mccnn_match_accurate_workspace_bytes(H, W, D, num_layers)


def match_pair(imageL, imageR, packed, D: int, num_layers: int = 5, params=None, stage_ms: np.ndarray | None = None,
               out=None, workspace: torch.Tensor | None = None, head: "FcHeadWeights | None" = None, mode=EXACT):
    """mccnn_match_pair on device u8 images -> (dispL filtered, dispR raw WTA); with `head` the matching cost is the
    MC-CNN-accurate decision head (mccnn_match_pair_accurate)."""
    lib = _lib.load()
    H, W = imageL.shape
    params = params or _lib.default_sgm_params()
    nws = (lib.mccnn_match_accurate_workspace_bytes if head is not None else lib.mccnn_match_workspace_bytes)(H, W, D, num_layers)
    ws = workspace if workspace is not None else _ws.get(nws)
    dl, dr = out if out is not None else (torch.empty((H, W), dtype=torch.float32, device="cuda"),
                                          torch.empty((H, W), dtype=torch.float32, device="cuda"))
    sm = stage_ms.ctypes.data if stage_ms is not None else None
    if head is not None:
        _lib.check(lib.mccnn_match_pair_accurate(_p(imageL), _p(imageR), _p(packed), C.byref(head.c), _p(dl), _p(dr), _p(ws),
                                                 ws.numel(), H, W, D, num_layers, C.byref(params), _mode(mode), sm, _stream()),
                   "mccnn_match_pair_accurate")
    else:
        _lib.check(lib.mccnn_match_pair(_p(imageL), _p(imageR), _p(packed), _p(dl), _p(dr), _p(ws), ws.numel(), H, W, D,
                                        num_layers, C.byref(params), _mode(mode), sm, _stream()), "mccnn_match_pair")
    return dl, dr
