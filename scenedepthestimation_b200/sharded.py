"""One stereo pair split over the GPUs of a node by image rows (BASELINE config 4 at N > 1).

Rank r owns rows [row0, row0 + rows). What crosses NVLink:
  * the u8 image bands are all-gathered once (11 MB for the full-res pair): every rank needs whole images for
    the SGM penalties, and the gather also provides the 5-row conv halos and the global mean / std;
  * conv tower, cost volume and the horizontal SGM paths are band-local;
  * vertical / diagonal SGM scanlines resume on the next rank from the fp64 path state the previous rank stored
    into its neighbour's exchange buffer from INSIDE the scan kernel (peer stores + a release flag; no collective,
    no host round trip): 5 passes x 2 sides x W scanlines x (D + 1) doubles per boundary;
  * the two raw WTA bands are all-gathered (23 MB each) and the cheap L-R check / fill / median run on the whole map.
The result is bit-identical to the single-GPU path.

mode="fused" (opt-in, csrc/sgm_fused.cu): the 4 sweeps of the fused SGM on the bands. The two row sweeps advance in lock step
on all ranks (the last row of a band streams one fp32 state per step into its neighbour over NVLink), the column and the
diagonal sweep hand their states over per column / scanline. Value-identical to the single-GPU fused mode.

torch.distributed provides the process group and torch's symmetric memory the peer-mapped exchange buffers.
`emulate_bands` runs the same band kernels for several "ranks" on ONE GPU in dependency order (tests).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import engine as eng

PASS_DOWNWARD = (True, None, None, True, False, True, False)  # pass -> sweeps down / band-local / sweeps up


def band_rows(H: int, world: int):
    """Even split of H rows: [(row0, rows)] per rank."""
    base, extra = divmod(H, world)
    out, r0 = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((r0, n))
        r0 += n
    return out


def _shard(rank, world, H, row0, rows, local, prev, nxt, epoch, go_flag=None, timeout_ms=0) -> _lib.Shard:
    return _lib.Shard(rank, world, H, row0, rows, local, prev, nxt, epoch, go_flag, timeout_ms)


def sgm_band(CLb, CRb, il_full, ir_full, D, shard: _lib.Shard, pass_mask=0x7F, keep_volumes=True, out=None, params=None, ws=None):
    """mccnn_sgm_sharded on this rank's band -> (SLb, SRb, dispLb, dispRb). `ws`: the 256-byte SGM workspace (holds the
    launch's status word, read back by band_status)."""
    lib = _lib.load()
    rows, W, _ = CLb.shape
    params = params or _lib.default_sgm_params()
    if out is None:
        out = (torch.empty_like(CLb), torch.empty_like(CRb),
               torch.empty((rows, W), dtype=torch.float32, device="cuda"), torch.empty((rows, W), dtype=torch.float32, device="cuda"))
    SLb, SRb, dl, dr = out
    nws = 256   # the exact-mode words only (counters + status); the fused mode's rings are not used by sharded launches
    if ws is None:
        ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
    _lib.check(lib.mccnn_sgm_sharded(CLb.data_ptr(), CRb.data_ptr(), il_full.data_ptr(), ir_full.data_ptr(), SLb.data_ptr(),
                                     SRb.data_ptr(), dl.data_ptr(), dr.data_ptr(), ws.data_ptr(), nws, W, D, C.byref(params),
                                     eng.EXACT, 1 if keep_volumes else 0, C.byref(shard), pass_mask,
                                     torch.cuda.current_stream().cuda_stream), "mccnn_sgm_sharded")
    return out


FUSED_SWEEP_DOWNWARD = (True, True, False, False)   # sweep -> runs through the bands top-down / bottom-up


def sgm_fused_band(CLb, CRb, il_full, ir_full, D, shard: _lib.Shard, ws, sweep_mask=15, keep_volumes=True, out=None, params=None):
    """mccnn_sgm_fused_sharded on this rank's band -> (SLb, SRb, dispLb, dispRb). `ws`: mccnn_sgm_workspace_bytes(H, W, D) bytes."""
    lib = _lib.load()
    rows, W, _ = CLb.shape
    params = params or _lib.default_sgm_params()
    if out is None:
        out = (torch.empty_like(CLb), torch.empty_like(CRb),
               torch.empty((rows, W), dtype=torch.float32, device="cuda"), torch.empty((rows, W), dtype=torch.float32, device="cuda"))
    SLb, SRb, dl, dr = out
    _lib.check(lib.mccnn_sgm_fused_sharded(CLb.data_ptr(), CRb.data_ptr(), il_full.data_ptr(), ir_full.data_ptr(), SLb.data_ptr(),
                                           SRb.data_ptr(), dl.data_ptr(), dr.data_ptr(), ws.data_ptr(), ws.numel(), W, D,
                                           C.byref(params), 1 if keep_volumes else 0, C.byref(shard), sweep_mask,
                                           torch.cuda.current_stream().cuda_stream), "mccnn_sgm_fused_sharded")
    return out


def emulate_fused_bands(CL, CR, il, ir, D, world: int, epoch: int = 1, keep_volumes: bool = True, params=None):
    """The fused mode's `world` row bands of one pair on ONE GPU, sweep by sweep in dependency order."""
    lib = _lib.load()
    H, W, _ = CL.shape
    bands = band_rows(H, world)
    nx = lib.mccnn_sgm_fused_shard_exchange_bytes(W, D)
    xchg = [torch.zeros(nx, dtype=torch.uint8, device="cuda") for _ in range(world)]
    ws = [torch.zeros(lib.mccnn_sgm_workspace_bytes(H, W, D), dtype=torch.uint8, device="cuda") for _ in range(world)]
    # the fused sweeps accumulate into the S volumes in place, starting from a copy of the cost volumes made by sweep 0
    SL, SR = torch.empty_like(CL), torch.empty_like(CR)
    dl = torch.empty((H, W), dtype=torch.float32, device="cuda")
    dr = torch.empty((H, W), dtype=torch.float32, device="cuda")
    shards = []
    for r, (r0, n) in enumerate(bands):
        shards.append(_shard(r, world, H, r0, n, xchg[r].data_ptr(), xchg[r - 1].data_ptr() if r > 0 else None,
                             xchg[r + 1].data_ptr() if r < world - 1 else None, epoch))
    for s in range(4):
        order = range(world) if FUSED_SWEEP_DOWNWARD[s] else range(world - 1, -1, -1)
        for r in order:
            r0, n = bands[r]
            sgm_fused_band(CL[r0:r0 + n], CR[r0:r0 + n], il, ir, D, shards[r], ws[r], sweep_mask=1 << s, keep_volumes=keep_volumes,
                           out=(SL[r0:r0 + n], SR[r0:r0 + n], dl[r0:r0 + n], dr[r0:r0 + n]), params=params)
    for w in ws:
        if band_status(w):
            raise RuntimeError(lib.mccnn_last_error().decode(errors="replace"))
    return SL, SR, dl, dr


def band_status(ws) -> int:
    """Synchronises the stream; 0 = every scanline of the last mccnn_sgm_sharded launch got its hand-over, 1 = a wait hit the
    deadline (a neighbour rank died / never launched): the outputs are invalid."""
    st = C.c_int(0)
    _lib.check(_lib.load().mccnn_sgm_shard_status(ws.data_ptr(), C.byref(st), torch.cuda.current_stream().cuda_stream),
               "mccnn_sgm_shard_status")
    return int(st.value)


def emulate_bands(CL, CR, il, ir, D, world: int, epoch: int = 1):
    """Run `world` row bands of one pair on ONE GPU, pass by pass in dependency order (a band never waits: the band
    it depends on has already finished its launch). Returns (SL, SR, dispL, dispR) assembled from the bands."""
    lib = _lib.load()
    H, W, _ = CL.shape
    bands = band_rows(H, world)
    nx = lib.mccnn_sgm_shard_exchange_bytes(W)
    xchg = [torch.zeros(nx, dtype=torch.uint8, device="cuda") for _ in range(world)]
    SL, SR = torch.empty_like(CL), torch.empty_like(CR)
    dl = torch.empty((H, W), dtype=torch.float32, device="cuda")
    dr = torch.empty((H, W), dtype=torch.float32, device="cuda")
    shards = []
    for r, (r0, n) in enumerate(bands):
        shards.append(_shard(r, world, H, r0, n, xchg[r].data_ptr(), xchg[r - 1].data_ptr() if r > 0 else None,
                             xchg[r + 1].data_ptr() if r < world - 1 else None, epoch))
    for p in range(7):
        order = range(world) if PASS_DOWNWARD[p] in (True, None) else range(world - 1, -1, -1)
        for r in order:
            r0, n = bands[r]
            sgm_band(CL[r0:r0 + n], CR[r0:r0 + n], il, ir, D, shards[r], pass_mask=1 << p,
                     out=(SL[r0:r0 + n], SR[r0:r0 + n], dl[r0:r0 + n], dr[r0:r0 + n]))
    return SL, SR, dl, dr


def gather_bands(send, recv, bands, left_full, right_full, group=None):
    """One all_gather of every rank's (padded) [left | right] band `send` [2, max_rows, W] into `recv` [world, 2, max_rows, W],
    then the valid rows of each band are laid out as whole maps. Bands may differ by one row, hence the padding."""
    import torch.distributed as dist

    dist.all_gather_into_tensor(recv.view(-1, *send.shape[1:]), send, group=group)   # (concatenated along dim 0: any backend)
    for r, (r0, n) in enumerate(bands):
        left_full[r0:r0 + n].copy_(recv[r, 0, :n])
        right_full[r0:r0 + n].copy_(recv[r, 1, :n])


class ShardedMatcher:
    """One pair per call, split by rows over the ranks of `group` (one process per GPU, NCCL)."""

    def __init__(self, H: int, W: int, D: int, weights: dict, num_layers: int = 5, group=None, timeout_ms: int = 2000,
                 mode: str = "exact"):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.dist = dist
        self.group = group or dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.H, self.W, self.D, self.nl = H, W, D, num_layers
        self.bands = band_rows(H, self.world)
        self.row0, self.rows = self.bands[self.rank]
        self.max_rows = max(n for _, n in self.bands)  # bands may differ by one row: gathers are padded to this
        self.packed = eng.pack_weights(weights, num_layers)
        lib = _lib.load()
        self.fused = eng._mode(mode) == eng.FUSED
        nx = lib.mccnn_sgm_fused_shard_exchange_bytes(W, D) if self.fused else lib.mccnn_sgm_shard_exchange_bytes(W)
        self.xchg = symm_mem.empty(nx, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        self.xchg.zero_()
        self.handle = symm_mem.rendezvous(self.xchg, self.group)
        ptrs = list(self.handle.buffer_ptrs)
        self.prev = ptrs[self.rank - 1] if self.rank > 0 else None
        self.next = ptrs[self.rank + 1] if self.rank < self.world - 1 else None
        self.epoch = 0
        self.timeout_ms = int(timeout_ms)
        self.go = torch.ones(1, dtype=torch.int32, device="cuda")       # all-reduced (MIN) before every SGM launch
        self.sgm_ws = torch.zeros(lib.mccnn_sgm_workspace_bytes(H, W, D) if self.fused else 256, dtype=torch.uint8, device="cuda")
        self.tc_ws = None
        if self.fused:
            self.tc_ws = torch.empty(lib.mccnn_cost_volume_fast_tc_workspace_bytes(self.max_rows, W), dtype=torch.uint8, device="cuda")
        elif D >= 512:   # the cost-volume variant mccnn_disparity_pipeline picks for wide bands (pipeline.cu)
            self.tc_ws = torch.empty(lib.mccnn_cost_volume_tc_workspace_bytes(self.max_rows, W), dtype=torch.uint8, device="cuda")
        Dp = eng.disp_pitch(D)
        self.S = (torch.empty((self.rows, W, Dp), dtype=torch.float32, device="cuda"),
                  torch.empty((self.rows, W, Dp), dtype=torch.float32, device="cuda"))
        self.il = torch.empty((H, W), dtype=torch.uint8, device="cuda")
        self.ir = torch.empty((H, W), dtype=torch.uint8, device="cuda")
        self.dl = torch.empty((H, W), dtype=torch.float32, device="cuda")
        self.dr = torch.empty((H, W), dtype=torch.float32, device="cuda")
        # [left | right] band of this rank, padded to max_rows, and the gathered stacks
        self.u8_send = torch.zeros((2, self.max_rows, W), dtype=torch.uint8, device="cuda")
        self.u8_recv = torch.empty((self.world, 2, self.max_rows, W), dtype=torch.uint8, device="cuda")
        self.f_send = torch.zeros((2, self.max_rows, W), dtype=torch.float32, device="cuda")
        self.f_recv = torch.empty((self.world, 2, self.max_rows, W), dtype=torch.float32, device="cuda")

    def _gather(self, send, recv, left_full, right_full):
        gather_bands(send, recv, self.bands, left_full, right_full, self.group)

    def _cost_volume(self, fl, fr):
        if self.tc_ws is None:
            return eng.cost_volume(fl, fr, self.D)
        lib = _lib.load()
        n, W = fl.shape[0], self.W
        Dp = eng.disp_pitch(self.D)
        CL = torch.empty((n, W, Dp), dtype=torch.float32, device="cuda")
        CR = torch.empty((n, W, Dp), dtype=torch.float32, device="cuda")
        if self.fused:
            _lib.check(lib.mccnn_cost_volume_fast_tc(fl.data_ptr(), fr.data_ptr(), CL.data_ptr(), CR.data_ptr(), self.tc_ws.data_ptr(),
                                                     self.tc_ws.numel(), n, W, self.D, 1.0, torch.cuda.current_stream().cuda_stream),
                       "mccnn_cost_volume_fast_tc")
            return CL, CR
        _lib.check(lib.mccnn_cost_volume_tc(fl.data_ptr(), fr.data_ptr(), CL.data_ptr(), CR.data_ptr(), self.tc_ws.data_ptr(),
                                            self.tc_ws.numel(), n, W, self.D, 1.0, torch.cuda.current_stream().cuda_stream),
                   "mccnn_cost_volume_tc")
        return CL, CR

    def match(self, il_band: torch.Tensor, ir_band: torch.Tensor, check: bool = True):
        """u8 bands [rows, W] of this rank -> (filtered left disparity, raw right WTA), whole maps on every rank.

        No host round trip between the ranks: an all-reduce (MIN) of each rank's "my inputs are fine, I will launch" word
        runs on the stream in front of the SGM launch. It is the device-side barrier that keeps a rank from overwriting
        exchange slots its neighbours still read for the previous pair, and the kernels return at once when it is 0. With
        check=True the status word of the launch (a hand-over that never came within timeout_ms) is read back at the end
        and raises; check=False leaves that to a later call of status()."""
        dist, nl = self.dist, self.nl
        r0, n = self.row0, self.rows
        ok = tuple(il_band.shape) == (n, self.W) and tuple(ir_band.shape) == (n, self.W) and il_band.dtype == torch.uint8 \
            and ir_band.dtype == torch.uint8 and il_band.is_cuda and ir_band.is_cuda
        self.go.fill_(1 if ok else 0)
        if not ok:   # take part in this pair's collectives with an empty band so that the other ranks are told, then raise
            il_band = ir_band = torch.zeros((n, self.W), dtype=torch.uint8, device="cuda")
        self.u8_send[0, :n].copy_(il_band)
        self.u8_send[1, :n].copy_(ir_band)
        self._gather(self.u8_send, self.u8_recv, self.il, self.ir)
        feats = []
        for img in (self.il, self.ir):
            padded = eng.standardize_pad(img, nl)          # global statistics, zero padding at the image borders
            feats.append(eng.conv_tower(padded[r0:r0 + n + 2 * nl], self.packed, nl))  # band + 5-row halos
        CL, CR = self._cost_volume(feats[0], feats[1])
        # on the stream, no host sync: neighbours have finished the previous pair's SGM (stream order) once this completes
        dist.all_reduce(self.go, op=dist.ReduceOp.MIN, group=self.group)
        self.epoch += 1
        shard = _shard(self.rank, self.world, self.H, r0, n, self.xchg.data_ptr(), self.prev, self.next, self.epoch,
                       self.go.data_ptr(), self.timeout_ms)
        out = (self.S[0], self.S[1], self.f_send[0, :n], self.f_send[1, :n])
        if self.fused:
            sgm_fused_band(CL, CR, self.il, self.ir, self.D, shard, self.sgm_ws, keep_volumes=False, out=out)
        else:
            sgm_band(CL, CR, self.il, self.ir, self.D, shard, keep_volumes=False, out=out, ws=self.sgm_ws)
        self._gather(self.f_send, self.f_recv, self.dl, self.dr)
        fl, _ = eng.lr_flags(self.dl, self.dr, right=False)
        filled = eng.lrc_fill(self.dl, fl)
        out = eng.median5(filled, self.dl), self.dr
        if not ok:
            raise ValueError(f"rank {self.rank}: bands must be CUDA u8 tensors of shape ({n}, {self.W}); the pair was abandoned on every rank")
        if check:
            self.status()
        return out

    def status(self):
        """Raise if the last pair was abandoned (a rank reported bad inputs) or a hand-over never arrived."""
        st = band_status(self.sgm_ws)   # synchronises the stream
        if int(self.go.item()) == 0:
            raise RuntimeError("sharded match: another rank reported invalid inputs; the pair was abandoned on every rank")
        if st != 0:
            raise RuntimeError(_lib.load().mccnn_last_error().decode(errors="replace"))
