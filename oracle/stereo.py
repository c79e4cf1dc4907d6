"""ORACLE (test infrastructure, never the product path): CPU restatement of the
reference's post-feature stereo pipeline, stage by stage, with the reference's exact
arithmetic ("quirks") and the disparity count D as a parameter.

Every function cites the lines of /root/reference/process_functional.py it restates.
The reference kernels are Numba ``@cuda.jit`` Python; this file is CPU ``@numba.njit``
so the SAME type-inference engine decides where the arithmetic is fp32, fp64 or
unsigned 64-bit. Where the reference relies on behaviour that
is formally undefined (the ``uint8()`` cast of an out-of-range double in the L-R
check), the oracle writes the hardware-observed behaviour out explicitly.

Pinned against tests/golden/ref_*.npz, which were produced by running the reference's
own kernels on a B200 with tools/ref_gpu_probe.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module. The product path never does.
"""
from __future__ import annotations

import numpy as np
from numba import njit, prange

# launch order of the 8 path kernels, process_functional.py:1166-1202
PATH_NAMES = ("UpToDown", "DownToUp", "LeftToRight", "RightToLeft",
              "UpToDownAndLeftToRight", "DownToUpAndLeftToRight",
              "UpToDownAndRightToLeft", "DownToUpAndRightToLeft")
#            dy, dx, P1 channel (P2 channel = +1)   -- kernels at :346-797
_PATH_DY = np.array([1, -1, 0, 0, 1, -1, 1, -1], np.int64)
_PATH_DX = np.array([0, 0, 1, -1, 1, 1, -1, -1], np.int64)
_PATH_CH = np.array([2, 0, 6, 4, 10, 12, 8, 14], np.int64)

SGM_P1 = 2.3        # process_functional.py:1141
SGM_P2 = 55.9       # :1142
SGM_THRESHOLD = 30  # :1143
SGM_LAMBDA = 4      # :1144
COST_FILL = 1.0     # host-initialised volumes, :1111-1112


# --------------------------------------------------------------------------- cost volume
@njit(parallel=True, cache=True)
def cost_volume(featuresl, featuresr, ndisp):
    """compute_cost_volume_kernel, process_functional.py:120-131 (+ np.ones fill :1111-1112).

    temp = 0.0 is fp64; each product is an fp32 multiply widened and added in fp64,
    i = 0..F-1 in order; the store rounds -temp to fp32. Entries the kernel never
    writes keep 1.0. CR[y, x-d, d] = CL[y, x, d].
    """
    rows, cols, nf = featuresl.shape
    cl = np.ones((rows, cols, ndisp), np.float32)
    cr = np.ones((rows, cols, ndisp), np.float32)
    for y in prange(rows):
        for x in range(cols):
            for d in range(ndisp):
                if x - d >= 0:
                    temp = 0.0
                    for i in range(nf):
                        temp += featuresl[y, x, i] * featuresr[y, x - d, i]
                    cl[y, x, d] = -temp
                    cr[y, x - d, d] = -temp
    return cl, cr


def cost_volume_cpu_reference(featuresl, featuresr, ndisp):
    """The reference's NumPy CPU path, compute_cost_volume (process_functional.py:48-73),
    minus its per-disparity print: layout [D,H,W], fp32 np.sum, invalid entries 0, negated."""
    height, width = featuresl.shape[:2]
    vol = np.zeros([ndisp, height, width], dtype=np.float32)
    xl, xr = featuresl, featuresr
    for d in range(ndisp):
        if xl.shape[1] == 0:
            break
        vol[d, :, d:] = np.sum(np.multiply(xl, xr), axis=-1)
        xl = xl[:, 1:]
        xr = xr[:, :xr.shape[1] - 1]
    for d in range(ndisp - 1, 0, -1):
        vol[d:ndisp, :, :d - 1] = 0
    return -1 * vol


# --------------------------------------------------------------------------- penalties
@njit(cache=True)
def _pen(img, y, x, yn, xn, inside, P1, P2, p1, p2, threshold):
    """One (P1,P2) pair of sgm_penelty_kernel (:144-262): uint8 - uint8 is typed uint64 by
    Numba and wraps, so the ``-diff if diff < 0`` branch is dead (App. A3)."""
    if inside:
        diff = img[yn, xn] - img[y, x]
        diff = -diff if diff < 0 else diff
        if diff > threshold:
            return p1, p2
    return P1, P2


@njit(parallel=True, cache=True)
def sgm_penalties(image, P1=SGM_P1, P2=SGM_P2, threshold=SGM_THRESHOLD, lamda=SGM_LAMBDA):
    """sgm_penelty_kernel for one image, process_functional.py:134-262 -> f32 [H,W,16].

    Channels 0,1 are never written (stay 0). Channels 2,3 are written for (y-1,x) and
    then overwritten for (y+1,x) (:144-172)."""
    rows, cols = image.shape
    out = np.zeros((rows, cols, 16), np.float32)
    p1 = P1 / lamda
    p2 = P2 / lamda
    for y in prange(rows):
        for x in range(cols):
            a, b = _pen(image, y, x, y - 1, x, y - 1 >= 0, P1, P2, p1, p2, threshold)
            out[y, x, 2] = a
            out[y, x, 3] = b
            a, b = _pen(image, y, x, y + 1, x, y + 1 < rows, P1, P2, p1, p2, threshold)
            out[y, x, 2] = a
            out[y, x, 3] = b
            a, b = _pen(image, y, x, y, x - 1, x - 1 >= 0, P1, P2, p1, p2, threshold)
            out[y, x, 4] = a
            out[y, x, 5] = b
            a, b = _pen(image, y, x, y, x + 1, x + 1 < cols, P1, P2, p1, p2, threshold)
            out[y, x, 6] = a
            out[y, x, 7] = b
            a, b = _pen(image, y, x, y + 1, x - 1, (y + 1 < rows) and (x - 1 >= 0), P1, P2, p1, p2, threshold)
            out[y, x, 8] = a
            out[y, x, 9] = b
            a, b = _pen(image, y, x, y + 1, x + 1, (y + 1 < rows) and (x + 1 < cols), P1, P2, p1, p2, threshold)
            out[y, x, 10] = a
            out[y, x, 11] = b
            a, b = _pen(image, y, x, y - 1, x + 1, (y - 1 >= 0) and (x + 1 < cols), P1, P2, p1, p2, threshold)
            out[y, x, 12] = a
            out[y, x, 13] = b
            a, b = _pen(image, y, x, y - 1, x - 1, (y - 1 >= 0) and (x - 1 >= 0), P1, P2, p1, p2, threshold)
            out[y, x, 14] = a
            out[y, x, 15] = b
    return out


def edge_is_full_penalty(i_prev: np.ndarray, i_cur: np.ndarray, threshold: int = SGM_THRESHOLD) -> np.ndarray:
    """Closed form of the rule above for a path step prev -> cur: full (P1,P2) iff
    0 <= I[cur] - I[prev] <= threshold as integers, reduced pair otherwise (App. A3)."""
    d = i_cur.astype(np.int64) - i_prev.astype(np.int64)
    return (d >= 0) & (d <= threshold)


# --------------------------------------------------------------------------- SGM
@njit(cache=True)
def _sgm_step(cvol, svol, row, col, ndisp, P1, P2, old, new, min_cost, min_cost_P2, restart, cal_min):
    """SGM_Interation, process_functional.py:265-343, for all ndisp disparities of one pixel.

    ``old``/``new`` are fp64 (the kernel's state starts from the literals 1.0, :359-364);
    P1 is the fp64 unification of int 0 and an fp32 load, P2 an fp32 load; the S update
    is fp32 += fp64, i.e. S = fp32(fp64(S) + c). The lane-edge clamps (:300-303) become
    d == 0 and d == ndisp-1.
    """
    if restart:
        for d in range(ndisp):
            c = 0.0
            c = cvol[row, col, d]
            new[d] = c
    else:
        for d in range(ndisp):
            c = 0.0
            c = cvol[row, col, d]
            pre = old[d - 1] if d > 0 else old[0]
            nxt = old[d + 1] if d < ndisp - 1 else old[ndisp - 1]
            m1 = min(pre + P1, old[d])
            m2 = min(nxt + P1, min_cost_P2)
            c += (min(m1, m2) - min_cost)
            new[d] = c
    for d in range(ndisp):
        svol[row, col, d] += new[d]
    if cal_min:
        m = new[0]
        for d in range(1, ndisp):
            m = min(m, new[d])
        min_cost = m
        min_cost_P2 = min_cost + P2
    return min_cost, min_cost_P2


@njit(parallel=True, cache=True)
def sgm_path(cvol, svol, pen, path):
    """One of the 8 path kernels (process_functional.py:346-797) on one volume, in place on svol.

    A scanline starts at row 0 (dy=+1), row rows-1 (dy=-1) or, for horizontal paths, at
    column 0 / cols-1, visits max_iter = rows-1 (resp. cols-1) pixels (first step, then
    ``range(1, max_iter-1)``, then a last step without the min: the last pixel of every
    scanline is skipped), and diagonal paths wrap around the columns with a restart
    (is_copy, :568-572). P1 is read from the previous pixel's entry, P2 from the current
    pixel's entry (used at the next step).
    """
    rows, cols, ndisp = cvol.shape
    dy, dx, ch = _PATH_DY[path], _PATH_DX[path], _PATH_CH[path]
    horizontal = dy == 0
    nlines = rows if horizontal else cols
    max_iter = (cols - 1) if horizontal else (rows - 1)
    for line in prange(nlines):
        old = np.ones(ndisp, np.float64)
        new = np.ones(ndisp, np.float64)
        min_cost = 1.0
        min_cost_P2 = 1.0
        if horizontal:
            row = np.int64(line)
            col = np.int64(0) if dx > 0 else np.int64(cols - 1)
        else:
            row = np.int64(0) if dy > 0 else np.int64(rows - 1)
            col = np.int64(line)
        for it in range(max_iter):
            restart = it == 0
            if it > 0:
                row += dy
                col += dx
                if col >= cols:
                    col = 0
                    restart = True
                if col < 0:
                    col = cols - 1
                    restart = True
            prow, pcol = row - dy, col - dx
            P1 = 0.0
            if prow >= 0 and prow < rows and pcol >= 0 and pcol < cols:
                P1 = pen[prow, pcol, ch]
            P2 = pen[row, col, ch + 1]
            min_cost, min_cost_P2 = _sgm_step(cvol, svol, row, col, ndisp, P1, P2, old, new,
                                              min_cost, min_cost_P2, restart, it < max_iter - 1)
            tmp = old
            old = new
            new = tmp
    return svol


def sgm_all_paths(cl, cr, pen_l, pen_r, keep_each=False):
    """The 8 launches of disparity_compute_by_gpu (:1166-1202); S volumes start at 0 (:1116-1117)."""
    sl = np.zeros_like(cl)
    sr = np.zeros_like(cr)
    each = []
    for p in range(8):
        sgm_path(cl, sl, pen_l, p)
        sgm_path(cr, sr, pen_r, p)
        if keep_each:
            each.append((sl.copy(), sr.copy()))
    return (sl, sr, each) if keep_each else (sl, sr)


# --------------------------------------------------------------------------- SGM, "fused" arithmetic (MCCNN_SGM_FUSED)
# Not a reference mode: the opt-in mode of this repo that gives up the reference's fp64 path state and its per-path fp32
# rounding ORDER of S to run the 8 paths in 4 sweeps (csrc/sgm_fused.cu). Same recurrence, same traversal extents, wraps,
# restarts and penalty rule as above (:265-343, :346-797); what changes is (a) the path state, the minimum and the
# hand-over min + P2 are fp32, (b) the contributions are added into the fp32 S in the order of FUSED_PATH_ORDER, each add
# rounded once. The kernels follow exactly this arithmetic, so they are compared with it value for value; how far the mode
# is from the reference (exact) mode is measured separately (tools/fused_census.py).
FUSED_PATH_ORDER = (0, 4, 1, 3, 6, 2, 5, 7)   # down, down-right, up (raw cost), left, down-left, right, up-right, up-left


@njit(cache=True)
def _sgm_step_f32(cvol, svol, row, col, ndisp, P1, P2, old, new, min_cost, min_cost_P2, restart):
    inf = np.float32(np.inf)
    if restart:
        for d in range(ndisp):
            new[d] = cvol[row, col, d]
    else:
        for d in range(ndisp):
            pre = old[d - 1] if d > 0 else inf
            nxt = old[d + 1] if d < ndisp - 1 else inf
            a = min(pre, nxt) + P1
            b = min(old[d], min_cost_P2)
            new[d] = cvol[row, col, d] + (min(a, b) - min_cost)
    for d in range(ndisp):
        svol[row, col, d] = svol[row, col, d] + new[d]
    m = new[0]
    for d in range(1, ndisp):
        m = min(m, new[d])
    m = m + np.float32(0.0)
    return m, m + P2


@njit(parallel=True, cache=True)
def sgm_path_f32(cvol, svol, pen, path):
    """sgm_path with fp32 state (the fused mode's arithmetic); the "up" path (P1 = P2 = 0) is the plain add it reduces to."""
    rows, cols, ndisp = cvol.shape
    dy, dx, ch = _PATH_DY[path], _PATH_DX[path], _PATH_CH[path]
    horizontal = dy == 0
    nlines = rows if horizontal else cols
    max_iter = (cols - 1) if horizontal else (rows - 1)
    if path == 1:
        for row in prange(1, rows):
            for col in range(cols):
                for d in range(ndisp):
                    svol[row, col, d] = svol[row, col, d] + cvol[row, col, d]
        return svol
    for line in prange(nlines):
        old = np.ones(ndisp, np.float32)
        new = np.ones(ndisp, np.float32)
        min_cost = np.float32(1.0)
        min_cost_P2 = np.float32(1.0)
        if horizontal:
            row = np.int64(line)
            col = np.int64(0) if dx > 0 else np.int64(cols - 1)
        else:
            row = np.int64(0) if dy > 0 else np.int64(rows - 1)
            col = np.int64(line)
        for it in range(max_iter):
            restart = it == 0
            if it > 0:
                row += dy
                col += dx
                if col >= cols:
                    col = 0
                    restart = True
                if col < 0:
                    col = cols - 1
                    restart = True
            prow, pcol = row - dy, col - dx
            P1 = np.float32(0.0)
            if prow >= 0 and prow < rows and pcol >= 0 and pcol < cols:
                P1 = pen[prow, pcol, ch]
            P2 = pen[row, col, ch + 1]
            min_cost, min_cost_P2 = _sgm_step_f32(cvol, svol, row, col, ndisp, P1, P2, old, new, min_cost, min_cost_P2, restart)
            tmp = old
            old = new
            new = tmp
    return svol


def sgm_all_paths_fused(cl, cr, pen_l, pen_r):
    """The fused mode's S volumes: fp32 state, contributions added in FUSED_PATH_ORDER."""
    sl = np.zeros_like(cl)
    sr = np.zeros_like(cr)
    for p in FUSED_PATH_ORDER:
        sgm_path_f32(cl, sl, pen_l, p)
        sgm_path_f32(cr, sr, pen_r, p)
    return sl, sr


# --------------------------------------------------------------------------- WTA
@njit(parallel=True, cache=True)
def wta(svol):
    """WTA_and_SupixelRefinement_kernel, :800-837: first strict minimum over d, stored as f32."""
    rows, cols, ndisp = svol.shape
    out = np.zeros((rows, cols), np.float32)
    for y in prange(rows):
        for x in range(cols):
            min_s = svol[y, x, 0]
            index = 0
            for i in range(1, ndisp):
                tmp = svol[y, x, i]
                if min_s > tmp:
                    min_s = tmp
                    index = i
            out[y, x] = index
    return out


@njit(parallel=True, cache=True)
def wta_subpixel(svol):
    """WTA + the parabola refinement the reference left commented out (:813-819): "parity unpinned". Typing as Numba
    would infer it: fp32 differences, `2 * (...)` promotes the denominator to fp64. Own definition where the
    reference is silent: border indices (the reference tests 0 < index < 127) and a flat parabola keep the index."""
    rows, cols, ndisp = svol.shape
    out = np.zeros((rows, cols), np.float32)
    for y in prange(rows):
        for x in range(cols):
            min_s = svol[y, x, 0]
            index = 0
            for i in range(1, ndisp):
                tmp = svol[y, x, i]
                if min_s > tmp:
                    min_s = tmp
                    index = i
            min_index = np.float64(index)
            if index > 0 and index < ndisp - 1:
                c = svol[y, x, index]
                _c = svol[y, x, index - 1]
                c_ = svol[y, x, index + 1]
                den = 2 * (_c + c_ - 2 * c)
                if den > 0:
                    min_index = index - (c_ - _c) / den
            out[y, x] = min_index
    return out


@njit(parallel=True, cache=True)
def wta_dhw(vol_dhw):
    """CPU WTA1, process_functional.py:96-113: argmin with strict <, layout [D,H,W]."""
    ndisp, rows, cols = vol_dhw.shape
    out = np.zeros((rows, cols), np.float32)
    for y in prange(rows):
        for x in range(cols):
            min_cost = np.inf
            idx = -1
            for d in range(ndisp):
                if vol_dhw[d, y, x] < min_cost:
                    min_cost = vol_dhw[d, y, x]
                    idx = d
            out[y, x] = idx
    return out


# --------------------------------------------------------------------------- L-R check / fill / median
@njit(cache=True)
def lr_flags(dl, dr, wrap_uint8=False):
    """is_error_match_kernel, :977-1000. The reference indexes with ``uint8(x - ld)``.
    SURVEY.md App. A5 predicted a modulo-256 wrap from CPU njit behaviour; the reference
    kernel actually run on a B200 (tests/golden/ref_wide_9x300.npz, W=300) shows NO wrap:
    the un-truncated column is read. The oracle follows the hardware; wrap_uint8=True
    keeps the CPU-njit behaviour for the record only."""
    rows, cols = dl.shape
    fl = np.zeros((rows, cols), np.uint8)
    fr = np.zeros((rows, cols), np.uint8)
    for y in range(rows):
        for x in range(cols):
            ld = dl[y, x]
            rd = x - ld
            if rd >= 0:
                idx = np.int64(rd)
                if wrap_uint8:
                    idx = idx % 256
                if idx < cols:
                    rdv = dr[y, idx]
                    minus = ld - rdv
                    fl[y, x] = 1 if (minus > 1 or minus < -1) else 0
            rdv = dr[y, x]
            ldx = x + rdv
            if ldx < cols:
                idx = np.int64(ldx)
                if wrap_uint8:
                    idx = idx % 256
                ldv = dl[y, idx]
                minus = rdv - ldv
                fr[y, x] = 1 if (minus > 1 or minus < -1) else 0
    return fl, fr


@njit(cache=True)
def lrc_fill(dl, flag_l):
    """LRC_kernel, :1003-1088 (left map only; the right output is never written, App. A6)."""
    rows, cols = dl.shape
    out = np.zeros((rows, cols), np.float32)
    for y in range(rows):
        for x in range(cols):
            if flag_l[y, x] == 1:
                number = 0
                sum_d = 0
                idy = y
                while idy >= 0 and flag_l[idy, x] == 1:
                    idy -= 1
                if idy >= 0:
                    number += 1
                    sum_d += dl[idy, x]
                idy = y
                while idy < rows and flag_l[idy, x] == 1:
                    idy += 1
                if idy < rows:
                    number += 1
                    sum_d += dl[idy, x]
                idx = x
                while idx < cols and flag_l[y, idx] == 1:
                    idx += 1
                if idx < cols:
                    number += 1
                    sum_d += dl[y, idx]
                idx = x
                while idx >= 0 and flag_l[y, idx] == 1:
                    idx -= 1
                if idx >= 0:
                    number += 1
                    sum_d += dl[y, idx]
                if number > 0:
                    out[y, x] = sum_d / number
                else:
                    out[y, x] = dl[y, x]
            else:
                out[y, x] = dl[y, x]
    return out


@njit(parallel=True, cache=True)
def median5(filled, wta_map):
    """Median_Filter_kernel, :840-879, launched as (filled -> WTA buffer) at :1250: interior
    pixels get the 13th smallest of the 5x5 window, the 2-pixel border keeps the raw WTA map."""
    rows, cols = filled.shape
    out = wta_map.copy()
    for idy in prange(2, rows - 2):
        win = np.zeros(25, np.float32)
        for idx in range(2, cols - 2):
            for i in range(-2, 3):
                for j in range(-2, 3):
                    win[(i + 2) * 5 + j + 2] = filled[idy + i, idx + j]
            current_min = np.float32(0.0)
            for i in range(13):
                current_min = win[i]
                k = i
                for j in range(i + 1, 25):
                    if current_min > win[j]:
                        current_min = win[j]
                        k = j
                win[k] = win[i]
            out[idy, idx] = current_min
    return out


_BILATERAL_W = np.array([0.167747, 0.165145, 0.157581, 0.145735, 0.130632,
                         0.113490, 0.095563, 0.077991, 0.061692, 0.047297], np.float32)


@njit(parallel=True, cache=True)
def bilateral9(image, disp, weights=_BILATERAL_W):
    """Bilateral_Filter_kernel, :882-974 (its launch is commented out at :1260, so this
    stage is "parity unpinned"). 9x9 window, zero padding outside the image, range-only
    weights; ``current - neighbour`` is uint8 - uint8 = wrapping uint64 (App. A7), so only
    neighbours with 0 <= Ic - In < 5 contribute. Sums are fp64, weights/disparities fp32."""
    rows, cols = image.shape
    out = np.zeros((rows, cols), np.float32)
    for y in prange(rows):
        for x in range(cols):
            weights_sum = 0.0
            disparity_sum = 0.0
            ic = image[y, x]
            for i in range(-4, 5):
                for j in range(-4, 5):
                    yy, xx = y + i, x + j
                    inside = yy >= 0 and yy < rows and xx >= 0 and xx < cols
                    tmp_intensity = image[yy, xx] if inside else np.uint8(0)
                    tmp_disparity = disp[yy, xx] if inside else np.float32(0.0)
                    minus = ic - tmp_intensity
                    absolute_minus = minus if minus >= 0 else -minus
                    if absolute_minus < 5:
                        weight = weights[absolute_minus]
                        weights_sum += weight
                        disparity_sum += (weight * tmp_disparity)
            out[y, x] = disparity_sum / weights_sum
    return out


# --------------------------------------------------------------------------- orchestration
# --------------------------------------------------------------------------- cross-based aggregation
# north_star names this stage; the reference has no code for it (only the label at match.py:98 and the
# parameter name at process_functional.py:347) => PARITY UNPINNED. Definition after the MC-CNN paper
# (Zbontar & LeCun 2016, sec. 5.1): arm rules, intersection of the supports of the pixel and of its match,
# mean over the support, rows first and columns second; stated in full in csrc/cbca.cu.
CBCA_L1 = 14   # maximum arm length (own default)
CBCA_TAU = 6   # intensity threshold on u8 grey levels (own default)


@njit(parallel=True, cache=True)
def cross_arms(image, L1=CBCA_L1, tau=CBCA_TAU):
    """u8 [H,W] -> u8 [H,W,4]: distance to the first excluded position to the left, right, up, down."""
    H, W = image.shape
    arms = np.zeros((H, W, 4), np.uint8)
    dxs = np.array([-1, 1, 0, 0])
    dys = np.array([0, 0, -1, 1])
    for y in prange(H):
        for x in range(W):
            c = np.int64(image[y, x])
            for k4 in range(4):
                k = 1
                while True:
                    xx = x + dxs[k4] * k
                    yy = y + dys[k4] * k
                    if xx < 0 or xx >= W or yy < 0 or yy >= H:
                        break
                    if k > 1:
                        if abs(np.int64(image[yy, xx]) - c) >= tau:
                            break
                        if k >= L1:
                            break
                    k += 1
                arms[y, x, k4] = k
    return arms


@njit(parallel=True, cache=True)
def cbca_iteration(vol, arms_self, arms_other, direction):
    """One aggregation pass of vol [H,W,D] (direction -1: left volume, other pixel x-d; +1: right volume).
    Row sums are rounded to fp32 (they are an fp32 volume on the device), the column sum and the mean are fp64,
    the result is rounded once to fp32."""
    H, W, D = vol.shape
    rows = np.zeros((H, W, D), np.float32)
    cnts = np.zeros((H, W, D), np.int64)
    for y in prange(H):
        for x in range(W):
            for d in range(D):
                xo = x + direction * d
                if xo < 0 or xo >= W:
                    continue
                lo = x - min(arms_self[y, x, 0], arms_other[y, xo, 0])
                hi = x + min(arms_self[y, x, 1], arms_other[y, xo, 1])
                acc = 0.0
                for xx in range(lo + 1, hi):
                    acc += np.float64(vol[y, xx, d])
                rows[y, x, d] = np.float32(acc)
                cnts[y, x, d] = hi - lo - 1
    out = np.empty((H, W, D), np.float32)
    for y in prange(H):
        for x in range(W):
            for d in range(D):
                xo = x + direction * d
                if xo < 0 or xo >= W:
                    out[y, x, d] = vol[y, x, d]
                    continue
                lo = y - min(arms_self[y, x, 2], arms_other[y, xo, 2])
                hi = y + min(arms_self[y, x, 3], arms_other[y, xo, 3])
                acc = 0.0
                n = 0
                for yy in range(lo + 1, hi):
                    acc += np.float64(rows[yy, x, d])
                    n += cnts[yy, x, d]
                out[y, x, d] = np.float32(acc / n)
    return out


def cbca(cl, cr, imagel, imager, iters=2, L1=CBCA_L1, tau=CBCA_TAU):
    """`iters` aggregation passes of both volumes."""
    al, ar = cross_arms(imagel, L1, tau), cross_arms(imager, L1, tau)
    for _ in range(iters):
        cl = cbca_iteration(cl, al, ar, -1)
        cr = cbca_iteration(cr, ar, al, 1)
    return cl, cr


def disparity_pipeline(imagel, imager, featuresl, featuresr, ndisp=128, keep=False):
    """disparity_compute_by_gpu, process_functional.py:1093-1267, on the CPU.

    Returns (left disparity after median, right raw-WTA disparity) and, with keep=True, a
    dict of every intermediate. The reference's returned right map is the median of an
    uninitialised buffer in the interior (App. A6); the defined part is the raw WTA map.
    """
    assert imagel.shape == imager.shape
    rows, cols = imagel.shape
    assert rows >= 3 and cols >= 3
    cl, cr = cost_volume(np.ascontiguousarray(featuresl), np.ascontiguousarray(featuresr), ndisp)
    pl = sgm_penalties(np.ascontiguousarray(imagel))
    pr = sgm_penalties(np.ascontiguousarray(imager))
    sl, sr = sgm_all_paths(cl, cr, pl, pr)
    dl, dr = wta(sl), wta(sr)
    fll, flr = lr_flags(dl, dr)
    filled = lrc_fill(dl, fll)
    final = median5(filled, dl)
    if keep:
        return final, dr, dict(CL=cl, CR=cr, PL=pl, PR=pr, SL=sl, SR=sr, dl_wta=dl, dr_wta=dr,
                               flag_l=fll, flag_r=flr, dl_fill=filled, dl_final=final)
    return final, dr


def bad_pixel_rate(disp_u8, true_disp_fullres, resize=True):
    """error_calculate.py:58-83: GT resized to the result size and halved; a pixel is bad iff
    GT is finite and non-zero and |disp - gt| > 1; the rate divides by ALL H*W pixels."""
    height, width = disp_u8.shape[0:2]
    gt = true_disp_fullres
    if resize and gt.shape[:2] != (height, width):
        import cv2

        gt = cv2.resize(gt, (width, height))
    gt = gt / 2
    valid = ~((gt == np.inf) | (gt == 0.0))
    bad = valid & (np.fabs(disp_u8.astype(np.float32) - gt) > 1)
    return float(bad.sum()) / float(height * width)
