"""numpy restatement of bilateral_filter.py:13-60 on the path the reference takes (mask None) — TEST INFRASTRUCTURE.

Closed form of one iteration (checked bit-identical to the reference's per-pixel Python loop by
tests/golden/make_golden.py and pinned in tests/golden/bilateral_*.npz):
  disc = interior 4-neighbour |1/d - 1/d'| > thr  (bilateral_filter.py:63-116), disc[depth_orig == 0] = 1 (:46);
  depth / disc: border ring <- edge-replicated interior, then edge pad by window//2 (:141-147);
  pixels whose window holds a discontinuity take the rank-k(n) smallest of the n non-discontinuity taps, centre if
  n == 0 (:167-198); k(n) from the float32 cumsum / digitize rule (:194-197).
"""
from __future__ import annotations

import numpy as np


def rank_table(nmax, coef_dtype=np.float32):
    """k(n) = #{m <= n : running sum (in coef_dtype) of m copies of 1/n <= 0.5}  (bilateral_filter.py:194-197).  The coefficients
    are float32 without a mask (:180) and float32 * mask.dtype with one (:182: float64 for float64 / integer masks)."""
    k = np.zeros(nmax + 1, np.int64)
    for n in range(1, nmax + 1):
        coef = np.ones(n, coef_dtype)
        coef = coef / coef.sum()
        k[n] = int(np.digitize(0.5, np.cumsum(coef)))
    return k


def discontinuity(depth, depth_orig, thr, mask=None):
    """bilateral_filter.py:63-116 + :45-49 (thr compared in the array dtype).  mask (binary): a neighbour difference only counts
    when both pixels are unmasked (:72-80), and masked pixels are never discontinuities (:48-49, after the depth == 0 rule)."""
    disp = 1.0 / depth
    H, W = depth.shape
    disc = np.zeros((H, W), bool)
    c = disp[1:-1, 1:-1]
    thr = depth.dtype.type(thr)
    with np.errstate(invalid="ignore"):
        tests = [np.abs(c - disp[:-2, 1:-1]) > thr, np.abs(c - disp[2:, 1:-1]) > thr,
                 np.abs(c - disp[1:-1, :-2]) > thr, np.abs(c - disp[1:-1, 2:]) > thr]
    if mask is not None:
        mk = mask != 0
        mc = mk[1:-1, 1:-1]
        pairs = [mc & mk[:-2, 1:-1], mc & mk[2:, 1:-1], mc & mk[1:-1, :-2], mc & mk[1:-1, 2:]]
        tests = [t & p for t, p in zip(tests, pairs)]
    disc[1:-1, 1:-1] = tests[0] | tests[1] | tests[2] | tests[3]
    disc[depth_orig == 0] = True
    if mask is not None:
        disc[mask == 0] = False
    return disc


def bilateral_iter(depth, depth_orig, window, thr, mask=None):
    """One iteration (bilateral_filter.py:33-58) -> new depth [H,W], same dtype.  mask: BINARY (0 / non-zero... the weighted case of
    a fractional mask is not restated); masked pixels keep their (ring-replicated) depth (:169-170), masked taps - and taps outside
    the image, the mask being zero-padded (:161) - are left out of the median (:180-182)."""
    if mask is not None:
        return _bilateral_iter_masked(depth, depth_orig, window, thr, mask)
    with np.errstate(divide="ignore"):
        disc = discontinuity(depth, depth_orig, thr)
    m = window // 2
    d = np.pad(depth[1:-1, 1:-1], 1, "edge")
    q = np.pad(disc[1:-1, 1:-1], 1, "edge")
    pd = np.pad(d, m, "edge")
    pq = np.pad(q, m, "edge")
    H, W = depth.shape
    out = d.copy()
    win_d = np.lib.stride_tricks.sliding_window_view(pd, (window, window)).reshape(H, W, -1)
    win_q = np.lib.stride_tricks.sliding_window_view(pq, (window, window)).reshape(H, W, -1)
    active = win_q.any(-1)
    ktab = rank_table(window * window)
    rr, cc = np.nonzero(active)
    if rr.size:
        vals = win_d[rr, cc].astype(depth.dtype, copy=True)
        bad = win_q[rr, cc]
        n = (~bad).sum(-1)
        big = np.array(np.inf, depth.dtype)
        vals[bad] = big
        vals.sort(-1)
        k = ktab[np.maximum(n, 1)]
        sel = vals[np.arange(rr.size), k]
        sel = np.where(n == 0, d[rr, cc], sel)
        out[rr, cc] = sel
    return out


def _bilateral_iter_masked(depth, depth_orig, window, thr, mask):
    with np.errstate(divide="ignore"):
        disc = discontinuity(depth, depth_orig, thr, mask)
    m = window // 2
    d = np.pad(depth[1:-1, 1:-1], 1, "edge")
    q = np.pad(disc[1:-1, 1:-1], 1, "edge")
    pd = np.pad(d, m, "edge")
    pq = np.pad(q, m, "edge")
    pm = np.pad(mask != 0, m, "constant")
    H, W = depth.shape
    out = d.copy()
    win_d = np.lib.stride_tricks.sliding_window_view(pd, (window, window)).reshape(H, W, -1)
    win_q = np.lib.stride_tricks.sliding_window_view(pq, (window, window)).reshape(H, W, -1)
    win_m = np.lib.stride_tricks.sliding_window_view(pm, (window, window)).reshape(H, W, -1)
    active = win_q.any(-1) & (mask != 0)
    coef_dtype = np.result_type(np.float32, np.asarray(mask).dtype)
    ktab = rank_table(window * window, coef_dtype)
    rr, cc = np.nonzero(active)
    if rr.size:
        vals = win_d[rr, cc].astype(depth.dtype, copy=True)
        bad = win_q[rr, cc] | ~win_m[rr, cc]
        n = (~bad).sum(-1)
        vals[bad] = np.array(np.inf, depth.dtype)
        vals.sort(-1)
        sel = vals[np.arange(rr.size), ktab[np.maximum(n, 1)]]
        out[rr, cc] = np.where(n == 0, d[rr, cc], sel)
    return out


def sparse_bilateral_filtering(depth, filter_size, depth_threshold=0.04, num_iter=None, mask=None):
    """bilateral_filter.py:13-60: returns the filtered depth (mask None, or a binary mask)."""
    cur = depth.copy()
    for i in range(num_iter):
        cur = bilateral_iter(cur, depth, filter_size[i], depth_threshold, mask)
    return cur
