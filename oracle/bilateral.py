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


def rank_table(nmax):
    """k(n) = #{m <= n : float32 running sum of m copies of float32(1)/float32(n) <= 0.5}  (bilateral_filter.py:194-197)."""
    k = np.zeros(nmax + 1, np.int64)
    for n in range(1, nmax + 1):
        coef = np.ones(n, np.float32)
        coef = coef / coef.sum()
        k[n] = int(np.digitize(0.5, np.cumsum(coef)))
    return k


def discontinuity(depth, depth_orig, thr):
    """bilateral_filter.py:63-116 + :45-46 (thr compared in the array dtype)."""
    disp = 1.0 / depth
    H, W = depth.shape
    disc = np.zeros((H, W), bool)
    c = disp[1:-1, 1:-1]
    thr = depth.dtype.type(thr)
    with np.errstate(invalid="ignore"):
        disc[1:-1, 1:-1] = ((np.abs(c - disp[:-2, 1:-1]) > thr) | (np.abs(c - disp[2:, 1:-1]) > thr)
                            | (np.abs(c - disp[1:-1, :-2]) > thr) | (np.abs(c - disp[1:-1, 2:]) > thr))
    disc[depth_orig == 0] = True
    return disc


def bilateral_iter(depth, depth_orig, window, thr):
    """One iteration (bilateral_filter.py:33-58) -> new depth [H,W], same dtype."""
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


def sparse_bilateral_filtering(depth, filter_size, depth_threshold=0.04, num_iter=None):
    """bilateral_filter.py:13-60 (mask None): returns the filtered depth."""
    cur = depth.copy()
    for i in range(num_iter):
        cur = bilateral_iter(cur, depth, filter_size[i], depth_threshold)
    return cur
