"""CPU restatement of Telea's fast-marching inpainting as OpenCV implements it - TEST INFRASTRUCTURE ONLY.

utils.inpaint (/root/reference/utils.py:136-151) calls cv2.inpaint(img_u8, mask, 3, cv2.INPAINT_TELEA); the algorithm lives in a
third-party dependency (opencv-python, unpinned by the reference; 4.13 in this image; opencv/modules/photo/src/inpaint.cpp), so it
is restated here from its published form and pinned against cv2.inpaint itself:

  * order="heap":  OpenCV's own order (a stable priority queue on T) - must reproduce cv2.inpaint BIT FOR BIT, which pins the
                   per-pixel arithmetic (tests/test_oracle_cpu.py::test_telea_heap_order_restatement_equals_cv2);
  * order="layer": the same arithmetic with the march advanced in layers (all hole pixels that have a 4-neighbour in an earlier
                   layer at once, every read seeing the state at the layer's start) - the order the CUDA kernel
                   (csrc/ofd_inpaint.cu) uses, so the kernel is checked bit for bit against this mode, and the difference between
                   the two modes is the (measured, reported) price of parallelism.

Pure-Python loops: small frames only.
"""
from __future__ import annotations

import heapq
import math

import numpy as np

KNOWN, BAND, INSIDE = 0, 1, 2
F = np.float32


def _solve(t, f_known, i1, j1, i2, j2):
    """inpaint.cpp FastMarching_solve (double arithmetic, float result)."""
    a11, a22 = float(t[i1, j1]), float(t[i2, j2])
    m12 = min(a11, a22)
    if f_known(i1, j1):
        if f_known(i2, j2):
            if abs(a11 - a22) >= 1.0:
                sol = 1 + m12
            else:
                sol = (a11 + a22 + math.sqrt(2 - (a11 - a22) * (a11 - a22))) * 0.5
        else:
            sol = 1 + a11
    elif f_known(i2, j2):
        sol = 1 + a22
    else:
        sol = 1 + m12
    return F(sol)


def _min4_solve(t, f_known, i, j):
    return min(_solve(t, f_known, i - 1, j, i, j - 1), _solve(t, f_known, i + 1, j, i, j - 1),
               _solve(t, f_known, i - 1, j, i, j + 1), _solve(t, f_known, i + 1, j, i, j + 1))


def _setup(mask, rng):
    H, W = mask.shape
    EH, EW = H + 2, W + 2
    m = np.zeros((EH, EW), bool)
    m[1:-1, 1:-1] = mask != 0
    cross = m.copy()
    cross[1:, :] |= m[:-1, :]
    cross[:-1, :] |= m[1:, :]
    cross[:, 1:] |= m[:, :-1]
    cross[:, :-1] |= m[:, 1:]
    band = cross & ~m
    band[0, :] = band[-1, :] = band[:, 0] = band[:, -1] = False
    rect = np.zeros_like(m)
    for di in range(-rng, rng + 1):
        for dj in range(-rng, rng + 1):
            src = m[max(0, -di):EH - max(0, di), max(0, -dj):EW - max(0, dj)]
            rect[max(0, di):EH - max(0, -di), max(0, dj):EW - max(0, -dj)] |= src
    ring = rect & ~m & ~band
    ring[0, :] = ring[-1, :] = ring[:, 0] = ring[:, -1] = False
    t = np.full((EH, EW), 1.0e6, F)
    t[band] = 0
    return m, band, ring, t


def _fill_pixel(i, j, t, dist, known, out, rng, EH, EW):
    """The colour of hole pixel (i, j) (extended coordinates), inpaint.cpp icvTeleaInpaintFMM's inner loops, float32 arithmetic in
    OpenCV's operand order.  `known(k, l)` = f != INSIDE; t[i, j] must already hold `dist`."""
    def gk(a, b):
        return known(a, b)

    if gk(i, j + 1):
        gx = F(F(t[i, j + 1] - t[i, j - 1]) * F(0.5)) if gk(i, j - 1) else F(t[i, j + 1] - dist)
    else:
        gx = F(dist - t[i, j - 1]) if gk(i, j - 1) else F(0)
    if gk(i + 1, j):
        gy = F(F(t[i + 1, j] - t[i - 1, j]) * F(0.5)) if gk(i - 1, j) else F(t[i + 1, j] - dist)
    else:
        gy = F(dist - t[i - 1, j]) if gk(i - 1, j) else F(0)
    Ia, Jx, Jy, s = [F(0)] * 3, [F(0)] * 3, [F(0)] * 3, [F(1.0e-20)] * 3
    H, W = EH - 2, EW - 2
    for k in range(i - rng, i + rng + 1):
        km, kp = k - 1 + (k == 1), k - 1 - (k == EH - 2)
        for l in range(j - rng, j + rng + 1):
            lm, lp = l - 1 + (l == 1), l - 1 - (l == EW - 2)
            if not (k > 0 and l > 0 and k < EH - 1 and l < EW - 1):
                continue
            if not gk(k, l) or (l - j) * (l - j) + (k - i) * (k - i) > rng * rng:
                continue
            ry, rx = F(i - k), F(j - l)
            len2 = F(F(rx * rx) + F(ry * ry))
            dst = F(1.0 / (float(len2) * math.sqrt(float(len2))))
            lev = F(1.0 / (1 + abs(float(F(t[k, l] - dist)))))
            dr = F(F(rx * gx) + F(ry * gy))
            if abs(dr) <= F(0.01):
                dr = F(0.000001)
            w = F(abs(F(F(dst * lev) * dr)))
            kr, kl, kd, ku = gk(k, l + 1), gk(k, l - 1), gk(k + 1, l), gk(k - 1, l)

            def px(c, r, col):
                r = min(max(r, 0), H - 1)
                col = min(max(col, 0), W - 1)
                return F(out[r, col, c])

            for c in range(3):
                if kr:
                    gix = F(F(px(c, km, lp + 1) - px(c, km, lm - 1)) * F(2.0)) if kl else F(px(c, km, lp + 1) - px(c, km, lm))
                else:
                    gix = F(px(c, km, lp) - px(c, km, lm - 1)) if kl else F(0)
                if kd:
                    giy = F(F(px(c, kp + 1, lm) - px(c, km - 1, lm)) * F(2.0)) if ku else F(px(c, kp + 1, lm) - px(c, km, lm))
                else:
                    giy = F(px(c, kp, lm) - px(c, km - 1, lm)) if ku else F(0)
                Ia[c] = F(Ia[c] + F(w * px(c, k - 1, l - 1)))  # the tap's own pixel (the km / lm shifts apply to the gradients only)
                Jx[c] = F(Jx[c] - F(w * F(gix * rx)))
                Jy[c] = F(Jy[c] - F(w * F(giy * ry)))
                s[c] = F(s[c] + w)
    res = []
    for c in range(3):
        nrm = F(np.sqrt(F(F(Jx[c] * Jx[c]) + F(Jy[c] * Jy[c]))) + F(1.0e-20))
        sat = F(F(F(Ia[c] / s[c]) + F(F(Jx[c] + Jy[c]) / nrm)) + F(0.5))
        v = int(np.rint(sat)) if np.isfinite(sat) else 0
        res.append(min(max(v, 0), 255))
    return res


def telea(img_u8, mask, radius=3, order="layer", return_stats=False):
    """cv2.inpaint(img_u8[H,W,3] uint8, mask[H,W] (!= 0 = fill), radius, cv2.INPAINT_TELEA) restated.
    order = "heap" (OpenCV's) or "layer" (the CUDA kernel's)."""
    img_u8 = np.ascontiguousarray(img_u8)
    H, W = mask.shape
    EH, EW = H + 2, W + 2
    m, band, ring, t = _setup(mask, radius)
    out = img_u8.copy()
    with np.errstate(over="ignore", invalid="ignore"):
        # ---- outward march over the ring: icvCalcFMM(out, t, Out, negate=true) -----------------------------------------------------------
        g = np.where(ring, INSIDE, KNOWN).astype(np.uint8)
        if order == "heap":
            heap, cnt = [], 0
            for i, j in zip(*np.nonzero(band)):
                heap.append((0.0, cnt, int(i), int(j)))
                cnt += 1
            heapq.heapify(heap)
            changed = np.zeros_like(m)
            while heap:
                _, _, ii, jj = heapq.heappop(heap)
                changed[ii, jj] = True
                for i, j in ((ii - 1, jj), (ii, jj - 1), (ii + 1, jj), (ii, jj + 1)):
                    if i <= 0 or j <= 0 or i > EH or j > EW or i >= EH or j >= EW:
                        continue
                    if g[i, j] == INSIDE:
                        d = _min4_solve(t, lambda a, b: g[a, b] != INSIDE, i, j)
                        t[i, j] = d
                        g[i, j] = BAND
                        heapq.heappush(heap, (float(d), cnt, i, j))
                        cnt += 1
            t[changed & ring] *= -1
        else:
            layer_of = np.where(band, 1, np.where(ring, 255, 0)).astype(np.int32)
            for k in range(1, 2 * radius + 1):
                todo = []
                for i, j in zip(*np.nonzero(layer_of == 255)):
                    if any(1 <= layer_of[a, b] <= k for a, b in ((i - 1, j), (i + 1, j), (i, j - 1), (i, j + 1))):
                        todo.append((i, j, _min4_solve(t, lambda a, b: layer_of[a, b] <= k, i, j)))
                for i, j, d in todo:
                    t[i, j] = d
                    layer_of[i, j] = k + 1
            t[(layer_of >= 2) & (layer_of != 255)] *= -1
        # ---- inward march: icvTeleaInpaintFMM ---------------------------------------------------------------------------------------------
        layers = filled = 0
        if order == "heap":
            f = np.where(m, INSIDE, np.where(band, BAND, KNOWN)).astype(np.uint8)
            known = lambda a, b: f[a, b] != INSIDE  # noqa: E731
            heap, cnt = [], 0
            for i, j in zip(*np.nonzero(band)):
                heap.append((0.0, cnt, int(i), int(j)))
                cnt += 1
            heapq.heapify(heap)
            while heap:
                _, _, ii, jj = heapq.heappop(heap)
                f[ii, jj] = KNOWN
                for i, j in ((ii - 1, jj), (ii, jj - 1), (ii + 1, jj), (ii, jj + 1)):
                    if i <= 0 or j <= 0 or i > EH - 1 or j > EW - 1:
                        continue
                    if f[i, j] == INSIDE:
                        dist = _min4_solve(t, known, i, j)
                        t[i, j] = dist
                        out[i - 1, j - 1] = _fill_pixel(i, j, t, dist, known, out, radius, EH, EW)
                        f[i, j] = BAND
                        heapq.heappush(heap, (float(dist), cnt, i, j))
                        cnt += 1
                        filled += 1
        else:
            L = np.where(m, 0xFFFF, 0).astype(np.int32)
            interior = np.zeros_like(m)
            interior[1:-1, 1:-1] = True
            frontier = [(int(i), int(j)) for i, j in zip(*np.nonzero(m))
                        if any(L[a, b] == 0 and interior[a, b] for a, b in ((i - 1, j), (i + 1, j), (i, j - 1), (i, j + 1)))]
            queued = np.zeros_like(m)
            for i, j in frontier:
                queued[i, j] = True
            layer = 1
            while frontier:
                known = lambda a, b, layer=layer: L[a, b] < layer  # noqa: E731
                staged = []
                for i, j in frontier:
                    dist = _min4_solve(t, known, i, j)
                    told = t[i, j]
                    t[i, j] = dist  # _fill_pixel reads t[i, j] only through `dist`; restore below so the layer sees its start state
                    col = _fill_pixel(i, j, t, dist, known, out, radius, EH, EW)
                    t[i, j] = told
                    staged.append((i, j, dist, col))
                nxt = []
                for i, j, dist, col in staged:
                    t[i, j] = dist
                    L[i, j] = layer
                    out[i - 1, j - 1] = col
                for i, j, _, _ in staged:
                    for a, b in ((i - 1, j), (i, j - 1), (i + 1, j), (i, j + 1)):
                        if a <= 0 or b <= 0 or a > EH - 1 or b > EW - 1:
                            continue
                        if L[a, b] == 0xFFFF and not queued[a, b]:
                            queued[a, b] = True
                            nxt.append((a, b))
                filled += len(staged)
                layers = layer
                layer += 1
                frontier = nxt
    return (out, (layers, filled)) if return_stats else out
