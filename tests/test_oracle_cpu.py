"""The oracle against the golden vectors generated from the reference's own Python (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import bilateral as obil
from oracle import flow as oflow

FW_TAGS = ["A", "B", "C", "D", "E", "F"]


@pytest.mark.parametrize("tag", FW_TAGS)
def test_fw_prologue_and_splat_match_reference(golden, tag):
    g = golden("fw_cases")
    obj, flow, depth = g[f"{tag}_obj"], g[f"{tag}_flow"], g[f"{tag}_depth"]
    sx, sy = oracle.fw_targets(flow)
    assert np.array_equal(sx, g[f"{tag}_safe_x"]) and np.array_equal(sy, g[f"{tag}_safe_y"])
    out, valid, coll, rc = oracle.splat_literal(obj[None], sy[None, None], sx[None, None], depth[None])
    assert rc == 0
    assert np.array_equal(out[0], g[f"{tag}_out"], equal_nan=True)
    assert np.array_equal(valid[0], g[f"{tag}_valid"]) and np.array_equal(coll[0], g[f"{tag}_coll"])
    out2, valid2, coll2, winner, dropped = oracle.fw_forward(obj, flow, depth)
    assert dropped == 0
    assert np.array_equal(out2, g[f"{tag}_out"], equal_nan=True)
    assert np.array_equal(valid2, g[f"{tag}_valid"]) and np.array_equal(coll2, g[f"{tag}_coll"])
    assert np.array_equal(winner, g[f"{tag}_winner"])


def _lexsort_splat(safe_x, safe_y, depth, H, W):
    """Closed form of SURVEY.md section 8c: sort sources by (target, depth, raster id); first of a group wins if depth < 1000."""
    t = (safe_y.astype(np.int64) * W + safe_x.astype(np.int64)).ravel()
    d = depth.ravel().astype(np.float32)
    ok = d < 1000  # NaN -> False
    winner = np.full(H * W, -1, np.int64)
    hit = np.zeros(H * W, bool)
    hit[t] = True
    src = np.arange(H * W)
    order = np.lexsort((src[ok], d[ok], t[ok]))
    ts, ss = t[ok][order], src[ok][order]
    first = np.ones(ts.size, bool)
    first[1:] = ts[1:] != ts[:-1]
    winner[ts[first]] = ss[first]
    winner[hit & (winner < 0)] = -2
    return winner.reshape(H, W)


@pytest.mark.parametrize("seed", range(12))
def test_literal_loop_equals_shared_lut_and_closed_form(seed):
    rng = np.random.default_rng(seed)
    H, W, C = int(rng.integers(1, 40)), int(rng.integers(1, 40)), int(rng.integers(1, 8))
    obj = rng.normal(0, 1, (C, H, W)).astype(np.float32)
    flow = rng.normal(0, rng.uniform(0.5, 30), (2, H, W)).astype(np.float32)
    depth = rng.integers(1, 4, (1, H, W)).astype(np.float32)  # heavy ties
    depth[rng.random((1, H, W)) < 0.05] = 1000.0
    depth[rng.random((1, H, W)) < 0.02] = np.nan
    depth[rng.random((1, H, W)) < 0.02] = -0.0
    sx, sy = oracle.fw_targets(flow)
    o1, v1, c1, rc = oracle.splat_literal(obj[None], sy[None, None], sx[None, None], depth[None])
    o2, v2, c2, win, _ = oracle.splat_frame(obj, sy, sx, depth)
    assert rc == 0
    assert np.array_equal(o1[0], o2) and np.array_equal(v1[0], v2) and np.array_equal(c1[0], c2)
    assert np.array_equal(win, _lexsort_splat(sx, sy, depth, H, W))


def test_convert_oracle_matches_reference(golden):
    g = golden("convert_cases")
    for tag in ("f32", "f64"):
        out = oflow.normalize_depth(torch.from_numpy(g[f"norm_{tag}_in"].copy())[None]).numpy()
        assert np.array_equal(out, g[f"norm_{tag}_out"])
        flow = oflow.disparity_flow(torch.from_numpy(out), torch.tensor(g[f"disp_{tag}_sBf"]))
        assert np.array_equal(flow.numpy(), g[f"disp_{tag}_flow"])
    assert np.array_equal(oflow.fix_warped_depth(torch.from_numpy(g["fix_in"])).numpy(), g["fix_out"])


@pytest.mark.parametrize("tag", ["f32", "f64", "f32b"])
def test_reproject_oracle_matches_reference(golden, tag):
    g = golden("reproject_cases")
    torch.set_num_threads(1)
    flow = oflow.reproject_flow(torch.from_numpy(g[f"{tag}_depth"]), torch.from_numpy(g[f"{tag}_T1"]))
    # same torch build, same ops: bit-identical on this machine; allow BLAS differences on another host
    assert np.allclose(flow.numpy(), g[f"{tag}_flow"], rtol=0, atol=2e-3)
    K, invK = oflow.intrinsics(*g[f"{tag}_depth"].shape[1:])
    assert np.array_equal(K.numpy(), g[f"{tag}_K"]) and np.allclose(invK.numpy(), g[f"{tag}_invK"], rtol=1e-6)


def test_special_flow_oracle_matches_reference(golden):
    g = golden("special_cases")
    for kind in (5, 6, 7):
        for rep, (h, w) in enumerate(((23, 31), (46, 62))):
            torch.manual_seed(1000 + 10 * kind + rep)
            f, b, params = oflow.special_flow(h, w, kind)
            assert np.allclose(f.numpy(), g[f"k{kind}_{rep}_flow"], rtol=0, atol=1e-4)
            assert np.allclose(b.numpy(), g[f"k{kind}_{rep}_back"], rtol=0, atol=1e-4)
            if params is not None:
                assert np.allclose(params, g[f"k{kind}_{rep}_params"])


@pytest.mark.parametrize("tag", ["f32", "f64", "f32w3"])
def test_bilateral_oracle_matches_reference(golden, tag):
    g = golden("bilateral_cases")
    fs = [int(v) for v in g[f"{tag}_fs"]]
    out = obil.sparse_bilateral_filtering(g[f"{tag}_in"].copy(), fs, 0.04, len(fs))
    assert out.dtype == g[f"{tag}_out"].dtype
    assert np.array_equal(out, g[f"{tag}_out"], equal_nan=True)


def test_rank_table_exceptions(golden):
    """k(n) = n//2 except where the float32 running sum overshoots 0.5."""
    k = obil.rank_table(49)
    assert np.array_equal(k, golden("bilateral_cases")["rank_table"][:50])
    odd = [n for n in range(1, 50) if k[n] != n // 2]
    assert all(n % 2 == 0 and k[n] == n // 2 - 1 for n in odd)
    assert odd == [20, 22, 28, 36, 40, 44, 48]


def test_pair_oracle_equals_composition(golden):
    """oracle.disparity_pair == disparity_flow -> FW -> mask -> fix_warped_depth composed from the pinned pieces."""
    rng = np.random.default_rng(3)
    B, H, W = 3, 21, 30
    img = rng.integers(0, 256, (B, 3, H, W)).astype(np.float32)
    depth = rng.integers(1, 30, (B, 1, H, W)).astype(np.float32)
    depth[:, :, 2, 3] = 100.0
    sBf = rng.uniform(40, 55, B).astype(np.float32)
    img1, d1, back, flow, valid, coll = oracle.disparity_pair(img, depth, sBf, nthreads=2)
    for b in range(B):
        f = oflow.disparity_flow(torch.from_numpy(depth[b]), torch.tensor(sBf[b])).numpy()
        assert np.array_equal(f, flow[b]) and np.all(np.signbit(flow[b, 1]))
        obj = np.concatenate([img[b], depth[b], f * -1.0])
        out, v, c, _, _ = oracle.fw_forward(obj, f, depth[b])
        assert np.array_equal(v, valid[b]) and np.array_equal(c, coll[b])
        assert np.array_equal(out[0:3] * v, img1[b])
        assert np.array_equal(oflow.fix_warped_depth(torch.from_numpy(out[3:4] * v)).numpy(), d1[b])
        assert np.array_equal(out[4:6] * v, back[b])


def test_pipeline_golden_is_consistent_with_oracle(golden):
    """Recompute pair 0->1 of the reference's group tensor (preprocess.py:437-447) with the oracle."""
    g = golden("pipeline_case")
    grp = g["group"]
    depth0 = oflow.normalize_depth(torch.from_numpy(g["raw_depth"].copy())[None]).numpy()
    assert np.array_equal(depth0, grp[3:4])
    img1, d1, back, flow, valid, _ = oracle.disparity_pair(g["img0"][None], depth0[None], np.array([g["sBf"]], np.float32))
    assert np.array_equal(img1[0], grp[4:7]) and np.array_equal(d1[0], grp[7:8])
    assert np.array_equal(flow[0], grp[24:26]) and np.array_equal(back[0], grp[26:28])


def test_inpaint_mask_oracle_matches_reference(golden):
    """utils.inpaint's hole-mask logic (utils.py:137-149): masks captured at the reference's cv2.inpaint call."""
    g = golden("inpaint_case")
    for k in range(5):
        assert np.array_equal(oflow.inpaint_mask(g[f"valid{k}"][0], g[f"collision{k}"][0]), g[f"mask{k}"])
    for b in range(g["mask_valid"].shape[0]):
        assert np.array_equal(oflow.inpaint_mask(g["mask_valid"][b, 0], g["mask_collision"][b, 0]), g["mask_out"][b, 0])
    assert g["mask_out"].any() and not g["mask_out"].all()


def test_bilateral_mask_path_oracle_matches_reference(golden):
    """oracle/bilateral.py with a binary mask == the reference's mask path (golden m_* cases: uint8, float64 and bool masks)."""
    import numpy as np

    from oracle import bilateral as obil

    g = golden("bilateral_cases")
    for tag in ("m_f32_u8", "m_f32_f64", "m_f64_bool"):
        fs = [int(v) for v in g[f"{tag}_fs"]]
        with np.errstate(divide="ignore", invalid="ignore"):
            got = obil.sparse_bilateral_filtering(g[f"{tag}_in"].copy(), fs, 0.04, len(fs), mask=g[f"{tag}_mask"])
        assert got.dtype == g[f"{tag}_out"].dtype and np.array_equal(got, g[f"{tag}_out"], equal_nan=True), tag
    assert np.array_equal(obil.rank_table(225, np.float64), g["rank_table_f64"])
    assert not np.array_equal(g["rank_table"], g["rank_table_f64"])


def _telea_case(rng, h, w, smooth):
    if smooth:
        y, x = np.mgrid[0:h, 0:w]
        img = np.stack([(x * 3 + y) % 256, (x + 2 * y) % 256, 128 + 40 * np.sin(x / 5.0) + 30 * np.cos(y / 7.0)], -1).astype(np.uint8)
    else:
        img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    mask = np.zeros((h, w), np.uint8)
    mask[5:9, 4:20] = 1
    mask[12:14, :] = 1                 # a band across the whole frame
    mask[h // 2:h // 2 + 8, w - 14:w - 11] = 1
    mask[0:3, 0:5] = 1                 # image corner / borders (OpenCV's km / lm index shifts)
    mask[h - 4:, w - 10:] = 1
    mask[1, w // 2:w // 2 + 4] = 1
    mask[rng.random((h, w)) > 0.97] = 1
    return img, mask


def test_telea_heap_order_restatement_equals_cv2():
    """oracle/inpaint.py pins OpenCV's Telea arithmetic: with OpenCV's own fill order (a stable priority queue on T) the restatement
    reproduces cv2.inpaint(..., 3, cv2.INPAINT_TELEA) - what utils.inpaint calls (utils.py:149) - bit for bit, holes at the image
    borders and corners included.  The CUDA kernel shares this arithmetic and differs only in the fill order (layers)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import inpaint as oinp

    rng = np.random.default_rng(0)
    for (h, w), smooth in (((36, 48), False), ((36, 48), True), ((20, 27), False)):
        img, mask = _telea_case(rng, h, w, smooth)
        ref = cv2.inpaint(img, mask, 3, cv2.INPAINT_TELEA)
        assert np.array_equal(oinp.telea(img, mask, 3, order="heap"), ref), (h, w, smooth)
    for radius in (1, 2, 5):
        img, mask = _telea_case(rng, 24, 30, True)
        assert np.array_equal(oinp.telea(img, mask, radius, order="heap"), cv2.inpaint(img, mask, radius, cv2.INPAINT_TELEA)), radius


def test_telea_layer_order_fills_every_hole_and_touches_nothing_else():
    """The layer order (the CUDA kernel's): every hole pixel reachable from known pixels is filled exactly once, known pixels are
    never modified, an empty mask is the identity and a hole-only frame is left alone (nothing known to march from)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import inpaint as oinp

    rng = np.random.default_rng(3)
    img, mask = _telea_case(rng, 30, 40, True)
    out, (layers, filled) = oinp.telea(img, mask, 3, order="layer", return_stats=True)
    hole = mask != 0
    assert filled == int(hole.sum()) and layers >= 2 and np.array_equal(out[~hole], img[~hole])
    ref = cv2.inpaint(img, mask, 3, cv2.INPAINT_TELEA)
    d = np.abs(out.astype(int) - ref.astype(int))[hole]
    print(f"[telea] layer vs heap order on a smooth frame: mean |d| = {d.mean():.2f} levels, > 1 level: {(d > 1).mean():.3f}, max {d.max()}")
    assert d.mean() < 4.0
    assert np.array_equal(oinp.telea(img, np.zeros_like(mask), 3), img)
    assert np.array_equal(oinp.telea(img, np.ones_like(mask), 3), img)
