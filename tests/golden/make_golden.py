"""Generate the golden fixtures of tests/golden/ by running the REFERENCE's own Python (imported from
/root/reference, read-only) on seeded inputs.  Runs only in the build container; the fixtures are committed.

The reference has no tests or golden vectors for this path, so these fixtures are what pins the
oracle and the CUDA path:
  * alt_cuda/fw.py is imported unmodified with a stand-in `fw_cuda` module whose forward_warping is the literal
    C restatement of fw_cuda_kernel.cu (oracle.splat_literal) and which records the safe_y/safe_x the reference's
    own prologue computed;
  * geometry.py, bilateral_filter.py, utils.py are imported unmodified;
  * preprocess.py is exec'd from source with the one-character syntax fix at line 463 and import shims
    (SURVEY.md section 8c); utils.inpaint is replaced by the identity (OpenCV inpaint is outside the path).
The script also asserts that oracle/flow.py and oracle/bilateral.py reproduce the reference bit for bit.
"""
from __future__ import annotations

import io
import sys
import types
from contextlib import redirect_stdout
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from oracle import bilateral as obil  # noqa: E402
from oracle import flow as oflow  # noqa: E402

CAPTURE = {}


def install_reference():
    """Make the reference importable: stand-in fw_cuda, matplotlib stub, patched preprocess."""
    fw_cuda = types.ModuleType("fw_cuda")

    def forward_warping(obj, safe_y, safe_x, depth):
        CAPTURE["safe_y"], CAPTURE["safe_x"] = safe_y.clone(), safe_x.clone()
        out, valid, coll, rc = oracle.splat_literal(obj.numpy(), safe_y.numpy(), safe_x.numpy(), depth.numpy())
        assert rc == 0
        return [torch.from_numpy(out), torch.from_numpy(valid), torch.from_numpy(coll)]

    fw_cuda.forward_warping = forward_warping
    sys.modules["fw_cuda"] = fw_cuda
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    sys.path.insert(0, str(REF))
    import dataloader  # noqa: F401
    import utils as ref_utils

    sys.modules["dataloader"].COCO = None
    ref_utils._real_inpaint = ref_utils.inpaint
    ref_utils.inpaint = lambda img, valid, collision: img
    src = (REF / "preprocess.py").read_text().splitlines()
    assert src[462].rstrip().endswith("axis=0"), src[462]
    src[462] = src[462] + ")"
    mod = types.ModuleType("ref_preprocess")
    mod.__file__ = str(REF / "preprocess.py")
    exec(compile("\n".join(src), mod.__file__, "exec"), mod.__dict__)
    import bilateral_filter as ref_bil
    import geometry as ref_geo
    from alt_cuda.fw import FW as RefFW

    return mod, ref_utils, ref_geo, ref_bil, RefFW


def diml_depth(rng, h, w, quant=True):
    """DIML-shaped depth: ground ramp + rectangles, 8-bit disparity -> depth."""
    y, x = np.mgrid[0:h, 0:w]
    disp = 20 + 15 * y / h + 3 * np.sin(2 * np.pi * x / w)
    for _ in range(6):
        r0, c0 = rng.integers(0, h - 4), rng.integers(0, w - 4)
        r1, c1 = r0 + rng.integers(3, max(4, h // 2)), c0 + rng.integers(3, max(4, w // 2))
        disp[r0:r1, c0:c1] = rng.uniform(30, 250)
    disp = disp + rng.normal(0, 0.7, disp.shape)
    if quant:
        disp = np.clip(np.round(disp), 0, 255)
    disp = disp * 63 / 255
    return 50.0 / (disp + 0.005)


def fw_cases(RefFW):
    out = {}
    rng = np.random.default_rng(1234)
    fw = RefFW(device="cpu")

    def run(tag, obj, flow, depth):
        o, v, c = fw(torch.from_numpy(obj), torch.from_numpy(flow), torch.from_numpy(depth))
        out[f"{tag}_obj"], out[f"{tag}_flow"], out[f"{tag}_depth"] = obj, flow, depth
        out[f"{tag}_safe_x"] = CAPTURE["safe_x"].numpy()[0, 0]
        out[f"{tag}_safe_y"] = CAPTURE["safe_y"].numpy()[0, 0]
        out[f"{tag}_out"], out[f"{tag}_valid"], out[f"{tag}_coll"] = o.numpy(), v.numpy(), c.numpy()
        # the oracle's own prologue must agree with the reference's
        sx, sy = oracle.fw_targets(flow)
        assert np.array_equal(sx, out[f"{tag}_safe_x"]) and np.array_equal(sy, out[f"{tag}_safe_y"]), tag
        o2, v2, c2, win, _ = oracle.fw_forward(obj, flow, depth)
        assert np.array_equal(o2, out[f"{tag}_out"]) and np.array_equal(v2, out[f"{tag}_valid"]) and np.array_equal(c2, out[f"{tag}_coll"]), tag
        out[f"{tag}_winner"] = win

    # A: random flow, quantised depths (ties), sentinel depths, signed zeros; C = 6
    h, w = 24, 32
    obj = rng.uniform(-5, 255, (6, h, w)).astype(np.float32)
    flow = rng.normal(0, 4, (2, h, w)).astype(np.float32)
    depth = rng.integers(1, 6, (1, h, w)).astype(np.float32)
    depth[0, 3, 4:9] = 1000.0
    depth[0, 5, 1:4] = 2000.0
    depth[0, 7, 7] = np.nan
    depth[0, 8, 2:6] = np.float32(-0.0)
    depth[0, 9, 2:6] = 0.0
    depth[0, 10, 3] = -np.inf
    depth[0, 11, 3] = -3.5
    run("A", obj, flow, depth)
    # B: float64 flow landing within 1e-9 of pixel boundaries; C = 2 (a flow payload)
    h, w = 16, 20
    yy, xx = np.mgrid[0:h, 0:w]
    tgt = rng.integers(-3, w + 3, (h, w)) + rng.choice([-1e-9, 0.0, 1e-9, 0.5, 0.999999999], (h, w))
    tgy = rng.integers(-3, h + 3, (h, w)) + rng.choice([-1e-9, 0.0, 1e-9, 0.25], (h, w))
    flow = np.stack([tgt - xx, tgy - yy]).astype(np.float64)
    obj = rng.normal(0, 10, (2, h, w)).astype(np.float32)
    depth = rng.uniform(1, 99, (1, h, w)).astype(np.float32)
    run("B", obj, flow, depth)
    # C: huge flows -> everything piles up on the clamped border; all depths >= 1000 in one corner; C = 7
    h, w = 20, 28
    flow = rng.normal(0, 40, (2, h, w)).astype(np.float32)
    flow[:, :6, :6] = -1e4
    flow[:, -5:, -5:] = np.inf
    obj = rng.uniform(0, 1, (7, h, w)).astype(np.float32)
    depth = rng.integers(1, 4, (1, h, w)).astype(np.float32)
    depth[0, -5:, -5:] = 1000.0
    run("C", obj, flow, depth)
    # D: degenerate shapes
    run("D", rng.uniform(0, 1, (1, 1, 1)).astype(np.float32), np.zeros((2, 1, 1), np.float32), np.ones((1, 1, 1), np.float32))
    run("E", rng.uniform(0, 1, (3, 1, 37)).astype(np.float32), rng.normal(0, 5, (2, 1, 37)).astype(np.float32),
        rng.integers(1, 3, (1, 1, 37)).astype(np.float32))
    run("F", rng.uniform(0, 1, (4, 33, 1)).astype(np.float32), rng.normal(0, 5, (2, 33, 1)).astype(np.float32),
        rng.integers(1, 3, (1, 33, 1)).astype(np.float32))
    return out


def convert_cases(pp, ref_utils):
    out = {}
    rng = np.random.default_rng(77)
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        raw = diml_depth(rng, 30, 44).astype(dt)
        raw[2, 3:9] = 0
        raw[5, 5:9] = 250.0
        ref = ref_utils.normalize_depth(torch.from_numpy(raw.copy())[None]).numpy()
        mine = oflow.normalize_depth(torch.from_numpy(raw.copy())[None]).numpy()
        assert np.array_equal(ref, mine)
        out[f"norm_{tag}_in"], out[f"norm_{tag}_out"] = raw, ref
        ref_utils.set_seed(12345 + 7)
        disp = pp.Convert.depth_to_disparity(torch.from_numpy(ref))
        flow = pp.Convert.disparity_to_flow(disp, device="cpu", random_sign=False)
        ref_utils.set_seed(12345 + 7)
        sBf = oflow.get_random(0.3, 0.8, random_sign=False) * 50 * 1
        mine = oflow.disparity_flow(torch.from_numpy(ref), sBf)
        assert flow.dtype == mine.dtype and np.array_equal(flow.numpy(), mine.numpy())
        assert np.all(np.signbit(flow.numpy()[1]))  # flow.y is -0.0
        out[f"disp_{tag}_sBf"] = np.float32(sBf.item())
        out[f"disp_{tag}_flow"] = flow.numpy()
    d = rng.uniform(-1, 120, (1, 9, 11)).astype(np.float32)
    d[0, 0, :3] = 0
    d[0, 1, :3] = 99.5
    out["fix_in"] = d
    out["fix_out"] = ref_utils.fix_warped_depth(torch.from_numpy(d.copy())).numpy()
    assert np.array_equal(out["fix_out"], oflow.fix_warped_depth(torch.from_numpy(d)).numpy())
    return out


def reproject_cases(pp, ref_utils):
    out = {}
    rng = np.random.default_rng(99)
    for tag, dt, (h, w) in (("f32", np.float32, (48, 64)), ("f64", np.float64, (36, 52)), ("f32b", np.float32, (120, 160))):
        depth = oflow.normalize_depth(torch.from_numpy(diml_depth(rng, h, w).astype(dt))[None])
        ref_utils.set_seed(12345 + 3)
        flow, T1 = pp.Convert.depth_to_random_flow(depth, device="cpu")
        ref_utils.set_seed(12345 + 3)
        T_mine, _, _ = oflow.random_motion()
        assert torch.equal(T1, T_mine)
        mine = oflow.reproject_flow(depth, T1)
        assert torch.equal(flow, mine), tag
        K, inv_K = pp.Plausible.K((h, w))
        out[f"{tag}_depth"], out[f"{tag}_flow"], out[f"{tag}_T1"] = depth.numpy(), flow.numpy(), T1.numpy()
        out[f"{tag}_K"], out[f"{tag}_invK"] = K.numpy(), inv_K.numpy()
    return out


def special_cases(pp, ref_utils):
    out = {}
    for kind in (5, 6, 7):
        for rep, (h, w) in enumerate(((23, 31), (46, 62))):
            ref_utils.set_seed(1000 + 10 * kind + rep)
            sf = pp.SpecialFlow(device="cpu")
            f, bfl = sf((h, w), float(kind))
            ref_utils.set_seed(1000 + 10 * kind + rep)
            mf, mb, params = oflow.special_flow(h, w, kind)
            assert torch.equal(f, mf) and torch.equal(bfl, mb), kind
            out[f"k{kind}_{rep}_flow"], out[f"k{kind}_{rep}_back"] = f.numpy(), bfl.numpy()
            out[f"k{kind}_{rep}_params"] = np.array(params if params is not None else [0.0] * 10, np.float64)
    # a REUSED instance alternates its branches (preprocess.py:49,83): second call = horizontal flip / the other shear matrix
    h, w = 23, 31
    sf = pp.SpecialFlow(device="cpu")
    sf((h, w), 5.0)
    f, bfl = sf((h, w), 5.0)
    out["k5_reuse_flow"], out["k5_reuse_back"] = f.numpy(), bfl.numpy()
    sf = pp.SpecialFlow(device="cpu")
    ref_utils.set_seed(2000)
    sf((h, w), 7.0)
    ref_utils.set_seed(2001)
    f, bfl = sf((h, w), 7.0)
    out["k7_reuse_flow"], out["k7_reuse_back"] = f.numpy(), bfl.numpy()
    return out


def bilateral_cases(ref_bil):
    out = {}
    rng = np.random.default_rng(5)
    for tag, dt, (h, w), fs in (("f32", np.float32, (40, 56), [7, 7, 5, 5, 5]), ("f64", np.float64, (32, 40), [7, 5, 5]),
                                ("f32w3", np.float32, (17, 19), [3, 9])):
        depth = diml_depth(rng, h, w).astype(dt)
        depth = (depth / depth.max()).astype(dt) * dt(8) + dt(0.5)
        depth[4:7, 5:8] = 0
        depth[0, 3] = 0
        img = np.zeros((h, w, 3), np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ref = ref_bil.sparse_bilateral_filtering(depth.copy(), img, fs, depth_threshold=0.04, num_iter=len(fs))
            mine = obil.sparse_bilateral_filtering(depth.copy(), fs, 0.04, len(fs))
        assert ref.dtype == mine.dtype and np.array_equal(ref, mine, equal_nan=True), tag
        out[f"{tag}_in"], out[f"{tag}_out"] = depth, ref
        out[f"{tag}_fs"] = np.array(fs)
    out["rank_table"] = obil.rank_table(225)
    # the mask path (bilateral_filter.py:48-49,72-80,161,169-170,180-182) with BINARY masks of three dtypes: the median
    # coefficients are float32 * mask.dtype, so the rank rule runs in float32 (uint8 / bool masks) or float64 (float64 masks)
    for tag, dt, mdt, (h, w), fs in (("m_f32_u8", np.float32, np.uint8, (36, 52), [7, 5, 5]), ("m_f32_f64", np.float32, np.float64, (30, 44), [7, 7, 5]),
                                     ("m_f64_bool", np.float64, np.bool_, (28, 33), [5, 3])):
        depth = diml_depth(rng, h, w).astype(dt)
        depth = (depth / depth.max()).astype(dt) * dt(8) + dt(0.5)
        depth[4:7, 5:8] = 0
        mask = (rng.random((h, w)) > 0.25)
        mask[10:16, 8:30] = False
        mask[0, :5] = False
        mask[5, 6] = False          # a masked zero-depth pixel: the depth == 0 rule is undone by the mask rule
        mask = mask.astype(mdt)
        img = np.zeros((h, w, 3), np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ref = ref_bil.sparse_bilateral_filtering(depth.copy(), img, fs, depth_threshold=0.04, num_iter=len(fs), mask=mask)
            mine = obil.sparse_bilateral_filtering(depth.copy(), fs, 0.04, len(fs), mask=mask)
        assert ref.dtype == mine.dtype and np.array_equal(ref, mine, equal_nan=True), tag
        with np.errstate(divide="ignore", invalid="ignore"):
            plain = obil.sparse_bilateral_filtering(depth.copy(), fs, 0.04, len(fs))
        assert not np.array_equal(ref, plain, equal_nan=True), "the mask must matter in this case"
        out[f"{tag}_in"], out[f"{tag}_mask"], out[f"{tag}_out"], out[f"{tag}_fs"] = depth, mask, ref, np.array(fs)
    out["rank_table_f64"] = obil.rank_table(225, np.float64)
    return out


def pipeline_case(pp, ref_utils):
    """The reference's PreprocessPlusAugment.forward up to group.npz (preprocess.py:341-447), inpaint = identity,
    float32 inputs, on a small frame; aborted before the augmentation loop."""
    rng = np.random.default_rng(2024)
    h, w = 40, 56
    img0 = rng.integers(0, 256, (3, h, w)).astype(np.float32)
    raw = diml_depth(rng, h, w).astype(np.float32)
    grabbed = {}

    class Stop(Exception):
        pass

    def fake_savez(path, **kw):
        grabbed.update(kw)
        raise Stop()

    ppa = pp.PreprocessPlusAugment(device="cpu")
    ref_utils.set_seed(12345 + 11)
    real = np.savez_compressed
    np.savez_compressed = fake_savez
    pp.os.makedirs = lambda *a, **k: None
    try:
        with redirect_stdout(io.StringIO()):
            ppa((torch.from_numpy(img0), torch.from_numpy(raw.copy())[None]), "/tmp/ofd_golden/0", is_stereo=False)
    except Stop:
        pass
    finally:
        np.savez_compressed = real
    group = grabbed["img_depth_flow"]
    assert group.shape == (44, h, w)
    # replay the RNG to record the scalars the product needs
    ref_utils.set_seed(12345 + 11)
    sBf = oflow.get_random(0.3, 0.8, random_sign=False) * 50 * 1
    T1, _, _ = oflow.random_motion()
    return dict(img0=img0, raw_depth=raw, group=group.astype(np.float32), sBf=np.float32(sBf.item()), T1=T1.numpy())


def pipeline_case_f64(pp, ref_utils):
    """pipeline_case with the depth the reference's loaders really deliver: float64 (cv2.imread(...).astype(float),
    utils.py:48,62).  The reference then keeps float64 in normalize_depth, the disparity flow, the FW target computation and the
    flow composition; the group tensor comes out float64."""
    rng = np.random.default_rng(2025)
    h, w = 40, 56
    img0 = rng.integers(0, 256, (3, h, w)).astype(np.float32)
    raw = diml_depth(rng, h, w, quant=False).astype(np.float64)  # continuous depth: float32 evaluation could move targets
    grabbed = {}

    class Stop(Exception):
        pass

    def fake_savez(path, **kw):
        grabbed.update(kw)
        raise Stop()

    ppa = pp.PreprocessPlusAugment(device="cpu")
    ref_utils.set_seed(12345 + 12)
    real = np.savez_compressed
    np.savez_compressed = fake_savez
    pp.os.makedirs = lambda *a, **k: None
    try:
        with redirect_stdout(io.StringIO()):
            ppa((torch.from_numpy(img0), torch.from_numpy(raw.copy())[None]), "/tmp/ofd_golden/0", is_stereo=False)
    except Stop:
        pass
    finally:
        np.savez_compressed = real
    group = grabbed["img_depth_flow"]
    assert group.shape == (44, h, w) and group.dtype == np.float64
    return dict(img0=img0, raw_depth=raw, group=group)


def inpaint_case(pp, ref_utils):
    """The same pipeline with the reference's REAL utils.inpaint (utils.py:136-151; OpenCV Telea on the CPU).  Records
    the group tensor and, for each of the 5 inpaint calls, (valid, collision) and the mask handed to cv2.inpaint."""
    import cv2

    rng = np.random.default_rng(2024)
    h, w = 40, 56
    img0 = rng.integers(0, 256, (3, h, w)).astype(np.float32)
    raw = diml_depth(rng, h, w).astype(np.float32)
    calls, grabbed = [], {}

    class Stop(Exception):
        pass

    real_cv_inpaint = cv2.inpaint

    def rec_cv_inpaint(src, mask, radius, flags):
        calls[-1]["mask"] = mask.copy()
        return real_cv_inpaint(src, mask, radius, flags)

    def rec_inpaint(img, valid, collision):
        calls.append({"valid": valid.numpy().copy(), "collision": collision.numpy().copy()})
        return ref_utils._real_inpaint(img, valid, collision)

    def fake_savez(path, **kw):
        grabbed.update(kw)
        raise Stop()

    saved = (ref_utils.inpaint, cv2.inpaint, np.savez_compressed, torch.Tensor.get_device)
    ref_utils.inpaint, cv2.inpaint, np.savez_compressed = rec_inpaint, rec_cv_inpaint, fake_savez
    torch.Tensor.get_device = lambda self: "cpu"  # utils.py:150 `.to(img.get_device())` fails for CPU tensors (Appendix B)
    pp.os.makedirs = lambda *a, **k: None
    ppa = pp.PreprocessPlusAugment(device="cpu")
    ref_utils.set_seed(12345 + 11)
    try:
        with redirect_stdout(io.StringIO()):
            ppa((torch.from_numpy(img0), torch.from_numpy(raw.copy())[None]), "/tmp/ofd_golden/0", is_stereo=False)
    except Stop:
        pass
    finally:
        ref_utils.inpaint, cv2.inpaint, np.savez_compressed, torch.Tensor.get_device = saved
    assert len(calls) == 5 and grabbed["img_depth_flow"].shape == (44, h, w)
    out = dict(img0=img0, raw_depth=raw, group=grabbed["img_depth_flow"].astype(np.float32))
    for k, c in enumerate(calls):
        out[f"valid{k}"], out[f"collision{k}"], out[f"mask{k}"] = c["valid"], c["collision"], c["mask"]
        assert np.array_equal(oflow.inpaint_mask(c["valid"][0], c["collision"][0]), c["mask"]), k
    # synthetic collision cases for the mask logic (collision is identically 0 in the pipeline)
    v = (rng.random((6, 1, 21, 29)) > 0.3).astype(np.float32)
    c = ((rng.random((6, 1, 21, 29)) > 0.8) * v).astype(np.float32)
    out["mask_valid"], out["mask_collision"] = v, c
    ms = []
    for b in range(6):
        rec = {}
        cv2.inpaint = lambda src, mask, radius, flags, rec=rec: rec.setdefault("m", mask.copy()) is None or src
        torch.Tensor.get_device = lambda self: "cpu"
        try:
            ref_utils._real_inpaint(torch.zeros(3, 21, 29), torch.from_numpy(v[b]), torch.from_numpy(c[b]))
        finally:
            cv2.inpaint, torch.Tensor.get_device = saved[1], saved[3]
        ms.append(rec["m"])
        assert np.array_equal(oflow.inpaint_mask(v[b, 0], c[b, 0]), rec["m"])
    out["mask_out"] = np.stack(ms)[:, None]
    return out


def concat_back_cases(pp):
    out = {}
    rng = np.random.default_rng(31)
    h, w = 26, 34
    fAB = rng.normal(0, 3, (2, h, w)).astype(np.float32)
    bAB = rng.normal(0, 3, (2, h, w)).astype(np.float32)
    fBC = rng.normal(0, 3, (2, h, w)).astype(np.float32)
    dB = rng.integers(1, 5, (1, h, w)).astype(np.float32)
    cf, bf = pp.ConcatFlow("cpu"), pp.BackFlow("cpu")
    c, cv = cf(*(torch.from_numpy(a) for a in (fAB, bAB, fBC, dB)))
    b, bv = bf(torch.from_numpy(fAB), torch.from_numpy(dB))
    out.update(fAB=fAB, bAB=bAB, fBC=fBC, dB=dB, concat=c.numpy(), concat_valid=cv.numpy(), back=b.numpy(), back_valid=bv.numpy())
    return out


def preprocess_case(pp, ref_utils):
    """The reference's WHOLE PreprocessPlusAugment.forward (preprocess.py:341-476): group.npz and the 5 x 12 x 2 augmented
    files, inpaint = identity, float32 inputs, small frame.  Every np.savez_compressed call is captured."""
    rng = np.random.default_rng(77)
    h, w = 30, 44
    img0 = rng.integers(0, 256, (3, h, w)).astype(np.float32)
    raw = diml_depth(rng, h, w).astype(np.float32)
    files = {}

    def fake_savez(path, **kw):
        files[Path(path).name] = {k: np.asarray(v) for k, v in kw.items()}

    ppa = pp.PreprocessPlusAugment(device="cpu")
    ref_utils.set_seed(12345 + 3)
    real = np.savez_compressed
    np.savez_compressed = fake_savez
    pp.os.makedirs = lambda *a, **k: None
    try:
        with redirect_stdout(io.StringIO()):
            try:
                ppa((torch.from_numpy(img0), torch.from_numpy(raw.copy())[None]), "/tmp/ofd_golden/7", is_stereo=False)
            except (NameError, UnboundLocalError) as e:  # the `del` block after the last file is written (Appendix B)
                files["__tail_error__"] = {"msg": np.array(str(e))}
    finally:
        np.savez_compressed = real
    assert "group.npz" in files and len([k for k in files if k.endswith(".npz")]) == 121, sorted(files)[:5]
    out = dict(img0=img0, raw_depth=raw)
    for name, kw in files.items():
        if not name.endswith(".npz"):
            continue
        stem = name[:-4]
        arr = kw["img_depth_flow"]
        assert arr.dtype == np.float32, (name, arr.dtype)
        out[f"{stem}__data"] = arr
        if "augment_flow_type" in kw:
            out[f"{stem}__type"] = np.asarray(kw["augment_flow_type"])
    return out


def reader_case(pp, ref_utils):
    """The reference's training reader (dataloader.AugmentedDataset.getitem_from_npz / DepthToFlowDataset, dataloader.py:79-232)
    on files built from the preprocess_case fixture, with the `augment_img` key its writer omits added (0 for *_1, 1 for *_2).
    Records (img0, img1, flow, img0_depth, label) for a few (file, group, seed) combinations."""
    import tempfile

    import dataloader as ref_dl

    g = np.load(HERE / "preprocess_case.npz")
    out = {}
    cases = [("1_5_1", 1, 3, None, True), ("2_3_2", 2, 4, (16, 24), True), ("0_0_2", 0, 5, (20, 20), False), ("0_9_1", 0, 6, None, True)]
    with tempfile.TemporaryDirectory() as tmp:
        np.savez(f"{tmp}/group.npz", img_depth_flow=g["group__data"])
        for k, (stem, grp, seed, crop, norm) in enumerate(cases):
            np.savez(f"{tmp}/{stem}.npz", img_depth_flow=g[f"{stem}__data"], augment_flow_type=g[f"{stem}__type"],
                     augment_img=int(stem[-1]) - 1)
            ds = ref_dl.AugmentedDataset(normalize_dataset=norm, crop_size=crop, do_flip=True)
            np.random.seed(seed)
            res = ds.getitem_from_npz(f"{tmp}/{stem}.npz", f"{tmp}/group.npz", grp, 0)
            assert len(res) == 5
            for name, t in zip(("img0", "img1", "flow", "depth", "label"), res):
                out[f"aug{k}_{name}"] = t.numpy()
            out[f"aug{k}_meta"] = np.array([grp, seed, -1 if crop is None else crop[0], -1 if crop is None else crop[1], int(norm)])
            out[f"aug{k}_stem"] = np.array(stem)
        # DepthToFlowDataset crops with undefined h, w (dataloader.py:221): only the crop-free call runs in the reference
        for k, (grp, seed) in enumerate([(0, 7), (1, 8), (2, 9)]):
            ds = ref_dl.DepthToFlowDataset(crop_size=None, do_flip=True)
            np.random.seed(seed)
            res = ds.getitem_from_npz(f"{tmp}/group.npz", grp, 0)
            for name, t in zip(("img0", "img1", "flow", "depth", "label"), res):
                out[f"d2f{k}_{name}"] = t.numpy()
            out[f"d2f{k}_meta"] = np.array([grp, seed])
    return out


def reader_size_case(pp, ref_utils):
    """The reference's reader with `size=` (T.Compose([ToTensor, Resize(size)]), dataloader.py:68-72,168-172): one upscaling and one
    downscaling AugmentedDataset case (the second with a crop, whose offsets the reference draws from the FILE's h and w) and one
    DepthToFlowDataset case.  torchvision's defaults of this image (0.26: antialiased bilinear)."""
    import tempfile

    import dataloader as ref_dl

    g = np.load(HERE / "preprocess_case.npz")
    out = {}
    cases = [("1_5_1", 1, 13, (45, 66), None, True), ("2_3_2", 2, 14, (20, 30), (12, 16), True)]
    with tempfile.TemporaryDirectory() as tmp:
        np.savez(f"{tmp}/group.npz", img_depth_flow=g["group__data"])
        for k, (stem, grp, seed, size, crop, norm) in enumerate(cases):
            np.savez(f"{tmp}/{stem}.npz", img_depth_flow=g[f"{stem}__data"], augment_flow_type=g[f"{stem}__type"],
                     augment_img=int(stem[-1]) - 1)
            ds = ref_dl.AugmentedDataset(normalize_dataset=norm, size=size, crop_size=crop, do_flip=True)
            np.random.seed(seed)
            res = ds.getitem_from_npz(f"{tmp}/{stem}.npz", f"{tmp}/group.npz", grp, 0)
            for name, t in zip(("img0", "img1", "flow", "depth", "label"), res):
                out[f"aug{k}_{name}"] = t.numpy()
            out[f"aug{k}_meta"] = np.array([grp, seed, size[0], size[1], -1 if crop is None else crop[0], -1 if crop is None else crop[1], int(norm)])
            out[f"aug{k}_stem"] = np.array(stem)
        ds = ref_dl.DepthToFlowDataset(size=(33, 50), crop_size=None, do_flip=True)
        np.random.seed(15)
        res = ds.getitem_from_npz(f"{tmp}/group.npz", 1, 0)
        for name, t in zip(("img0", "img1", "flow", "depth", "label"), res):
            out[f"d2f0_{name}"] = t.numpy()
        out["d2f0_meta"] = np.array([1, 15, 33, 50])
    return out


def main():
    torch.set_num_threads(1)
    pp, ref_utils, ref_geo, ref_bil, RefFW = install_reference()
    jobs = {
        "fw_cases": lambda: fw_cases(RefFW),
        "convert_cases": lambda: convert_cases(pp, ref_utils),
        "reproject_cases": lambda: reproject_cases(pp, ref_utils),
        "special_cases": lambda: special_cases(pp, ref_utils),
        "bilateral_cases": lambda: bilateral_cases(ref_bil),
        "concat_back_cases": lambda: concat_back_cases(pp),
        "pipeline_case": lambda: pipeline_case(pp, ref_utils),
        "pipeline_case_f64": lambda: pipeline_case_f64(pp, ref_utils),
        "inpaint_case": lambda: inpaint_case(pp, ref_utils),
        "preprocess_case": lambda: preprocess_case(pp, ref_utils),
        "reader_case": lambda: reader_case(pp, ref_utils),
        "reader_size_case": lambda: reader_size_case(pp, ref_utils),
    }
    only = set(sys.argv[1:])
    for name, job in jobs.items():
        if only and name not in only:
            continue
        data = job()
        np.savez_compressed(HERE / f"{name}.npz", **data)
        print(f"{name}: {len(data)} arrays, {(HERE / (name + '.npz')).stat().st_size / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
