"""Run the REFERENCE's own preprocess.py (staged unmodified in baseline/_ref by __graft_entry__.build()) on top of this
repository's drop-in modules, on a GPU - the zero-edit drop-in proof of BASELINE.json's north_star ("preprocess.py and
dataloader.py use it unchanged"; VERDICT r1 missing #3).

    python tools/run_reference_on_dropin.py --case dropin --out DIR      dropin/ first on sys.path: `import fw_cuda`, `import
                                                                         geometry`, `from bilateral_filter import ...`,
                                                                         `from alt_cuda.fw import FW` all resolve to this repo
    python tools/run_reference_on_dropin.py --case ref_fw --out DIR      the reference's own alt_cuda/fw.py (unmodified, its torch
                                                                         prologue included) over dropin/fw_cuda.py

The reference's `utils.py`, `dataloader.py`, `flow_colors.py` and `preprocess.py` are ALWAYS the reference's own files.  The only
edits are the load-time shims SURVEY.md section 8c lists (none touches the flow-synthesis code): the missing `)` on
preprocess.py:463, a stub for the absent `matplotlib`, `dataloader.COCO = None`, the module-level `device` the stereo branch
reads.  `--inpaint identity` replaces utils.inpaint by the identity (the goldens were made that way); `reference` keeps OpenCV.
Input: the `preprocess_case` golden frame (tests/golden) with utils.set_seed(12345 + 3), exactly as make_golden.py ran the
reference on the CPU.  Writes the reference's 121 .npz files into DIR and prints a one-line JSON summary.
"""
from __future__ import annotations

import argparse
import importlib.util
import io
import json
import sys
import time
import types
from contextlib import redirect_stdout
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
STAGE = ROOT / "baseline" / "_ref"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", choices=["dropin", "ref_fw"], default="dropin")
    ap.add_argument("--out", required=True)
    ap.add_argument("--inpaint", choices=["identity", "reference"], default="identity")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--golden", default="preprocess_case")
    args = ap.parse_args()
    if not (STAGE / "preprocess.py").exists():
        raise SystemExit(f"{STAGE} is missing: run __graft_entry__.build() where /root/reference exists")

    import numpy as np
    import torch

    # resolution order: this repo's package, then dropin/ (the module names the reference imports), then the reference's own files
    sys.path[:0] = [str(ROOT), str(ROOT / "dropin"), str(STAGE)]
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    if args.case == "ref_fw":
        # keep the reference's alt_cuda/fw.py (its torch prologue + `import fw_cuda`), with fw_cuda = dropin/fw_cuda.py
        import fw_cuda  # noqa: F401  (dropin/fw_cuda.py)

        pkg = types.ModuleType("alt_cuda")
        pkg.__path__ = [str(STAGE / "alt_cuda")]
        sys.modules["alt_cuda"] = pkg
        spec = importlib.util.spec_from_file_location("alt_cuda.fw", str(STAGE / "alt_cuda" / "fw.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["alt_cuda.fw"] = mod
        spec.loader.exec_module(mod)
    import dataloader  # the reference's (imports the reference's utils)
    import utils as ref_utils

    dataloader.COCO = None
    if args.inpaint == "identity":
        ref_utils.inpaint = lambda img, valid, collision: img
    src = (STAGE / "preprocess.py").read_text().splitlines()
    assert src[462].rstrip().endswith("axis=0"), src[462]
    src[462] = src[462] + ")"
    pp = types.ModuleType("ref_preprocess")
    pp.__file__ = str(STAGE / "preprocess.py")
    exec(compile("\n".join(src), pp.__file__, "exec"), pp.__dict__)
    pp.device = args.device

    def origin(name):  # where the reference's `import <name>` resolves in this interpreter
        if name in sys.modules:
            return getattr(sys.modules[name], "__file__", "?")
        spec = importlib.util.find_spec(name)
        return spec.origin if spec else None

    where = {name: origin(name) for name in ("fw_cuda", "geometry", "bilateral_filter", "alt_cuda.fw", "utils", "dataloader")}
    g = np.load(ROOT / "tests" / "golden" / f"{args.golden}.npz")
    img0, raw = torch.from_numpy(g["img0"]), torch.from_numpy(g["raw_depth"].copy())[None]
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    ppa = pp.PreprocessPlusAugment(device=args.device)
    ref_utils.set_seed(12345 + 3)
    tail = None
    t0 = time.time()
    with redirect_stdout(io.StringIO()):
        try:
            ppa((img0, raw), str(out), is_stereo=False)
        except (NameError, UnboundLocalError) as e:  # the reference's `del` block after the last file
            tail = str(e)
    torch.cuda.synchronize()
    files = sorted(p.name for p in out.glob("*.npz"))
    print(json.dumps({"case": args.case, "files": len(files), "seconds": round(time.time() - t0, 2), "tail_error": tail,
                      "modules": where, "device": torch.cuda.get_device_name(0)}))


if __name__ == "__main__":
    main()
