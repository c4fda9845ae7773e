"""Writes profiles/r2/parity_report.json on a GPU box: the CUDA path against the reference's own full-size outputs (SHA-256 of every
bit-exact plane, 6-DoF flow error and truncated-target mismatch fraction, tie / collision counts, per-plane differing fraction of the
5-pair group).  Thin launcher: the comparison code lives with the tests (tests/parity_fullsize.py)."""
import runpy
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.argv[0] = "parity_fullsize"
runpy.run_path(str(Path(__file__).resolve().parent.parent / "tests" / "parity_fullsize.py"), run_name="__main__")
