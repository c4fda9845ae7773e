import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(GOLDEN / f"{name}.npz"))
        return cache[name]

    return load


def _try_build():
    """Build the product library and the C oracle (cheap no-op when up to date).  Returns (lib_ok, oracle_ok, why)."""
    import shutil

    lib_ok = oracle_ok = True
    why = ""
    try:
        from opticalflowfromdepth_b200 import _build

        try:
            _build.build()
        except Exception as e:  # no nvcc on this box: a prebuilt, up-to-date-or-not library is still loadable
            if not _build.LIB.exists():
                lib_ok, why = False, f"libofd_b200.so cannot be built here ({type(e).__name__}: {str(e)[:80]})"
    except Exception as e:
        lib_ok, why = False, repr(e)
    try:
        import oracle

        oracle.build(ref=False)
    except Exception as e:
        oracle_ok = False
        why += f" oracle: {type(e).__name__}: {str(e)[:80]}"
        if shutil.which("gcc") is None:
            why += " (no gcc)"
    return lib_ok, oracle_ok, why


_BUILD_STATE = None

# test modules that only exercise the oracle (numpy / C restatement) and need neither nvcc nor the CUDA library
ORACLE_ONLY = {"test_oracle_cpu.py"}


def pytest_collection_modifyitems(config, items):
    """A box without the CUDA toolkit can still run the oracle-level CPU tests: everything that needs libofd_b200.so is skipped
    (with the reason) instead of failing in a session-wide fixture."""
    global _BUILD_STATE
    if _BUILD_STATE is None:
        _BUILD_STATE = _try_build()
    lib_ok, oracle_ok, why = _BUILD_STATE
    for item in items:
        name = Path(str(item.fspath)).name
        if not lib_ok and name not in ORACLE_ONLY:
            item.add_marker(pytest.mark.skip(reason=f"needs libofd_b200.so: {why}"))
        if not oracle_ok and name in ORACLE_ONLY:
            item.add_marker(pytest.mark.skip(reason=f"needs the C oracle: {why}"))
