import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def lib_built():
    """Build libtvc_b200.so if it is missing or stale (nvcc cross-compiles without a GPU)."""
    from tvc_ai_b200 import build
    return build.build()


# ---- parity record: the GPU tests put their measured numbers here; printed with -q and written to disk ----
PARITY = {}


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Print the parity numbers the GPU tests measured (pytest -q shows only dots otherwise) and write them to
    gpurun_out/parity_r02.json (travels back from the GPU box) and profiles/parity_r02.json."""
    if not PARITY:
        return
    import json
    tr = terminalreporter
    tr.write_sep("=", "parity record (device vs fp64 oracle, identical fp32 inputs)")
    for name, rec in PARITY.items():
        tr.write_line(f"[parity] {name}: " + json.dumps(rec, sort_keys=True))
    for d in ("gpurun_out", "profiles"):
        try:
            os.makedirs(os.path.join(ROOT, d), exist_ok=True)
            with open(os.path.join(ROOT, d, "parity_r02.json"), "w") as f:
                json.dump(PARITY, f, indent=1, sort_keys=True)
        except OSError:
            pass


@pytest.fixture(scope="session")
def parity_record():
    return PARITY
