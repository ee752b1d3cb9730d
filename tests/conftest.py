import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_NPZ = sorted(f[:-4] for f in os.listdir(os.path.join(ROOT, "tests", "golden")) if f.endswith(".npz"))
# "...h" fixtures (full-size cases of BASELINE.json configs 3 and 4) hold the
# reference frame as per-row CRC-32s instead of pixels
GOLDEN_HASHED = [g for g in _NPZ if g.endswith("h")]
# "...c" fixtures (config 5 at size) hold row CRC-32s only; the scene comes from
# the generator inside the harness binaries
GOLDEN_CRC = [g for g in _NPZ if g.endswith("c")]
# "..._T" fixtures hold the reference's own primary hit distances (dump mode),
# written by the patched scratch build oracle/Makefile `tdump` makes
GOLDEN_T = [g for g in _NPZ if g.endswith("_T")]
# "..._pt" fixtures: path tracer (blob with QR_BLOB_PT, the reference's frame after N accumulated frames)
GOLDEN_PT = [g for g in _NPZ if g.endswith("_pt")]
GOLDEN_ALL = [g for g in _NPZ if not g.endswith("h") and not g.endswith("c") and not g.endswith("_T")
              and not g.endswith("_pt")]
# the 1080p frame is the bench workload; CPU tests use the 800x480 cases
GOLDEN_SMALL = [g for g in GOLDEN_ALL if "1080p" not in g]


@pytest.fixture(scope="session")
def entry():
    import __graft_entry__ as ge
    return ge


@pytest.fixture(scope="session")
def pkg(entry):
    return entry.load_package()
