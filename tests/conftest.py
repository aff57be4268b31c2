import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artifacts (they are git-ignored): build the CUDA library and the oracle
    when they are missing and a compiler is at hand.  Never rebuilds what is already there."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "dexterous_rl_manipulation_b200", "libdexsim_b200.so")
    ora = os.path.join(ROOT, "oracle", "_build", "libdexsim_oracle.so")
    if (not os.path.exists(lib) or not os.path.exists(ora)) and shutil.which("nvcc") and shutil.which("make"):
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=False)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
