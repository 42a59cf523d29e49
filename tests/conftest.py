import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The C restatement is test infrastructure: (re)build it on demand so a fresh checkout works."""
    import subprocess
    odir = os.path.join(ROOT, "oracle")
    so = os.path.join(odir, "liblbm_oracle.so")
    src = [os.path.join(odir, "lbm_oracle.c"), os.path.join(odir, "lbm_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", odir, "liblbm_oracle.so"], stdout=subprocess.DEVNULL)
    yield
