"""Multi-GPU parity over NCCL (needs >= 2 devices; skipped on a 1-GPU box where tests/test_gpu_decomposed.py covers the same
kernels through the loopback world): one process per GPU via torch.distributed.run on 127.0.0.1."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    from panslbm2_b200 import _lib
    return _lib.lib().pl_device_count()


def _port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("m,size,nt", [("1x1x2", "12x10x16", 9), ("2x1x1", "16x10x8", 8), ("2x2x1", "16x12x8", 7), ("2x2x2", "16x12x8", 6)])
def test_heatsink3d_over_nccl_equals_single_block(m, size, nt):
    n = eval(m.replace("x", "*"))
    if _ngpu() < n:
        pytest.skip(f"needs {n} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port", str(_port()),
           os.path.join(HERE, "nccl_worker.py"), m, size, str(nt)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK") == n, r.stdout
