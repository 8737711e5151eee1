"""GPU tier, N > 1: the one-process-per-GPU driver (lrbinner_b200/dist.py) on real devices, against the oracle.

tests/gpu_dist_worker.py is launched with torch.distributed.run on 2 (and, when the box has them, 4 and 8) GPUs; every
plan incl. the NVLink peer-memory exchange is checked row for row and table entry for table entry, two steps in a row.
Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu -x -q`; the log is kept under profiles/.
"""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_every_multi_gpu_plan_row_for_row_vs_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"{world} GPUs not available (gpurun --gpus {world})")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "gpu_dist_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    tail = "\n".join(p.stdout.splitlines()[-40:])
    log_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(log_dir, exist_ok=True)
    with open(os.path.join(log_dir, f"gpu_dist_world{world}.log"), "w") as f:
        f.write(p.stdout)
    assert p.returncode == 0, tail
    assert "ALL ROWS AND TABLES BIT-EXACT vs oracle" in p.stdout, tail
