"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped on a single-GPU box): launches tools/dist_check.py under
torchrun with NCCL — both plans (broadcast build, radix partition + all-to-all) against the oracle."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def test_two_gpu_plans_match_oracle(lib):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("single-GPU box: nothing to launch (the N>1 host logic is covered by tests/test_dist_cpu.py, gloo)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tools" / "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)         # >= 2 GPUs: this must RUN and pass, never skip
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("parity=OK") == 3 * 8 and "FAIL" not in r.stdout, r.stdout[-3000:]
