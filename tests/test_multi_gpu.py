"""
Parity of the NCCL multi-GPU path on hardware (needs >= 2 GPUs; the driver's one-GPU box skips).

`tools/check_sharded.py` runs under torchrun: the row-sharded evaluation (`dftd3_sharded`: cached
engine, in-place all-gathers / all-reduces, CUDA graph) against the same structure on one GPU and
against the blocked fp64 C oracle, unit and random cotangents, several steps with moving positions.
"""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(world, *extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "check_sharded.py"), *extra]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line)


def _world():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    return 8 if n >= 8 else (4 if n >= 4 else 2)


@pytest.mark.timeout(900)
def test_sharded_two_body_matches_single_gpu_and_oracle():
    out = _run(_world(), "--nwater", "1500", "--oracle")
    assert out["max_rel_err_vs_single_gpu"] <= 1e-10, out
    assert out["max_rel_err_vs_oracle"] <= 1e-10, out


@pytest.mark.timeout(900)
def test_sharded_three_body_matches_single_gpu_and_oracle():
    out = _run(_world(), "--nwater", "600", "--atm", "--cutoff", "18", "--oracle")
    assert out["max_rel_err_vs_single_gpu"] <= 1e-10, out
    assert out["max_rel_err_vs_oracle"] <= 1e-10, out
