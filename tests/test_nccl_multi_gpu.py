"""Multi-GPU parity as a pytest: spawns `torchrun --nproc-per-node N tests/nccl_dist_check.py` for every N in {2, 4, 8}
the box has GPUs for.  The script shards seeded data over the ranks, runs the sharded reduce / scan (carried offset) /
running max / fused reduce+total / grouped aggregation (both exchanges) / find / differentiable loss / 1brc ingest over
NCCL and compares every rank's slice with the oracle evaluated on the FULL data; it exits non-zero on any mismatch.
(The same logic runs over gloo at world size 2 and 3 in the CPU suite: tests/test_dist_gloo.py.)"""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [2, 4, 8])
def test_sharded_primitives_over_nccl(nproc):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs, box has {torch.cuda.device_count() if torch.cuda.is_available() else 0}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join("tests", "nccl_dist_check.py")]
    env = dict(os.environ)
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-4000:]
    assert "MISMATCH" not in out, out[-4000:]
    assert out.count(": ok") >= nproc * 4, out[-4000:]  # 3 sizes + the ingest line, every rank
