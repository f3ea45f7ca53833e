"""Row-sharded mode inside `pytest -m gpu` (which the driver runs on ONE GPU): two ranks are spawned
with torch.distributed.run; with a single visible device they share cuda:0 (gloo rendezvous,
peer-memory transport over same-device CUDA IPC), with two or more they take one GPU each (NCCL
rendezvous, NVLink peer memory).  The worker (tests/multirank_worker.py) checks sharded SpMV in all
three storage formats, the halo push + flag protocol (including a structurally one-sided pattern),
the LL all-reduce, the fused Arnoldi step, sharded Gram-Schmidt, a full Lanczos run against the
reference golden, the LINDEP abort position against the CPU oracle, and node-distributed FEAST.
"""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _spawn(world, cases, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multirank_worker.py")] + list(cases)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    proc = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            start_new_session=True)
    try:
        out, _ = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, 9)      # the exact process group this test started
        out, _ = proc.communicate()
        pytest.fail(f"multirank worker timed out after {timeout}s\n{out[-4000:]}")
    return proc.returncode, out


@pytest.mark.parametrize("cases", [("kernels", "onesided"), ("lanczos", "lindep", "feast")])
def test_two_ranks(rt, cases):
    rc, out = _spawn(2, cases, timeout=900)
    assert rc == 0, out[-6000:]
    assert out.count("PASS (all ranks: PASS") == 2, out[-6000:]
