"""Multi-GPU checks (need >= 2 visible GPUs: `gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import umgap_b200.capi as c
    return c.device_count()


def test_key_range_sharded_index_over_peer_memory():
    n = _ngpus()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
                        os.path.join(ROOT, "tests", "dist_sharded_check.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, timeout=600)
    out = p.stdout.decode()
    assert p.returncode == 0, p.stderr.decode()[-3000:]
    assert "sharded ok rank 0/2" in out and "sharded ok rank 1/2" in out


def _torchrun(script, nproc):
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                        "--master-addr", "127.0.0.1", "--master-port", str(29900 + os.getpid() % 300 + nproc),
                        os.path.join(ROOT, "tests", script)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-3000:] + p.stdout.decode()[-2000:]
    return p.stdout.decode()


def test_routed_exchange_single_rank():
    """The exchange path with one rank (one shard = the whole table, the exchange degenerates to the local bucket copy):
    pack / lookup / scatter kernels and the two-round sampled form against the fused path, on any GPU box."""
    if _ngpus() < 1:
        pytest.skip("needs a GPU")
    assert "routed ok rank 0/1" in _torchrun("dist_routed_check.py", 1)


def test_routed_exchange_two_ranks():
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _torchrun("dist_routed_check.py", 2)
    assert "routed ok rank 0/2" in out and "routed ok rank 1/2" in out
