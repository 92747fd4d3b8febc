"""Multi-GPU parity (-m gpu): one process per GPU over NCCL.  Needs >= 2 visible devices; on a
single-GPU box these tests are skipped (the single-GPU parity suite still runs)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    from spin_ed_b200 import ffi

    return ffi.deviceCount()


@pytest.mark.parametrize("shard_build", [False, True])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_build_matvec_and_eigh_match_oracle(world, shard_build):
    """shard_build: force the enumeration to be sharded over the ranks (small sectors are otherwise
    enumerated redundantly by every rank, without communication) and force two exchange rounds with
    three source classes (shards this small would otherwise take one NCCL all-gather)."""
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ)
    if shard_build:
        env["SPED_BUILD_SHARD_MIN"] = "0"
        env["SPED_REMOTE_GROUPS"] = "2"
    names = ["heisenberg_chain_10", "heisenberg_square_4x4", "chain_8_k1_complex", "heisenberg_kagome_12",
             "heisenberg_square_5x5", "ring_4site_nosym"]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mp_worker.py")] + names
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)  # a collective mismatch hangs: fail fast
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MP_WORKER_OK" in out.stdout


@pytest.mark.parametrize("fail_rank", [None, 1])
def test_copy_engine_exchange_and_its_fallback(fail_rank):
    """SPED_EXCHANGE=ce: the shard exchange runs on the copy engines over IPC-mapped peer memory; when
    one rank cannot set its side up (SPED_EXCHANGE_TEST_FAIL) every rank falls back to NCCL together."""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, SPED_EXCHANGE="ce", SPED_LOG="1")
    if fail_rank is not None:
        env["SPED_EXCHANGE_TEST_FAIL"] = str(fail_rank)
    names = ["heisenberg_square_4x4", "chain_8_k1_complex", "heisenberg_square_5x5"]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "mp_worker.py")] + names
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MP_WORKER_OK" in out.stdout
    log = out.stdout + out.stderr
    if fail_rank is None:
        assert "copy-engine exchange: two send buffers" in log
    else:
        assert "copy-engine exchange not available" in log


def test_two_traversal_cache_build_still_works():
    """Several-class caches are built by one staging traversal plus a placement pass; the counting +
    filling traversals remain as the fallback for when the staging area does not fit (SPED_FILL_STAGED=0)."""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, SPED_FILL_STAGED="0", SPED_FILL_CHUNK_BYTES="400")
    names = ["heisenberg_square_4x4", "chain_8_k1_complex", "heisenberg_square_5x5"]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "mp_worker.py")] + names
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MP_WORKER_OK" in out.stdout
