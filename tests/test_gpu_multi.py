"""-m gpu, needs >= 2 GPUs (skipped otherwise): CreateVersionIndex + WriteContent sharded over two GPUs
(tests/tools_multi_gpu_write.py: NCCL allgather of the chunk tables, blocks sharded by owner, foreign chunks of straddling blocks
moved point to point) — every StoredBlock byte-identical to the single-process CPU upsync."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("codec", ["lz4", "zstd"])
def test_write_content_two_gpus(codec):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "tools_multi_gpu_write.py"), "--gib", "0.25", "--codec", codec, "--verify"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "VERIFY OK" in r.stderr
