"""-m gpu, needs >= 2 GPUs (skipped otherwise): the multi-GPU verbs of the C ABI (lt_b200_index_sharded, lt_b200_write_blocks_sharded: NCCL
all-gather of the chunk tables, dedup split by hash, global block plan, chunk exchange point to point) over two GPUs under torchrun —
VersionIndex and every StoredBlock byte-identical to the single-process reference upsync (tests/tools_multi_gpu_upsync.py)."""
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


@pytest.mark.parametrize("codec,extra,env", [("lz4", [], {}), ("zstd", [], {}), ("none", [], {}), ("zstd", ["--single-file"], {}),
                                             # block boundaries pulled back to the ownership boundaries (at most 1 MiB imported per boundary)
                                             ("lz4", [], {"LT_B200_EXCHANGE_CAP_MB": "1"}), ("lz4", [], {"LT_B200_EXCHANGE_CAP_MB": "0"})])
def test_sharded_upsync_two_gpus(codec, extra, env):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "tools_multi_gpu_upsync.py"), "--gib", "0.25", "--codec", codec] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **env))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "VERIFY OK" in r.stderr
