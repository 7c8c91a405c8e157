"""Multi-GPU parity of the sharded search behind the C ABI, on real NCCL ranks (one process per GPU, torchrun).
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_sharded_nccl_gpu.py -m gpu`); on a 1-GPU box the test is
skipped and the same collective is asserted inside bench.py's warm-up instead (planted neighbours across shards)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_search_and_live_merge_on_nccl_ranks(native_lib, cuda_device, world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, torch.cuda.device_count()))
    env = dict(os.environ)
    env.setdefault("NCCL_DEBUG", "WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + world), os.path.join(ROOT, "tests", "sharded_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "sharded parity ok: world %d" % world in r.stdout
