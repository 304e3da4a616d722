"""Row-sharded frames on real GPUs against the single-GPU frame, with both exchange paths of the C++
graph: peer-memory stores from the downsample kernel (default) and NCCL broadcasts + all-reduce."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("fxaa", [0, 1])
def test_sharded_frame_is_bit_identical(cuda, fxaa, exchange):
    n = _gpu_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the box")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29511 + fxaa + (2 if exchange == "nccl" else 0)), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), "1280", "768", "300", str(fxaa)]
    env = dict(os.environ, GRB_SHARD_EXCHANGE=exchange)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    if exchange == "peer":
        assert "peer-memory exchange unavailable" not in r.stderr, "the box has NVLink peers: the peer path must be the one that ran"
