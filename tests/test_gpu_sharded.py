"""-m gpu: one ChaCha20 trace proved by 2 GPUs together (BASELINE cfg-5 mechanism at a test size); skipped on 1-GPU boxes.
Both exchanges are covered: peer-window stores fused into the last transform pass (the default when CUDA IPC works) and the
NCCL grouped send/recv all-to-all (S2C_NO_P2P=1), plus the single-stream variant of the former."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env,port", [({}, 29541), ({"S2C_NO_P2P": "1"}, 29542), ({"S2C_P2P_1STREAM": "1"}, 29543)])
def test_two_rank_sharded_proof_equals_single_gpu_proof(env, port):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_proof_worker.py"), "16"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, **env))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED_OK" in out.stdout
