"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the gathered containers, rank by rank.

Launches tests/multi_gpu_worker.py with one process per GPU.  Every rank decodes every other rank's shard out of
the gathered container -- SlotGather (ctr_gather_*), ctr_gather_compressed_nccl and all_gather_compressed -- and
compares words and symbols with the CPU oracle."""
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


def _gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_every_rank_decodes_every_other_ranks_shard(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0 and "MULTI_GPU_OK" in proc.stdout, proc.stdout[-2000:] + proc.stderr[-4000:]
