"""Multi-GPU paths on real GPUs (skipped on a single-GPU box; the 2-rank CPU version is in
test_synth_and_sharding.py): flight shards and owned row bands + NCCL all-gather, see multigpu_worker.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_shards_and_row_bands():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
