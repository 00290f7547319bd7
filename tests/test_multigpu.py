"""N-rank vs 1-rank parity of the CUDA path itself (the gloo test only proves that the ORACLE shards
correctly).  Needs >= 2 GPUs on the box: skipped otherwise (the driver's 1-GPU `pytest -m gpu` run skips it; it is run
with `gpurun --gpus 2`, and bench.py prints the same comparison as `multi_gpu_parity` at every N > 1)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_ranks_match_one_rank():
    here = os.path.dirname(os.path.abspath(__file__))
    n = 2 if _ngpu() < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(here, "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("MGPU_RESULT ")][-1]
    res = json.loads(line[len("MGPU_RESULT "):])
    assert res.pop("world") == n
    for k, v in res.items():
        # sums over k-points in a different order: rounding only.  Integer-valued entries (k-point numbers, multiplicities)
        # and the Fermi levels are exact.
        assert v <= 1e-12, (k, v)
