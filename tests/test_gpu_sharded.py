"""GPU: the sharded KRR feed (aqml_b200/sharded.py -- what bench.py runs) against the same problem done directly on one GPU:
kernel widths, every K(train,train) tile, every K(test,train) tile and the prediction (tools/check_sharded.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(cmd, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "-> OK" in out.stdout and "FAILED" not in out.stdout, out.stdout[-2000:]


def test_sharded_feed_one_rank():
    _run([sys.executable, os.path.join("tools", "check_sharded.py")], {"N_TRAIN": "700"})


def test_sharded_feed_two_ranks_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
          "--master-port", str(29600 + os.getpid() % 300), os.path.join("tools", "check_sharded.py")], {"N_TRAIN": "700"})
