"""The process-per-GPU multi-GPU layer on REAL peers (skips on a one-GPU box): tools/multigpu_check.py under torchrun with one
rank per visible GPU — replica mode is the ground truth on every rank, the fused partitioned path (records and ids as
peer-memory stores over NVLink, ordering by device-side flags, CUDA IPC) and the plain NCCL exchange must return exactly the
same ids for every rank's reads, counters equal."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ret", ["stream", "pull", "stream+ahead", "pull+overlap", "direct", "legacy"])
def test_partition_ids_equal_replica_on_real_peers(ret):
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    env = dict(os.environ, BLIGHT_CHECK_GENOME="20000000", BLIGHT_CHECK_READS="400000", BLIGHT_CHECK_SUB=str(8 << 20), BLIGHT_CHECK_REPS="2",
               **({"BLIGHT_PART_PIPELINE": "legacy"} if ret == "legacy" else {"BLIGHT_PART_RETURN": ret.split("+")[0]}))
    if "+" in ret:
        env["BLIGHT_PART_ORDER"] = ret.split("+")[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == n
    assert out["fused_ids_equal_replica"] and out["fused_counters_equal_replica"] and out["plain_ids_equal_replica"], out
