"""Multi-GPU parity gates that the driver can run: W ranks on W GPUs of one box, launched with torchrun
from inside the test, skipped on a single-GPU box.

Every mode of the gradient exchange must leave the W replicas bit-identical and equal to single-GPU training on the
global batches up to summation order (scripts/dp_parity.py asserts 1e-4 relative after 8 Adam steps, observed 1e-7):
  * fused reduce-scatter + Adam + all-gather kernel over P2P pointers          (MRI_DP_MULTIMEM=0)
  * the same kernel on NVSwitch multimem.ld_reduce / multimem.st (what SCALE runs at N >= 4)   (MRI_DP_MULTIMEM=1)
  * NCCL all-reduce + full Adam                                                (MRI_DP_SHARDED=0)
The fused kernel is bracketed by two symmetric-memory barrier launches (the default); the variant that synchronises the
ranks itself (flag stores / polls in symmetric memory, MRI_DP_INKERNEL_SYNC=1) is covered too.
and the ranks' dense-sweep slabs must tile the single-GPU sweep exactly."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


MODES = {
    "sharded_p2p": {"MRI_DP_SHARDED": "1", "MRI_DP_MULTIMEM": "0", "MRI_DP_INKERNEL_SYNC": "0"},
    "sharded_multimem": {"MRI_DP_SHARDED": "1", "MRI_DP_MULTIMEM": "1", "MRI_DP_INKERNEL_SYNC": "0"},
    "sharded_p2p_inkernel_sync": {"MRI_DP_SHARDED": "1", "MRI_DP_MULTIMEM": "0", "MRI_DP_INKERNEL_SYNC": "1"},
    "sharded_multimem_inkernel_sync": {"MRI_DP_SHARDED": "1", "MRI_DP_MULTIMEM": "1", "MRI_DP_INKERNEL_SYNC": "1"},
    "nccl_allreduce": {"MRI_DP_SHARDED": "0", "MRI_DP_OVERLAP": "0"},
}


@pytest.mark.parametrize("mode", list(MODES))
def test_two_gpu_training_equals_single_gpu_training(mode, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box (gpurun --gpus 2)")
    env = dict(os.environ, **MODES[mode], MRI_DP_PARITY_OUT=str(tmp_path))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "scripts", "dp_parity.py")]
    run = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-4000:]
    res = json.load(open(tmp_path / "dp_parity_w2.json"))
    assert res["world"] == 2 and res["max_abs_replica_diff"] == 0.0 and res["sweep_slabs_tile_exactly"]
    assert res["max_rel_param_diff_vs_single_gpu"] < 1e-5, res
    if mode.startswith("sharded_multimem"):
        assert res["sharded_p2p_adam"] and res["multimem"], res  # the NVLS branch really ran
    elif mode.startswith("sharded_p2p"):
        assert res["sharded_p2p_adam"] and not res["multimem"], res
    else:
        assert not res["sharded_p2p_adam"], res
    if mode.startswith("sharded"):
        assert res["inkernel_sync"] == mode.endswith("inkernel_sync"), res
