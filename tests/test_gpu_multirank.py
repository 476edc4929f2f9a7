"""Multi-GPU parity: a 2-rank NCCL run of the sharded path equals the single-rank
result -- confusion matrices and histograms bit for bit, loss within 1e-4 relative.  Reference loop being
sharded: test.py:52-115 (images), utils/evaluate.py:158-174 (one coverage injection per evaluation),
utils/profile.py:98-111 (tiles), models/model.py:317-325 (loss of one training step).
Skipped on a box with fewer than two GPUs (the default `gpurun` box has one; `gpurun --gpus 2`)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_sharded_path_equals_single_rank(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29100 + os.getpid() % 1500
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multirank_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert res["world"] == world
    # integers: bit-identical for any GPU count, and exactly ONE coverage injection (sum = pixel count)
    assert res["conf_equal"] and res["conf_contig_equal"]
    assert res["conf_sum"] == res["conf_expected_sum"]
    assert res["hist_equal"] and res["probs_equal"] and res["mean_std_close"]
    # floating point: loss within 1e-4 relative (north_star), gradient of the local logits likewise
    assert res["loss_rel"] <= 1e-4
    assert res["grad_rel"] <= 2e-3 and res["grad_ddp_rel"] <= 2e-3
    # one DDP training step: averaged parameter gradients == the single large batch's (fp32 convolutions)
    assert res["n_param_grads"] > 300 and res["param_grad_rel"] <= 1e-3
    assert res["launches"] > 0
    # the single-launch data-parallel loss (partials all-reduced inside the kernel over peer memory) equals the
    # NCCL route: f64 sums in another order, identical bits on every rank
    if res["dp_fused_available"]:
        assert res["dp_partials_rel"] <= 1e-12 and res["dp_loss_rel"] <= 1e-6 and res["dp_grad_rel"] <= 1e-5
        assert res["dp_partials_same_on_all_ranks"]
