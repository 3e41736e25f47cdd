"""-m gpu: the N > 1 path on real GPUs (needs >= 2 devices; skipped on a 1-GPU box).
Launches tests/dist_worker.py under torch.distributed.run, one rank per GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("gemm_mode", [0, 1])
def test_sharded_step_matches_single_gpu(gemm_mode):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    env = dict(os.environ, NVSM_TEST_GEMM_MODE=str(gemm_mode))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + gemm_mode), os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTI_GPU_OK" in res.stdout


def _run_worker(world, port, env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    return res.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_allgather_sparse_mode_matches_single_gpu_trajectory(world):
    """NVSM_SPARSE_ALLGATHER: `world` ranks x 3 steps == 1 GPU x 3 steps on the whole batch, all five optimisers. Strict: no
    retry. The intermittent failure of round 1 was root-caused with the worker's soak mode (NVSM_TEST_SOAK, 12 passes x 2
    ranks plus a 1-rank control without any exchange, gpurun_out/soak_r2b.log / soak_control_r2c.log): with hard_tanh in
    the Adam combinations 3-5 of 12 passes mismatch -- and 4 of 8 with ONE rank, where nothing is exchanged -- while
    with tanh 0 of 24 do. It is the clip boundary's 0/1 derivative flipping under a different batch-norm summation
    order, amplified by Adam's normalised step; not an ordering race. The Adam combinations run with tanh."""
    if _ngpus() < world:
        pytest.skip("needs at least %d GPUs" % world)
    env = dict(os.environ, NVSM_TEST_GEMM_MODE="0", NVSM_TEST_SPARSE_MODE="1")
    out = _run_worker(world, 29610 + world, env)
    assert "MULTI_GPU_OK world=%d" % world in out and "sparse=allgather" in out


def test_allgather_trajectory_with_grad_transform_over_nvlink_inboxes():
    """Same trajectory check with the opt-in NCCL-free exchange of grad_transform in the fused step (NVSM_FUSED_GT=1:
    gt_reduce_push_kernel -> transform_update_kernel's flag wait + rank-ordered inbox sum)."""
    if _ngpus() < 2:
        pytest.skip("needs at least 2 GPUs")
    env = dict(os.environ, NVSM_TEST_GEMM_MODE="0", NVSM_TEST_SPARSE_MODE="1", NVSM_FUSED_GT="1")
    out = _run_worker(2, 29630, env)
    assert "MULTI_GPU_OK world=2" in out and "sparse=allgather" in out


@pytest.mark.parametrize("world", [4, 8])
def test_sharded_step_matches_single_gpu_wide(world):
    """The sharded-step check of test_sharded_step_matches_single_gpu on 4 and 8 ranks (fp32 GEMMs)."""
    if _ngpus() < world:
        pytest.skip("needs at least %d GPUs" % world)
    out = _run_worker(world, 29620 + world, dict(os.environ, NVSM_TEST_GEMM_MODE="0"))
    assert "MULTI_GPU_OK world=%d" % world in out
