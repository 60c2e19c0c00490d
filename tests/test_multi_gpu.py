"""The N > 1 path on real GPUs, inside the `-m gpu` suite: spawns one process per GPU (torchrun, NCCL) when the box shows at least two
devices and checks the sharded build against the oracle (tests/multi_gpu_check.py).  On a one-GPU box the test is skipped; the host
logic of the same path is covered on CPU by tests/test_multi_rank_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_build_matches_the_oracle(world):
    n = _devices()
    if n < world:
        pytest.skip(f"{n} CUDA device(s) visible, {world} needed")
    env = dict(os.environ)
    if world > 2:
        env["AC_MULTI_CHECK_QUICK"] = "1"      # cfg2 at full size once (world 2) is enough
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29650 + world), os.path.join(ROOT, "tests", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_GPU_CHECK OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
