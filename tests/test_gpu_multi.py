"""Multi-GPU parity through torchrun, collected only where >= 2 GPUs are visible (the driver's `-m gpu` box has one:
skipped there; `scripts/gpu_multi_check.sh` / `gpu_multi2.sh` run the same script under `gpurun --gpus N`).
See tests/multigpu/check_sharded_update.py for what is checked; the CPU twin is tests/test_distributed_cpu.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_update_parity_all_visible_gpus():
    import torch

    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip("needs >= 2 GPUs on one box")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "tests", "multigpu", "check_sharded_update.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MULTIGPU PARITY OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
