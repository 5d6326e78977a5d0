"""GPU probe: run the tcgen05 GEMM engine in one operand mode over a few shapes, print errors.
Each mode runs in its own process (scripts/gpu_round.sh) so that a trap in one cannot poison the rest."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_gemm import _run

mode = int(sys.argv[1])
dev = torch.device("cuda:0")
for shape in [(128, 256, 64, 1), (128, 256, 256, 1), (256, 256, 512, 2), (128, 64, 64, 1), (384, 128, 1024, 4), (128, 192, 192, 3)]:
    M, N, K, S = shape
    try:
        err = _run(mode, M, N, K, S, dev, seed=1)
        print(f"mode {mode} M={M} N={N} K={K} splits={S}: rel err {err:.3e}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"mode {mode} shape {shape}: EXCEPTION {type(e).__name__}: {e}", flush=True)
        break
