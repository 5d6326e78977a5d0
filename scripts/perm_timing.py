"""Timing of the on-device permutation generator (Threefry bits + stable radix sort) at the sizes the env-sharded
learner meets: E epochs x B elements sorted by ONE rank (rank e % W sorts epoch e).  CUDA events, 20 repetitions."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minppo_b200.learner import permutations

dev = torch.device("cuda:0")
rng = torch.tensor([0, 1337], dtype=torch.int32, device=dev)
for E, B in ((4, 262144), (2, 524288), (1, 1 << 20), (1, 1 << 21)):
    for _ in range(3):
        permutations(rng, B, E)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        permutations(rng, B, E)
    b.record()
    torch.cuda.synchronize()
    print(f"perm E={E} B={B}: {a.elapsed_time(b) / 20 * 1e3:.1f} us")
