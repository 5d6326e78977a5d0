"""ncu target: GAE at two bandwidth-relevant shapes (128 x 2^20 and 128 x 2^16), 3 calls each."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minppo_b200.learner import Memory, calculate_gae
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for T, N in ((128, 1 << 20), (128, 1 << 16)):
    m = Memory(torch.rand(T, N, device=dev, generator=g) < 0.01, None, torch.randn(T, N, device=dev, generator=g),
               torch.randn(T, N, device=dev, generator=g), None, None)
    lv = torch.randn(N, device=dev, generator=g)
    for _ in range(3):
        calculate_gae(m, lv, 0.99, 0.95)
    torch.cuda.synchronize()
