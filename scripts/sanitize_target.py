"""Target of the compute-sanitizer runs (scripts/sanitize.sh): small learner updates with use_graph = 0 (kernels
enqueued directly: the sanitizer sees every launch), the layer-wise path, the rollout's policy step, GAE in both
kernel variants and one permutation.  Results are compared with nothing here -- parity is tests/'s job; this
exercises every kernel under memcheck / racecheck / initcheck / synccheck."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from minppo_b200.config import Config                                  # noqa: E402
from minppo_b200.learner import Learner, Memory, TrainState, calculate_gae, permutations   # noqa: E402
from minppo_b200.params import param_count                             # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "all"


def run(name, N, T, M, E, D, A, H=256, L=2, fused=True, tanh=True):
    c = Config()
    c.model.hidden_size, c.model.num_layers, c.model.use_tanh = H, L, tanh
    c.rl.num_env_steps = T
    c.training.num_envs, c.training.num_steps, c.training.num_minibatches, c.training.update_epochs = N, T, M, E
    c.training.anneal_lr = False
    c.learner.use_graph, c.learner.fused = False, fused
    lrn = Learner(c, D, A, dev)
    g = torch.Generator(device=dev).manual_seed(1)
    P = param_count(D, A, H, L)
    ts = TrainState.create((0.05 * np.random.default_rng(0).standard_normal(P)).astype(np.float32), dev)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    mem = Memory(done=torch.rand(T, N, device=dev, generator=g) < 0.05, action=r(T, N, A), value=r(T, N), reward=r(T, N),
                 log_prob=r(T, N) * 0.1 - 10.0, obs=r(T, N, D))
    rng = torch.tensor([0, 1337], dtype=torch.int32, device=dev)
    ts, rng2, losses = lrn.update(ts, mem, r(N), rng)
    lrn.check()
    act, logp, val, rng3, _ = lrn.policy_step(ts.params, r(N, D), rng2)
    bv = lrn.bootstrap_value(ts.params, r(N, D), weights_current=True)
    torch.cuda.synchronize()
    print(name, "ok: last loss", [round(float(x), 5) for x in losses[-1, -1].cpu()], "launches", lrn.launches_per_update(), flush=True)
    lrn.close()


if which in ("all", "fused"):
    run("fused c1-like (5-row minibatches)", 16, 10, 8, 1, 225, 10)
    run("fused 2 tiles", 32, 16, 2, 1, 225, 10)
    run("fused wide (D=415: streamed X slots, A=20: 32-wide heads, 2 optimizer units per thread)", 32, 16, 2, 1, 415, 20)
    run("fused narrow (H=128, D=600, A=17, relu actor)", 32, 8, 2, 1, 600, 17, H=128, tanh=False)
if which in ("all", "layerwise"):
    run("layer-wise ragged", 24, 16, 3, 1, 37, 3, H=128, L=1, fused=False)
    run("layer-wise deep relu", 32, 16, 2, 1, 256, 16, H=192, L=3, fused=False, tanh=False)
if which in ("all", "small"):
    g = torch.Generator(device=dev).manual_seed(2)
    for T, N, ch in ((16, 1024, 0), (64, 4096, 4), (7, 1001, 0)):
        m = Memory(torch.rand(T, N, device=dev, generator=g) < 0.1, None, torch.randn(T, N, device=dev, generator=g),
                   torch.randn(T, N, device=dev, generator=g), None, None)
        calculate_gae(m, torch.randn(N, device=dev, generator=g), 0.99, 0.95, chunks=ch)
    permutations(torch.tensor([0, 1337], dtype=torch.int32, device=dev), 5000, 2)
    torch.cuda.synchronize()
    print("gae + permutations ok", flush=True)
