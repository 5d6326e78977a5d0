"""use_graph = 0 inside a CALLER's stream capture (include/minppo_b200.h): the mode the XLA-FFI shim depends on -- the
reference's whole training step is one jax.jit (/root/reference/minppo/train.py:306), so XLA may capture the custom
calls into its own command buffer.  minppo_update (with its forked staging branch) and minppo_policy_step are captured
into a torch CUDA graph, replayed, and must reproduce the directly enqueued results bit for bit."""
import numpy as np
import pytest

from oracle import ppo_numpy as P
from oracle import synth
from tests.helpers import hyper_to_config

pytestmark = pytest.mark.gpu


def _state(problem, hp, device):
    import torch

    from minppo_b200.learner import Memory, TrainState

    t = lambda x: torch.as_tensor(np.ascontiguousarray(x)).to(device)
    tr = problem["traj"]
    mem = Memory(done=t(tr["done"]), action=t(tr["action"]), value=t(tr["value"]), reward=t(tr["reward"]),
                 log_prob=t(tr["log_prob"]), obs=t(tr["obs"]))
    ts = TrainState.create(P.flatten_params(problem["params"], hp.num_layers), device)
    rng = torch.as_tensor(problem["rng"].view(np.int32)).to(device)
    return ts, mem, t(problem["last_val"]), rng


@pytest.mark.parametrize("fused", [True, False])
def test_update_and_policy_step_replay_from_a_callers_capture(fused, cuda_device):
    import torch

    from minppo_b200.learner import Learner

    hp = P.Hyper(num_envs=64, num_steps=32, num_minibatches=4, update_epochs=2, anneal_lr=True)
    pr = synth.make_problem(hp, 225, 10, seed=5, done_p=0.02)
    cfg = hyper_to_config(hp, use_graph=False, fused=fused)
    obs = torch.as_tensor(np.random.default_rng(0).standard_normal((hp.num_envs, 225)).astype(np.float32)).to(cuda_device)

    # reference: directly enqueued
    lrn = Learner(cfg, 225, 10, cuda_device)
    ts, mem, lv, rng = _state(pr, hp, cuda_device)
    losses = torch.empty((hp.update_epochs, hp.num_minibatches, 4), device=cuda_device)
    rng_out = torch.empty_like(rng)
    lrn.update(ts, mem, lv, rng, losses, rng_out)
    act, logp, val, rng2, _ = lrn.policy_step(ts.params, obs, rng_out)
    lrn.check()
    want = [x.clone() for x in (ts.params, ts.mu, ts.nu, ts.step, losses, rng_out, act, logp, val, rng2)]
    lrn.close()

    # the same calls captured into the caller's graph, then replayed
    lrn = Learner(cfg, 225, 10, cuda_device)
    ts, mem, lv, rng = _state(pr, hp, cuda_device)
    p0, m0, n0, s0 = ts.params.clone(), ts.mu.clone(), ts.nu.clone(), ts.step.clone()
    losses = torch.zeros((hp.update_epochs, hp.num_minibatches, 4), device=cuda_device)
    rng_out = torch.zeros_like(rng)
    side = torch.cuda.Stream(device=cuda_device)
    side.wait_stream(torch.cuda.current_stream(cuda_device))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            lrn.update(ts, mem, lv, rng, losses, rng_out)
            act, logp, val, rng2, _ = lrn.policy_step(ts.params, obs, rng_out)
    torch.cuda.current_stream(cuda_device).wait_stream(side)
    torch.cuda.synchronize(cuda_device)
    assert torch.equal(ts.params, p0) and int(ts.step.item()) == 0, "capture must not execute anything"
    for rep in range(2):                                   # two replays from the same initial state: identical results
        ts.params.copy_(p0); ts.mu.copy_(m0); ts.nu.copy_(n0); ts.step.copy_(s0)
        graph.replay()
        torch.cuda.synchronize(cuda_device)
        lrn.check()
        got = (ts.params, ts.mu, ts.nu, ts.step, losses, rng_out, act, logp, val, rng2)
        for i, (g, w) in enumerate(zip(got, want)):
            assert torch.equal(g, w), (rep, i)
    lrn.close()


def test_host_pipeline_equals_sequential_updates(cuda_device):
    """HostPipeline (pinned-host trajectories in, params / losses out, two updates in flight) produces exactly what the
    same updates give when called one after the other on device-resident inputs."""
    import torch

    from minppo_b200.learner import HostPipeline, Learner

    hp = P.Hyper(num_envs=64, num_steps=32, num_minibatches=4, update_epochs=2, anneal_lr=True)
    pr = synth.make_problem(hp, 225, 10, seed=6, done_p=0.02)
    cfg = hyper_to_config(hp)
    lrn = Learner(cfg, 225, 10, cuda_device)
    ts, mem, lv, rng = _state(pr, hp, cuda_device)
    want = []
    r = rng
    for _ in range(3):
        ts, r_out, losses = lrn.update(ts, mem, lv, r)
        r = r_out.clone()
        want.append((losses.cpu().numpy().copy(), ts.params.cpu().numpy().copy()))
    lrn.check()
    lrn.close()

    lrn = Learner(cfg, 225, 10, cuda_device)
    ts, mem, lv, rng = _state(pr, hp, cuda_device)
    host = {"obs": mem.obs, "action": mem.action, "value": mem.value, "reward": mem.reward, "log_prob": mem.log_prob,
            "done": mem.done.view(torch.uint8), "last_val": lv}
    host = {k: v.cpu().pin_memory() for k, v in host.items()}
    pipe = HostPipeline(lrn, ts, rng)
    got = []
    for i in range(3):
        pipe.submit(host)
        if i >= 1:
            l, p = pipe.result()
            got.append((l.copy(), p.copy()))
    l, p = pipe.result()
    got.append((l.copy(), p.copy()))
    lrn.check()
    for i, ((lw, pw), (lg, pg)) in enumerate(zip(want, got)):
        assert np.array_equal(lw, lg), i
        assert np.array_equal(pw, pg), i
    with pytest.raises(RuntimeError):
        pipe.result()
    lrn.close()
