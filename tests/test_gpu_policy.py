"""Policy / value inference for the rollout through the C ABI (minppo_policy_step) against the oracle
(train.py:157-160, 182-183; oracle/ppo_numpy.py policy_step).

Bit-exact: rng' (= split(rng)[0]); the Threefry bits behind the normal draw (checked through the recovered noise).
Tolerances (BF16 operands / FP32 accumulate hidden layers, fp32 heads, as in tests/test_gpu_update.py):
  mean, value        max |d| <= 5e-3 * max|ref|, mean |d| <= 2e-4 * max|ref|  vs the oracle with the same rounding points
  noise eps          |(action - mean) / scale - eps_oracle| <= 2e-6 * max(1, |eps|)   (XLA erf_inv polynomial, ~1 ulp)
  log_prob           <= 1e-4 absolute vs the float64 closed form on the GPU's own (action, mean); <= 1e-4 vs the oracle
"""
import numpy as np
import pytest

from oracle import ppo_numpy as P
from oracle import synth, threefry
from tests.helpers import hyper_to_config

pytestmark = pytest.mark.gpu

CASES = {
    "c1": dict(hp=dict(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4), D=225, A=10),
    "c2_envs": dict(hp=dict(num_envs=2048, num_steps=2, num_minibatches=4, update_epochs=1), D=225, A=10),
    "ragged": dict(hp=dict(num_envs=24, num_steps=16, num_minibatches=3, update_epochs=2, hidden_size=128, num_layers=1),
                   D=37, A=3),
    "deep": dict(hp=dict(num_envs=300, num_steps=4, num_minibatches=2, update_epochs=1, hidden_size=192, num_layers=3,
                         use_tanh=False, prng_mode=threefry.PARTITIONABLE), D=256, A=16),
}


def _setup(case, device, log_std=None):
    import torch

    from minppo_b200.learner import Learner

    c = CASES[case]
    hp = P.Hyper(anneal_lr=False, **c["hp"])
    pr = synth.make_problem(hp, c["D"], c["A"], seed=11)
    if log_std is not None:
        pr["params"]["params"]["log_std"] = np.asarray(log_std, np.float32)
    learner = Learner(hyper_to_config(hp), c["D"], c["A"], device)
    flat = torch.as_tensor(P.flatten_params(pr["params"], hp.num_layers)).to(device)
    g = np.random.default_rng(77)
    obs = g.standard_normal((hp.num_envs, c["D"])).astype(np.float32)
    return hp, pr, learner, flat, obs


@pytest.mark.parametrize("case", list(CASES))
def test_policy_step_vs_oracle(case, cuda_device):
    import torch

    A = CASES[case]["A"]
    hp, pr, learner, flat, obs = _setup(case, cuda_device, log_std=np.linspace(-0.7, 0.3, A))
    rng = torch.as_tensor(pr["rng"].view(np.int32)).to(cuda_device)
    action, log_prob, value, rng2, mean = learner.policy_step(flat, torch.as_tensor(obs).to(cuda_device), rng, want_mean=True)
    learner.check()
    action, log_prob, value, mean = (x.cpu().numpy() for x in (action, log_prob, value, mean))
    rng2 = rng2.cpu().numpy().view(np.uint32)
    p32 = P.tree_like(pr["params"], lambda x: x.astype(np.float32))
    a_o, lp_o, v_o, rng_o, m_o = P.policy_step(p32, obs, pr["rng"], hp, hp.prng_mode, gemm="bf16")
    assert np.array_equal(rng2, rng_o)
    for got, ref, name in ((mean, m_o, "mean"), (value, v_o, "value")):
        scale = np.abs(ref).max()
        d = np.abs(got - ref)
        assert d.max() <= 5e-3 * scale and d.mean() <= 2e-4 * scale, (name, d.max(), d.mean(), scale)
    ls = pr["params"]["params"]["log_std"].astype(np.float64)
    _, akey = threefry.split(pr["rng"], 2, hp.prng_mode)
    eps_o = threefry.normal_f32(akey, action.size, hp.prng_mode, erfinv="xla").reshape(action.shape)
    eps = (action.astype(np.float64) - mean.astype(np.float64)) / np.exp(ls)
    assert np.all(np.abs(eps - eps_o) <= 2e-6 * np.maximum(1.0, np.abs(eps_o))), np.abs(eps - eps_o).max()
    z = (action.astype(np.float64) - mean) / np.exp(ls)
    lp64 = (-0.5 * z * z - 0.5 * np.log(2 * np.pi)).sum(-1) - ls.sum()
    assert np.abs(log_prob - lp64).max() <= 1e-4
    assert np.abs(log_prob - lp_o).max() <= 1e-4
    learner.close()


def test_policy_step_modes(cuda_device):
    """No sampling (action = mode), critic-only bootstrap value, and the weights-current fast path."""
    import torch

    hp, pr, learner, flat, obs = _setup("c1", cuda_device)
    dobs = torch.as_tensor(obs).to(cuda_device)
    rng = torch.as_tensor(pr["rng"].view(np.int32)).to(cuda_device)
    a1, lp1, v1, r1, m1 = learner.policy_step(flat, dobs, rng, want_mean=True)
    a2, lp2, v2, r2, m2 = learner.policy_step(flat, dobs, rng, weights_current=True, want_mean=True)
    for x, y in ((a1, a2), (lp1, lp2), (v1, v2), (r1, r2), (m1, m2)):
        assert torch.equal(x, y)
    a0, lp0, v0, r0, m0 = learner.policy_step(flat, dobs, None, want_mean=True)
    assert r0 is None and torch.equal(a0, m0) and torch.equal(m0, m1) and torch.equal(v0, v1)
    A = pr["act_dim"]
    ls = pr["params"]["params"]["log_std"].astype(np.float64)
    assert np.allclose(lp0.cpu().numpy(), -0.5 * A * np.log(2 * np.pi) - ls.sum(), atol=1e-5)
    vb = learner.bootstrap_value(flat, dobs)
    assert torch.equal(vb, v1)
    # a different key gives a different draw; the same key the same one (pure function of its inputs)
    rng_b = torch.as_tensor(np.array([0, 7], np.uint32).view(np.int32)).to(cuda_device)
    a3 = learner.policy_step(flat, dobs, rng_b)[0]
    assert not torch.equal(a3, a1)
    assert torch.equal(learner.policy_step(flat, dobs, rng)[0], a1)
    learner.check()
    learner.close()


def test_policy_step_rejects_bad_arguments(cuda_device):
    import torch

    from minppo_b200 import _lib

    hp, pr, learner, flat, obs = _setup("c1", cuda_device)
    dobs = torch.as_tensor(obs).to(cuda_device)
    with pytest.raises(ValueError):
        learner.policy_step(flat, dobs[:-1].contiguous(), None)
    rng = torch.as_tensor(pr["rng"].view(np.int32)).to(cuda_device)
    val = torch.empty(hp.num_envs, device=cuda_device)
    s = torch.cuda.current_stream(cuda_device).cuda_stream
    # key_out aliasing key_in, and a call that asks for nothing
    assert learner.lib.minppo_policy_step(learner._h, flat.data_ptr(), dobs.data_ptr(), rng.data_ptr(), rng.data_ptr(),
                                          None, None, val.data_ptr(), None, 0, s) == _lib.ERR_ARG
    assert learner.lib.minppo_policy_step(learner._h, flat.data_ptr(), dobs.data_ptr(), None, None, None, None, None,
                                          None, 0, s) == _lib.ERR_ARG
    learner.close()


def test_rollout_log_prob_matches_first_epoch(cuda_device):
    """A trajectory whose value / log_prob come from policy_step is seen by the learner's first minibatch with
    ratio == 1 and v == v_old (same weight images, same rounding points): the clipped and unclipped terms coincide,
    so value_loss = 0.5 * mean((v_old - tgt)^2) and actor_loss = -mean(normalised adv) ~ 0."""
    import torch

    from minppo_b200.learner import Learner, Memory, TrainState

    hp = P.Hyper(num_envs=256, num_steps=8, num_minibatches=1, update_epochs=1, anneal_lr=False)
    D, A = 225, 10
    pr = synth.make_problem(hp, D, A, seed=4)
    learner = Learner(hyper_to_config(hp), D, A, cuda_device)
    ts = TrainState.create(P.flatten_params(pr["params"], hp.num_layers), cuda_device)
    obs = torch.as_tensor(pr["traj"]["obs"]).to(cuda_device)
    rng = torch.as_tensor(pr["rng"].view(np.int32)).to(cuda_device)
    acts, lps, vals = [], [], []
    for t in range(hp.num_steps):
        a, lp, v, rng, _ = learner.policy_step(ts.params, obs[t].contiguous(), rng, weights_current=t > 0)
        acts.append(a); lps.append(lp); vals.append(v)
    mem = Memory(done=torch.as_tensor(pr["traj"]["done"]).to(cuda_device), action=torch.stack(acts), value=torch.stack(vals),
                 reward=torch.as_tensor(pr["traj"]["reward"]).to(cuda_device), log_prob=torch.stack(lps), obs=obs)
    last_val = learner.bootstrap_value(ts.params, obs[-1].contiguous(), weights_current=True)
    ts, _, losses = learner.update(ts, mem, last_val, rng)
    learner.check()
    total, value_loss, actor_loss, _ = losses.cpu().numpy()[0, 0]
    tgt = learner.read("targets").cpu().numpy()
    v_old = torch.stack(vals).cpu().numpy()
    assert abs(actor_loss) < 1e-5
    assert abs(value_loss - 0.5 * np.mean((v_old - tgt) ** 2)) <= 1e-5 * max(1.0, value_loss)
    learner.close()


def test_inference_policy_from_checkpoint(cuda_device, tmp_path):
    """infer.py:17-27: pickle written in the reference's layout -> InferencePolicy.act == Learner.policy_step."""
    import torch

    from minppo_b200.infer import InferencePolicy
    from minppo_b200.params import save_model

    hp, pr, learner, flat, obs = _setup("c1", cuda_device)
    dobs = torch.as_tensor(obs).to(cuda_device)
    ref_a, _, ref_v, _, _ = learner.policy_step(flat, dobs, None)
    path = str(tmp_path / "trained_model.pkl")
    save_model(P.tree_like(pr["params"], lambda x: x.astype(np.float32)), path)
    pol = InferencePolicy.from_checkpoint(path, hyper_to_config(hp), hp.num_envs, cuda_device)
    assert (pol.obs_dim, pol.act_dim) == (CASES["c1"]["D"], CASES["c1"]["A"])
    a, v = pol.act(dobs)
    assert torch.equal(a, ref_a) and torch.equal(v, ref_v)
    a2, v2 = pol.act(dobs)                                     # second call takes the weights-current fast path
    assert torch.equal(a2, ref_a) and torch.equal(v2, ref_v)
    rng = torch.as_tensor(pr["rng"].view(np.int32)).to(cuda_device)
    sa, slp, sv, rng2 = pol.act(dobs, rng)
    ra, rlp, rv, rrng, _ = learner.policy_step(flat, dobs, rng)
    assert torch.equal(sa, ra) and torch.equal(slp, rlp) and torch.equal(rng2, rrng)
    pol.close()
    learner.close()
