"""Pin the PPO-math oracle (oracle/ppo_numpy.py) on independent anchors: torch.autograd for every
gradient, closed forms for GAE / Gaussian log-prob / Adam step 1 / clip, and the reference's
own control flow (train.py line numbers in the test names' docstrings)."""
import math

import numpy as np
import pytest
import torch

from oracle import ppo_numpy as P
from oracle import ppo_torch as PT
from oracle import synth, threefry


def _mb(n, D, A, seed, dtype=np.float64):
    g = np.random.default_rng(seed)
    return {"obs": g.standard_normal((n, D)).astype(dtype), "action": g.standard_normal((n, A)).astype(dtype),
            "value": g.standard_normal(n).astype(dtype), "log_prob": (-14 + 0.3 * g.standard_normal(n)).astype(dtype),
            "adv": g.standard_normal(n).astype(dtype), "tgt": g.standard_normal(n).astype(dtype)}


@pytest.mark.parametrize("L,H,tanh,ent", [(2, 64, True, 0.0), (1, 32, True, 0.01), (3, 48, False, 0.02)])
def test_hand_gradients_equal_autograd_fp64(L, H, tanh, ent):
    """jax.value_and_grad(_loss_fn) (train.py:246) restated by hand == torch.autograd, float64."""
    D, A = 19, 5
    hp = P.Hyper(hidden_size=H, num_layers=L, use_tanh=tanh, ent_coef=ent)
    params = P.init_params(D, A, H, L, seed=2)
    mb = _mb(96, D, A, 3)
    # make log_prob_old consistent enough that both clip branches occur
    mean, log_std, _, _ = P.actor_critic_forward(params, mb["obs"], hp)
    lp, _, _ = P.gaussian_log_prob(mean, log_std, mb["action"])
    mb["log_prob"] = lp + 0.4 * np.random.default_rng(0).standard_normal(lp.shape)
    ls, gr = P.loss_and_grads(params, mb, hp)
    ls_t, gr_t = PT.loss_and_grads(PT.to_torch(params, torch.float64), {k: torch.tensor(v) for k, v in mb.items()}, hp)
    np.testing.assert_allclose(np.array(ls, np.float64), np.array([float(x) for x in ls_t]), rtol=1e-12, atol=1e-14)
    for pth in P.leaf_order(L):
        a, b = P.get_leaf(gr, pth), P.get_leaf(gr_t, pth).numpy()
        assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max()), pth
    ratio = np.exp(lp - mb["log_prob"])
    assert (ratio > 1.2).any() and (ratio < 0.8).any() and ((ratio > 0.8) & (ratio < 1.2)).any()


def test_gae_closed_forms():
    """train.py:185-205.  done == 0: adv_t = sum_k (gamma*lam)^k delta_{t+k}; done == 1: adv = r - v."""
    g = np.random.default_rng(0)
    T, N, gam, lam = 12, 5, 0.99, 0.95
    r, v, lv = g.standard_normal((T, N)), g.standard_normal((T, N)), g.standard_normal(N)
    adv, tgt = P.gae(r, v, np.zeros((T, N), bool), lv, gam, lam)
    nxt = np.concatenate([v[1:], lv[None]], 0)
    delta = r + gam * nxt - v
    ref = np.zeros((T, N))
    for t in range(T):
        ref[t] = sum((gam * lam) ** k * delta[t + k] for k in range(T - t))
    np.testing.assert_allclose(adv, ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(tgt, adv + v)
    adv1, tgt1 = P.gae(r, v, np.ones((T, N), bool), lv, gam, lam)
    np.testing.assert_allclose(adv1, r - v)
    np.testing.assert_allclose(tgt1, r)
    # a done at step t cuts the bootstrap and the carry exactly there
    done = np.zeros((T, N), bool)
    done[6] = True
    adv2, _ = P.gae(r, v, done, lv, gam, lam)
    np.testing.assert_allclose(adv2[6], r[6] - v[6])
    np.testing.assert_allclose(adv2[7:], adv[7:])


def test_gae_torch_equals_numpy():
    g = np.random.default_rng(1)
    r, v, lv = g.standard_normal((9, 7)), g.standard_normal((9, 7)), g.standard_normal(7)
    d = g.random((9, 7)) < 0.2
    a, t = P.gae(r, v, d, lv, 0.99, 0.95)
    a2, t2 = PT.gae(torch.tensor(r), torch.tensor(v), torch.tensor(d), torch.tensor(lv), 0.99, 0.95)
    np.testing.assert_allclose(a, a2.numpy(), rtol=1e-13)
    np.testing.assert_allclose(t, t2.numpy(), rtol=1e-13)


def test_gaussian_logprob_and_entropy_closed_form():
    """distrax.MultivariateNormalDiag.log_prob / .entropy (train.py:223, 240) vs scipy."""
    from scipy.stats import multivariate_normal

    g = np.random.default_rng(2)
    A = 6
    mean, ls, a = g.standard_normal((4, A)), 0.3 * g.standard_normal(A), g.standard_normal((4, A))
    lp, _, _ = P.gaussian_log_prob(mean, ls, a)
    for i in range(4):
        ref = multivariate_normal(mean[i], np.diag(np.exp(2 * ls))).logpdf(a[i])
        assert abs(lp[i] - ref) < 1e-10
    assert abs(P.gaussian_entropy(ls, A) - multivariate_normal(np.zeros(A), np.diag(np.exp(2 * ls))).entropy()) < 1e-10


def test_adam_first_step_closed_form_and_clip_branches():
    """optax: step 1 bias correction gives update = -lr * g / (|g| + eps) (SURVEY.md 8c anchor 4);
    clip_by_global_norm scales only when ||g|| >= max_norm (train.py:117)."""
    hp = P.Hyper(anneal_lr=False, opt_lr=1e-3, max_grad_norm=0.5, hidden_size=8, num_layers=1)
    params = P.init_params(3, 2, 8, 1, seed=0)
    g = np.random.default_rng(3)
    small = P.tree_like(params, lambda x: 1e-3 * g.standard_normal(x.shape))          # ||g|| < 0.5
    p1, o1, gn = P.clip_adam_step(params, small, P.init_opt_state(params), hp)
    assert gn < 0.5 and o1["count"] == 1
    for pth in P.leaf_order(1):
        gg = P.get_leaf(small, pth)
        np.testing.assert_allclose(P.get_leaf(p1, pth) - P.get_leaf(params, pth), -1e-3 * gg / (np.abs(gg) + 1e-5), rtol=1e-9)
    big = P.tree_like(params, lambda x: 10.0 * g.standard_normal(x.shape))
    p2, o2, gn2 = P.clip_adam_step(params, big, P.init_opt_state(params), hp)
    assert gn2 > 0.5
    for pth in P.leaf_order(1):
        gg = P.get_leaf(big, pth) / gn2 * 0.5
        np.testing.assert_allclose(P.get_leaf(o2["mu"], pth), 0.1 * gg, rtol=1e-9)
        np.testing.assert_allclose(P.get_leaf(p2, pth) - P.get_leaf(params, pth), -1e-3 * gg / (np.abs(gg) + 1e-5), rtol=1e-9)


def test_adam_chain_equals_torch_optim_adam_over_many_steps():
    """Independent-library anchor for the optimizer beyond step 1: optax.adam(lr, eps=1e-5) (train.py:118-121) is the
    Kingma-Ba update with both bias corrections and eps OUTSIDE the square root -- which is also what torch.optim.Adam
    implements.  25 steps with a changing gradient, clip never triggering (||g|| < max_grad_norm), fp64: the oracle's
    params / mu / nu / count follow torch's to 1e-12."""
    import torch

    hp = P.Hyper(anneal_lr=False, opt_lr=3e-4, max_grad_norm=1e9, hidden_size=8, num_layers=2)
    params = P.init_params(5, 3, 8, 2, seed=2)
    order = P.leaf_order(2)
    tparams = [torch.tensor(P.get_leaf(params, pth), dtype=torch.float64, requires_grad=True) for pth in order]
    topt = torch.optim.Adam(tparams, lr=3e-4, betas=(0.9, 0.999), eps=1e-5)
    opt = P.init_opt_state(params)
    g = np.random.default_rng(5)
    for step in range(25):
        grads = P.tree_like(params, lambda x: (0.3 + 0.1 * step) * g.standard_normal(x.shape))
        for tp, pth in zip(tparams, order):
            tp.grad = torch.tensor(P.get_leaf(grads, pth), dtype=torch.float64)
        topt.step()
        params, opt, _ = P.clip_adam_step(params, grads, opt, hp)
    assert opt["count"] == 25
    for tp, pth in zip(tparams, order):
        np.testing.assert_allclose(P.get_leaf(params, pth), tp.detach().numpy(), rtol=1e-12, atol=1e-14)
        st = topt.state[tp]
        np.testing.assert_allclose(P.get_leaf(opt["mu"], pth), st["exp_avg"].numpy(), rtol=1e-12, atol=1e-16)
        np.testing.assert_allclose(P.get_leaf(opt["nu"], pth), st["exp_avg_sq"].numpy(), rtol=1e-12, atol=1e-18)


def test_gaussian_logprob_equals_torch_distributions():
    """Second independent anchor for distrax.MultivariateNormalDiag(loc, exp(log_std)) (train.py:80-83, 223, 240):
    torch.distributions.Independent(Normal(loc, scale), 1) -- log_prob of a batch and the entropy, fp64."""
    import torch

    g = np.random.default_rng(9)
    mean, log_std, act = g.standard_normal((7, 6)), 0.3 * g.standard_normal(6), g.standard_normal((7, 6))
    d = torch.distributions.Independent(torch.distributions.Normal(torch.tensor(mean), torch.tensor(np.exp(log_std))), 1)
    np.testing.assert_allclose(P.gaussian_log_prob(mean, log_std, act)[0], d.log_prob(torch.tensor(act)).numpy(), rtol=1e-12)
    np.testing.assert_allclose(P.gaussian_entropy(log_std, 6), d.entropy().numpy()[0], rtol=1e-12)


def test_lr_schedule_quirk_F8():
    """train.py:98-101 divides the Adam count by minibatch_size * update_epochs (NOT num_minibatches):
    with "one update" (total_timesteps = 160) the LR reaches zero at step 20 and goes negative."""
    hp = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=True, total_timesteps=160)
    assert hp.num_updates == 1 and hp.minibatch_size == 5
    lr = [P.learning_rate(c, hp) for c in range(128)]
    assert lr[0] == hp.training_lr and lr[19] == hp.training_lr and lr[20] == 0.0 and lr[40] == -hp.training_lr
    assert lr[127] == hp.training_lr * (1 - 6)
    hd = P.Hyper(num_envs=2048, num_steps=10, anneal_lr=True)        # defaults: 1e9 timesteps
    assert hd.num_updates == 48828
    assert P.learning_rate(0, hd) == hd.training_lr
    assert P.learning_rate(10**6, hd) == hd.training_lr * (1 - (10**6 // (640 * 4)) / 48828)
    assert P.learning_rate(5, P.Hyper(anneal_lr=False, opt_lr=7e-4)) == 7e-4    # constant path reads opt.lr (train.py:123)


def test_advantage_normalisation_is_per_minibatch_F7():
    """train.py:235 normalises inside the loss, over the gathered minibatch (population std)."""
    hp = P.Hyper(hidden_size=16, num_layers=1)
    params = P.init_params(4, 2, 16, 1, seed=1)
    mb = _mb(32, 4, 2, 5)
    ls0, _ = P.loss_and_grads(params, mb, hp)
    mb2 = dict(mb)
    mb2["adv"] = 3.0 * mb["adv"] + 11.0          # affine change leaves the normalised advantage unchanged
    ls1, _ = P.loss_and_grads(params, mb2, hp)
    assert abs(ls0[2] - ls1[2]) < 1e-7 * max(1.0, abs(ls0[2]))      # the +1e-8 on std is the only non-invariant term


def test_batch_size_guard():
    """train.py:253-255."""
    hp = P.Hyper(num_envs=10, num_steps=3, num_minibatches=4, update_epochs=1)
    pr = synth.make_problem(P.Hyper(num_envs=10, num_steps=3, num_minibatches=1, update_epochs=1), 5, 2)
    with pytest.raises(ValueError, match="batch_size"):
        P.update(pr["params"], P.init_opt_state(pr["params"]), pr["traj"], pr["last_val"], pr["rng"], hp)


def test_numpy_and_torch_full_update_agree_fp64():
    hp = P.Hyper(num_envs=8, num_steps=6, num_minibatches=4, update_epochs=2, anneal_lr=True, hidden_size=32)
    pr = synth.make_problem(hp, 11, 3, seed=2, done_p=0.1)
    p0 = P.tree_like(pr["params"], lambda x: x.astype(np.float64))
    p1, o1, r1, l1, aux = P.update(p0, P.init_opt_state(p0), pr["traj"], pr["last_val"], pr["rng"], hp)
    pt = PT.to_torch(p0, torch.float64)
    tr = {k: torch.tensor(v if v.dtype == bool else v.astype(np.float64)) for k, v in pr["traj"].items()}
    opt = {"count": 0, "mu": P.tree_like(pt, torch.zeros_like), "nu": P.tree_like(pt, torch.zeros_like)}
    p2, o2, r2, l2, _ = PT.update(pt, opt, tr, torch.tensor(pr["last_val"].astype(np.float64)), pr["rng"], hp)
    assert np.array_equal(r1, r2) and o1["count"] == o2["count"] == 8
    np.testing.assert_allclose(l1, l2.numpy(), rtol=1e-10, atol=1e-12)
    for pth in P.leaf_order(hp.num_layers):
        np.testing.assert_allclose(P.get_leaf(p1, pth), P.get_leaf(p2, pth).numpy(), rtol=1e-9, atol=1e-12)


def test_minibatch_k_is_perm_slice_k():
    """train.py:260-265: flat = t*N + n, minibatch k = perm[k*mb:(k+1)*mb] in that order."""
    hp = P.Hyper(num_envs=4, num_steps=3, num_minibatches=2, update_epochs=1, hidden_size=8, num_layers=1)
    pr = synth.make_problem(hp, 3, 2, seed=0)
    flat = P.flatten_traj({"obs": pr["traj"]["obs"]})["obs"]
    assert np.array_equal(flat[2 * 4 + 1], pr["traj"]["obs"][2, 1])
    _, sub = threefry.split(pr["rng"], 2)
    perm = threefry.permutation(sub, 12)
    _, _, _, _, aux = P.update(pr["params"], P.init_opt_state(pr["params"]), pr["traj"], pr["last_val"], pr["rng"], hp)
    assert np.array_equal(aux["perms"][0], perm)


def test_bf16_rounding_helper():
    x = np.array([1.0, 1.00390625, 1.005859375, -3.1415927, 0.0, 1e-40], np.float32)
    ref = torch.tensor(x).bfloat16().float().numpy()
    assert np.array_equal(P.bf16_round(x), ref)
    g = np.random.default_rng(0).standard_normal(10000).astype(np.float32)
    assert np.array_equal(P.bf16_round(g), torch.tensor(g).bfloat16().float().numpy())


def test_param_layout_and_count():
    """P = 250,133 at D=225, A=10, H=256, L=2; 13 leaves, sorted order."""
    shapes = P.leaf_shapes(225, 10, 256, 2)
    assert len(shapes) == 13 and sum(int(np.prod(s)) for s in shapes) == 250133
    order = P.leaf_order(2)
    assert order[0] == ("MLP_0", "Dense_0", "bias") and order[1] == ("MLP_0", "Dense_0", "kernel")
    assert order[6] == ("MLP_1", "Dense_0", "bias") and order[-1] == ("log_std",)
    params = P.init_params(225, 10, 256, 2)
    flat = P.flatten_params(params, 2)
    back = P.unflatten_params(flat, 225, 10, 256, 2)
    for pth in order:
        assert np.array_equal(P.get_leaf(back, pth), P.get_leaf(params, pth).astype(np.float32))
    assert math.isclose(float(flat[:256].sum()), float(params["params"]["MLP_0"]["Dense_0"]["bias"].astype(np.float32).sum()), rel_tol=1e-6)
