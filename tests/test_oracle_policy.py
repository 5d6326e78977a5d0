"""The oracle's policy step (train.py:157-160; oracle/ppo_numpy.py policy_step) against external anchors and closed
forms, and against the committed golden vectors (tests/golden/policy_golden.npz).

jax.random.normal is third-party arithmetic (JAX, unpinned): bits -> uniform -> sqrt(2) * erf_inv(u) with XLA's
single-precision erf_inv.  Anchors: the four normal() values printed in JAX's documentation -- the XLA polynomial
reproduces ALL printed digits, a correctly rounded erf_inv does not (so the anchors really pin the polynomial)."""
import os

import numpy as np
import pytest

from oracle import ppo_numpy as P
from oracle import synth, threefry as tf

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_xla_erfinv_reproduces_jax_doc_values_to_the_printed_digit():
    k = tf.prng_key(0)
    got = [tf.normal_f32(k, 1, erfinv="xla")[0]]
    k, sub = tf.split(k, 2)
    got.append(tf.normal_f32(sub, 1, erfinv="xla")[0])
    k, sub = tf.split(k, 2)
    got.append(tf.normal_f32(sub, 1, erfinv="xla")[0])
    got.append(tf.normal_f32(tf.prng_key(42), 1, erfinv="xla")[0])
    want = ["-0.20584226", "-1.2515389", "-0.58665055", "-0.18471177"]        # as printed by JAX (numpy float32 repr)
    assert [str(np.float32(g)) for g in got] == want


def test_xla_erfinv_close_to_exact_and_handles_edges():
    x = np.linspace(-0.9999999, 0.9999999, 200001).astype(np.float32)
    a, b = tf.erfinv_xla_f32(x), tf._erfinv_f32(x)
    assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)) < 1e-5          # Giles: ~6e-6 relative in the tails
    assert tf.erfinv_xla_f32(np.float32([0.0]))[0] == 0.0
    e = tf.erfinv_xla_f32(np.float32([1.0, -1.0]))
    assert np.isposinf(e[0]) and np.isneginf(e[1])


@pytest.mark.parametrize("mode", [tf.LEGACY, tf.PARTITIONABLE])
def test_normal_moments_and_range(mode):
    z = tf.normal_f32(tf.prng_key(3), 200000, mode, erfinv="xla")
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert np.isfinite(z).all() and np.abs(z).max() < 6.0


@pytest.mark.parametrize("mode", [tf.LEGACY, tf.PARTITIONABLE])
def test_policy_step_closed_forms(mode):
    hp = P.Hyper(num_envs=32, num_steps=4, num_minibatches=4, update_epochs=1, anneal_lr=False, prng_mode=mode)
    pr = synth.make_problem(hp, seed=5)
    p64 = P.tree_like(pr["params"], lambda x: x.astype(np.float64))
    p64["params"]["log_std"] = np.linspace(-1.0, 0.5, pr["act_dim"])
    obs = pr["traj"]["obs"][0].astype(np.float64)
    a, lp, v, rng2, mean = P.policy_step(p64, obs, pr["rng"], hp, mode)
    # rng' is the first half of split(rng); the draw comes from the second (train.py:158)
    r2, akey = tf.split(pr["rng"], 2, mode)
    assert np.array_equal(rng2, r2)
    eps = tf.normal_f32(akey, a.size, mode, erfinv="xla").reshape(a.shape).astype(np.float64)
    ls = p64["params"]["log_std"]
    assert np.allclose(a, mean + np.exp(ls) * eps, rtol=0, atol=1e-12)
    # log N(a; mean, sigma) of a = mean + sigma eps is -sum(eps^2)/2 - A log(2 pi)/2 - sum(log sigma)
    want = -0.5 * (eps ** 2).sum(-1) - 0.5 * a.shape[1] * np.log(2 * np.pi) - ls.sum()
    assert np.allclose(lp, want, rtol=1e-9, atol=1e-9)
    # value and mean are the plain forward pass; no sampling -> action = mean, rng unchanged
    m2, _, v2, _ = P.actor_critic_forward(p64, obs, hp)
    assert np.array_equal(mean, m2) and np.array_equal(v, v2)
    a0, lp0, _, rng0, _ = P.policy_step(p64, obs, pr["rng"], hp, mode, sample=False)
    assert np.array_equal(a0, mean) and np.array_equal(rng0, pr["rng"])
    assert np.allclose(lp0, -0.5 * a.shape[1] * np.log(2 * np.pi) - ls.sum())


def test_policy_golden():
    z = np.load(os.path.join(G, "policy_golden.npz"))
    for mode, tag in ((tf.LEGACY, "legacy"), (tf.PARTITIONABLE, "partitionable")):
        key = tf.prng_key(1337)
        assert np.array_equal(tf.normal_f32(key, 160, mode, erfinv="xla"), z[f"normal_{tag}_160"])
        assert np.array_equal(tf.normal_f32(key, 7, mode, erfinv="xla"), z[f"normal_{tag}_7"])
        hp = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=False, prng_mode=mode)
        pr = synth.make_problem(hp, seed=1)
        a, lp, v, rng2, mean = P.policy_step(pr["params"], pr["traj"]["obs"][0], pr["rng"], hp, mode)
        assert np.array_equal(rng2, z[f"step_{tag}_rng"])
        for got, k in ((a, "action"), (lp, "log_prob"), (v, "value"), (mean, "mean")):
            assert np.allclose(got, z[f"step_{tag}_{k}"], rtol=1e-6, atol=1e-6), k
    assert not np.array_equal(z["normal_legacy_160"], z["normal_partitionable_160"])
