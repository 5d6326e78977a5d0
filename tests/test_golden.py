"""The oracle against the committed golden vectors (tests/golden/*.npz, made by
tests/golden/make_golden.py): freezes the oracle's numbers so that the GPU parity tests always
compare against the same thing."""
import os

import numpy as np

from oracle import ppo_numpy as P
from oracle import synth, threefry

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_threefry_golden():
    z = np.load(os.path.join(G, "threefry_golden.npz"))
    for mode, tag in ((threefry.LEGACY, "legacy"), (threefry.PARTITIONABLE, "partitionable")):
        key = threefry.prng_key(1337)
        assert np.array_equal(threefry.split(key, 2, mode), z[f"split_{tag}"])
        sub = threefry.split(key, 2, mode)[1]
        for B in (10, 160, 1000, 4097):
            assert np.array_equal(threefry.permutation(sub, B, mode), z[f"perm_{tag}_{B}"])
        rng, keys = threefry.epoch_key_chain(key, 4, mode)
        assert np.array_equal(rng, z[f"chain_rng_{tag}"]) and np.array_equal(np.stack(keys), z[f"chain_keys_{tag}"])
        assert np.array_equal(threefry.random_bits(key, 7, mode), z[f"bits_{tag}_7"])
        assert np.array_equal(threefry.random_bits(key, 8, mode), z[f"bits_{tag}_8"])
    # the two bit-stream modes really differ
    assert not np.array_equal(z["perm_legacy_160"], z["perm_partitionable_160"])


def test_c1_update_golden_fp64():
    z = np.load(os.path.join(G, "c1_update_golden.npz"))
    hp = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=False)
    pr = synth.make_problem(hp, seed=int(z["seed"]))
    for k in ("obs", "action", "value", "log_prob", "reward", "done"):
        assert np.array_equal(pr["traj"][k], z[k]), k
    assert np.array_equal(pr["last_val"], z["last_val"]) and np.array_equal(pr["rng"], z["rng_in"])
    flat_in = P.flatten_params(pr["params"], hp.num_layers, np.float32).astype(np.float64)
    assert np.isclose(flat_in.sum(), float(z["params_in_sum"]), rtol=1e-12)
    p0 = P.tree_like(pr["params"], lambda x: x.astype(np.float64))
    p1, o1, rng, losses, aux = P.update(p0, P.init_opt_state(p0), pr["traj"], pr["last_val"], pr["rng"], hp)
    assert np.array_equal(rng, z["rng_out"]) and np.array_equal(aux["perms"], z["perms"]) and o1["count"] == int(z["count"]) == 128
    np.testing.assert_allclose(aux["advantages"], z["advantages"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(aux["targets"], z["targets"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(losses, z["losses"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(aux["grad_norms"], z["grad_norms"], rtol=1e-9)
    flat = P.flatten_params(p1, hp.num_layers, np.float64)
    np.testing.assert_allclose(flat[::53], z["params_out_stride53"], rtol=1e-9, atol=1e-12)
    assert np.isclose(flat.sum(), float(z["params_out_sum"]), rtol=1e-9)
    assert np.isclose((flat ** 2).sum(), float(z["params_out_sumsq"]), rtol=1e-9)


def test_c1_update_fp32_and_bf16_modes_stay_close_to_golden():
    """The float32 oracle and the bf16-emulating oracle (what the GPU path is compared with) against
    the float64 golden losses: documents the precision cost of each mode on config 1."""
    z = np.load(os.path.join(G, "c1_update_golden.npz"))
    hp = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=False)
    pr = synth.make_problem(hp, seed=int(z["seed"]))
    scale = np.abs(z["losses"]).max()
    p0 = P.tree_like(pr["params"], lambda x: x.astype(np.float32))
    _, _, _, l32, _ = P.update(p0, P.init_opt_state(p0), pr["traj"], pr["last_val"], pr["rng"], hp, dtype=np.float32)
    assert np.abs(l32 - z["losses"]).max() < 1e-5 * scale
    _, _, _, lbf, _ = P.update(p0, P.init_opt_state(p0), pr["traj"], pr["last_val"], pr["rng"], hp, dtype=np.float32, gemm="bf16")
    assert np.abs(lbf - z["losses"]).max() < 2e-2 * scale


def test_lr_schedule_golden():
    z = np.load(os.path.join(G, "lr_schedule_golden.npz"))
    hq = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=True, total_timesteps=160)
    hd = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=True)
    assert np.array_equal(np.array([P.learning_rate(c, hq, np.float32) for c in range(128)], np.float32), z["one_update"])
    assert np.array_equal(np.array([P.learning_rate(c, hd, np.float32) for c in range(128)], np.float32), z["default"])
    assert z["one_update"][20] == 0.0 and z["one_update"][127] < 0
