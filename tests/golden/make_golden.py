"""Generate the committed golden vectors under tests/golden/.

The reference cannot be imported in this image (JAX, flax, optax, distrax are not installed;
SURVEY.md F3) and ships no golden vectors of its own (F2), so these are produced by the ORACLE
(oracle/threefry.py, oracle/ppo_numpy.py in float64) after it has been pinned on the external
anchors checked in tests/test_oracle_threefry.py (Random123 KATs, JAX's documented split /
normal values) and tests/test_oracle_ppo.py (torch.autograd, closed forms).  They freeze the
oracle: any later edit that changes its numbers fails tests/test_golden.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ppo_numpy as P  # noqa: E402
from oracle import synth, threefry  # noqa: E402


def main():
    # ---- PRNG / permutation vectors -----------------------------------------------------------
    out = {}
    for mode, tag in ((threefry.LEGACY, "legacy"), (threefry.PARTITIONABLE, "partitionable")):
        key = threefry.prng_key(1337)
        out[f"split_{tag}"] = threefry.split(key, 2, mode)
        for B in (10, 160, 1000, 4097):
            out[f"perm_{tag}_{B}"] = threefry.permutation(threefry.split(key, 2, mode)[1], B, mode)
        rng, keys = threefry.epoch_key_chain(key, 4, mode)
        out[f"chain_rng_{tag}"] = rng
        out[f"chain_keys_{tag}"] = np.stack(keys)
        out[f"bits_{tag}_7"] = threefry.random_bits(key, 7, mode)
        out[f"bits_{tag}_8"] = threefry.random_bits(key, 8, mode)
    np.savez_compressed(os.path.join(HERE, "threefry_golden.npz"), **out)

    # ---- config-1 update (N=16, T=10, M=32, E=4; stand-in D=225, A=10; anneal_lr=False) ---------
    hp = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=False)
    pr = synth.make_problem(hp, seed=1)
    p0 = P.tree_like(pr["params"], lambda x: x.astype(np.float64))
    p1, o1, rng, losses, aux = P.update(p0, P.init_opt_state(p0), pr["traj"], pr["last_val"], pr["rng"], hp)
    flat_out = P.flatten_params(p1, hp.num_layers, np.float64)
    flat_in = P.flatten_params(pr["params"], hp.num_layers, np.float32)
    np.savez_compressed(
        os.path.join(HERE, "c1_update_golden.npz"),
        seed=np.array(1), rng_out=rng, losses=losses, perms=aux["perms"], advantages=aux["advantages"],
        targets=aux["targets"], grad_norms=aux["grad_norms"], count=np.array(o1["count"]),
        # parameters after the update: every 53rd element plus two moments of the whole arena (keeps the fixture small)
        params_out_stride53=flat_out[::53], params_out_sum=np.array(flat_out.sum()), params_out_sumsq=np.array((flat_out ** 2).sum()),
        params_in_sum=np.array(flat_in.astype(np.float64).sum()), params_in_sumsq=np.array((flat_in.astype(np.float64) ** 2).sum()),
        # trajectory inputs, so that the fixture detects a change in oracle/synth.py too
        obs=pr["traj"]["obs"], action=pr["traj"]["action"], value=pr["traj"]["value"], log_prob=pr["traj"]["log_prob"],
        reward=pr["traj"]["reward"], done=pr["traj"]["done"], last_val=pr["last_val"], rng_in=pr["rng"])

    # ---- the annealed-LR quirk: lr per step for "one update" and for the default run ----
    hq = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=True, total_timesteps=160)
    hd = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=True)
    np.savez_compressed(os.path.join(HERE, "lr_schedule_golden.npz"),
                        one_update=np.array([P.learning_rate(c, hq, np.float32) for c in range(128)], np.float32),
                        default=np.array([P.learning_rate(c, hd, np.float32) for c in range(128)], np.float32))

    # ---- policy step (train.py:157-160): normal draws (XLA erf_inv) and one sampled step at N=16 -----------------
    pol = {}
    for mode, tag in ((threefry.LEGACY, "legacy"), (threefry.PARTITIONABLE, "partitionable")):
        key = threefry.prng_key(1337)
        pol[f"normal_{tag}_160"] = threefry.normal_f32(key, 160, mode, erfinv="xla")
        pol[f"normal_{tag}_7"] = threefry.normal_f32(key, 7, mode, erfinv="xla")
        hq = P.Hyper(num_envs=16, num_steps=10, num_minibatches=32, update_epochs=4, anneal_lr=False, prng_mode=mode)
        pq = synth.make_problem(hq, seed=1)
        obs0 = pq["traj"]["obs"][0]
        a, lp, v, rng2, mean = P.policy_step(pq["params"], obs0, pq["rng"], hq, mode)
        pol[f"step_{tag}_action"], pol[f"step_{tag}_log_prob"], pol[f"step_{tag}_value"] = a, lp, v
        pol[f"step_{tag}_rng"], pol[f"step_{tag}_mean"] = rng2, mean
    np.savez_compressed(os.path.join(HERE, "policy_golden.npz"), **pol)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
