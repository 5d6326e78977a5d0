"""ORACLE-side synthetic inputs (test infrastructure).

Synthetic trajectories of the learner's input contract (/root/reference/minppo/train.py:26-33,
``Memory``), produced the way SURVEY.md section 8d prescribes: obs ~ N(0,1); action sampled
from the policy under the SAME parameters the update starts from, so ratio == 1 at step 0 as
in real PPO; value / log_prob from that same forward; reward ~ N(0,1); done ~ Bernoulli(p).

The stompy_pro observation / action sizes are fetched at run time by the reference
(env.py:27-29, 99, 245-261) and are unknown offline: D=225, A=10 is a
DECLARED STAND-IN, not the real robot's shape.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import ppo_numpy as P

STANDIN_OBS_DIM = 225
STANDIN_ACT_DIM = 10


def make_problem(hp: P.Hyper, obs_dim: int = STANDIN_OBS_DIM, act_dim: int = STANDIN_ACT_DIM,
                 seed: int = 0, done_p: float = 0.01, dtype=np.float32) -> Dict:
    g = np.random.default_rng(seed + 1000)
    T, N = hp.num_steps, hp.num_envs
    params = P.init_params(obs_dim, act_dim, hp.hidden_size, hp.num_layers, seed, np.float64)
    obs = g.standard_normal((T, N, obs_dim))
    mean, log_std, value, _ = P.actor_critic_forward(params, obs.reshape(T * N, obs_dim), hp)
    action = mean + np.exp(log_std) * g.standard_normal(mean.shape)
    log_prob, _, _ = P.gaussian_log_prob(mean, log_std, action)
    last_obs = g.standard_normal((N, obs_dim))
    _, _, last_val, _ = P.actor_critic_forward(params, last_obs, hp)
    traj = {
        "obs": obs.astype(dtype),
        "action": action.reshape(T, N, act_dim).astype(dtype),
        "value": value.reshape(T, N).astype(dtype),
        "log_prob": log_prob.reshape(T, N).astype(dtype),
        "reward": g.standard_normal((T, N)).astype(dtype),
        "done": g.random((T, N)) < done_p,
    }
    return {
        "params": P.tree_like(params, lambda x: x.astype(dtype)),
        "traj": traj,
        "last_val": last_val.astype(dtype),
        "rng": np.array([0, 1337], dtype=np.uint32),     # PRNGKey(1337), config.py:77
        "obs_dim": obs_dim,
        "act_dim": act_dim,
    }
