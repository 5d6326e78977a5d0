"""Shared glue for the parity tests: oracle Hyper <-> product Config, running one update through
the C ABI (via minppo_b200.learner) and collecting everything as NumPy."""
from __future__ import annotations

import numpy as np

from oracle import ppo_numpy as P
from oracle import threefry


def hyper_to_config(hp: P.Hyper, fast_tanh: bool = True, use_graph: bool = True, dw_splits: int = 0, fused: bool = True):
    from minppo_b200.config import Config

    c = Config()
    c.model.hidden_size, c.model.num_layers, c.model.use_tanh = hp.hidden_size, hp.num_layers, hp.use_tanh
    c.opt.lr, c.opt.max_grad_norm = hp.opt_lr, hp.max_grad_norm
    c.rl.num_env_steps, c.rl.gamma, c.rl.gae_lambda = hp.num_steps, hp.gamma, hp.gae_lambda
    c.rl.clip_eps, c.rl.ent_coef, c.rl.vf_coef = hp.clip_eps, hp.ent_coef, hp.vf_coef
    c.training.lr, c.training.num_envs, c.training.total_timesteps = hp.training_lr, hp.num_envs, hp.total_timesteps
    c.training.num_minibatches, c.training.num_steps = hp.num_minibatches, hp.num_steps
    c.training.update_epochs, c.training.anneal_lr = hp.update_epochs, hp.anneal_lr
    c.learner.prng_mode = "legacy" if hp.prng_mode == threefry.LEGACY else "partitionable"
    c.learner.fast_tanh, c.learner.use_graph, c.learner.dw_splits = fast_tanh, use_graph, dw_splits
    c.learner.fused = fused
    return c


def run_gpu_update(hp: P.Hyper, problem, device, opt=None, **cfg_kw):
    """One update on the GPU.  Returns dict of NumPy arrays: params (flat), mu, nu, step, rng,
    losses [E,M,4], advantages, targets, perms, grad (last minibatch, flat P+4), grad_norms."""
    import torch

    from minppo_b200.learner import Learner, Memory, TrainState

    cfg = hyper_to_config(hp, **cfg_kw)
    D, A = problem["obs_dim"], problem["act_dim"]
    learner = Learner(cfg, D, A, device)
    flat = P.flatten_params(problem["params"], hp.num_layers)
    ts = TrainState.create(flat, device)
    if opt is not None:
        ts.mu.copy_(torch.as_tensor(P.flatten_params(opt["mu"], hp.num_layers)))
        ts.nu.copy_(torch.as_tensor(P.flatten_params(opt["nu"], hp.num_layers)))
        ts.step.fill_(int(opt["count"]))
    tr = problem["traj"]
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x)).to(device)
    mem = Memory(done=t(tr["done"]), action=t(tr["action"]), value=t(tr["value"]), reward=t(tr["reward"]),
                 log_prob=t(tr["log_prob"]), obs=t(tr["obs"]))
    rng = torch.as_tensor(problem["rng"].view(np.int32)).to(device)
    ts, rng_out, losses = learner.update(ts, mem, t(problem["last_val"]), rng)
    learner.check()
    out = {
        "params": ts.params.cpu().numpy(), "mu": ts.mu.cpu().numpy(), "nu": ts.nu.cpu().numpy(),
        "step": int(ts.step.item()), "rng": rng_out.cpu().numpy().view(np.uint32), "losses": losses.cpu().numpy(),
        "advantages": learner.read("advantages").cpu().numpy(), "targets": learner.read("targets").cpu().numpy(),
        "perms": learner.read("perms").cpu().numpy(), "grad": learner.read("grad").cpu().numpy(),
        "grad_norms": learner.read("grad_norms").cpu().numpy().reshape(hp.update_epochs, hp.num_minibatches),
        "launches": learner.launches_per_update(),
    }
    learner.close()
    return out


def flat_grads(grads, num_layers: int) -> np.ndarray:
    return P.flatten_params(grads, num_layers, np.float64)


def rel_err(a, b) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
