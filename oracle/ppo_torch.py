"""ORACLE (test infrastructure, never shipped, never on the product path).

PyTorch-CPU restatement of /root/reference/minppo/train.py:181-281, written line by line
against the reference and differentiated with ``torch.autograd`` the way the reference
uses ``jax.value_and_grad`` (train.py:246).  Two jobs:

* independent check of the hand-derived gradients in oracle/ppo_numpy.py (float64);
* the timed CPU baseline of bench.py (float32, all host threads) -- labelled
  "CPU restatement (PyTorch), not JAX" wherever it is reported, because JAX cannot be
  installed in this image.

PARITY STATUS: "parity unpinned" (see oracle/ppo_numpy.py header).
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch

from . import threefry
from .ppo_numpy import Hyper, leaf_order, get_leaf, set_leaf, tree_like, learning_rate

LOG_2PI = math.log(2.0 * math.pi)


def to_torch(tree: Dict, dtype=torch.float32, requires_grad=False) -> Dict:
    return tree_like(tree, lambda x: torch.tensor(np.asarray(x), dtype=dtype, requires_grad=requires_grad))


def to_numpy(tree: Dict) -> Dict:
    return tree_like(tree, lambda x: x.detach().cpu().numpy())


def gae(reward, value, done, last_val, gamma: float, lam: float):
    """train.py:185-205 with torch tensors ([T,N]); vectorised over envs, loop over T."""
    T = reward.shape[0]
    nd = (1 - done.to(torch.int32)).to(reward.dtype)     # (1 - done): int -> float
    adv = torch.empty_like(reward)
    carry = torch.zeros_like(last_val)
    nxt = last_val
    for t in range(T - 1, -1, -1):
        delta = reward[t] + gamma * nxt * nd[t] - value[t]
        carry = delta + gamma * lam * nd[t] * carry
        adv[t] = carry
        nxt = value[t]
    return adv, adv + value


def mlp(m: Dict, x, num_layers: int, tanh: bool):
    """train.py:61-68."""
    for i in range(num_layers):
        d = m[f"Dense_{i}"]
        x = x @ d["kernel"] + d["bias"]
        x = torch.tanh(x) if tanh else torch.relu(x)
    d = m[f"Dense_{num_layers}"]
    return x @ d["kernel"] + d["bias"]


def loss_fn(params: Dict, mb: Dict, hp: Hyper):
    """train.py:218-243."""
    p = params["params"]
    mean = mlp(p["MLP_0"], mb["obs"], hp.num_layers, hp.use_tanh)           # train.py:79
    scale = torch.exp(p["log_std"])                                          # train.py:81
    value = mlp(p["MLP_1"], mb["obs"], hp.num_layers, False).squeeze(-1)     # train.py:82-83
    z = (mb["action"] - mean) * (1.0 / scale)
    log_prob = (-0.5 * z * z - 0.5 * LOG_2PI).sum(-1) - torch.log(torch.abs(scale)).sum()
    # train.py:226-231
    value_pred_clipped = mb["value"] + (value - mb["value"]).clamp(-hp.clip_eps, hp.clip_eps)
    value_losses = torch.square(value - mb["tgt"])
    value_losses_clipped = torch.square(value_pred_clipped - mb["tgt"])
    value_loss = 0.5 * torch.maximum(value_losses, value_losses_clipped).mean()
    # train.py:234-239
    ratio = torch.exp(log_prob - mb["log_prob"])
    g = mb["adv"]
    g = (g - g.mean()) / (g.std(unbiased=False) + 1e-8)
    loss_actor1 = ratio * g
    loss_actor2 = torch.clamp(ratio, 1.0 - hp.clip_eps, 1.0 + hp.clip_eps) * g
    loss_actor = (-torch.minimum(loss_actor1, loss_actor2)).mean()
    A = mb["action"].shape[1]
    entropy = A * (0.5 + 0.5 * LOG_2PI) + torch.log(torch.abs(scale)).sum()  # train.py:240
    total = loss_actor + hp.vf_coef * value_loss - hp.ent_coef * entropy      # train.py:242
    return total, (value_loss, loss_actor, entropy)


_COMPILED = {"on": False, "fn": None}


def set_compiled(on: bool) -> None:
    """bench.py's GPU proxy arm only: evaluate loss_fn through torch.compile (inductor).  Off by default."""
    _COMPILED["on"] = bool(on)
    if on and _COMPILED["fn"] is None:
        _COMPILED["fn"] = torch.compile(loss_fn)


def loss_and_grads(params: Dict, mb: Dict, hp: Hyper):
    """jax.value_and_grad(_loss_fn, has_aux=True) (train.py:246-247)."""
    paths = leaf_order(hp.num_layers)
    leaves = [get_leaf(params, p).detach().requires_grad_(True) for p in paths]
    tree = tree_like(params, lambda x: x)
    for p, l in zip(paths, leaves):
        set_leaf(tree, p, l)
    total, (vl, al, ent) = (_COMPILED["fn"] if _COMPILED["on"] else loss_fn)(tree, mb, hp)
    gs = torch.autograd.grad(total, leaves)
    grads = tree_like(params, lambda x: x)
    for p, g in zip(paths, gs):
        set_leaf(grads, p, g)
    return (total.detach(), vl.detach(), al.detach(), ent.detach()), grads


@torch.no_grad()
def clip_adam_step(params: Dict, grads: Dict, opt: Dict, hp: Hyper):
    """optax.chain(clip_by_global_norm, adam) + apply_updates (train.py:116-123, 248)."""
    paths = leaf_order(hp.num_layers)
    dt = get_leaf(params, paths[0]).dtype
    npdt = np.float32 if dt == torch.float32 else np.float64
    g_norm = torch.sqrt(sum((get_leaf(grads, p) ** 2).sum() for p in paths))
    on_gpu = g_norm.device.type != "cpu"
    # CPU: a Python bool like optax's lax.select predicate; accelerator (bench.py's GPU proxy): a 0-d tensor and
    # torch.where below, so that the eager stream is never synchronised mid-update
    trigger = (g_norm < hp.max_grad_norm) if on_gpu else bool(g_norm < hp.max_grad_norm)
    count = opt["count"]
    lr = float(learning_rate(count, hp, npdt))
    c1 = float(npdt(1) - npdt(hp.b1) ** npdt(count + 1))
    c2 = float(npdt(1) - npdt(hp.b2) ** npdt(count + 1))
    new_p, new_mu, new_nu = (tree_like(params, lambda x: x) for _ in range(3))
    for p in paths:
        g = get_leaf(grads, p)
        if on_gpu:
            g = torch.where(trigger, g, (g / g_norm) * hp.max_grad_norm)
        elif not trigger:
            g = (g / g_norm) * hp.max_grad_norm
        mu = (1 - hp.b1) * g + hp.b1 * get_leaf(opt["mu"], p)
        nu = (1 - hp.b2) * (g * g) + hp.b2 * get_leaf(opt["nu"], p)
        u = (mu / c1) / (torch.sqrt(nu / c2 + hp.eps_root) + hp.eps)
        set_leaf(new_p, p, get_leaf(params, p) + (-lr) * u)
        set_leaf(new_mu, p, mu)
        set_leaf(new_nu, p, nu)
    return new_p, {"count": count + 1, "mu": new_mu, "nu": new_nu}, g_norm


def update(params: Dict, opt: Dict, traj: Dict, last_val, rng, hp: Hyper, perms=None,
           epochs: int | None = None, minibatches: int | None = None):
    """One learner update on torch CPU tensors (train.py:181-281).  ``epochs`` /
    ``minibatches`` bound the work for the timed CPU baseline sample (bench.py): the
    first ``minibatches`` minibatches of the first ``epochs`` epochs are run, nothing else
    changes."""
    adv, tgt = gae(traj["reward"], traj["value"], traj["done"], last_val, hp.gamma, hp.gae_lambda)
    flat = {k: v.reshape((-1,) + tuple(v.shape[2:])) for k, v in
            {"obs": traj["obs"], "action": traj["action"], "value": traj["value"],
             "log_prob": traj["log_prob"], "adv": adv, "tgt": tgt}.items()}
    B, M, mbs = hp.batch_size, hp.num_minibatches, hp.minibatch_size
    if mbs * M != B:
        raise ValueError("`batch_size` must be equal to `num_steps * num_envs`")
    E = hp.update_epochs if epochs is None else epochs
    Mrun = M if minibatches is None else minibatches
    rng = np.asarray(rng, np.uint32)
    losses: List = []
    for e in range(E):
        rng, sub = threefry.split(rng, 2, hp.prng_mode)
        perm = threefry.permutation(sub, B, hp.prng_mode) if perms is None else perms[e]
        perm_t = perm if torch.is_tensor(perm) else torch.from_numpy(np.asarray(perm).astype(np.int64))
        perm_t = perm_t.to(flat["obs"].device)
        # train.py:261: a full shuffled copy of every leaf, then [M, mb, ...] (262-265)
        shuffled = {k: v.index_select(0, perm_t) for k, v in flat.items()}
        for k in range(Mrun):
            mb = {name: arr[k * mbs:(k + 1) * mbs] for name, arr in shuffled.items()}
            ls, grads = loss_and_grads(params, mb, hp)
            params, opt, _ = clip_adam_step(params, grads, opt, hp)
            losses.append(torch.stack(ls))
    out = torch.stack(losses).reshape(E, Mrun, 4) if losses else torch.zeros((E, 0, 4))
    return params, opt, rng, out, {"advantages": adv, "targets": tgt}
