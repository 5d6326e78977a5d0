"""ORACLE (test infrastructure, never shipped, never on the product path).

NumPy restatement of the learner hot path of /root/reference/minppo/train.py:181-281:
GAE, per-epoch shuffle + minibatching, ActorCritic forward, the PPO loss, its gradient
(hand-derived; cross-checked against torch.autograd in oracle/ppo_torch.py) and the
optax ``chain(clip_by_global_norm, adam)`` update.

Runs in float64 (ground truth for tolerances) or float32.  ``gemm="bf16"`` additionally
rounds every GEMM operand to bfloat16 exactly where the CUDA path does, so that the
tensor-core path can be checked to accumulation-order noise instead of to bf16 noise.

The arithmetic the reference delegates to flax / distrax / optax (un-vendored, unpinned:
/root/reference/requirements.txt:8-13) is restated from their published semantics:

* ``nn.Dense``: ``y = x @ kernel + bias`` (train.py:63, 68)
* ``distrax.MultivariateNormalDiag(loc, scale).log_prob / .entropy`` (train.py:81, 223, 240)
* ``optax.clip_by_global_norm`` -> ``optax.adam(eps=1e-5)`` (train.py:116-123)
* ``TrainState.apply_gradients`` (train.py:248)

PARITY STATUS: "parity unpinned" -- the reference has no tests or golden vectors and JAX
cannot be imported in this image.  Anchors: torch.autograd (fp64) for
every gradient, closed forms for GAE / Gaussian log-prob / Adam step 1 (tests/).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Tuple

import numpy as np

from . import threefry

LOG_2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------
# hyper-parameters: the rl.* / training.* / opt.* / model.* keys the learner reads
# (/root/reference/minppo/config.py:50-84)
# ----------------------------------------------------------------------------------------
@dataclasses.dataclass
class Hyper:
    num_envs: int = 2048            # training.num_envs      config.py:78
    num_steps: int = 10             # training.num_steps == rl.num_env_steps
    num_minibatches: int = 32       # training.num_minibatches
    update_epochs: int = 4          # training.update_epochs
    total_timesteps: int = 1_000_000_000
    anneal_lr: bool = True          # training.anneal_lr
    training_lr: float = 3e-4       # training.lr  (annealed path)
    opt_lr: float = 3e-4            # opt.lr       (constant path)
    max_grad_norm: float = 0.5      # opt.max_grad_norm
    gamma: float = 0.99
    gae_lambda: float = 0.95
    clip_eps: float = 0.2
    ent_coef: float = 0.0
    vf_coef: float = 0.5
    hidden_size: int = 256
    num_layers: int = 2
    use_tanh: bool = True
    prng_mode: int = threefry.LEGACY
    b1: float = 0.9                 # optax.adam defaults
    b2: float = 0.999
    eps: float = 1e-5               # train.py:118
    eps_root: float = 0.0

    @property
    def batch_size(self) -> int:
        return self.num_envs * self.num_steps

    @property
    def minibatch_size(self) -> int:  # train.py:94
        return self.num_envs * self.num_steps // self.num_minibatches

    @property
    def num_updates(self) -> int:     # train.py:93
        return self.total_timesteps // self.num_steps // self.num_envs


# ----------------------------------------------------------------------------------------
# parameter tree: the pickle layout of train.py:86-89
# ----------------------------------------------------------------------------------------
def leaf_order(num_layers: int) -> List[Tuple[str, ...]]:
    """Leaf paths in JAX's sorted-key flatten order: MLP_0 < MLP_1 < log_std, Dense_i in
    order, bias < kernel."""
    out: List[Tuple[str, ...]] = []
    for mlp in ("MLP_0", "MLP_1"):
        for i in range(num_layers + 1):
            out.append((mlp, f"Dense_{i}", "bias"))
            out.append((mlp, f"Dense_{i}", "kernel"))
    out.append(("log_std",))
    return out


def get_leaf(params: Dict, path: Tuple[str, ...]):
    node = params["params"]
    for k in path:
        node = node[k]
    return node


def set_leaf(params: Dict, path: Tuple[str, ...], value) -> None:
    node = params["params"]
    for k in path[:-1]:
        node = node[k]
    node[path[-1]] = value


def tree_like(params: Dict, fn) -> Dict:
    """New tree with fn(leaf) at every leaf."""
    def rec(n):
        return {k: rec(v) for k, v in n.items()} if isinstance(n, dict) else fn(n)
    return rec(params)


def init_params(obs_dim: int, act_dim: int, hidden: int, num_layers: int, seed: int = 0,
                dtype=np.float64) -> Dict:
    """Synthetic parameters with the init GAINS of train.py:63/68 (sqrt(2) hidden, 0.01
    last layer) but plain scaled normals -- flax's orthogonal init is off the hot path and
    not reproducible here."""
    g = np.random.default_rng(seed)
    p: Dict = {"params": {}}
    for mlp, out_dim in (("MLP_0", act_dim), ("MLP_1", 1)):
        d: Dict = {}
        fan_in = obs_dim
        for i in range(num_layers):
            d[f"Dense_{i}"] = {
                "kernel": (g.standard_normal((fan_in, hidden)) * math.sqrt(2.0 / fan_in)).astype(dtype),
                "bias": (0.01 * g.standard_normal(hidden)).astype(dtype),
            }
            fan_in = hidden
        d[f"Dense_{num_layers}"] = {
            "kernel": (g.standard_normal((fan_in, out_dim)) * (0.01 / math.sqrt(fan_in))).astype(dtype),
            "bias": (0.01 * g.standard_normal(out_dim)).astype(dtype),
        }
        p["params"][mlp] = d
    p["params"]["log_std"] = (0.05 * g.standard_normal(act_dim)).astype(dtype)
    return p


# ----------------------------------------------------------------------------------------
# bfloat16 rounding (round-to-nearest-even), for gemm="bf16"
# ----------------------------------------------------------------------------------------
def bf16_round(x: np.ndarray) -> np.ndarray:
    x32 = np.ascontiguousarray(x, dtype=np.float32)
    u = x32.view(np.uint32)
    with np.errstate(over="ignore"):
        r = (u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)
    return r.view(np.float32).astype(x.dtype if x.dtype in (np.float32, np.float64) else np.float32)


def _q(x, gemm):
    return bf16_round(x) if gemm == "bf16" else x


# ----------------------------------------------------------------------------------------
# GAE  (train.py:185-205)
# ----------------------------------------------------------------------------------------
def gae(reward, value, done, last_val, gamma: float, lam: float, dtype=np.float64):
    """Reverse scan, carry (gae, next_value) init (0, last_val) (train.py:200):
    delta = r + gamma*next_value*(1-done) - value; gae = delta + gamma*lam*(1-done)*gae
    (train.py:193-194); targets = advantages + value (train.py:205)."""
    reward = np.asarray(reward, dtype)
    value = np.asarray(value, dtype)
    nd = (1 - np.asarray(done).astype(np.int32)).astype(dtype)
    T = reward.shape[0]
    g = dtype(gamma)
    gl = dtype(gamma) * dtype(lam)
    adv = np.empty_like(reward)
    carry = np.zeros_like(np.asarray(last_val, dtype))
    nxt = np.asarray(last_val, dtype)
    for t in range(T - 1, -1, -1):
        delta = reward[t] + g * nxt * nd[t] - value[t]
        carry = delta + gl * nd[t] * carry
        adv[t] = carry
        nxt = value[t]
    return adv, adv + value


# ----------------------------------------------------------------------------------------
# model  (train.py:56-83)
# ----------------------------------------------------------------------------------------
def _act(z, tanh: bool):
    return np.tanh(z) if tanh else np.maximum(z, 0)


def _dact_from_out(h, tanh: bool):
    # tanh' = 1 - tanh^2 ; relu' = [z > 0] == [h > 0]  (jax.nn.relu: zero gradient at 0)
    return (1 - h * h) if tanh else (h > 0).astype(h.dtype)


def mlp_forward(mlp: Dict, x, num_layers: int, tanh: bool, gemm="exact"):
    """Returns (out, [a_0 .. a_L]) with a_0 = x (train.py:61-68).

    gemm="bf16" mirrors the CUDA path's rounding points (DESIGN.md "precision"): the
    observation and hidden kernels are rounded to bf16 for the tensor-core GEMMs (fp32
    accumulate), every hidden activation is STORED as bf16 (so all its consumers see the
    rounded value), the output head runs in fp32 on the stored activation."""
    a = _q(x, gemm)
    acts = [a]
    for i in range(num_layers):
        d = mlp[f"Dense_{i}"]
        a = _q(_act(a @ _q(d["kernel"], gemm) + d["bias"], tanh), gemm)
        acts.append(a)
    d = mlp[f"Dense_{num_layers}"]
    out = a @ d["kernel"] + d["bias"]
    return out, acts


def mlp_backward(mlp: Dict, acts, dout, num_layers: int, tanh: bool, gemm="exact"):
    """Gradient of sum(out * dout) w.r.t. every kernel / bias of one MLP.  gemm="bf16": dz of
    every hidden layer is stored as bf16 once; dW, db and the next dA all see that value."""
    grads: Dict = {}
    d = mlp[f"Dense_{num_layers}"]
    grads[f"Dense_{num_layers}"] = {"kernel": acts[num_layers].T @ dout, "bias": dout.sum(0)}
    da = dout @ d["kernel"].T
    for i in range(num_layers - 1, -1, -1):
        dz = _q(da * _dact_from_out(acts[i + 1], tanh), gemm)
        dd = mlp[f"Dense_{i}"]
        grads[f"Dense_{i}"] = {"kernel": acts[i].T @ dz, "bias": dz.sum(0)}
        if i > 0:
            da = dz @ _q(dd["kernel"], gemm).T
    return grads


def actor_critic_forward(params: Dict, obs, hp: Hyper, gemm="exact"):
    """(mean, log_std, value, caches) -- train.py:78-83.  Critic is relu always (82)."""
    p = params["params"]
    mean, acts_a = mlp_forward(p["MLP_0"], obs, hp.num_layers, hp.use_tanh, gemm)
    v, acts_c = mlp_forward(p["MLP_1"], obs, hp.num_layers, False, gemm)
    return mean, p["log_std"], v[:, 0], (acts_a, acts_c)


def gaussian_log_prob(mean, log_std, action):
    """distrax.MultivariateNormalDiag(mean, exp(log_std)).log_prob(action) (train.py:223):
    z = (a - loc) * (1/scale); sum(-0.5 z^2 - 0.5 log 2pi) - sum(log|scale|)."""
    scale = np.exp(log_std)
    z = (action - mean) * (1.0 / scale)
    dt = mean.dtype.type
    return (dt(-0.5) * z * z - dt(0.5 * LOG_2PI)).sum(-1) - np.log(np.abs(scale)).sum(), z, scale


def gaussian_entropy(log_std, act_dim: int):
    """distrax Transformed entropy: A*(0.5 + 0.5 log 2pi) + sum(log|scale|) (train.py:240)."""
    scale = np.exp(log_std)
    dt = log_std.dtype.type
    return dt(act_dim * (0.5 + 0.5 * LOG_2PI)) + np.log(np.abs(scale)).sum()


def policy_step(params: Dict, obs, rng, hp: Hyper, mode: int = 0, gemm="exact", sample: bool = True):
    """One env step's network evaluation, train.py:157-160:

        pi, value = network.apply(params, last_obs); rng, action_rng = jax.random.split(rng)
        action = pi.sample(seed=action_rng); log_prob = pi.log_prob(action)

    distrax.MultivariateNormalDiag.sample = loc + scale * jax.random.normal(action_rng, (N, A)) (Normal(0,1)._sample_n
    pushed through the Shift / ScalarAffine bijectors).  Returns (action, log_prob, value, rng', mean).
    ``sample=False``: action = mean, rng unchanged.  ``obs`` is the GLOBAL [N, D] batch."""
    from . import threefry as tf

    mean, log_std, value, _ = actor_critic_forward(params, obs, hp, gemm)
    dt = mean.dtype.type
    if sample:
        rng2, action_rng = tf.split(np.asarray(rng, np.uint32), 2, mode)
        n, a = mean.shape
        eps = tf.normal_f32(action_rng, n * a, mode, erfinv="xla").reshape(n, a).astype(mean.dtype)
        action = mean + np.exp(log_std).astype(mean.dtype) * eps
    else:
        rng2, action = np.asarray(rng, np.uint32), mean.copy()
    logp, _, _ = gaussian_log_prob(mean, log_std, action)
    return action.astype(mean.dtype), logp.astype(mean.dtype), value, rng2, mean


# ----------------------------------------------------------------------------------------
# loss and gradient  (train.py:218-247)
# ----------------------------------------------------------------------------------------
def _dclip(x, lo, hi):
    """d/dx of minimum(maximum(x, lo), hi) with JAX's 0.5/0.5 tie split."""
    d = ((x > lo) & (x < hi)).astype(x.dtype)
    return d + 0.5 * ((x == lo) | (x == hi))


def loss_and_grads(params: Dict, mb: Dict, hp: Hyper, gemm="exact", n_total=None, adv_mean_std=None):
    """mb: obs[mb,D], action[mb,A], value[mb], log_prob[mb], adv[mb], tgt[mb].
    Returns ((total, value_loss, actor_loss, entropy), grads-tree).

    ``n_total`` / ``adv_mean_std`` describe the env-sharded case: ``mb``
    holds only this rank's rows of a minibatch of ``n_total`` rows whose advantage mean / std
    are given; means become sums over local rows divided by ``n_total`` so that SUMMING the
    returned losses (except entropy) and gradients over ranks gives the global values."""
    obs = mb["obs"]
    dt = obs.dtype.type
    n = obs.shape[0] if n_total is None else n_total
    A = mb["action"].shape[1]
    eps = dt(hp.clip_eps)
    mean, log_std, v, (acts_a, acts_c) = actor_critic_forward(params, obs, hp, gemm)
    logp, z, scale = gaussian_log_prob(mean, log_std, mb["action"])

    # value loss (train.py:226-231)
    v_old, tgt = mb["value"], mb["tgt"]
    dv = v - v_old
    v_clip = v_old + np.clip(dv, -eps, eps)
    vl = (v - tgt) ** 2
    vlc = (v_clip - tgt) ** 2
    value_loss = dt(0.5) * np.maximum(vl, vlc).sum() / dt(n)

    # actor loss (train.py:234-239); advantage normalised over THIS minibatch (235)
    ratio = np.exp(logp - mb["log_prob"])
    adv = mb["adv"]
    a_mean, a_std = (adv.mean(), adv.std()) if adv_mean_std is None else (dt(adv_mean_std[0]), dt(adv_mean_std[1]))
    adv_n = (adv - a_mean) / (a_std + dt(1e-8))
    l1 = ratio * adv_n
    l2 = np.clip(ratio, dt(1.0) - eps, dt(1.0) + eps) * adv_n
    actor_loss = (-np.minimum(l1, l2)).sum() / dt(n)
    entropy = gaussian_entropy(log_std, A)          # batch-independent (train.py:240)
    total = actor_loss + dt(hp.vf_coef) * value_loss - dt(hp.ent_coef) * entropy

    # ---- backward -------------------------------------------------------------------
    inv_n = dt(1.0 / n)
    # d total / d v
    wa = (vl > vlc).astype(v.dtype) + 0.5 * (vl == vlc)
    wb = 1 - wa
    g_v = dt(hp.vf_coef) * dt(0.5) * inv_n * (
        wa * 2 * (v - tgt) + wb * 2 * (v_clip - tgt) * _dclip(dv, -eps, eps))
    # d total / d logp
    w1 = (l1 < l2).astype(v.dtype) + 0.5 * (l1 == l2)
    w2 = 1 - w1
    dmin_dr = w1 * adv_n + w2 * _dclip(ratio, dt(1.0) - eps, dt(1.0) + eps) * adv_n
    g_logp = -inv_n * dmin_dr * ratio
    # logp -> mean, log_std
    g_mean = g_logp[:, None] * (z / scale)
    g_log_std = (g_logp[:, None] * (z * z - 1)).sum(0) - dt(hp.ent_coef)

    p = params["params"]
    grads = {"params": {
        "MLP_0": mlp_backward(p["MLP_0"], acts_a, g_mean, hp.num_layers, hp.use_tanh, gemm),
        "MLP_1": mlp_backward(p["MLP_1"], acts_c, g_v[:, None], hp.num_layers, False, gemm),
        "log_std": g_log_std,
    }}
    return (total, value_loss, actor_loss, entropy), grads


# ----------------------------------------------------------------------------------------
# optimizer  (train.py:98-101, 114-124, 248; optax semantics)
# ----------------------------------------------------------------------------------------
def learning_rate(count: int, hp: Hyper, dtype=np.float64):
    """train.py:98-101 (annealed; note SURVEY F8: divides by minibatch_size*E) or opt.lr."""
    if hp.anneal_lr:
        frac = dtype(1.0) - dtype(count // (hp.minibatch_size * hp.update_epochs)) / dtype(hp.num_updates)
        return dtype(hp.training_lr) * frac
    return dtype(hp.opt_lr)


def init_opt_state(params: Dict) -> Dict:
    return {"count": 0,
            "mu": tree_like(params, lambda x: np.zeros_like(x)),
            "nu": tree_like(params, lambda x: np.zeros_like(x))}


def clip_adam_step(params: Dict, grads: Dict, opt: Dict, hp: Hyper):
    """optax.chain(clip_by_global_norm(max_norm), adam(lr, eps)) then apply_updates.
    Returns (params', opt', grad_norm).  ``count`` seen by the schedule is the value
    BEFORE the increment; bias correction uses count+1."""
    L = hp.num_layers
    paths = leaf_order(L)
    dt = get_leaf(params, paths[0]).dtype.type
    sq = dt(0)
    for pth in paths:
        g = get_leaf(grads, pth)
        sq = sq + (g * g).sum()
    g_norm = np.sqrt(sq)
    trigger = g_norm < dt(hp.max_grad_norm)
    count = opt["count"]
    lr = learning_rate(count, hp, dt)
    c1 = dt(1) - dt(hp.b1) ** dt(count + 1)
    c2 = dt(1) - dt(hp.b2) ** dt(count + 1)
    new_p = tree_like(params, lambda x: x)
    new_mu = tree_like(params, lambda x: x)
    new_nu = tree_like(params, lambda x: x)
    for pth in paths:
        g = get_leaf(grads, pth)
        if not trigger:
            g = (g / g_norm) * dt(hp.max_grad_norm)
        mu = (dt(1) - dt(hp.b1)) * g + dt(hp.b1) * get_leaf(opt["mu"], pth)
        nu = (dt(1) - dt(hp.b2)) * (g * g) + dt(hp.b2) * get_leaf(opt["nu"], pth)
        u = (mu / c1) / (np.sqrt(nu / c2 + dt(hp.eps_root)) + dt(hp.eps))
        set_leaf(new_p, pth, get_leaf(params, pth) + (-lr) * u)
        set_leaf(new_mu, pth, mu)
        set_leaf(new_nu, pth, nu)
    return new_p, {"count": count + 1, "mu": new_mu, "nu": new_nu}, g_norm


# ----------------------------------------------------------------------------------------
# one full update  (train.py:181-281, minus the bootstrap forward which is an input)
# ----------------------------------------------------------------------------------------
def flatten_traj(traj: Dict) -> Dict:
    """[T,N,...] -> [B,...] row-major, flat = t*N + n (train.py:260)."""
    return {k: np.reshape(v, (-1,) + v.shape[2:]) for k, v in traj.items()}


def update(params: Dict, opt: Dict, traj: Dict, last_val, rng, hp: Hyper, dtype=np.float64,
           gemm="exact", perms=None):
    """traj: obs[T,N,D], action[T,N,A], value/reward/log_prob[T,N], done bool[T,N].
    Returns (params', opt', rng', losses[E,M,4], aux) where losses columns are
    (total, value_loss, actor_loss, entropy) and aux has advantages/targets/perms."""
    cast = lambda x: np.asarray(x, dtype)
    params = tree_like(params, cast)
    opt = {"count": int(opt["count"]), "mu": tree_like(opt["mu"], cast), "nu": tree_like(opt["nu"], cast)}
    adv, tgt = gae(traj["reward"], traj["value"], traj["done"], last_val, hp.gamma, hp.gae_lambda, dtype)
    flat = flatten_traj({"obs": cast(traj["obs"]), "action": cast(traj["action"]),
                         "value": cast(traj["value"]), "log_prob": cast(traj["log_prob"]),
                         "adv": adv, "tgt": tgt})
    B, M, mbs = hp.batch_size, hp.num_minibatches, hp.minibatch_size
    if mbs * M != B:
        raise ValueError("`batch_size` must be equal to `num_steps * num_envs`")  # train.py:254
    rng = np.asarray(rng, np.uint32)
    losses = np.zeros((hp.update_epochs, M, 4), dtype)
    used_perms = []
    gnorms = np.zeros((hp.update_epochs, M), dtype)
    for e in range(hp.update_epochs):
        rng, sub = threefry.split(rng, 2, hp.prng_mode)                # train.py:252
        perm = threefry.permutation(sub, B, hp.prng_mode) if perms is None else perms[e]
        used_perms.append(perm)
        for k in range(M):                                             # train.py:262-268
            idx = perm[k * mbs:(k + 1) * mbs]
            mb = {name: arr[idx] for name, arr in flat.items()}
            ls, grads = loss_and_grads(params, mb, hp, gemm)
            params, opt, gn = clip_adam_step(params, grads, opt, hp)
            losses[e, k] = ls
            gnorms[e, k] = gn
    aux = {"advantages": adv, "targets": tgt, "perms": np.stack(used_perms), "grad_norms": gnorms}
    return params, opt, rng, losses, aux


# ----------------------------------------------------------------------------------------
# flat parameter arena helpers (the C-ABI's view of the tree; see include/minppo_b200.h)
# ----------------------------------------------------------------------------------------
def leaf_shapes(obs_dim: int, act_dim: int, hidden: int, num_layers: int):
    shapes = []
    for mlp, out_dim in (("MLP_0", act_dim), ("MLP_1", 1)):
        fan_in = obs_dim
        for i in range(num_layers + 1):
            o = hidden if i < num_layers else out_dim
            shapes.append((o,))
            shapes.append((fan_in, o))
            fan_in = o
    shapes.append((act_dim,))
    return shapes


def flatten_params(params: Dict, num_layers: int, dtype=np.float32) -> np.ndarray:
    return np.concatenate([np.asarray(get_leaf(params, p), dtype).ravel() for p in leaf_order(num_layers)])


def unflatten_params(flat: np.ndarray, obs_dim: int, act_dim: int, hidden: int, num_layers: int) -> Dict:
    out: Dict = {"params": {"MLP_0": {}, "MLP_1": {}}}
    off = 0
    for pth, shp in zip(leaf_order(num_layers), leaf_shapes(obs_dim, act_dim, hidden, num_layers)):
        n = int(np.prod(shp))
        node = out["params"]
        for k in pth[:-1]:
            node = node.setdefault(k, {})
        node[pth[-1]] = np.array(flat[off:off + n]).reshape(shp)
        off += n
    assert off == flat.size
    return out
