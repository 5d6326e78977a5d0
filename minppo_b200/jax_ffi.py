"""JAX host side of the drop-in: XLA FFI custom calls over libminppo_b200_xla.so.

STATUS: NOT RUNNABLE IN THIS IMAGE -- JAX / jaxlib are not installed here,
so this module is written against the public ``jax.ffi`` API and has never been executed.
It is the file a maintainer of kscalelabs/minppo imports where JAX exists; INTEGRATION.md
shows the splice into /root/reference/minppo/train.py:185-281.  Everything that IS tested
here goes through the same C ABI (include/minppo_b200.h) via ctypes (minppo_b200/_lib.py).

Three custom calls (minppo_b200/csrc/xla_ffi_shim.cc):

* ``minppo_gae``          replaces ``_calculate_gae``                     (train.py:185-207)
* ``minppo_update``       replaces GAE + the epoch / minibatch scans      (train.py:185-281)
* ``minppo_policy_step``  replaces network.apply / split / sample / log_prob of the rollout (train.py:157-160)

The params pytree of the reference (checkpoint layout, train.py:86-89) is carried as ONE flat
fp32 arena in JAX's sorted flatten order; ``ravel`` / ``unravel`` below convert at the seam
(outside the hot loop: once after ``TrainState.create`` and once before ``save_model``).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

try:  # pragma: no cover - JAX is absent in the build image
    import jax
    import jax.numpy as jnp
except ImportError as e:  # pragma: no cover
    raise ImportError("minppo_b200.jax_ffi needs jax >= 0.4.38 (jax.ffi); the rest of minppo_b200 does not") from e

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "lib", "libminppo_b200_xla.so")
_registered = False


def register() -> None:
    """dlopen the shim and register both handlers for platform "CUDA" (idempotent)."""
    global _registered
    if _registered:
        return
    if not os.path.exists(_SHIM):
        raise ImportError(f"{_SHIM} not found: build it with the command at the top of csrc/xla_ffi_shim.cc")
    lib = ctypes.CDLL(_SHIM)
    jax.ffi.register_ffi_target("minppo_gae", jax.ffi.pycapsule(lib.MinppoGae), platform="CUDA")
    jax.ffi.register_ffi_target("minppo_update", jax.ffi.pycapsule(lib.MinppoUpdate), platform="CUDA")
    jax.ffi.register_ffi_target("minppo_policy_step", jax.ffi.pycapsule(lib.MinppoPolicyStep), platform="CUDA")
    jax.ffi.register_ffi_target("minppo_bootstrap_value", jax.ffi.pycapsule(lib.MinppoBootstrapValue), platform="CUDA")
    _registered = True


def calculate_gae(mem_batch, last_val, gamma: float, gae_lambda: float):
    """Drop-in for ``_calculate_gae(mem_batch, last_val)`` (train.py:185-207)."""
    register()
    out = jax.ShapeDtypeStruct(mem_batch.reward.shape, jnp.float32)
    return jax.ffi.ffi_call("minppo_gae", (out, out))(
        mem_batch.reward, mem_batch.value, mem_batch.done, last_val,
        gamma=np.float32(gamma), gae_lambda=np.float32(gae_lambda))


def ravel(params) -> "jax.Array":
    """Checkpoint tree -> flat fp32 arena (JAX's sorted flatten order == minppo_param_layout)."""
    return jnp.concatenate([jnp.ravel(x) for x in jax.tree_util.tree_leaves(params)]).astype(jnp.float32)


def unravel(params_like, flat):
    leaves, treedef = jax.tree_util.tree_flatten(params_like)
    out, off = [], 0
    for x in leaves:
        out.append(flat[off:off + x.size].reshape(x.shape))
        off += x.size
    return jax.tree_util.tree_unflatten(treedef, out)


def learner_update(flat_params, mu, nu, count, mem_batch, last_val, rng, config, prng_mode: int | None = None):
    """The seam of ``_update_step`` (train.py:181-281) as one custom call.

    flat_params / mu / nu: f32[P]; count: i32[1]; mem_batch: the reference's ``Memory`` (time-major);
    rng: the raw uint32[2] key data (``jax.random.key_data(rng)`` for new-style keys).
    Returns (flat_params', mu', nu', count', rng', losses[E, M, 4]).  params / mu / nu / count are
    updated in place through input_output_aliases."""
    register()
    if prng_mode is None:
        prng_mode = int(bool(jax.config.jax_threefry_partitionable))
    E, M = config.training.update_epochs, config.training.num_minibatches
    f32 = jnp.float32
    outs = (
        jax.ShapeDtypeStruct(flat_params.shape, f32), jax.ShapeDtypeStruct(mu.shape, f32),
        jax.ShapeDtypeStruct(nu.shape, f32), jax.ShapeDtypeStruct(count.shape, jnp.int32),
        jax.ShapeDtypeStruct((2,), jnp.uint32), jax.ShapeDtypeStruct((E, M, 4), f32),
    )
    call = jax.ffi.ffi_call("minppo_update", outs, input_output_aliases={0: 0, 1: 1, 2: 2, 3: 3})
    return call(
        flat_params, mu, nu, count, mem_batch.obs, mem_batch.action, mem_batch.value, mem_batch.reward,
        mem_batch.log_prob, mem_batch.done, last_val, rng,
        num_minibatches=np.int32(M), update_epochs=np.int32(E),
        total_timesteps=np.int64(config.training.total_timesteps), anneal_lr=bool(config.training.anneal_lr),
        hidden_size=np.int32(config.model.hidden_size), num_layers=np.int32(config.model.num_layers),
        use_tanh=bool(config.model.use_tanh), prng_mode=np.int32(prng_mode),
        training_lr=np.float32(config.training.lr), opt_lr=np.float32(config.opt.lr),
        max_grad_norm=np.float32(config.opt.max_grad_norm), gamma=np.float32(config.rl.gamma),
        gae_lambda=np.float32(config.rl.gae_lambda), clip_eps=np.float32(config.rl.clip_eps),
        ent_coef=np.float32(config.rl.ent_coef), vf_coef=np.float32(config.rl.vf_coef))


def _static_attrs(config, prng_mode):
    """The attribute block shared by ``minppo_update`` and ``minppo_policy_step`` (one context for both)."""
    return dict(
        num_minibatches=np.int32(config.training.num_minibatches), update_epochs=np.int32(config.training.update_epochs),
        total_timesteps=np.int64(config.training.total_timesteps), anneal_lr=bool(config.training.anneal_lr),
        hidden_size=np.int32(config.model.hidden_size), num_layers=np.int32(config.model.num_layers),
        use_tanh=bool(config.model.use_tanh), prng_mode=np.int32(prng_mode),
        training_lr=np.float32(config.training.lr), opt_lr=np.float32(config.opt.lr),
        max_grad_norm=np.float32(config.opt.max_grad_norm), gamma=np.float32(config.rl.gamma),
        gae_lambda=np.float32(config.rl.gae_lambda), clip_eps=np.float32(config.rl.clip_eps),
        ent_coef=np.float32(config.rl.ent_coef), vf_coef=np.float32(config.rl.vf_coef))


def policy_step(flat_params, last_obs, rng, config, act_dim: int, weights_current: bool = False,
                prng_mode: int | None = None):
    """Drop-in for the network evaluation of ``_env_step`` (train.py:157-160).

    flat_params: f32[P]; last_obs: f32[N, D]; rng: raw uint32[2] key data.
    Returns (action f32[N, A], log_prob f32[N], value f32[N], rng' uint32[2]) with rng' = split(rng)[0]."""
    register()
    if prng_mode is None:
        prng_mode = int(bool(jax.config.jax_threefry_partitionable))
    n = last_obs.shape[0]
    f32 = jnp.float32
    outs = (jax.ShapeDtypeStruct((n, act_dim), f32), jax.ShapeDtypeStruct((n,), f32), jax.ShapeDtypeStruct((n,), f32),
            jax.ShapeDtypeStruct((2,), jnp.uint32))
    return jax.ffi.ffi_call("minppo_policy_step", outs)(
        flat_params, last_obs, rng, num_steps=np.int32(config.training.num_steps), act_dim=np.int32(act_dim),
        weights_current=bool(weights_current), **_static_attrs(config, prng_mode))


def bootstrap_value(flat_params, last_obs, config, act_dim: int, weights_current: bool = True,
                    prng_mode: int | None = None):
    """Drop-in for ``_, last_val = network.apply(params, last_obs)`` (train.py:182-183): the critic only, no key.

    flat_params: f32[P]; last_obs: f32[N, D].  Returns value f32[N].  ``weights_current`` defaults to True: the call follows
    the rollout's policy steps on the same, unmodified arena."""
    register()
    if prng_mode is None:
        prng_mode = int(bool(jax.config.jax_threefry_partitionable))
    n = last_obs.shape[0]
    return jax.ffi.ffi_call("minppo_bootstrap_value", jax.ShapeDtypeStruct((n,), jnp.float32))(
        flat_params, last_obs, num_steps=np.int32(config.training.num_steps), act_dim=np.int32(act_dim),
        weights_current=bool(weights_current), **_static_attrs(config, prng_mode))
