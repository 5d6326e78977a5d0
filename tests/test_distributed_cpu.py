"""Env-sharded data parallelism on CPU: world_size-2 `gloo` processes run
the sharding ALGORITHM the CUDA path implements (minibatch.cu compact_rows + adv_stats, per-rank
gradient sums, all-reduce, identical Adam on every rank) with the oracle's arithmetic, and must
reproduce the single-process oracle update on the same global batch and the same permutation.

What this pins: the row-ownership rule (flat = t*N + n; rank owns envs [r*N/G, (r+1)*N/G)), local
flat index t*Nl + (n - n0), two-pass global advantage statistics, sums (not means) being reduced,
-ent_coef added once, that params stay identical on all ranks with no broadcast, and that sorting each epoch's
permutation on one rank (e % world) and broadcasting it reproduces the single-process permutations."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo_numpy as P
from oracle import synth, threefry


def shard_rows(perm_slice: np.ndarray, N: int, n0: int, Nl: int) -> np.ndarray:
    """compact_rows_kernel: entries of one minibatch owned by this rank, in permutation order,
    as local flat indices."""
    t, n = perm_slice // N, perm_slice % N
    own = (n >= n0) & (n < n0 + Nl)
    return (t[own] * Nl + (n[own] - n0)).astype(np.int64)


def sharded_update(rank: int, world: int, pr, hp: P.Hyper):
    N, T, M, E, mbs = hp.num_envs, hp.num_steps, hp.num_minibatches, hp.update_epochs, hp.minibatch_size
    Nl, n0 = N // world, rank * (N // world)
    dt = np.float64
    tr = {k: v[:, n0:n0 + Nl] for k, v in pr["traj"].items()}              # this rank's env shard
    adv, tgt = P.gae(tr["reward"], tr["value"], tr["done"], pr["last_val"][n0:n0 + Nl], hp.gamma, hp.gae_lambda, dt)
    flat = P.flatten_traj({"obs": tr["obs"].astype(dt), "action": tr["action"].astype(dt), "value": tr["value"].astype(dt),
                           "log_prob": tr["log_prob"].astype(dt), "adv": adv, "tgt": tgt})
    params = P.tree_like(pr["params"], lambda x: x.astype(dt))
    opt = P.init_opt_state(params)
    rng = pr["rng"]
    paths = P.leaf_order(hp.num_layers)
    losses = np.zeros((E, M, 4))

    def allreduce(x: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    # The GLOBAL permutations depend only on the key chain: rank e % world sorts epoch e and broadcasts it
    # (learner.cu enqueue_update, share_perm), so every rank holds all of them without sorting all of them.
    perms = []
    for e in range(E):
        rng, sub = threefry.split(rng, 2, hp.prng_mode)
        mine = e % world == rank
        pt = torch.from_numpy(threefry.permutation(sub, hp.batch_size, hp.prng_mode).astype(np.int64)) if mine \
            else torch.zeros(hp.batch_size, dtype=torch.int64)
        dist.broadcast(pt, src=e % world)
        perms.append(pt.numpy())
    rows = [[shard_rows(perms[e][k * mbs:(k + 1) * mbs], N, n0, Nl) for k in range(M)] for e in range(E)]
    # advantage statistics of all E*M minibatches up front: sum, then centred second moment
    sums = allreduce(np.array([[flat["adv"][rows[e][k]].sum() for k in range(M)] for e in range(E)]))
    means = sums / mbs
    sq = allreduce(np.array([[((flat["adv"][rows[e][k]] - means[e, k]) ** 2).sum() for k in range(M)] for e in range(E)]))
    stds = np.sqrt(sq / mbs)
    counts = np.array([[len(rows[e][k]) for k in range(M)] for e in range(E)])
    for e in range(E):
        for k in range(M):
            idx = rows[e][k]
            mb = {name: arr[idx] for name, arr in flat.items()}
            ls, gr = P.loss_and_grads(params, mb, hp, n_total=mbs, adv_mean_std=(means[e, k], stds[e, k]))
            vec = np.concatenate([np.asarray(P.get_leaf(gr, p), dt).ravel() for p in paths] + [np.array(ls[1:3], dt)])
            if rank != 0:                              # -ent_coef on log_std is added once (rank 0)
                vec[-2 - pr["act_dim"]:-2] += hp.ent_coef
            vec = allreduce(vec)
            off = 0
            g = P.tree_like(params, lambda x: x)
            for p in paths:
                leaf = P.get_leaf(params, p)
                P.set_leaf(g, p, vec[off:off + leaf.size].reshape(leaf.shape))
                off += leaf.size
            value_loss, actor_loss = vec[-2], vec[-1]
            losses[e, k] = (actor_loss + hp.vf_coef * value_loss - hp.ent_coef * ls[3], value_loss, actor_loss, ls[3])
            params, opt, _ = P.clip_adam_step(params, g, opt, hp)
    return P.flatten_params(params, hp.num_layers, np.float64), losses, rng, counts


def _worker(rank, world, port, hp_kw, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hp = P.Hyper(**hp_kw)
    pr = synth.make_problem(hp, 13, 3, seed=seed, done_p=0.05)
    flat, losses, rng, counts = sharded_update(rank, world, pr, hp)
    gathered = [None] * world
    dist.all_gather_object(gathered, (flat, losses, rng, counts))
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2])
def test_env_sharded_update_equals_single_process(world):
    hp_kw = dict(num_envs=8, num_steps=6, num_minibatches=4, update_epochs=2, anneal_lr=False, hidden_size=16,
                 num_layers=2, ent_coef=0.01)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, hp_kw, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hp = P.Hyper(**hp_kw)
    pr = synth.make_problem(hp, 13, 3, seed=5, done_p=0.05)
    p0 = P.tree_like(pr["params"], lambda x: x.astype(np.float64))
    p1, o1, rng, losses, aux = P.update(p0, P.init_opt_state(p0), pr["traj"], pr["last_val"], pr["rng"], hp)
    ref = P.flatten_params(p1, hp.num_layers, np.float64)
    for flat, ls, r, counts in gathered:
        np.testing.assert_allclose(flat, ref, rtol=1e-9, atol=1e-12)       # every rank: identical params, equal to 1-process
        np.testing.assert_allclose(ls, losses, rtol=1e-9, atol=1e-12)
        assert np.array_equal(r, rng)
    assert np.array_equal(gathered[0][0], gathered[1][0])                  # bitwise identical across ranks
    total = sum(g[3] for g in gathered)
    assert np.all(total == hp.minibatch_size)                               # every row owned exactly once


def test_shard_rows_partition_property():
    g = np.random.default_rng(0)
    N, T, G = 12, 5, 4
    perm = g.permutation(N * T).astype(np.int32)
    seen = []
    for r in range(G):
        loc = shard_rows(perm, N, r * (N // G), N // G)
        t, n = loc // (N // G), loc % (N // G)
        seen.append(t * N + n + r * (N // G))
    allr = np.concatenate(seen)
    assert sorted(allr.tolist()) == list(range(N * T))
    # order within a rank follows the permutation order
    r0 = seen[0]
    pos = {v: i for i, v in enumerate(perm.tolist())}
    assert all(pos[a] < pos[b] for a, b in zip(r0[:-1].tolist(), r0[1:].tolist()))
