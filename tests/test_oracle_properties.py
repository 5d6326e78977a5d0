"""Size-independent properties of the oracle that the CUDA kernels rely on (checked here on CPU with hypothesis;
the `-m gpu` tests check the same properties of the kernels at full sizes).

* GAE is an affine recurrence gae_t = delta_t + c_t * gae_{t+1}: composing per-segment maps (A = prod c, B = gae|0)
  reproduces the sequential scan -- the algebra behind gae_chunked_kernel (minppo_b200/csrc/gae.cu).
* GAE is linear in (reward, value, last_val) jointly for fixed `done`.
* jax.random.normal(key, (N, A)) restricted to a rank's rows is what policy_head_kernel draws for an env shard.
* env-sharded row ownership partitions every minibatch exactly (no row lost or duplicated)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import ppo_numpy as P
from oracle import threefry as tf

GAMMA, LAM = 0.99, 0.95


def _rand(seed, T, N, p_done):
    g = np.random.default_rng(seed)
    return (g.standard_normal((T, N)), g.standard_normal((T, N)), g.random((T, N)) < p_done, g.standard_normal(N))


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), T=st.integers(2, 40), N=st.integers(1, 7), chunks=st.integers(1, 6),
       p_done=st.sampled_from([0.0, 0.05, 0.5, 1.0]))
def test_gae_segment_composition_equals_sequential_scan(seed, T, N, chunks, p_done):
    r, v, d, lv = _rand(seed, T, N, p_done)
    adv_ref, tgt_ref = P.gae(r, v, d, lv, GAMMA, LAM)
    seg = -(-T // chunks)
    nd = 1.0 - d.astype(np.float64)
    nxt = np.vstack([v[1:], lv[None]])                           # next_value of step t (train.py:193-195)
    delta = r + GAMMA * nxt * nd - v
    c = GAMMA * LAM * nd
    # pass 1: per segment, the affine map gae_in -> gae at the segment's first step:  gae_first = B + A * gae_in
    A, B = [], []
    for s0 in range(0, T, seg):
        a, b = np.ones(N), np.zeros(N)
        for t in range(min(T, s0 + seg) - 1, s0 - 1, -1):
            b = delta[t] + c[t] * b
            a = c[t] * a
        A.append(a); B.append(b)
    # carries: gae entering each segment from the later ones (right to left), then pass 2 reruns the segment
    adv = np.zeros((T, N))
    carry = np.zeros(N)
    for k in range(len(A) - 1, -1, -1):
        s0 = k * seg
        g = carry.copy()
        for t in range(min(T, s0 + seg) - 1, s0 - 1, -1):
            g = delta[t] + c[t] * g
            adv[t] = g
        assert np.allclose(g, B[k] + A[k] * carry, rtol=1e-12, atol=1e-12)
        carry = g
    assert np.allclose(adv, adv_ref, rtol=1e-12, atol=1e-12)
    assert np.allclose(adv + v, tgt_ref, rtol=1e-12, atol=1e-12)


@settings(max_examples=20, deadline=None)
@given(seed=st.integers(0, 10_000), a=st.floats(-3, 3), b=st.floats(-3, 3))
def test_gae_is_linear_for_fixed_done(seed, a, b):
    r1, v1, d, lv1 = _rand(seed, 12, 5, 0.1)
    r2, v2, _, lv2 = _rand(seed + 1, 12, 5, 0.1)
    x1, _ = P.gae(r1, v1, d, lv1, GAMMA, LAM)
    x2, _ = P.gae(r2, v2, d, lv2, GAMMA, LAM)
    x, _ = P.gae(a * r1 + b * r2, a * v1 + b * v2, d, a * lv1 + b * lv2, GAMMA, LAM)
    assert np.allclose(x, a * x1 + b * x2, rtol=1e-9, atol=1e-9)


@settings(max_examples=20, deadline=None)
@given(N=st.integers(2, 40), A=st.integers(1, 12), world=st.sampled_from([2, 4, 8]),
       mode=st.sampled_from([tf.LEGACY, tf.PARTITIONABLE]), seed=st.integers(0, 2**31 - 1))
def test_sharded_normal_draw_is_a_row_slice_of_the_global_one(N, A, world, mode, seed):
    N = N * world
    key = tf.prng_key(seed)
    full_bits = tf.random_bits(key, N * A, mode)
    Nl = N // world
    for rank in range(world):
        idx = np.arange(rank * Nl * A, (rank + 1) * Nl * A, dtype=np.uint64)
        # element i of the global draw, computed alone (what random_bits_at does on the device)
        if mode == tf.LEGACY:
            n = N * A
            h = (n + 1) // 2
            lo = idx < h
            c0 = np.where(lo, idx, idx - h).astype(np.uint32)
            c1 = np.where(lo, np.where(idx + h < n, idx + h, 0), idx).astype(np.uint32)
            o0, o1 = tf.threefry2x32(key[0], key[1], c0, c1)
            bits = np.where(lo, o0, o1)
        else:
            o0, o1 = tf.threefry2x32(key[0], key[1], np.zeros(idx.size, np.uint32), idx.astype(np.uint32))
            bits = o0 ^ o1
        assert np.array_equal(bits, full_bits[rank * Nl * A:(rank + 1) * Nl * A])


@settings(max_examples=20, deadline=None)
@given(T=st.integers(1, 9), Nl=st.integers(1, 6), world=st.sampled_from([1, 2, 4]), M=st.sampled_from([1, 2, 4]),
       seed=st.integers(0, 1000))
def test_env_sharding_partitions_every_minibatch(T, Nl, world, M, seed):
    N = Nl * world
    B = T * N
    if B % M:
        return
    perm = tf.permutation(tf.prng_key(seed), B)
    mb = B // M
    for k in range(M):
        sl = perm[k * mb:(k + 1) * mb]
        seen = []
        for rank in range(world):
            n = sl % N
            own = (n >= rank * Nl) & (n < (rank + 1) * Nl)
            local = (sl[own] // N) * Nl + (n[own] - rank * Nl)
            assert local.min(initial=0) >= 0 and local.max(initial=0) < T * Nl
            seen.append(sl[own])
        allrows = np.concatenate(seen)
        assert allrows.size == mb and np.array_equal(np.sort(allrows), np.sort(sl))
