"""Pin the PRNG oracle (oracle/threefry.py) on anchors EXTERNAL to this repo.

The reference delegates to jax.random (train.py:252, 258); JAX is unpinned and not installable
here, so the anchors are: the Random123 known-answer vectors for Threefry-2x32/20, the
``split(PRNGKey(0))`` values of both JAX bit-stream modes, and values printed in JAX's own
documentation ("JAX - The Sharp Bits": the PRNGKey(0) split/normal chain; "Pseudorandom numbers":
normal(key(42)))."""
import numpy as np
import pytest

from oracle import threefry as tf


def _h(a, b, c, d):
    o = tf.threefry2x32(np.uint32(a), np.uint32(b), np.uint32(c), np.uint32(d))
    return int(o[0]), int(o[1])


def test_random123_known_answers():
    assert _h(0, 0, 0, 0) == (0x6B200159, 0x99BA4EFE)
    assert _h(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == (0x1CB996FC, 0xBB002BE7)
    assert _h(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3) == (0xC4923A9C, 0x483DF7A0)


def test_prng_key_layout():
    assert tf.prng_key(1337).tolist() == [0, 1337]            # config.py:77 default seed
    assert tf.prng_key(0).tolist() == [0, 0]


def test_split_key0_both_modes():
    k = tf.prng_key(0)
    assert tf.split(k, 2, tf.LEGACY).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert tf.split(k, 2, tf.PARTITIONABLE).tolist() == [[1797259609, 2579123966], [928981903, 3453687069]]


def test_jax_sharp_bits_chain_legacy():
    """key = PRNGKey(0): normal(key,(1,)) = -0.20584226; key, subkey = split(key):
    key = [4146024105 967050713], subkey = [2718843009 1272950319], normal(subkey,(1,)) = -1.2515389;
    again: key = [2384771982 3928867769], subkey = [1278412471 2182328957], normal = -0.58665055."""
    k = tf.prng_key(0)
    assert abs(float(tf.normal_f32(k, 1)[0]) - (-0.20584226)) < 2e-6
    k, sub = tf.split(k, 2)
    assert k.tolist() == [4146024105, 967050713] and sub.tolist() == [2718843009, 1272950319]
    assert abs(float(tf.normal_f32(sub, 1)[0]) - (-1.2515389)) < 2e-6
    k, sub = tf.split(k, 2)
    assert k.tolist() == [2384771982, 3928867769] and sub.tolist() == [1278412471, 2182328957]
    assert abs(float(tf.normal_f32(sub, 1)[0]) - (-0.58665055)) < 2e-6


def test_jax_docs_normal_key42():
    assert abs(float(tf.normal_f32(tf.prng_key(42), 1)[0]) - (-0.18471177)) < 2e-6


def test_shuffle_round_counts():
    # 1 round for config 1, 2 for configs 2/4/5
    assert tf.shuffle_rounds(1) == 0
    assert tf.shuffle_rounds(160) == 1
    assert tf.shuffle_rounds(1625) == 1 and tf.shuffle_rounds(1626) == 2      # 3 ln B > ln(2^32 - 1) from B = 1626
    assert tf.shuffle_rounds(81920) == 2 and tf.shuffle_rounds(262144) == 2 and tf.shuffle_rounds(1 << 20) == 2


@pytest.mark.parametrize("mode", [tf.LEGACY, tf.PARTITIONABLE])
@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 160, 1001])
def test_permutation_properties(n, mode):
    p = tf.permutation(tf.prng_key(5), n, mode)
    assert p.dtype == np.int32 and sorted(p.tolist()) == list(range(n))
    # deterministic
    assert np.array_equal(p, tf.permutation(tf.prng_key(5), n, mode))
    if n >= 160:
        assert not np.array_equal(p, tf.permutation(tf.prng_key(6), n, mode))


def test_legacy_bits_odd_length_is_prefix_consistent():
    """Legacy random_bits pads odd lengths with one zero counter and drops the last output."""
    k = tf.prng_key(9)
    b7 = tf.random_bits(k, 7, tf.LEGACY)
    c = np.concatenate([np.arange(7, dtype=np.uint32), np.zeros(1, np.uint32)])
    o0, o1 = tf.threefry2x32(k[0], k[1], c[:4], c[4:])
    assert np.array_equal(b7, np.concatenate([o0, o1])[:7])


def test_stability_matters_with_duplicate_keys():
    """A stable sort on duplicate keys keeps input order -- the property the device radix sort must keep."""
    keys = np.array([5, 1, 5, 1, 5], np.uint32)
    x = np.arange(5, dtype=np.int32)
    assert x[np.argsort(keys, kind="stable")].tolist() == [1, 3, 0, 2, 4]


def test_epoch_key_chain_matches_sequential_splits():
    rng = tf.prng_key(1337)
    out, keys = tf.epoch_key_chain(rng, 4)
    r = rng
    for e in range(4):
        r, s = tf.split(r, 2)
        assert np.array_equal(s, keys[e])
    assert np.array_equal(r, out)
