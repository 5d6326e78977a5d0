"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement, in NumPy uint32 arithmetic, of the PRNG pieces minppo's learner
reaches through ``jax.random``:

* ``jax.random.split(rng)``            -- /root/reference/minppo/train.py:252
* ``jax.random.permutation(_rng, B)``  -- /root/reference/minppo/train.py:258

The arithmetic itself lives in JAX (``jax/_src/prng.py``, ``jax/_src/random.py``), an
UNPINNED, un-vendored dependency (/root/reference/requirements.txt:11 is the bare name
``jax``).  JAX is not installed in this image, so this file restates the published
algorithm (Threefry-2x32, 20 rounds, Random123 / Salmon et al. 2011, as JAX lowers it) and
is anchored on:

* the three Random123 known-answer vectors for threefry2x32,
* ``split(PRNGKey(0))`` in both bit-stream modes,
* three values printed in JAX's own documentation ("The Sharp Bits", legacy mode):
  ``normal(PRNGKey(0), (1,)) == -0.20584226`` and the split/normal chain that follows.

PARITY STATUS: "parity unpinned" by the reference itself -- it ships no tests and no golden
vectors.  The anchors above are external to the reference.

Two bit-stream modes exist because ``jax_threefry_partitionable`` flipped default in JAX
0.5.0 and the reference does not pin JAX:

* ``LEGACY`` (0): counters are ``iota(n)``, split in halves and hashed pairwise.
* ``PARTITIONABLE`` (1): element ``i`` hashes the 64-bit counter ``(hi, lo) = (0, i)``;
  32-bit outputs are ``out0 ^ out1``.
"""
from __future__ import annotations

import math

import numpy as np

LEGACY = 0
PARTITIONABLE = 1

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_PARITY = np.uint32(0x1BD11BDA)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """Threefry-2x32 with 20 rounds; all arguments broadcastable uint32.

    Follows the round/key-injection schedule of jax/_src/prng.py ``_threefry2x32_lowering``:
    five groups of four rounds, rotation constants alternating between the two rows of
    ``_ROT``, key schedule ``ks = [k0, k1, k0 ^ k1 ^ 0x1BD11BDA]``.
    """
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    x0 = np.asarray(x0, dtype=np.uint32).copy()
    x1 = np.asarray(x1, dtype=np.uint32).copy()
    ks = (k0, k1, k0 ^ k1 ^ _PARITY)
    with np.errstate(over="ignore"):
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for r in _ROT[g % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x0 ^ x1
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + np.uint32(g + 1)
    return x0, x1


def prng_key(seed: int) -> np.ndarray:
    """``jax.random.PRNGKey(seed)`` for a seed that fits 32 bits -> ``[0, seed]``
    (/root/reference/minppo/train.py:303; default seed 1337, config.py:77)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def _legacy_hash(key: np.ndarray, count: np.ndarray) -> np.ndarray:
    """jax/_src/prng.py ``threefry_2x32(keypair, count)``: pad to even length, hash the
    two halves against each other, concatenate the two output halves, drop the pad."""
    count = np.asarray(count, dtype=np.uint32).ravel()
    odd = count.size % 2
    if odd:
        count = np.concatenate([count, np.zeros(1, np.uint32)])
    h = count.size // 2
    o0, o1 = threefry2x32(key[0], key[1], count[:h], count[h:])
    out = np.concatenate([o0, o1])
    return out[:-1] if odd else out


def split(key: np.ndarray, num: int = 2, mode: int = LEGACY) -> np.ndarray:
    """``jax.random.split(key, num)`` -> uint32[num, 2]."""
    key = np.asarray(key, dtype=np.uint32)
    if mode == LEGACY:
        return _legacy_hash(key, np.arange(2 * num, dtype=np.uint32)).reshape(num, 2)
    lo = np.arange(num, dtype=np.uint32)
    o0, o1 = threefry2x32(key[0], key[1], np.zeros(num, np.uint32), lo)
    return np.stack([o0, o1], axis=-1)


def random_bits(key: np.ndarray, n: int, mode: int = LEGACY) -> np.ndarray:
    """``jax.random.bits``-style ``_random_bits(key, 32, (n,))`` -> uint32[n]."""
    key = np.asarray(key, dtype=np.uint32)
    if mode == LEGACY:
        return _legacy_hash(key, np.arange(n, dtype=np.uint32))
    lo = np.arange(n, dtype=np.uint32)
    o0, o1 = threefry2x32(key[0], key[1], np.zeros(n, np.uint32), lo)
    return o0 ^ o1


def shuffle_rounds(n: int) -> int:
    """Number of sort rounds in jax/_src/random.py ``_shuffle``:
    ``ceil(3 * ln(max(1, n)) / ln(2**32 - 1))``."""
    return int(np.ceil(3 * np.log(max(1, n)) / np.log(np.iinfo(np.uint32).max)))


def permutation(key: np.ndarray, n: int, mode: int = LEGACY) -> np.ndarray:
    """``jax.random.permutation(key, n)`` for integer ``n`` -> int32[n]
    (/root/reference/minppo/train.py:258).  ``_shuffle``: per round split the key, draw
    32 random bits per element and do a STABLE key-value sort (lax.sort_key_val)."""
    x = np.arange(n, dtype=np.int32)
    key = np.asarray(key, dtype=np.uint32)
    for _ in range(shuffle_rounds(n)):
        key, sub = split(key, 2, mode)
        bits = random_bits(sub, n, mode)
        x = x[np.argsort(bits, kind="stable")]
    return x


def epoch_key_chain(rng: np.ndarray, epochs: int, mode: int = LEGACY):
    """Key chain of one update: per epoch ``rng, _rng = split(rng)``
    (/root/reference/minppo/train.py:252).  Returns (rng_out, [perm keys])."""
    rng = np.asarray(rng, dtype=np.uint32)
    keys = []
    for _ in range(epochs):
        rng, sub = split(rng, 2, mode)
        keys.append(sub)
    return rng, keys


# --- helpers used only to check the documented JAX anchors -------------------------------

def _erfinv_f32(x: np.ndarray) -> np.ndarray:
    from scipy.special import erfinv

    return erfinv(x.astype(np.float64)).astype(np.float32)


_GILES_LT = (2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087, -0.00125372503,
             -0.00417768164, 0.246640727, 1.50140941)
_GILES_GE = (-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773, -0.0076224613,
             0.00943887047, 1.00167406, 2.83297682)


def erfinv_xla_f32(x: np.ndarray) -> np.ndarray:
    """XLA's single-precision ``erf_inv`` as ``jax.random.normal`` reaches it (``lax.erf_inv``): M. Giles'
    polynomial -- ``w = -log1p(-x*x)``; ``w < 5``: degree-8 polynomial in ``w - 2.5``; else in ``sqrt(w) - 3``;
    result ``p * x``.  float32 throughout (third-party arithmetic restated from the published algorithm;
    anchored by the JAX-documentation normal() values in tests/test_oracle_threefry.py)."""
    x = np.asarray(x, np.float32)
    f = np.float32
    with np.errstate(divide="ignore", invalid="ignore"):
        w = -np.log1p(-(x * x)).astype(np.float32)
        lt = w < f(5.0)
        wl = (w - f(2.5)).astype(np.float32)
        wg = (np.sqrt(np.where(lt, f(9.0), w)).astype(np.float32) - f(3.0)).astype(np.float32)
        ww = np.where(lt, wl, wg).astype(np.float32)
        p = np.where(lt, f(_GILES_LT[0]), f(_GILES_GE[0])).astype(np.float32)
        for cl, cg in zip(_GILES_LT[1:], _GILES_GE[1:]):
            p = (np.where(lt, f(cl), f(cg)) + p * ww).astype(np.float32)
        out = (p * x).astype(np.float32)
    return np.where(np.abs(x) == f(1.0), np.copysign(f(np.inf), x), out).astype(np.float32)


def normal_from_bits_f32(bits: np.ndarray, erfinv: str = "xla") -> np.ndarray:
    """``jax.random.normal`` element-wise from its 32 random bits (jax/_src/random.py ``_uniform`` + ``_normal_real``):
    mantissa trick -> [0,1) -> uniform(nextafter(-1,0), 1) -> sqrt(2) * erf_inv(u)."""
    bits = np.asarray(bits, np.uint32)
    f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    hi = np.float32(1.0)
    u = np.maximum(lo, f * (hi - lo) + lo).astype(np.float32)
    e = erfinv_xla_f32(u) if erfinv == "xla" else _erfinv_f32(u)
    return (np.float32(math.sqrt(2.0)) * e).astype(np.float32)


def normal_f32(key: np.ndarray, n: int, mode: int = LEGACY, erfinv: str = "scipy") -> np.ndarray:
    """``jax.random.normal(key, (n,))`` in float32.  ``erfinv="scipy"``: correctly rounded erf_inv (checks the anchors
    quoted from JAX's documentation); ``"xla"``: the polynomial XLA evaluates (what the policy sampler mirrors)."""
    return normal_from_bits_f32(random_bits(key, n, mode), erfinv)
