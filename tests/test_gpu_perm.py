"""Device permutation (threefry + stable radix sort, minppo_b200/csrc/prng_sort.cu) against the
NumPy restatement of jax.random.split / jax.random.permutation (train.py:252, 258).
Bar: BIT-EXACT indices, both JAX bit-stream modes, and the key chain across epochs."""
import numpy as np
import pytest

from oracle import threefry

pytestmark = pytest.mark.gpu


def _gpu_perms(key, B, epochs, mode, device):
    import torch

    from minppo_b200.learner import permutations

    rng = torch.as_tensor(np.asarray(key, np.uint32).view(np.int32)).to(device)
    rng_out, perms = permutations(rng, B, epochs, mode)
    torch.cuda.synchronize(device)
    return rng_out.cpu().numpy().view(np.uint32), perms.cpu().numpy()


def _oracle_perms(key, B, epochs, mode):
    rng = np.asarray(key, np.uint32)
    out = []
    for _ in range(epochs):
        rng, sub = threefry.split(rng, 2, mode)
        out.append(threefry.permutation(sub, B, mode))
    return rng, np.stack(out)


@pytest.mark.parametrize("mode", [threefry.LEGACY, threefry.PARTITIONABLE])
@pytest.mark.parametrize("B", [1, 2, 5, 160, 1000, 2047, 2048, 2049, 4097, 81920, 262144])
def test_permutation_bit_exact(B, mode, cuda_device):
    key = threefry.prng_key(1337)
    epochs = 4 if B <= 81920 else 2
    rng_ref, ref = _oracle_perms(key, B, epochs, mode)
    rng_out, got = _gpu_perms(key, B, epochs, mode, cuda_device)
    assert np.array_equal(rng_out, rng_ref)
    assert got.dtype == np.int32 and np.array_equal(got, ref)


def test_permutation_other_keys(cuda_device):
    for seed in (0, 1, 42, 2**31 + 5):
        key = threefry.prng_key(seed)
        rng_ref, ref = _oracle_perms(key, 3000, 3, threefry.LEGACY)
        rng_out, got = _gpu_perms(key, 3000, 3, threefry.LEGACY, cuda_device)
        assert np.array_equal(rng_out, rng_ref) and np.array_equal(got, ref)


def test_permutation_is_permutation_at_2pow20(cuda_device):
    """Config-4 batch (B = 2^20): every row is a permutation of arange(B); duplicate sort keys
    exist at this size (birthday bound ~128), so also cross-check against the oracle."""
    B = 1 << 20
    key = threefry.prng_key(1337)
    _, got = _gpu_perms(key, B, 2, threefry.LEGACY, cuda_device)
    for row in got:
        assert np.array_equal(np.sort(row), np.arange(B, dtype=np.int32))
    _, ref = _oracle_perms(key, B, 2, threefry.LEGACY)
    assert np.array_equal(got, ref)
