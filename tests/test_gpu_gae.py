"""GAE kernel (minppo_b200/csrc/gae.cu) against the fp64 oracle of train.py:185-205.
north_star tolerance: advantages and targets within 1e-5 relative in fp32 -- taken here as
max |x - ref| <= 1e-5 * max |ref| against the float64 recurrence."""
import numpy as np
import pytest

from oracle import ppo_numpy as P

pytestmark = pytest.mark.gpu

GAMMA, LAM = 0.99, 0.95


def _gpu_gae(reward, value, done, last_val, device, chunks=0):
    import torch

    from minppo_b200.learner import Memory, calculate_gae

    t = lambda x: torch.as_tensor(np.ascontiguousarray(x)).to(device)
    mem = Memory(done=t(done), action=None, value=t(value), reward=t(reward), log_prob=None, obs=None)
    adv, tgt = calculate_gae(mem, t(last_val), GAMMA, LAM, chunks)
    torch.cuda.synchronize(device)
    return adv.cpu().numpy(), tgt.cpu().numpy()


def _case(T, N, done_p, seed):
    g = np.random.default_rng(seed)
    reward = g.standard_normal((T, N)).astype(np.float32)
    value = g.standard_normal((T, N)).astype(np.float32)
    done = g.random((T, N)) < done_p
    last_val = g.standard_normal(N).astype(np.float32)
    return reward, value, done, last_val


@pytest.mark.parametrize("T,N", [(1, 4), (1, 7), (10, 16), (10, 2048), (128, 2048), (64, 16384), (17, 1001), (33, 5),
                                 (256, 4096), (1024, 1024)])
@pytest.mark.parametrize("done_p", [0.0, 0.01, 0.5, 1.0])
def test_gae_matches_fp64_oracle(T, N, done_p, cuda_device):
    reward, value, done, last_val = _case(T, N, done_p, seed=T * 131 + N)
    ref_adv, ref_tgt = P.gae(reward, value, done, last_val, GAMMA, LAM, np.float64)
    adv, tgt = _gpu_gae(reward, value, done, last_val, cuda_device)
    assert np.abs(adv - ref_adv).max() <= 1e-5 * np.abs(ref_adv).max()
    assert np.abs(tgt - ref_tgt).max() <= 1e-5 * np.abs(ref_tgt).max()


@pytest.mark.parametrize("chunks", [1, 2, 4, 16])
def test_gae_chunked_scan_equals_sequential(chunks, cuda_device):
    reward, value, done, last_val = _case(128, 2048, 0.01, seed=3)
    ref_adv, ref_tgt = P.gae(reward, value, done, last_val, GAMMA, LAM, np.float64)
    adv, tgt = _gpu_gae(reward, value, done, last_val, cuda_device, chunks=chunks)
    assert np.abs(adv - ref_adv).max() <= 1e-5 * np.abs(ref_adv).max()
    assert np.abs(tgt - ref_tgt).max() <= 1e-5 * np.abs(ref_tgt).max()


def test_gae_sequential_path_is_bitwise_fp32_recurrence(cuda_device):
    """chunks == 1 runs the reference's scan order; compare with the same recurrence in float32
    NumPy.  FMA contraction on the GPU differs from NumPy's separate multiply/add by at most a
    few ulp per step, so this is a tight (1e-6) check, not bit-equality."""
    reward, value, done, last_val = _case(64, 512, 0.05, seed=11)
    ref_adv, _ = P.gae(reward, value, done, last_val, GAMMA, LAM, np.float32)
    adv, _ = _gpu_gae(reward, value, done, last_val, cuda_device, chunks=1)
    assert np.abs(adv - ref_adv).max() <= 2e-6 * np.abs(ref_adv).max()


def test_gae_closed_forms_large(cuda_device):
    """Size-independent properties at a bandwidth-relevant size (T=256, N=262144: 67M transitions):
    done == 1 everywhere  =>  adv = reward - value, tgt = reward;  tgt - adv == value always."""
    import torch

    from minppo_b200.learner import Memory, calculate_gae

    T, N = 256, 262144
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(5)
    reward = torch.randn(T, N, device=dev, generator=g)
    value = torch.randn(T, N, device=dev, generator=g)
    last_val = torch.randn(N, device=dev, generator=g)
    done = torch.ones(T, N, dtype=torch.bool, device=dev)
    adv, tgt = calculate_gae(Memory(done, None, value, reward, None, None), last_val, GAMMA, LAM)
    assert torch.equal(adv, reward - value)
    assert torch.allclose(tgt, reward, atol=1e-6)
    done = torch.rand(T, N, device=dev, generator=g) < 0.01
    adv, tgt = calculate_gae(Memory(done, None, value, reward, None, None), last_val, GAMMA, LAM)
    assert torch.equal(tgt, adv + value)
    # linearity in (reward, value, last_val): GAE(2x) == 2 GAE(x) exactly in binary floating point
    adv2, _ = calculate_gae(Memory(done, None, 2 * value, 2 * reward, None, None), 2 * last_val, GAMMA, LAM)
    assert torch.equal(adv2, 2 * adv)
