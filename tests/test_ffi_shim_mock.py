"""The XLA-FFI shim (minppo_b200/csrc/xla_ffi_shim.cc) against a MOCK of the FFI C++ API (tests/stubs/xla/ffi/api/ffi.h).

jaxlib is not installable here, so the shim cannot meet the real headers.  What these
tests do prove: the shim is valid C++ against the API shape it targets; its handler bodies forward to the C ABI correctly
(bit-identical with `Learner.update`); its context cache is keyed on the WHOLE `minppo_config` (two calls with equal shapes
but different hyper-parameters get two contexts -- the round-1 advisor finding); the in-place aliasing contract is
enforced.  They prove nothing about binary compatibility with a real jaxlib."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_SO = os.path.join(ROOT, "minppo_b200", "lib", "libminppo_ffi_mock.so")


def build_mock():
    cmd = ["g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-I" + os.path.join(ROOT, "tests", "stubs"),
           "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "stubs", "ffi_mock_driver.cc"),
           "-L" + os.path.join(ROOT, "minppo_b200", "lib"), "-lminppo_b200", "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath,$ORIGIN", "-o", MOCK_SO]
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)


def test_shim_compiles_against_the_mock_ffi_header():
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "stubs"),
                        "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "minppo_b200", "csrc", "xla_ffi_shim.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_mock_build_exports_every_handler_the_python_side_registers():
    """minppo_b200/jax_ffi.py registers MinppoGae, MinppoUpdate, MinppoPolicyStep and MinppoBootstrapValue (the critic-only
    bootstrap value, train.py:182-183): the shim defines each of them (through the mock's handler macro: <name>_mock_symbol)."""
    import re

    r = build_mock()
    assert r.returncode == 0, r.stderr[-3000:]
    src = open(os.path.join(ROOT, "minppo_b200", "jax_ffi.py")).read()
    wanted = sorted(set(re.findall(r"pycapsule\(lib\.(\w+)\)", src)))
    assert wanted == ["MinppoBootstrapValue", "MinppoGae", "MinppoPolicyStep", "MinppoUpdate"]
    nm = subprocess.run(["nm", "-D", "--defined-only", MOCK_SO], capture_output=True, text=True).stdout
    for name in wanted:
        assert f"{name}_mock_symbol" in nm, name


@pytest.mark.gpu
def test_shim_handlers_forward_and_key_contexts_on_the_whole_config(cuda_device):
    import torch

    from minppo_b200.learner import Learner
    from oracle import ppo_numpy as P
    from oracle import synth
    from tests.helpers import hyper_to_config

    if not os.path.exists(MOCK_SO):
        r = build_mock()
        assert r.returncode == 0, r.stderr[-3000:]
    lib = C.CDLL(MOCK_SO)
    lib.mock_last_message.restype = C.c_char_p
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.mock_update.argtypes = [vp] * 15 + [i64] * 5 + [i32] * 5 + [f32, f32, i32]
    lib.mock_gae.argtypes = [vp] * 7 + [i64, i64, f32, f32]

    hp = P.Hyper(num_envs=64, num_steps=32, num_minibatches=4, update_epochs=2, anneal_lr=False)
    D, A = 225, 10
    pr = synth.make_problem(hp, D, A, seed=21, done_p=0.02)
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x)).to(cuda_device)
    tr = pr["traj"]
    dev_in = {k: t(tr[k]) for k in ("obs", "action", "value", "reward", "log_prob", "done")}
    lv = t(pr["last_val"])
    flat = P.flatten_params(pr["params"], hp.num_layers)
    rng = torch.as_tensor(pr["rng"].view(np.int32)).to(cuda_device)
    stream = torch.cuda.current_stream(cuda_device).cuda_stream

    def run_shim(lr, alias_ok=1):
        params = torch.as_tensor(flat).to(cuda_device)
        mu, nu = torch.zeros_like(params), torch.zeros_like(params)
        count = torch.zeros(1, dtype=torch.int32, device=cuda_device)
        rng_out = torch.zeros_like(rng)
        losses = torch.zeros((hp.update_epochs, hp.num_minibatches, 4), device=cuda_device)
        rc = lib.mock_update(stream, params.data_ptr(), mu.data_ptr(), nu.data_ptr(), count.data_ptr(), dev_in["obs"].data_ptr(),
                             dev_in["action"].data_ptr(), dev_in["value"].data_ptr(), dev_in["reward"].data_ptr(),
                             dev_in["log_prob"].data_ptr(), dev_in["done"].data_ptr(), lv.data_ptr(), rng.data_ptr(),
                             rng_out.data_ptr(), losses.data_ptr(), hp.num_steps, hp.num_envs, D, A, flat.size,
                             hp.num_minibatches, hp.update_epochs, hp.hidden_size, hp.num_layers, 0, lr, hp.clip_eps, alias_ok)
        torch.cuda.synchronize(cuda_device)
        return rc, params.cpu().numpy(), losses.cpu().numpy(), rng_out.cpu().numpy()

    with torch.cuda.device(cuda_device):
        n0 = lib.mock_ctx_count()
        rc, p_a, l_a, r_a = run_shim(3e-4)
        assert rc == 0, lib.mock_last_message()
        assert lib.mock_ctx_count() == n0 + 1
        rc, p_a2, l_a2, _ = run_shim(3e-4)                      # same config: same context, same result
        assert rc == 0 and lib.mock_ctx_count() == n0 + 1
        assert np.array_equal(p_a, p_a2) and np.array_equal(l_a, l_a2)
        rc, p_b, l_b, _ = run_shim(1e-3)                        # same shapes, other learning rate: its OWN context
        assert rc == 0 and lib.mock_ctx_count() == n0 + 2
        assert np.abs(p_b - p_a).max() > 1e-4                   # ... and the other learning rate really was used
        rc, *_ = run_shim(3e-4, alias_ok=0)                     # params result not aliased to the operand
        assert rc != 0 and b"aliased" in lib.mock_last_message()

        # bit-identical with the Python host on the same inputs (same library underneath)
        from minppo_b200.learner import Memory, TrainState

        hp_cfg = hyper_to_config(hp, use_graph=False)
        lrn = Learner(hp_cfg, D, A, cuda_device)
        ts = TrainState.create(flat, cuda_device)
        mem = Memory(dev_in["done"], dev_in["action"], dev_in["value"], dev_in["reward"], dev_in["log_prob"], dev_in["obs"])
        ts, r_out, losses = lrn.update(ts, mem, lv, rng)
        lrn.check()
        assert np.array_equal(ts.params.cpu().numpy(), p_a)
        assert np.array_equal(losses.cpu().numpy(), l_a)
        assert np.array_equal(r_out.cpu().numpy(), r_a)
        lrn.close()

        # GAE handler
        adv = torch.zeros_like(dev_in["reward"]); tgt = torch.zeros_like(dev_in["reward"])
        rc = lib.mock_gae(stream, dev_in["reward"].data_ptr(), dev_in["value"].data_ptr(), dev_in["done"].data_ptr(), lv.data_ptr(),
                          adv.data_ptr(), tgt.data_ptr(), hp.num_steps, hp.num_envs, hp.gamma, hp.gae_lambda)
        torch.cuda.synchronize(cuda_device)
        assert rc == 0
        a_o, t_o = P.gae(tr["reward"], tr["value"], tr["done"], pr["last_val"], hp.gamma, hp.gae_lambda, np.float64)
        assert np.abs(adv.cpu().numpy() - a_o).max() <= 1e-5 * np.abs(a_o).max()
