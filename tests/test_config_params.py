"""Host logic that mirrors the reference's config keys and checkpoint layout (no GPU)."""
import os
import pickle

import numpy as np
import pytest

from minppo_b200 import config, params


def test_defaults_match_reference_config():
    """/root/reference/minppo/config.py:50-84."""
    c = config.Config()
    assert (c.model.hidden_size, c.model.num_layers, c.model.use_tanh) == (256, 2, True)
    assert (c.opt.lr, c.opt.max_grad_norm) == (3e-4, 0.5)
    assert (c.rl.num_env_steps, c.rl.gamma, c.rl.gae_lambda, c.rl.clip_eps, c.rl.ent_coef, c.rl.vf_coef) == (
        10, 0.99, 0.95, 0.2, 0.0, 0.5)
    t = c.training
    assert (t.lr, t.seed, t.num_envs, t.total_timesteps, t.num_minibatches, t.num_steps, t.update_epochs, t.anneal_lr,
            t.model_save_path) == (3e-4, 1337, 2048, 1_000_000_000, 32, 10, 4, True, "trained_model.pkl")


def test_dotlist_overrides_and_unknown_keys():
    c = config.load_config(["training.num_envs=16", "rl.gamma=0.9", "training.anneal_lr=false", "training.total_timesteps=1e6",
                            "model.use_tanh=False"])
    assert c.training.num_envs == 16 and c.rl.gamma == 0.9 and c.training.anneal_lr is False
    assert c.training.total_timesteps == 1_000_000 and c.model.use_tanh is False
    # BASELINE.json config 1 says rl.num_envs=16; that key does not exist in the reference's structured
    # config and raises there too -- num_envs lives under training.
    with pytest.raises(KeyError):
        config.load_config(["rl.num_envs=16"])
    with pytest.raises(KeyError):
        config.load_config(["nosuch.key=1"])
    with pytest.raises(ValueError):
        config.load_config(["training.num_envs"])


def test_num_env_steps_must_equal_num_steps():
    """the reference flattens [rl.num_env_steps, N] as training.num_steps * N (train.py:260)."""
    c = config.load_config(["rl.num_env_steps=1000"])
    with pytest.raises(ValueError, match="num_env_steps"):
        config.to_c_config(c, 8, 2)
    c = config.load_config(["rl.num_env_steps=128", "training.num_steps=128"])
    cc = config.to_c_config(c, 225, 10, world_size=2, rank=1)
    assert (cc.num_steps, cc.num_envs, cc.obs_dim, cc.act_dim, cc.world_size, cc.rank) == (128, 2048, 225, 10, 2, 1)
    assert cc.adam_eps == 1e-5 and cc.adam_b1 == 0.9 and cc.adam_b2 == 0.999       # train.py:118 + optax defaults
    assert cc.training_lr == 3e-4 and cc.opt_lr == 3e-4 and cc.prng_mode == 0


def test_pickle_layout_roundtrip(tmp_path):
    """train.py:86-89 / infer.py:17-19: {'params': {'MLP_0': {'Dense_i': {'kernel','bias'}}, 'MLP_1': ..., 'log_std'}}."""
    D, A, H, L = 11, 3, 64, 2
    P = params.param_count(D, A, H, L)
    flat = np.arange(P, dtype=np.float32) / P
    tree = params.unflatten_params(flat, D, A, H, L)
    assert set(tree) == {"params"} and set(tree["params"]) == {"MLP_0", "MLP_1", "log_std"}
    assert set(tree["params"]["MLP_0"]) == {"Dense_0", "Dense_1", "Dense_2"}
    assert tree["params"]["MLP_0"]["Dense_0"]["kernel"].shape == (D, H)          # [in, out]: y = x @ kernel + bias
    assert tree["params"]["MLP_0"]["Dense_2"]["kernel"].shape == (H, A)
    assert tree["params"]["MLP_1"]["Dense_2"]["kernel"].shape == (H, 1)
    assert tree["params"]["log_std"].shape == (A,)
    # bias precedes kernel (sorted keys), MLP_0 < MLP_1 < log_std
    assert np.array_equal(tree["params"]["MLP_0"]["Dense_0"]["bias"], flat[:H])
    assert np.array_equal(tree["params"]["log_std"], flat[-A:])
    assert np.array_equal(params.flatten_params(tree, L), flat)
    # the reference's default path is a bare filename, for which its os.makedirs("") raises; guarded here
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        params.save_model(tree, "trained_model.pkl")
        params.save_model(tree, os.path.join("sub", "dir", "m.pkl"))
        back = params.load_model("trained_model.pkl")
        with open(os.path.join("sub", "dir", "m.pkl"), "rb") as f:
            back2 = pickle.load(f)
    finally:
        os.chdir(cwd)
    assert np.array_equal(params.flatten_params(back, L), flat) and np.array_equal(params.flatten_params(back2, L), flat)
    with pytest.raises(ValueError):
        params.unflatten_params(flat[:-1], D, A, H, L)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without the built .so the loader raises with the build command."""
    from minppo_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_learner_needs_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from minppo_b200.learner import Learner

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Learner(config.Config(), 8, 2)
