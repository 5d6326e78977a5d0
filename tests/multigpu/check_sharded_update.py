"""Multi-GPU parity check (run under torchrun on a box with >= 2 B200s; scripts/gpu_multi.sh):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu/check_sharded_update.py

Every rank runs the env-sharded learner update on its shard of ONE global batch; rank 0 also
runs the single-GPU update on the whole batch.  Checks: global permutations and rng bit-exact; parameters bit-identical
on all ranks (no broadcast is ever done); sharded losses / parameters equal the single-GPU ones up to fp32 summation
order (different tile composition) amplified by the bf16 GEMMs -- the same tolerances as tests/test_gpu_update.py.
Not collected by pytest (the driver's `-m gpu` run has one GPU); the CPU twin is tests/test_distributed_cpu.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from minppo_b200.learner import Learner, Memory, TrainState, nccl_unique_id  # noqa: E402
from oracle import ppo_numpy as P  # noqa: E402
from oracle import synth  # noqa: E402
from tests.helpers import hyper_to_config, rel_err  # noqa: E402


def run(hp, pr, dev, world, rank, nccl_id, n0, Nl):
    learner = Learner(hyper_to_config(hp), pr["obs_dim"], pr["act_dim"], dev, world, rank, nccl_id)
    ts = TrainState.create(P.flatten_params(pr["params"], hp.num_layers), dev)
    t = lambda x: torch.as_tensor(np.ascontiguousarray(x)).to(dev)
    tr = {k: v[:, n0:n0 + Nl] for k, v in pr["traj"].items()}
    mem = Memory(done=t(tr["done"]), action=t(tr["action"]), value=t(tr["value"]), reward=t(tr["reward"]),
                 log_prob=t(tr["log_prob"]), obs=t(tr["obs"]))
    rng = torch.as_tensor(pr["rng"].view(np.int32)).to(dev)
    # one env step of the rollout policy on this rank's envs (train.py:157-160): the normal draw is the GLOBAL [N, A] one
    pa, plp, pv, prng, _ = learner.policy_step(ts.params, mem.obs[0].contiguous(), rng)
    pol = torch.cat([pa, plp[:, None], pv[:, None]], dim=1).clone()
    for _ in range(2):                                   # two updates: the second replays the captured graph
        ts, rng_out, losses = learner.update(ts, mem, t(pr["last_val"][n0:n0 + Nl]), rng)
    learner.check()
    out = {"params": ts.params.cpu().numpy(), "losses": losses.cpu().numpy(), "rng": rng_out.cpu().numpy(),
           "perms": learner.read("perms").cpu().numpy(), "step": int(ts.step.item()),
           "counts": learner.read("counts").cpu().numpy(), "pol": pol, "pol_rng": prng.cpu().numpy()}
    learner.close()
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for name, kw in (("small", dict(num_envs=64 * world, num_steps=32, num_minibatches=4, update_epochs=2)),
                     ("wide", dict(num_envs=512 * world, num_steps=64, num_minibatches=8, update_epochs=2))):
        hp = P.Hyper(anneal_lr=False, **kw)
        pr = synth.make_problem(hp, seed=3)
        Nl = hp.num_envs // world
        ids = [nccl_unique_id() if rank == 0 else None]      # one NCCL unique id per communicator
        dist.broadcast_object_list(ids, src=0)
        got = run(hp, pr, dev, world, rank, ids[0], rank * Nl, Nl)
        # parameters bit-identical on every rank
        mine = torch.from_numpy(got["params"]).to(dev)
        allp = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        same = all(torch.equal(allp[0], x) for x in allp)
        allpol = [torch.empty_like(got["pol"]) for _ in range(world)]
        dist.all_gather(allpol, got["pol"])
        if rank == 0:
            ref = run(hp, pr, dev, 1, 0, None, 0, hp.num_envs)
            lr = hp.opt_lr
            nsteps = 2 * hp.update_epochs * hp.num_minibatches
            checks = {
                "params identical on all ranks": same,
                "perms bit-exact": bool(np.array_equal(got["perms"], ref["perms"])),
                "policy step (action, log_prob, value) of the shards == single GPU, bit-exact":
                    bool(torch.equal(torch.cat(allpol, dim=0), ref["pol"]) and np.array_equal(got["pol_rng"], ref["pol_rng"])),
                "rng bit-exact": bool(np.array_equal(got["rng"], ref["rng"])),
                "step": got["step"] == ref["step"] == nsteps,
                "losses": rel_err(got["losses"], ref["losses"]) < 2e-3,
                "params": float(np.abs(got["params"] - ref["params"]).max()) < lr * (2.0 + 0.15 * nsteps),
            }
            if name == "small":
                # ... and against the ORACLE at world > 1 (not only against the single-GPU run of the same kernels): the two
                # updates of run() restated on the global batch with the same bf16 rounding points
                p_o = P.tree_like(pr["params"], lambda x: x.astype(np.float32))
                o_o = P.init_opt_state(p_o)
                for _ in range(2):
                    p_o, o_o, rng_o, l_o, aux_o = P.update(p_o, o_o, pr["traj"], pr["last_val"], pr["rng"], hp, dtype=np.float32, gemm="bf16")
                flat_o = P.flatten_params(p_o, hp.num_layers, np.float64)
                checks["losses vs oracle (bf16 rounding emulated)"] = rel_err(got["losses"], l_o) < 2e-3
                checks["perms vs oracle bit-exact"] = bool(np.array_equal(got["perms"], aux_o["perms"]))
                checks["params vs oracle rms <= 1 lr"] = float(np.sqrt(np.mean((got["params"] - flat_o) ** 2))) < lr
            print(f"[{name}] world={world} rows/rank/minibatch={got['counts'][:4].tolist()} "
                  f"loss rel err {rel_err(got['losses'], ref['losses']):.2e} "
                  f"max |dparam| {np.abs(got['params'] - ref['params']).max():.2e}  {checks}", flush=True)
            ok = ok and all(checks.values())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU PARITY", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
