"""Multi-GPU host logic on CPU: world_size-2 gloo processes shard the env range, step their shards
with the oracle standing in for the device (test infrastructure), and reduce metrics.  The sharded
run must equal the single-process run env for env (no cross-env coupling)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT, pkg, start_states

sh = pkg("sharding")


def test_shard_range_partitions():
    for n in (1, 7, 4096, 4097, 8192):
        for w in (1, 2, 3, 4, 8):
            spans = [sh.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sh.shard_range(4096, 3, 8) == (1536, 2048)  # num_envs // device_count (ppo_imitation/train.py:195)
    with pytest.raises(ValueError):
        sh.shard_range(8, 2, 2)


def _worker(rank, world, port, B, tmp):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    import conftest
    r, _, w = sh.init_process_group("gloo")
    rod = conftest.pkg("envs.rodent"); mb = conftest.pkg("model_blob")
    model, clip = rod.packaged_rodent()
    args = {k: rod.RODENT_ENV_ARGS[k] for k in ("end_eff_names", "appendage_names", "walker_body_names", "joint_names",
                                                "center_of_mass", "clip_length", "sub_clip_length", "ref_traj_length",
                                                "termination_threshold")}
    task, fclip, idx, obs_size, traj_size = rod.rodent_task_tables(model, clip, **args)
    blob = mb.build_model_blob(model)
    rd = dict(fclip=fclip)
    qpos, qvel, start = conftest.start_states(rd, B, seed=0)
    act = np.random.default_rng(1).uniform(-1, 1, size=(B, 30))
    lo, hi = sh.shard_range(B, r, w)
    kw = dict(precision=32, dims=mb.read_dims(blob), obs_size=obs_size, traj_size=traj_size, nthreads=1)
    s, _ = oracle.reset(blob, task, qpos[lo:hi], qvel[lo:hi], start[lo:hi], **kw)
    s, o = oracle.step(blob, task, s, act[lo:hi], **kw)
    sh.barrier()
    tot = sh.reduce_scalars(dict(reward=float(o["reward"].sum()), n=float(hi - lo)))
    mx = sh.reduce_scalars(dict(t=float(rank + 1)), op="max")
    np.savez(os.path.join(tmp, f"r{rank}.npz"), qpos=s["qpos"], reward=o["reward"], lo=lo, hi=hi, tot=tot["reward"], n=tot["n"], mx=mx["t"])
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_sharding_equals_single_process(tmp_path, rodent, oracle_mod):
    B, world, port = 6, 2, 29533 + os.getpid() % 200
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    qpos, qvel, start = start_states(rodent, B, seed=0)
    act = np.random.default_rng(1).uniform(-1, 1, size=(B, 30))
    kw = dict(precision=32, dims=rodent["dims"], obs_size=232, traj_size=795, nthreads=1)
    s, _ = oracle_mod.reset(rodent["model_blob"], rodent["task_blob"], qpos, qvel, start, **kw)
    s, o = oracle_mod.step(rodent["model_blob"], rodent["task_blob"], s, act, **kw)
    got_q = np.concatenate([p["qpos"] for p in parts]); got_r = np.concatenate([p["reward"] for p in parts])
    assert np.array_equal(got_q, s["qpos"]) and np.array_equal(got_r, o["reward"])  # bit-identical, order preserved
    assert [int(p["lo"]) for p in parts] == [0, 3] and [int(p["hi"]) for p in parts] == [3, 6]
    for p in parts:
        assert abs(float(p["tot"]) - float(o["reward"].sum())) < 1e-12 and float(p["n"]) == B and float(p["mx"]) == world
