"""BASELINE.json configs[3] as a whole: the reference's `training_step` (ppo_imitation/train.py:296-349) -- rollout of ENVS envs x
unroll 20 (policy kernel + fused env step), normaliser update, SGD phase of UPDATES x MINIBATCHES minibatch updates (tcgen05 TF32
GEMMs, Adam) with the gradient all-reduce over NCCL at N > 1 -- timed phase by phase with CUDA events, max over ranks.

    python tools/train_bench.py                                             # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py
  env: ENVS (8192) MINIBATCHES (32) UPDATES (16) STEPS (2) GRAPH (1) X3 (0)
  PROFILE=1 (with GRAPH=0): after the warm-up step, ONE eager minibatch update between cudaProfilerStart / Stop and exit -- the region
  `ncu --profile-from-start off --metrics gpu__time_duration.sum` lists (tools/make_profiles.py summarises the csv)."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def build(dev, rank, world, B, unroll, nmb, nup, graph=True, x3=False):
    P = lambda n: importlib.import_module("vnl-brax-imitation_b200." + n)
    sh, pol, lrn = P("sharding"), P("policy"), P("learner")
    env = bench.make_env("rodent", str(dev))
    eng = env.engine
    qpos, qvel, start = bench.workload_draws("rodent", env, B * world, *sh.shard_range(B * world, rank, world))
    s0 = env.reset_from(qpos, qvel, start)
    rng = np.random.default_rng(0)  # the same initial networks on every rank (key_policy / key_value are global, train.py:190-192)
    pp = pol.init_params(rng, pol.param_shapes(eng.traj_size, eng.obs_size, env.action_size))
    vp = lrn.init_value_params(rng, lrn.value_param_shapes(eng.obs_size))
    stats = P("normalizer").RunningStatistics(eng.obs_size, str(dev))
    policy = (pol.IntentionPolicy if os.environ.get("POLICY") == "bf16" else pol.PrecisePolicy)(pp, str(dev), stats.mean, stats.std)
    ro = P("rollout").Rollout(env, policy, s0, unroll, 150.0, use_graph=True)
    learner = lrn.PPOLearner(pp, vp, unroll, B // nmb, device=str(dev), x3=x3, clipping_epsilon=0.2, kl_weight=1e-4, learning_rate=6e-4)
    tr = P("trainer").Trainer(env, policy, learner, ro, stats, nmb, nup, use_graph=graph, seed=rank)
    return tr


def main():
    sh = importlib.import_module("vnl-brax-imitation_b200.sharding")
    rank, local_rank, world = sh.env_info()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        sh.init_process_group("nccl")
    B, unroll = int(os.environ.get("ENVS", 8192)), 20
    nmb, nup, steps = int(os.environ.get("MINIBATCHES", 32)), int(os.environ.get("UPDATES", 16)), int(os.environ.get("STEPS", 2))
    graph, x3 = os.environ.get("GRAPH", "1") == "1", os.environ.get("X3", "0") == "1"
    tr = build(dev, rank, world, B, unroll, nmb, nup, graph, x3)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    m = tr.training_step()  # warm-up: graph captures
    torch.cuda.synchronize()
    if os.environ.get("PROFILE") == "1":
        data = tr.rollout.generate_unroll()
        data = dict(data, state_extras_traj_in=tr.rollout.traj[:unroll])
        tr.discount_buf.copy_(data["discount"])
        tr.idx.copy_(torch.randperm(B, device=dev)[:tr.Bm].to(torch.int32))
        tr.eps.normal_(generator=tr.gen)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        tr._minibatch(data)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profiled": "one eager minibatch update", "rows": unroll * tr.Bm}), flush=True)
        return
    sh.barrier()
    t_ro = t_sgd = 0.0
    w0, w1 = ev(), ev()
    w0.record()
    for _ in range(steps):
        a, b, c = ev(), ev(), ev()
        a.record()
        tr.rollout.eps_z.normal_(generator=tr.gen); tr.rollout.eps_a.normal_(generator=tr.gen)
        data = tr.rollout.generate_unroll()
        tr.stats.update(data["observation"])
        tr.learner.set_normalizer(tr.stats.mean, tr.stats.std)
        b.record()
        tr.sgd_phase(data)
        tr.policy.load_params(tr.learner.policy_params())
        c.record()
        torch.cuda.synchronize()
        t_ro += a.elapsed_time(b); t_sgd += b.elapsed_time(c)
    w1.record()
    torch.cuda.synchronize()
    m = tr.learner.metrics_dict()
    red = sh.reduce_scalars(dict(ms=w0.elapsed_time(w1), ro=t_ro, sgd=t_sgd), op="max", device=dev)
    res = {"config": "rodent PPO training_step: %d envs/GPU x unroll %d; SGD phase %d updates x %d minibatches of %d rows (%s, %s)" % (
               B, unroll, nup, nmb, unroll * B // nmb, "3xTF32" if x3 else "TF32", "CUDA graph per minibatch update" if graph else "eager"),
           "n_gpus": world, "training_env_steps_per_s": world * B * unroll * steps / (red["ms"] * 1e-3), "ms_per_training_step": red["ms"] / steps,
           "rollout_ms": red["ro"] / steps, "sgd_phase_ms": red["sgd"] / steps, "ms_per_minibatch_update": red["sgd"] / steps / (nup * nmb),
           "minibatch_updates_per_training_step": nup * nmb, "last_metrics": m}
    flops = 6.0 * 1630397 * (unroll * B // nmb)  # fwd + bwd of both networks per minibatch update
    res["sgd_tflops_dense"] = flops * nup * nmb / (red["sgd"] / steps * 1e-3) / 1e12
    if world > 1:
        import torch.distributed as dist
        g = tr.learner.grads
        for _ in range(5):
            dist.all_reduce(g[:tr.learner.n_policy]); dist.all_reduce(g[tr.learner.n_policy:])
        torch.cuda.synchronize()
        a0, a1 = ev(), ev()
        a0.record()
        for _ in range(50):
            dist.all_reduce(g[:tr.learner.n_policy]); dist.all_reduce(g[tr.learner.n_policy:])
        a1.record()
        torch.cuda.synchronize()
        us = sh.reduce_scalars(dict(us=a0.elapsed_time(a1) / 50 * 1e3), op="max", device=dev)["us"]
        res["grad_allreduce_us_standalone"] = us
        res["grad_allreduce_bytes"] = int(g.numel() * 4)
        res["allreduce_share_of_sgd_phase_if_exposed"] = us * 1e-3 / res["ms_per_minibatch_update"]
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
