"""BASELINE.json configs[4]: rodent_pair.xml (two replicated rodents, nv 146, 114 floor contacts, nefc 590) physics-only
throughput sweep.  The reference has no env class for this model (render-only overlay, train.py:295-320), so one "env step"
here is PipelineEnv.pipeline_step = 5 x mjx.step through `vnl_pipeline_step`, random U(-1,1) controls, no task logic.

    python tools/sweep_pair.py [--sizes 1024 2048 ...] [--steps 20]        (one GPU; under torchrun: env shards per rank)"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="*", default=[1024, 2048, 4096, 8192, 16384, 32768, 65536])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--model", default="rodent_pair")
    args = ap.parse_args()
    mjcf = importlib.import_module("vnl-brax-imitation_b200.mjcf")
    mb = importlib.import_module("vnl-brax-imitation_b200.model_blob")
    lib = importlib.import_module("vnl-brax-imitation_b200._lib")
    sh = importlib.import_module("vnl-brax-imitation_b200.sharding")
    rank, local_rank, world = sh.env_info()
    torch.cuda.set_device(local_rank)
    if world > 1:
        sh.init_process_group("nccl")
    model = mjcf.load_model(os.path.join(ROOT, "vnl-brax-imitation_b200", "data", args.model + "_model.npz"))
    eng = lib.Engine(mb.build_model_blob(model), None, device="cuda:%d" % local_rank)
    rng = np.random.default_rng(rank)
    for B in args.sizes:
        qpos = np.tile(model.arrays["qpos0"], (B, 1)).astype(np.float32)
        qpos += (1e-3 * rng.standard_normal(qpos.shape)).astype(np.float32)
        st = dict(qpos=torch.tensor(qpos, device=eng.device), qvel=torch.zeros(B, model.nv, device=eng.device),
                  act=torch.zeros(B, model.na, device=eng.device), qacc_warmstart=torch.zeros(B, model.nv, device=eng.device))
        a, b = eng.alloc_state(B), eng.alloc_state(B)
        for k in ("qpos", "qvel", "act", "qacc_warmstart"):
            a[k].copy_(st[k])
        ctrl = torch.rand(args.warmup + args.steps, B, model.nu, device=eng.device) * 2 - 1
        stats = torch.zeros(B, 4, dtype=torch.int32, device=eng.device)
        for i in range(args.warmup):
            eng.pipeline_step(a, ctrl[i], b, 5, stats)
            a, b = b, a
        torch.cuda.synchronize()
        sh.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.warmup, args.warmup + args.steps):
            eng.pipeline_step(a, ctrl[i], b, 5, stats)
            a, b = b, a
        e1.record()
        torch.cuda.synchronize()
        ms = sh.reduce_scalars(dict(ms=e0.elapsed_time(e1) / args.steps), op="max", device=eng.device)["ms"]
        finite = bool(torch.isfinite(a["qpos"]).all())
        if rank == 0:
            st_mean = stats.float().mean(0).tolist()
            print(json.dumps({"model": args.model, "envs_per_gpu": B, "n_gpus": world, "ms_per_env_step": ms,
                              "env_steps_per_s": world * B / ms * 1e3, "physics_substeps_per_s": 5 * world * B / ms * 1e3,
                              "envs_per_cta": eng.envs_per_cta, "finite": finite,
                              "mean_solver_iters_ls_contacts_limits_per_env_step": st_mean}), flush=True)


if __name__ == "__main__":
    main()
