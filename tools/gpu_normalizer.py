"""Observation-normaliser update on the GPU box: HBM bandwidth of the one-pass statistics kernel at the rollout's batch
(20 x 8192 x 232 fp32 = 152 MB), and — under torchrun — that the sharded update with its single NCCL all-reduce equals
the one-process update of the concatenated batch.
    python tools/gpu_normalizer.py            python -m torch.distributed.run --nproc-per-node 2 tools/gpu_normalizer.py"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nz = importlib.import_module("vnl-brax-imitation_b200.normalizer")
sh = importlib.import_module("vnl-brax-imitation_b200.sharding")


def main():
    rank, local_rank, world = sh.env_info()
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        sh.init_process_group("nccl")
    W, T, B = 232, 20, 8192
    res = {"n_gpus": world}
    # sharded == concatenated (every rank builds the same full batch, updates on its shard with the all-reduce, and on the
    # full batch without)
    g = torch.Generator(device=dev).manual_seed(3)
    full = torch.randn(world * 4096, W, device=dev, generator=g) * 2 + 0.3
    a, b = nz.RunningStatistics(W, dev), nz.RunningStatistics(W, dev)
    for it in range(3):
        x = full + it
        lo, hi = sh.shard_range(x.shape[0], rank, world)
        a.update(x[lo:hi].contiguous())
        b.update(x, group_reduce=False)
    torch.cuda.synchronize()
    res["sharded_vs_whole_max_rel"] = {k: float(((getattr(a, k) - getattr(b, k)).abs() / (getattr(b, k).abs() + 1e-6)).max())
                                       for k in ("count", "mean", "summed_variance", "std")}
    # bandwidth
    x = torch.randn(T, B, W, device=dev)
    st = nz.RunningStatistics(W, dev)
    for _ in range(3):
        st.update(x)
    torch.cuda.synchronize()
    sh.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        st.update(x)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    res.update({"update_us": us, "batch_bytes": x.numel() * 4, "GBps_incl_allreduce_and_epilogue": x.numel() * 4 / us * 1e-3,
                "l2_note": "152 MB batch > 126 MB L2: re-read from HBM every update"})
    # the statistics kernel alone (the HBM pass): algorithmic bytes = the batch, once
    stream = torch.cuda.current_stream().cuda_stream
    e0.record()
    for _ in range(20):
        st.lib.vnl_obs_stats_partial(x.data_ptr(), x.numel() // W, W, st.mean.data_ptr(), st.workspace.data_ptr(), st.sums.data_ptr(), stream)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    res.update({"partial_kernel_us": us, "partial_kernel_GBps": x.numel() * 4 / us * 1e-3,
                "partial_kernel_frac_of_measured_hbm_peak_6547": x.numel() * 4 / us * 1e-3 / 6547.2})
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
