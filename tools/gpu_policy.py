"""Developer check of the policy-forward kernel (row f1) on the GPU box: per-layer accumulator errors against the torch
restatement (bf16-rounded operands, fp32 accumulation), output errors against it and against plain fp32, and launch timing
at 8192 envs beside the torch / cuBLAS forward.   python tools/gpu_policy.py [B]"""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pol = importlib.import_module("vnl-brax-imitation_b200.policy")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    shapes = pol.param_shapes(795, 232, 30)
    params = pol.init_params(rng, shapes, perturb=0.1)
    g = torch.Generator(device="cpu").manual_seed(1)
    mk = lambda *s: torch.randn(*s, generator=g).to(dev)
    traj, obs, eps_z, eps_a = mk(B, 795), mk(B, 232) * 2 + 0.5, mk(B, 64), mk(B, 30)
    rand_a = (torch.rand(B, 30, generator=g) * 2 - 1).to(dev)
    mean, std = mk(232) * 0.3, torch.rand(232, generator=g).to(dev) + 0.5
    p = pol.IntentionPolicy(params, str(dev), mean, std)
    ref = pol.reference_forward(params, traj, obs, eps_z, eps_a, rand_a, mean, std, operand_dtype=torch.bfloat16)
    ref32 = pol.reference_forward(params, traj, obs, eps_z, eps_a, rand_a, mean, std)
    res = {"B": B, "desc_mode": os.environ.get("VNL_POLICY_DESC_MODE", "0")}
    bias = ["encoder/hidden_0/bias", "encoder/hidden_1/bias", None, "decoder/hidden_0/bias", "decoder/hidden_1/bias", "decoder/hidden_2/bias"]
    for layer in range(6):
        d = p.debug_layer(traj, obs, eps_z, layer)
        torch.cuda.synchronize()
        n = min(B, 128)
        r = ref["pre"][layer][:n]
        if bias[layer]:
            r = r - torch.as_tensor(params[bias[layer]], device=dev)
        else:
            r = r - torch.cat([torch.as_tensor(params["encoder/fc2_mean/bias"]), torch.as_tensor(params["encoder/fc2_logvar/bias"])]).to(dev)
        w = r.shape[1]
        res["layer%d_err" % layer] = float((d[:n, :w] - r).abs().max())
        res["layer%d_ref_absmax" % layer] = float(r.abs().max())
    act, out = p(traj, obs, eps_z, eps_a, rand_a, heads=True)
    torch.cuda.synchronize()
    for k in ("action", "raw_action", "logits", "log_prob", "rand_log_prob", "z_mean", "z_logvar"):
        res["out_%s_err_vs_bf16ref" % k] = float((out[k] - ref[k]).abs().max())
        res["out_%s_err_vs_fp32" % k] = float((out[k] - ref32[k]).abs().max())
    print(json.dumps(res), flush=True)

    # timing at the rollout batch
    Bt = 8192
    traj, obs, eps_z, eps_a = mk(Bt, 795), mk(Bt, 232), mk(Bt, 64), mk(Bt, 30)
    outs = p.alloc_outputs(Bt)
    for _ in range(5):
        p(traj, obs, eps_z, eps_a, out=outs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        p(traj, obs, eps_z, eps_a, out=outs)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    flops = 2 * Bt * (795 * 256 + 256 * 128 + 128 * 128 + 296 * 128 + 128 * 256 + 256 * 60)
    tim = {"policy_forward_us_8192": us, "dense_tflops": flops / us * 1e-6}
    import ctypes
    stamps = torch.zeros(32, device=dev)
    ptr = lambda x: None if x is None else x.data_ptr()
    p.lib.vnl_policy_debug(p.blob_dev.data_ptr(), ctypes.byref(p.dims), Bt, ptr(traj), ptr(obs), ptr(p.obs_mean), ptr(p.obs_std),
                           ptr(eps_z), -1, stamps.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    st = stamps.cpu().numpy()
    names = ["setup", "L0c0", "L0c1", "L0c2", "L0c3", "L0_issued"] + [x for n in range(1, 6) for x in ("mma%d_done" % (n - 1), "epi%d_done" % (n - 1), "w%d_sync" % n)] + ["mma5_done", "end"]
    names += ["(unused)"] + ["mma%d_issued" % n for n in range(1, 6)]
    tim["stamps_cycles"] = {n: int(v) for n, v in zip(names, st[:29]) if n != "(unused)"}
    P = {k: torch.as_tensor(v, device=dev) for k, v in params.items()}
    for _ in range(3):
        pol.reference_forward(P, traj, obs, eps_z, eps_a, None, mean, std, operand_dtype=None)
    torch.cuda.synchronize()
    print(json.dumps(tim), flush=True)


if __name__ == "__main__":
    main()
