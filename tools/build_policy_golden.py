"""Golden vectors of the policy / GAE / normaliser CHECKERS (tests/golden/policy_golden.npz).

The generators are this repo's restatements (policy.reference_forward in fp32, tests' brax_update / reference_gae in fp64):
flax / brax are not installable here, so these vectors do not pin the restatements against the libraries — they pin them
against accidental change, and give the GPU tests fixed inputs with recorded outputs.  Parameters and inputs are regenerated
from seeds (numpy default_rng / torch.Generator), only outputs are stored.      python tools/build_policy_golden.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pol = importlib.import_module("vnl-brax-imitation_b200.policy")


def policy_case(seed=0, B=16):
    params = pol.init_params(np.random.default_rng(seed), pol.param_shapes(795, 232, 30), perturb=0.1)
    g = torch.Generator().manual_seed(seed)
    mk = lambda *s: torch.randn(*s, generator=g)
    x = dict(traj=mk(B, 795), obs=mk(B, 232) * 2 + 0.5, eps_z=mk(B, 64), eps_a=mk(B, 30), rand=torch.rand(B, 30, generator=g) * 2 - 1,
             mean=mk(232) * 0.3, std=torch.rand(232, generator=g) + 0.5)
    return params, x


def main():
    params, x = policy_case()
    r = pol.reference_forward(params, x["traj"], x["obs"], x["eps_z"], x["eps_a"], x["rand"], x["mean"], x["std"])
    rb = pol.reference_forward(params, x["traj"], x["obs"], x["eps_z"], x["eps_a"], x["rand"], x["mean"], x["std"], operand_dtype=torch.bfloat16)
    out = {"fp32_" + k: r[k].numpy() for k in ("logits", "action", "raw_action", "log_prob", "rand_log_prob", "z_mean", "z_logvar")}
    out.update({"bf16ops_" + k: rb[k].numpy() for k in ("logits", "action", "log_prob")})
    from test_ppo import _case, reference_gae
    c = _case(20, 64, seed=7)
    vs, adv = reference_gae(*[t.double() for t in c], lambda_=0.95, discount=0.9)
    out["gae_vs"], out["gae_adv"] = vs.numpy(), adv.numpy()
    from test_normalizer import brax_update
    g = torch.Generator().manual_seed(11)
    st = (torch.zeros((), dtype=torch.float64), torch.zeros(232, dtype=torch.float64), torch.zeros(232, dtype=torch.float64))
    for i in range(3):
        st, std = brax_update(st, [(torch.randn(500, 232, generator=g) * (1 + i) + 0.5 * i).double()])
    out["norm_mean"], out["norm_sv"], out["norm_std"] = st[1].numpy(), st[2].numpy(), std.numpy()
    path = os.path.join(ROOT, "tests", "golden", "policy_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
