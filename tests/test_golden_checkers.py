"""tests/golden/policy_golden.npz (made by tools/build_policy_golden.py from the repo's restatements at fixed seeds):
CPU — the restatements still reproduce it (guards the checkers against accidental change);
GPU — the kernels reproduce the recorded outputs from the regenerated inputs, through the C ABI."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, pkg

sys.path.insert(0, os.path.join(ROOT, "tools"))
gen = importlib.import_module("build_policy_golden")
pol = pkg("policy")
G = np.load(os.path.join(ROOT, "tests", "golden", "policy_golden.npz"))


def test_policy_restatement_reproduces_golden():
    params, x = gen.policy_case()
    r = pol.reference_forward(params, x["traj"], x["obs"], x["eps_z"], x["eps_a"], x["rand"], x["mean"], x["std"])
    for k in ("logits", "action", "raw_action", "z_mean", "z_logvar"):
        np.testing.assert_allclose(r[k].numpy(), G["fp32_" + k], atol=2e-5, rtol=0)
    np.testing.assert_allclose(r["log_prob"].numpy(), G["fp32_log_prob"], atol=5e-4)


def test_gae_and_normaliser_restatements_reproduce_golden():
    from test_normalizer import brax_update
    from test_ppo import _case, reference_gae
    vs, adv = reference_gae(*[t.double() for t in _case(20, 64, seed=7)], lambda_=0.95, discount=0.9)
    np.testing.assert_allclose(vs.numpy(), G["gae_vs"], atol=1e-12)
    np.testing.assert_allclose(adv.numpy(), G["gae_adv"], atol=1e-12)
    g = torch.Generator().manual_seed(11)
    st = (torch.zeros((), dtype=torch.float64), torch.zeros(232, dtype=torch.float64), torch.zeros(232, dtype=torch.float64))
    for i in range(3):
        st, std = brax_update(st, [(torch.randn(500, 232, generator=g) * (1 + i) + 0.5 * i).double()])
    np.testing.assert_allclose(st[1].numpy(), G["norm_mean"], atol=1e-12)
    np.testing.assert_allclose(std.numpy(), G["norm_std"], atol=1e-12)


@pytest.mark.gpu
def test_kernels_reproduce_golden():
    params, x = gen.policy_case()
    c = {k: v.cuda() for k, v in x.items()}
    p = pol.IntentionPolicy(params, "cuda:0", c["mean"], c["std"])
    _, out = p(c["traj"], c["obs"], c["eps_z"], c["eps_a"], c["rand"], heads=True)
    torch.cuda.synchronize()
    for k, tol in (("logits", 2e-2), ("action", 2e-2), ("log_prob", 0.25)):   # vs the bf16-operand restatement
        assert float(np.abs(out[k].cpu().numpy() - G["bf16ops_" + k]).max()) < tol, k
    for k, tol in (("logits", 0.12), ("action", 0.12), ("z_mean", 0.06), ("z_logvar", 0.06)):  # vs plain fp32
        assert float(np.abs(out[k].cpu().numpy() - G["fp32_" + k]).max()) < tol, k
    from test_ppo import _case
    vs, adv = pkg("ppo").compute_gae(*[t.cuda() for t in _case(20, 64, seed=7)], lambda_=0.95, discount=0.9)
    np.testing.assert_allclose(vs.cpu().numpy(), G["gae_vs"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(adv.cpu().numpy(), G["gae_adv"], rtol=2e-5, atol=2e-5)
    st = pkg("normalizer").RunningStatistics(232)
    g = torch.Generator().manual_seed(11)
    for i in range(3):
        st.update((torch.randn(500, 232, generator=g) * (1 + i) + 0.5 * i).float().cuda())
    torch.cuda.synchronize()
    np.testing.assert_allclose(st.mean.cpu().numpy(), G["norm_mean"], atol=2e-5)
    np.testing.assert_allclose(st.std.cpu().numpy(), G["norm_std"], rtol=2e-4)
