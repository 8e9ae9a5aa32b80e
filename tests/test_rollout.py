"""rollout.Rollout == `generate_unroll` / `actor_step` of ppo_imitation/acting.py:30-80 around the wrapped env: field
alignment of the transition, equality with a step-by-step loop over the two C entry points, and CUDA-graph replay ==
eager launches (both kernels are capturable: no allocation, no sync, static geometry)."""
import numpy as np
import pytest

from conftest import pkg, start_states

pytestmark = pytest.mark.gpu

T, B, EP = 6, 200, 4.0  # episode_length 4 < unroll: truncation + restore happen inside an unroll


def _setup(gpu_env, rodent, seed):
    import torch
    pol = pkg("policy")
    eng = gpu_env.engine
    qpos, qvel, start = start_states(rodent, B, seed=seed)
    s0 = gpu_env.reset_from(qpos, qvel, start)
    params = pol.init_params(np.random.default_rng(seed), pol.param_shapes(eng.traj_size, eng.obs_size, gpu_env.action_size), perturb=0.1)
    g = torch.Generator(device="cuda").manual_seed(seed)
    mean = torch.randn(eng.obs_size, device="cuda", generator=g) * 0.1
    std = torch.rand(eng.obs_size, device="cuda", generator=g) + 0.5
    policy = pol.IntentionPolicy(params, "cuda:0", mean, std)
    draws = [(torch.randn(T, B, 64, device="cuda", generator=g), torch.randn(T, B, 30, device="cuda", generator=g)) for _ in range(3)]
    return s0, policy, draws


def test_unroll_equals_stepwise_loop(gpu_env, rodent):
    import torch
    s0, policy, draws = _setup(gpu_env, rodent, 51)
    eng = gpu_env.engine
    ro = pkg("rollout").Rollout(gpu_env, policy, s0, T, EP, use_graph=False)
    # the reference loop, one call at a time with fresh buffers
    first, first_obs = dict(s0.pipeline_state), s0.obs
    st = dict({k: v.clone() for k, v in first.items()}, cur_frame=s0.info["cur_frame"].clone(), sub_clip_frame=s0.info["sub_clip_frame"].clone())
    obs, traj = s0.obs.clone(), s0.info["traj"].clone()
    steps, done = torch.zeros(B, device="cuda"), torch.zeros(B, device="cuda")
    seen_trunc = False
    for u in range(2):
        tr = ro.generate_unroll(*draws[u])
        torch.cuda.synchronize()
        assert torch.equal(tr["next_observation"][:-1], tr["observation"][1:])
        for t in range(T):
            act, ex = policy(traj, obs, draws[u][0][t], draws[u][1][t])
            nxt, out, trunc = eng.alloc_state(B), eng.alloc_outputs(B), torch.zeros(B, device="cuda")
            eng.step_training(st, act, nxt, out, first, first_obs, steps, done, steps, trunc, EP)
            torch.cuda.synchronize()
            assert torch.equal(tr["observation"][t], obs) and torch.equal(tr["action"][t], act), (u, t)
            assert torch.equal(tr["policy_extras"]["log_prob"][t], ex["log_prob"])
            assert torch.equal(tr["policy_extras"]["raw_action"][t], ex["raw_action"])
            assert torch.equal(tr["policy_extras"]["logits"][t], ex["logits"])
            assert torch.equal(tr["reward"][t], out["reward"]) and torch.equal(tr["discount"][t], 1 - out["done"])
            assert torch.equal(tr["next_observation"][t], out["obs"]) and torch.equal(tr["state_extras"]["traj"][t], out["traj"])
            assert torch.equal(tr["state_extras"]["truncation"][t], trunc)
            seen_trunc = seen_trunc or bool((trunc > 0).any())
            st, obs, traj, done = nxt, out["obs"], out["traj"], out["done"].clone()
        for k, v in st.items():
            assert torch.equal(ro.env_state[k], v), k
    assert seen_trunc and float(steps.max()) <= EP and ro.launches == 4 * T


def test_graph_replay_equals_eager(gpu_env, rodent):
    import torch
    s0, policy, draws = _setup(gpu_env, rodent, 52)
    eager = pkg("rollout").Rollout(gpu_env, policy, s0, T, EP, use_graph=False)
    graph = pkg("rollout").Rollout(gpu_env, policy, s0, T, EP, use_graph=True)
    for u in range(3):
        a = eager.generate_unroll(*draws[u])
        b = graph.generate_unroll(*draws[u])
        torch.cuda.synchronize()
        for k in ("observation", "next_observation", "action", "reward", "discount", "metrics"):
            assert torch.equal(a[k], b[k]), (u, k)
        for grp in ("policy_extras", "state_extras"):
            for k in a[grp]:
                assert torch.equal(a[grp][k], b[grp][k]), (u, grp, k)
        for k, v in eager.env_state.items():
            assert torch.equal(graph.env_state[k], v), (u, k)
    assert graph.graph is not None


def test_graph_replay_sees_reloaded_params_and_normaliser(gpu_env, rodent):
    """after a PPO update `load_params` / `set_normalizer` must reach a rollout graph that was captured
    before them — the device blob and the mean / std operands keep their addresses and are overwritten in place."""
    import torch
    pol = pkg("policy")
    s0, policy, draws = _setup(gpu_env, rodent, 53)
    eng = gpu_env.engine
    graph = pkg("rollout").Rollout(gpu_env, policy, s0, T, EP, use_graph=True)
    graph.generate_unroll(*draws[0])  # capture + first replay with the old weights
    torch.cuda.synchronize()
    ptr = policy.blob_dev.data_ptr(), policy.obs_mean.data_ptr(), policy.obs_std.data_ptr()
    new = pol.init_params(np.random.default_rng(530), pol.param_shapes(eng.traj_size, eng.obs_size, gpu_env.action_size), perturb=0.2)
    policy.load_params(new)
    policy.set_normalizer(policy.obs_mean * 0.5 + 0.01, policy.obs_std * 1.5)
    assert ptr == (policy.blob_dev.data_ptr(), policy.obs_mean.data_ptr(), policy.obs_std.data_ptr())
    b = {k: v.clone() for k, v in graph.generate_unroll(*draws[1]).items() if k in ("action", "reward")}
    # an eager twin that saw the same first unroll with the OLD weights and the second with the NEW ones
    policy2 = pol.IntentionPolicy(new, "cuda:0", policy.obs_mean.clone(), policy.obs_std.clone())
    old = _setup(gpu_env, rodent, 53)[1]
    eager = pkg("rollout").Rollout(gpu_env, old, s0, T, EP, use_graph=False)
    eager.generate_unroll(*draws[0])
    eager.policy = policy2
    a = eager.generate_unroll(*draws[1])
    torch.cuda.synchronize()
    assert torch.equal(a["action"], b["action"]) and torch.equal(a["reward"], b["reward"])
    with pytest.raises(ValueError):
        policy.set_normalizer(None, None)  # captured operands cannot silently become the identity


def test_odd_unroll_length_carries_state(gpu_env, rodent):
    import torch
    s0, policy, draws = _setup(gpu_env, rodent, 53)
    a = pkg("rollout").Rollout(gpu_env, policy, s0, 3, EP, use_graph=False)
    b = pkg("rollout").Rollout(gpu_env, policy, s0, 6, EP, use_graph=False)
    z, e = draws[0]
    a.generate_unroll(z[:3], e[:3])
    ta = {k: v.clone() for k, v in a.generate_unroll(z[3:], e[3:]).items() if k in ("reward", "action")}
    tb = b.generate_unroll(z, e)
    torch.cuda.synchronize()
    assert torch.equal(ta["reward"], tb["reward"][3:]) and torch.equal(ta["action"], tb["action"][3:])


def test_long_rollout_with_normaliser_stays_finite(gpu_env, rodent):
    """200 env steps of 512 envs through the packaged loop (graph replay), the obs normaliser updated every unroll and feeding
    the policy: everything stays finite, actions stay in (-1, 1), the normaliser counts every observation, episodes end."""
    import torch
    pol, nzm = pkg("policy"), pkg("normalizer")
    eng = gpu_env.engine
    Bn, Tn, U = 512, 20, 10
    qpos, qvel, start = start_states(rodent, Bn, seed=61)
    s0 = gpu_env.reset_from(qpos, qvel, start)
    params = pol.init_params(np.random.default_rng(61), pol.param_shapes(eng.traj_size, eng.obs_size, gpu_env.action_size))
    stats = nzm.RunningStatistics(eng.obs_size)
    policy = pol.IntentionPolicy(params, "cuda:0", stats.mean, stats.std)
    assert policy.obs_mean.data_ptr() == stats.mean.data_ptr()  # shared: updates reach the next launch
    ro = pkg("rollout").Rollout(gpu_env, policy, s0, Tn, 150.0, use_graph=True)
    g = torch.Generator(device="cuda").manual_seed(61)
    ended = 0.0
    for u in range(U):
        ro.eps_z.normal_(generator=g)
        ro.eps_a.normal_(generator=g)
        tr = ro.generate_unroll()
        stats.update(tr["observation"])
        torch.cuda.synchronize()
        for k in ("observation", "action", "reward"):
            assert torch.isfinite(tr[k]).all(), (u, k)
        assert torch.isfinite(tr["policy_extras"]["log_prob"]).all() and float(tr["action"].abs().max()) <= 1.0
        ended += float((1 - tr["discount"]).sum())
    assert float(stats.count) == U * Tn * Bn and torch.isfinite(stats.std).all() and float(stats.std.min()) > 0
    assert ended > 0  # sub-clips of 10 frames end inside 200 steps (Q7: and stay ended)
