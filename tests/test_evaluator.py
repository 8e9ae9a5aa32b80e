"""Evaluator path of SURVEY 8 row f3 (VERDICT r1 missing #4): brax `EvalWrapper` episode metrics + `Evaluator.run_evaluation`
(ppo_imitation/acting.py:83-156) over the fused rollout.  Checker: a line-by-line numpy restatement of EvalWrapper.step folded over
the recorded transitions."""
import numpy as np
import pytest

from conftest import pkg

pytestmark = pytest.mark.gpu


def _eval_wrapper_numpy(metrics, reward, done, steps_per_t):
    """brax EvalWrapper.step (envs/wrappers/training.py), one env at a time: metrics [T, B, 7], reward / done [T, B]."""
    T, B, nm = metrics.shape
    ep = np.zeros((B, nm + 1)); active = np.ones(B); ep_steps = np.zeros(B)
    for t in range(T):
        m = np.concatenate([metrics[t], reward[t][:, None]], 1)   # nstate.metrics['reward'] = nstate.reward
        ep_steps = np.where(active > 0, steps_per_t[t], ep_steps)
        ep = ep + m * active[:, None]
        active = active * (1 - done[t])
    return ep, active, ep_steps


def test_evaluator_matches_eval_wrapper_restatement(gpu_env, rodent):
    import torch
    pol, ev = pkg("policy"), pkg("evaluator")
    eng = gpu_env.engine
    B, T = 64, 30
    params = pol.init_params(np.random.default_rng(7), pol.param_shapes(eng.traj_size, eng.obs_size, gpu_env.action_size))
    policy = pol.PrecisePolicy(params, "cuda:0")
    e = ev.Evaluator(gpu_env, policy, num_eval_envs=B, episode_length=T, seed=7)
    for epoch in range(2):  # the second evaluation replays the captured graph from a fresh reset
        out = e.run_evaluation({"training/sps": 1.0})
        ro = e.rollout
        met, rew, done = ro.metrics.cpu().numpy().astype(np.float64), ro.reward.cpu().numpy().astype(np.float64), ro.done.cpu().numpy().astype(np.float64)
        # info["steps"] of EpisodeWrapper: counts up until the env's first done (AutoReset zeroes it after)
        steps = np.zeros((T, B)); cur = np.zeros(B); prev_done = np.zeros(B)
        for t in range(T):
            cur = np.where(prev_done > 0, 0, cur) + 1
            steps[t] = cur; prev_done = done[t]
        ep, active, ep_steps = _eval_wrapper_numpy(met, rew, done, steps)
        assert np.abs(e.episode_metrics.cpu().numpy() - ep).max() < 1e-5
        assert np.array_equal(e.active.cpu().numpy(), active) and np.array_equal(e.episode_steps.cpu().numpy(), ep_steps)
        assert (ep_steps <= 10).all() and ep_steps.min() >= 1  # sub-clips of 10 frames end every episode (rodent.py:207-215)
        names = list(pkg("envs.rodent").METRIC_KEYS) + ["reward"]
        for i, n in enumerate(names):
            assert abs(out[f"eval/episode_{n}"] - ep[:, i].mean()) < 1e-5 and abs(out[f"eval/episode_{n}_std"] - ep[:, i].std()) < 1e-5
        assert out["eval/avg_episode_length"] == ep_steps.mean() and out["training/sps"] == 1.0
        assert out["eval/sps"] > 0 and out["eval/walltime"] >= out["eval/epoch_eval_time"]
    raw = e.run_evaluation({}, aggregate_episodes=False)
    assert raw["eval/episode_reward"].shape == (B,)
    # deterministic_eval: the policy's mode
    d = ev.Evaluator(gpu_env, policy, num_eval_envs=B, episode_length=T, seed=7, deterministic=True)
    d.run_evaluation({})
    nu = gpu_env.action_size
    assert torch.allclose(d.rollout.action, torch.tanh(d.rollout.logits[..., :nu]), atol=1e-6)
