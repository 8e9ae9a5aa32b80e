"""`training_step` of the reference (ppo_imitation/train.py:296-349) on the GPU: rollout -> normaliser update -> SGD phase.

    (state, _), data = scan(generate_unroll, ...)                      rollout.Rollout (policy kernel + fused env step, CUDA graph)
    normalizer_params = running_statistics.update(..., pmap_axis_name)  normalizer.RunningStatistics (one pass + one all-reduce)
    for _ in range(num_updates_per_batch):                              sgd_step (train.py:270-294)
        permutation of the env axis, reshape to num_minibatches
        for each minibatch: gradient_update_fn(loss, adam, pmean)       learner.PPOLearner.update (tcgen05 TF32 GEMMs + row kernels)
    policy <- new parameters                                            policy.load_params (same device blob: captured graphs keep working)

The transition buffers stay on the device from the env kernel to the loss: a minibatch is an index list over the env axis
(`vnl_gather_rows`, which also pads traj rows 795 -> 796 floats for TMA).  One minibatch update (gathers + ~80 launches + the two
gradient exchange + Adam) is captured once into CUDA graphs (collective-free: with a process group the NCCL all-reduce runs eagerly
between the gradient graph and the Adam graph) and replayed with fresh indices / noise.  torch = memory, streams,
RNG for the noise operands and `torch.distributed`; no torch math on the data path, no CPU fallback.
"""
from __future__ import annotations

from typing import Dict

from . import learner as lrn
from . import train_kernels as tk


class Trainer:
    def __init__(self, env, policy, learner: "lrn.PPOLearner", rollout, stats, num_minibatches: int, num_updates_per_batch: int,
                 use_graph: bool = True, seed: int = 0):
        import torch

        self.torch = t = torch
        self.env, self.policy, self.learner, self.rollout, self.stats = env, policy, learner, rollout, stats
        self.B, self.T = rollout.B, rollout.T
        if self.B % num_minibatches:
            raise ValueError("num_envs must be divisible by num_minibatches")
        self.Bm = self.B // num_minibatches
        if (learner.T, learner.Bm) != (self.T, self.Bm):
            raise ValueError("learner was built for another minibatch shape")
        self.num_minibatches, self.num_updates = int(num_minibatches), int(num_updates_per_batch)
        dev = learner.device
        L = learner
        f = lambda *s: t.zeros(*s, dtype=t.float32, device=dev)
        R = self.T * self.Bm
        # static minibatch operands (graph-captured addresses)
        self.idx = t.zeros(self.Bm, dtype=t.int32, device=dev)
        self.eps = f(R * (L.L + L.nu))  # both noise operands of a minibatch in one buffer: one generator launch per minibatch
        self.mb = dict(traj=f(R, L.ld_traj), observation=f(R, L.obs), next_observation_last=f(self.Bm, L.obs), reward=f(R), discount=f(R),
                       truncation=f(R), log_prob=f(R), raw_action=f(R, L.nu), eps_z=self.eps[:R * L.L].view(R, L.L),
                       eps_ent=self.eps[R * L.L:].view(R, L.nu))
        import ctypes
        self._scalar_keys = ("reward", "discount", "truncation", "log_prob")
        self._sdst = (ctypes.c_void_p * 4)(*[self.mb[k].data_ptr() for k in self._scalar_keys])
        self.gen = t.Generator(device=dev).manual_seed(seed)
        self.use_graph, self.graph = bool(use_graph), None
        self.discount_buf = f(self.T, self.B)
        self.env_steps = 0
        self.sgd_launches = 0

    # ---- one minibatch ----------------------------------------------------------------------------------------------------
    def _gather(self, tr):
        """`convert_data` (train.py:279-283) for one minibatch: rows of the time-major transition buffers picked by self.idx."""
        L_, st = tk.lib(), tk.stream(self.idx)
        T, B, Bm, mb = self.T, self.B, self.Bm, self.mb
        g = lambda src, width, dst, ld, T_=T: tk.check(L_.vnl_gather_rows(src.data_ptr(), T_, B, width, self.idx.data_ptr(), Bm, dst.data_ptr(), ld, st), "vnl_gather_rows")
        lr = self.learner
        g(tr["state_extras_traj_in"], lr.traj, mb["traj"], lr.ld_traj)
        g(tr["observation"], lr.obs, mb["observation"], lr.obs)
        g(tr["next_observation"][T - 1:T], lr.obs, mb["next_observation_last"], lr.obs, 1)
        g(tr["policy_extras"]["raw_action"], lr.nu, mb["raw_action"], lr.nu)
        import ctypes
        srcs = (tr["reward"], self.discount_buf, tr["state_extras"]["truncation"], tr["policy_extras"]["log_prob"])  # order of _scalar_keys
        for s_ in srcs:
            assert s_.is_contiguous() and s_.shape == (T, B) and s_.dtype == self.torch.float32
        tk.check(L_.vnl_gather_scalars(4, (ctypes.c_void_p * 4)(*[s_.data_ptr() for s_ in srcs]), self._sdst, T, B, self.idx.data_ptr(), Bm, st),
                 "vnl_gather_scalars")
        self.sgd_launches += 5

    def _minibatch(self, tr):
        n0 = self.learner.launches
        self._gather(tr)
        self.learner.update(self.mb)
        self.sgd_launches += self.learner.launches - n0

    def _capture(self, fn):
        """Capture `fn` into a CUDA graph.  The graphs of the SGD phase never contain a collective: with a process group the gradient
        all-reduce runs eagerly BETWEEN two graphs (loss + gradients | Adam), so NCCL's watchdog thread and the capture never meet
        (`capture_error_mode="thread_local"`: calls from other threads -- the watchdog's event queries -- do not invalidate it)."""
        t = self.torch
        t.cuda.synchronize(self.idx.device)
        g = t.cuda.CUDAGraph()
        with t.cuda.graph(g, capture_error_mode="thread_local"):
            fn()
        return g

    def sgd_phase(self, tr) -> None:
        """num_updates_per_batch x num_minibatches gradient updates on the unroll `tr` (train.py:336-341)."""
        t = self.torch
        lr = self.learner
        dist = lr._dist()
        self.discount_buf.copy_(tr["discount"])  # `1 - done` view materialised once per unroll
        # the policy's INPUT trajectory of step t is traj[t] (acting.py:47: state.info["traj"]), i.e. rollout.traj[:T]
        tr = dict(tr, state_extras_traj_in=self.rollout.traj[:self.T])
        for _ in range(self.num_updates):
            perm = t.randperm(self.B, device=self.idx.device, generator=self.gen).to(t.int32)
            for mbi in range(self.num_minibatches):
                self.idx.copy_(perm[mbi * self.Bm:(mbi + 1) * self.Bm])
                self.eps.normal_(generator=self.gen)
                if not self.use_graph:
                    self._minibatch(tr)
                    continue
                if self.graph is None:
                    side = t.cuda.Stream(device=self.idx.device)  # warm-up outside capture (lazy module loads, attribute sets)
                    side.wait_stream(t.cuda.current_stream(self.idx.device))
                    l0 = self.sgd_launches
                    with t.cuda.stream(side):
                        self._minibatch(tr)
                    t.cuda.current_stream(self.idx.device).wait_stream(side)
                    self._launches_per_update = self.sgd_launches - l0  # kernels one replay of the captured update launches
                    n0, l1 = lr.updates, self.sgd_launches
                    if dist is None:
                        self.graph = (self._capture(lambda: self._minibatch(tr)),)
                    else:
                        def grads_only():
                            self._gather(tr)
                            lr.loss_and_grads(self.mb, exchange=False)
                        self.graph = (self._capture(grads_only), self._capture(lambda: lr.apply_gradients(exchange=False, scale=1.0 / dist.get_world_size())))
                    lr.updates, self.sgd_launches = n0, l1  # the capture passes went through the host-side counters without executing
                    continue  # this minibatch was the warm-up run
                self.graph[0].replay()
                if dist is not None:
                    dist.all_reduce(lr.grads, op=dist.ReduceOp.SUM)  # lax.pmean of the gradients: one NCCL call, 6.5 MB; 1 / world in Adam
                    self.graph[1].replay()
                lr.updates += 1  # host-side counter (the device-side step counter advanced inside the graph)
                self.sgd_launches += self._launches_per_update

    def training_step(self) -> Dict[str, float]:
        """One `training_step`: unroll, normaliser update, SGD phase, policy refresh.  Returns the last minibatch's loss metrics."""
        ro = self.rollout
        ro.eps_z.normal_(generator=self.gen)
        ro.eps_a.normal_(generator=self.gen)
        tr = ro.generate_unroll()
        self.stats.update(tr["observation"])  # running_statistics.update with its psum (train.py:330-334)
        self.learner.set_normalizer(self.stats.mean, self.stats.std)
        self.sgd_phase(tr)
        self.policy.load_params(self.learner.policy_params())
        self.env_steps += self.B * self.T
        return self.learner.metrics_dict()
