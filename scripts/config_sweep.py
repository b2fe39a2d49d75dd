#!/usr/bin/env python
"""scripts/config_sweep.py -- the other BASELINE.json configs, one JSON object
per line: device-resident env-steps/s for each shape, plus config #2 (1024 envs with a MAPPO-style
actor in the loop, models.py:106-122) eagerly and with the whole {actor -> step} iteration captured
in a CUDA graph.  Not the bench contract (bench.py is); evidence for profiles/."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marlnav_b200 as mb          # noqa: E402
from bench import algorithmic_bytes, env_params, make_action_pool, measured_peak   # noqa: E402


def time_steps(env, pool, steps, warmup=20):
    out = env._alloc_outputs()
    for i in range(warmup):
        env.step_fused(pool[i % len(pool)], out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        env.step_fused(pool[i % len(pool)], out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def shape_line(B, A, O, steps):
    env = mb.Env(env_params(B, A, O, 'cuda:0'))
    pool = make_action_pool(B, A, 16, 'cuda:0')
    ms = time_steps(env, pool, steps)
    peak, _ = measured_peak()
    gbs = B * algorithmic_bytes(A, O) / (ms * 1e-3) / 1e9
    ws = B * algorithmic_bytes(A, O)
    return {"config": f"{B}x{A}x{O}", "ms_per_step": ms, "env_steps_per_sec": B / (ms * 1e-3),
            "agent_steps_per_sec": B * A / (ms * 1e-3), "algorithmic_GBps": gbs,
            "frac_of_hbm_peak": gbs / peak if ws > 126e6 else None,
            "note": None if ws > 126e6 else "working set fits in the 126 MB L2: not an HBM fraction",
            "launch": dict(zip(("grid", "block", "smem", "envs_per_cta"), env.launch_info()))}


class Actor(torch.nn.Module):
    """Same shape as the reference's Actor (models.py:14-36): 12 -> 50 -> (2, 2)."""
    def __init__(self, obs_size, hidden=50):
        super().__init__()
        self.fc1 = torch.nn.Linear(obs_size, hidden)
        self.mu = torch.nn.Linear(hidden, 2)
        self.std = torch.nn.Linear(hidden, 2)

    def forward(self, x):
        h = self.fc1(x)
        return torch.tanh(self.mu(h)), torch.nn.functional.softplus(self.std(h))


def rollout_line(B=1024, A=3, O=3, steps=1000):
    import math
    env = mb.Env(env_params(B, A, O, 'cuda:0'))
    max_d = math.sqrt(1500.0 ** 2 + 750.0 ** 2)
    lo = [-math.pi, 0.] + O * [-math.pi] + O * [0.] + (A - 1) * [-math.pi] + (A - 1) * [0.]
    hi = [math.pi, max_d] + O * [math.pi] + O * [max_d] + (A - 1) * [math.pi] + (A - 1) * [max_d]
    env.fuse_io(dict(min_obs=lo, max_obs=hi), dict(min_action=[-math.pi, -0.5], max_action=[math.pi, 0.5]))
    actor = Actor(env.obs_size).cuda()
    out = env._alloc_outputs()
    obs = out[0]
    env.step_fused(torch.zeros(B, A, 2, device='cuda'), out=out)

    @torch.no_grad()
    def iteration():
        mu, std = actor(obs.view(B * A, -1))
        act = mu + std * torch.randn_like(mu)             # diagonal Gaussian sample
        env.step_fused(act.view(B, A, 2), out=out)        # raw [-1,1] actions, normalised obs back

    res = {}
    for i in range(50):
        iteration()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        iteration()
    torch.cuda.synchronize()
    res["eager_env_steps_per_sec"] = B * steps / (time.perf_counter() - t0)

    # the same iteration captured once in a CUDA graph (reset counter advanced per replay is not
    # possible inside a graph, so the graph holds `unroll` steps with consecutive counters)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    unroll = 50
    with torch.cuda.stream(s):
        iteration()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(unroll):
                iteration()
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = max(1, steps // unroll)
    for _ in range(reps):
        g.replay()
    torch.cuda.synchronize()
    res["cuda_graph_env_steps_per_sec"] = B * reps * unroll / (time.perf_counter() - t0)
    res["note"] = ("graph = 50 {actor -> fused step} iterations; replays reuse the captured Philox step "
                   "counters, fine for throughput, not for training")
    # fused actor kernel + device-resident rollout buffers (marlnav_b200.rollout)
    fa = mb.FusedActor(actor.state_dict().__class__(
        {'fc1.weight': actor.fc1.weight, 'fc1.bias': actor.fc1.bias, 'fc_mu.weight': actor.mu.weight,
         'fc_mu.bias': actor.mu.bias, 'fc_std.weight': actor.std.weight, 'fc_std.bias': actor.std.bias}), seed=1)
    critic = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(A * env.obs_size, 50), torch.nn.ReLU(),
                                 torch.nn.Linear(50, 1)).cuda()
    fcrit = mb.FusedCritic({'fc1.weight': critic[1].weight, 'fc1.bias': critic[1].bias,
                            'fc2.weight': critic[3].weight, 'fc2.bias': critic[3].bias})
    for name, cr in (("fused_actor_rollout_env_steps_per_sec", None), ("fused_actor_plus_torch_critic_env_steps_per_sec", critic),
                     ("fused_actor_plus_fused_critic_env_steps_per_sec", fcrit)):
        mb.collect_rollout(env, fa, 50, critic=cr)
        best = 0.0
        for _ in range(3):                                     # host-bound leg: best of three
            torch.cuda.synchronize(); t0 = time.perf_counter()
            buf = mb.collect_rollout(env, fa, steps, critic=cr)
            torch.cuda.synchronize(); best = max(best, B * steps / (time.perf_counter() - t0))
        res[name] = best
    # the whole T-step rollout as ONE CUDA-graph launch (device-resident counters: every replay
    # continues the eager random streams, tests/test_gpu_rollout.py)
    rg = mb.RolloutGraph(env, fa, steps, critic=fcrit)
    rg.replay(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        rg.replay()
    torch.cuda.synchronize(); res["rollout_graph_actor_critic_env_steps_per_sec"] = 5 * B * steps / (time.perf_counter() - t0)
    best_ours, best_ref = float('inf'), float('inf')
    for rep in range(3):                                   # host-timed: best of three
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ret = mb.discounted_returns(buf['rewards'], buf['done'], 0.9, normalize=True)
        torch.cuda.synchronize(); best_ours = min(best_ours, 1e3 * (time.perf_counter() - t0))
        t0 = time.perf_counter()
        curr = torch.zeros(B, dtype=float, device='cuda')
        for i in range(steps - 1, -1, -1):                 # the reference's loop, models.py:135-139
            curr = torch.where(buf['done'][i], 0., buf['rewards'][i] + 0.9 * curr)
        torch.cuda.synchronize(); best_ref = min(best_ref, 1e3 * (time.perf_counter() - t0))
    res["discounted_returns_ms_T1000"], res["reference_return_loop_ms_T1000"] = best_ours, best_ref
    return {"config": f"rollout {B}x{A}x{O} with actor in the loop (BASELINE configs[1])", **res}


if __name__ == "__main__":
    lines = [shape_line(2, 3, 3, 2000), shape_line(1024, 3, 3, 2000), shape_line(65536, 3, 3, 1000),
             shape_line(1048576, 3, 3, 300), shape_line(262144, 8, 16, 100), shape_line(65536, 4, 2, 300),
             rollout_line()]
    for l in lines:
        print(json.dumps(l), flush=True)
