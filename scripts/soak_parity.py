#!/usr/bin/env python
"""One-off soak: the CUDA step against the C oracle, free-running, at a batch and a length the test
suite does not afford -- every tensor compared bit for bit every `--every` steps, flags / rewards /
counters on every step.  Prints one JSON line.  (Test infrastructure: uses oracle/.)
    python scripts/soak_parity.py --envs 262144 --steps 1000 [--agents 8 --obstacles 16]"""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import marlnav_b200 as mb
from oracle import oracle as orc
from helpers import action_pool, cpu_params

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=262144); ap.add_argument("--steps", type=int, default=1000)
ap.add_argument("--agents", type=int, default=3); ap.add_argument("--obstacles", type=int, default=3)
ap.add_argument("--every", type=int, default=50); ap.add_argument("--noisy", action="store_true")
a = ap.parse_args()
B, A, O = a.envs, a.agents, a.obstacles
p = mb.default_env_params(B, A, O, sampling_style='policy') if A == 3 else mb.template_env_params(B, A, O)
if a.noisy:
    p['init']['noisy_ags'] = True
env = mb.Env(dict(p, seed=17))
oe = orc.OracleEnv(cpu_params(p), seed=17)
pool = action_pool(B, A, n=16)
dev_pool = [x.cuda() for x in pool]
eq = lambda x, y: np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8))
t0 = time.time(); full = 0; resets = 0
for t in range(a.steps):
    obs, rew, term, trunc = env.step_fused(dev_pool[t % 16])
    o_obs, o_rew, o_term, o_trunc = oe.step_fused(pool[t % 16].numpy())
    assert eq(rew.cpu().numpy(), o_rew) and eq(term.cpu().numpy(), o_term) and eq(trunc.cpu().numpy(), o_trunc), f"step {t}"
    resets += int((o_term | o_trunc).sum())
    if t % a.every == 0 or t == a.steps - 1:
        assert eq(obs.cpu().numpy(), o_obs), f"obs step {t}"
        assert eq(env.states.cpu().numpy(), oe.states) and eq(env.obstacles.cpu().numpy(), oe.obstacles), f"state step {t}"
        assert eq(env._step_num.cpu().numpy(), oe.step_num), f"step_num step {t}"
        full += 1
stats = (env._num_trunc, env._num_col, env._num_tar)
assert stats == tuple(int(v) for v in oe.stats)
print(json.dumps({"soak": "cuda == oracle, bit for bit", "envs": B, "agents": A, "obstacles": O, "steps": a.steps,
                  "noisy_ags": bool(a.noisy), "rewards_flags_checked_every_step": True, "full_tensor_checks": full,
                  "resets": resets, "episode_stats": stats, "kernel": "mn::step_env_kernel" if (A == 3 and B > 32768) else "mn::step_team_kernel" if (A == 3 or (A, O) == (8, 16)) else "mn::step_kernel",
                  "seconds": round(time.time() - t0, 1), "device": torch.cuda.get_device_name(0)}))
