#!/usr/bin/env python
"""scripts/dropin_demo.py -- ONE-OFF evidence run (not a test, not the bench): the UNMODIFIED
reference's own callers driving marlnav_b200.Env on a B200.

Needs a copy of the reference package under the git-ignored baseline/_ref/ (made by hand in the
build container with `cp -r /root/reference/marlnav baseline/_ref/`; it is never committed and
nothing in tests/, smoke() or bench.py reads it).  Runs
  1. `check_rews` (marlnav/utils.py:579-666), i.e. `python -m marlnav -rc -sn {-1,0,1}`, 1000 steps;
  2. `MAPPO.get_data` + `train_actor` + `train_critic` (marlnav/models.py:106-198), one rollout of
     buffer_len steps at -np 1024 -- the training loop of marlnav/__main__.py:21-27;
with `marlnav.__main__.Env` swapped for marlnav_b200.Env, and the same with the reference Env on the
same GPU for a wall-clock comparison.  Prints one JSON line per run."""
import argparse
import contextlib
import io
import json
import os
import sys
import time
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
if not os.path.isdir(os.path.join(REF, "marlnav")):
    print(json.dumps({"dropin_demo": "skipped", "why": "baseline/_ref/marlnav not present"}))
    sys.exit(0)
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
    sys.modules.setdefault(name, mock.MagicMock())
_axs = mock.MagicMock(); _axs.flat = []
sys.modules["matplotlib.pyplot"].subplots.return_value = (mock.MagicMock(), _axs)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

import torch                                   # noqa: E402
import marlnav.utils as U                      # noqa: E402
import marlnav.environment as E                # noqa: E402
import marlnav.models as M                     # noqa: E402
import marlnav_b200 as mb                      # noqa: E402


def ref_args(**over):
    d = dict(seed=0, max_x_value=1500.0, max_y_value=750.0, fig_size_x=10.0, fig_size_y=5.0, parallel_index=0,
             agent_index=0, interval=10, random=False, weights_file=None, num_parallel=2, num_agents=3,
             num_obstacles=3, max_step=1000, episode_len=200, min_speed=3., max_speed=10., min_accel=-0.5,
             max_accel=0.5, risk_factor=0., distance_factor=0., heading_factor=500., target_factor=500.,
             soft_factor=500., bond_factor=10., hidden_size=50, learning_rate=0.001, ent_const=0.001,
             epsilon=0.01, gamma=0.9, num_total=1024000, buffer_len=1000, num_epochs=2, batch_size=1000,
             rendering=False, sampling_style='sampler', reward_check=False, sampler_num=-1)
    d.update(over)
    return argparse.Namespace(**d)


def run_check_rews(env_cls, sn):
    args = ref_args(reward_check=True, sampler_num=sn, num_obstacles=3 if sn == -1 else 1)
    U.set_all_seeds(0)
    params = U.set_params(args)
    env = env_cls(params['env'])
    rewards = []
    orig_step = env.step

    def tap(actions):
        out = orig_step(actions)
        rewards.append(out[1].detach().float().cpu().clone())
        return out
    env.step = tap
    t0 = time.perf_counter()
    U.check_rews(env, params['animation']['max_step'], 0, 0)        # the reference's own harness
    dt = time.perf_counter() - t0
    r = torch.stack(rewards)
    return dict(mode=f"check_rews -sn {sn}", env=env_cls.__module__, steps=len(rewards), seconds=round(dt, 3),
                reward_sum=float(r.double().sum()), first_rewards=[round(float(x), 4) for x in r[:3, 0]],
                stats=[int(env._num_trunc), int(env._num_col), int(env._num_tar)])


def run_training_rollout(env_cls, buffer_len, fast=False):
    """fast=True: the reference's MAPPO object with marlnav_b200.MappoRollout attached -- ONE added line
    (`mb.MappoRollout(mappo).attach()`); get_data / train_actor / train_critic are then called as
    marlnav/__main__.py:21-27 does."""
    args = ref_args(num_parallel=1024, sampling_style='policy', buffer_len=buffer_len, batch_size=buffer_len,
                    num_total=1024 * buffer_len)
    U.set_all_seeds(0)
    params = U.set_params(args)
    env = env_cls(params['env'])
    cwd = os.getcwd()
    os.makedirs("/tmp/dropin", exist_ok=True); os.chdir("/tmp/dropin")
    try:
        mappo = M.MAPPO(params['model'], env)
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink):        # the reference prints every step
            if fast:
                mb.MappoRollout(mappo, seed=0).attach()
                mappo.get_data(); torch.cuda.synchronize()      # builds the CUDA graph (one-off)
            t0 = time.perf_counter(); mappo.get_data(); torch.cuda.synchronize(); t_roll = time.perf_counter() - t0
            t0 = time.perf_counter(); mappo.train_actor(); mappo.train_critic(); torch.cuda.synchronize()
            t_train = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    logs = mappo._logs
    return dict(mode=f"MAPPO rollout -np 1024 -bl {buffer_len}" + (" + MappoRollout.attach()" if fast else ""), env=env_cls.__module__,
                rollout_seconds=round(t_roll, 3), env_steps_per_sec=round(1024 * buffer_len / t_roll),
                train_seconds=round(t_train, 3), mean_rew=logs['mean_rews'][-1],
                epi_stats={k: v[-1] for k, v in logs['epi_stats'].items()},
                actor_loss_finite=bool(all(x == x for x in logs['actor'])))


if __name__ == "__main__":
    assert torch.cuda.is_available()
    for sn in (-1, 0, 1):
        print(json.dumps(run_check_rews(mb.Env, sn)), flush=True)
    print(json.dumps(run_check_rews(E.Env, 0)), flush=True)            # reference Env on the same GPU
    print(json.dumps(run_training_rollout(mb.Env, 1000)), flush=True)
    print(json.dumps(run_training_rollout(mb.Env, 1000, fast=True)), flush=True)
    print(json.dumps(run_training_rollout(E.Env, 100)), flush=True)    # reference Env: 10x shorter rollout
