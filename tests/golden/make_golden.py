#!/usr/bin/env python
"""tests/golden/make_golden.py -- generate the committed golden vectors by running the REAL
reference (/root/reference/marlnav, unmodified, imported in the build container).

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz

Two families (see DESIGN.md "Oracle and parity"):

  patched_*.npz  The reference with exactly two substitutions injected from outside:
                 (1) env._init_sampler := the addressed-Philox sampler of SURVEY.md
                 Appendix D (same agent template / target, obstacle draws addressed by
                 (seed, env id, call counter) instead of the global mt19937 stream);
                 (2) torch.sin/cos/acos := oracle/marlnav_trig.h (the reference's own bits
                 for these come from MKL VML, which cannot be restated).
                 Every other torch op the reference executes is stock.  FREE-RUNNING
                 traces; the C oracle and the CUDA kernel must reproduce them BIT FOR BIT.
  stock_*.npz    The reference with stock torch trig (only the Philox sampler injected):
                 per-step snapshots for teacher-forced tolerance tests (flags exact,
                 floats 1e-5 where the arithmetic is well-conditioned).

The reference's own `-rc -sn 0/1` scenarios (mock initialiser + scripted sampler) need no
sampler injection at all: they are run through the reference's set_params() untouched.
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refload                                   # noqa: E402  (also puts the repo root on sys.path)
from oracle import oracle as orc                 # noqa: E402

torch.set_num_threads(4)
env_mod, utils_mod = refload.load()

SNAP_EVERY = 20
FACTORS = ('risk_factor', 'distance_factor', 'heading_factor', 'target_factor', 'soft_factor', 'bond_factor')


def actions_for(B, A, steps, seed, angle=0.2, accel=0.5):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        ang = (torch.rand(B, A, generator=g) * 2 - 1) * angle
        acc = (torch.rand(B, A, generator=g) * 2 - 1) * accel
        out.append(torch.stack([ang, acc], dim=2).contiguous())
    return out


def fused(obs):
    return torch.cat(list(obs), dim=2).numpy().copy()


def checksum(a):
    return np.ascontiguousarray(a).view(np.uint32).astype(np.uint64).sum()


def make_env(B, A, O, seed, template=None, **over):
    """Reference Env built through the reference's own set_params, Philox sampler injected."""
    args = refload.reference_args(num_parallel=B, num_agents=A, num_obstacles=O, **over)
    params = copy.deepcopy(utils_mod.set_params(args)['env'])
    env = env_mod.Env(params)
    tmpl = orc.triangle_template(params['init']) if template is None else np.asarray(template, np.float32)
    smp = orc.PhiloxTemplateSampler(params, tmpl, seed=seed)
    env._init_sampler = smp
    env.states, env.obstacles, env.target = smp()        # call 0 = construction-time sample
    return env


def run_trace(env, actions, record_pre=False):
    """Free-running trace of a reference Env.  Returns dict of arrays."""
    T = len(actions)
    B = env.num_parallel
    rec = dict(rewards=np.empty((T, B), np.float32), terminated=np.empty((T, B), bool),
               truncated=np.empty((T, B), bool), obs_sum=np.empty(T, np.uint64),
               states_sum=np.empty(T, np.uint64), snap_steps=[], snap_states=[], snap_obstacles=[],
               snap_obs=[], snap_step_num=[], snap_terminates=[])
    pre = dict(states=[], obstacles=[], target=[], step_num=[], terminates=[]) if record_pre else None
    for t, act in enumerate(actions):
        if record_pre and t % 4 == 0:
            pre['states'].append(env.states.numpy().copy()); pre['obstacles'].append(env.obstacles.numpy().copy())
            pre['target'].append(env.target.numpy().copy()); pre['step_num'].append(env._step_num.numpy().copy())
            pre['terminates'].append(env._terminates.numpy().copy())
        obs, rew, term, trunc = env.step(act.clone())
        o = fused(obs)
        rec['rewards'][t], rec['terminated'][t], rec['truncated'][t] = rew.numpy(), term.numpy(), trunc.numpy()
        rec['obs_sum'][t], rec['states_sum'][t] = checksum(o), checksum(env.states.numpy())
        if t % SNAP_EVERY == 0 or t == T - 1 or (record_pre and t % 4 == 0):
            rec['snap_steps'].append(t); rec['snap_states'].append(env.states.numpy().copy())
            rec['snap_obstacles'].append(env.obstacles.numpy().copy()); rec['snap_obs'].append(o)
            rec['snap_step_num'].append(env._step_num.numpy().copy())
            rec['snap_terminates'].append(env._terminates.numpy().copy())
    out = {k: (np.stack(v) if isinstance(v, list) else v) for k, v in rec.items()}
    out['stats'] = np.array([env._num_trunc, env._num_col, env._num_tar], np.int64)
    if record_pre:
        for k, v in pre.items():
            out['pre_' + k] = np.stack(v)
    return out


def save(name, meta, arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **{'meta_' + k: np.asarray(v) for k, v in meta.items()}, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, stats {arrays.get('stats')}")


def random_case(name, B, A, O, steps, seed, act_seed, patched, template=None, angle=0.2, **over):
    acts = actions_for(B, A, steps, act_seed, angle=angle)
    ctx = refload.oracle_trig() if patched else torch.no_grad()
    with ctx:
        env = make_env(B, A, O, seed, template, **over)
        init = dict(init_states=env.states.numpy().copy(), init_obstacles=env.obstacles.numpy().copy(),
                    init_obs=fused(env.observations()))
        arr = run_trace(env, acts, record_pre=not patched)
    arr.update(init)
    arr['actions'] = np.stack([a.numpy() for a in acts])
    meta = dict(B=B, A=A, O=O, steps=steps, seed=seed, act_seed=act_seed, angle=angle,
                episode_len=over.get('episode_len', 200))
    for k in FACTORS:                                # non-default reward factors travel with the trace
        if k in over:
            meta[k] = float(over[k])
    if template is not None:
        meta['template'] = np.asarray(template, np.float32)
    save(name, meta, arr)


def scenario_case(name, sn, patched):
    """`python -m marlnav -rc -sn {0,1}`: reference initialiser + reference sampler, untouched."""
    args = refload.reference_args(sampler_num=sn, num_obstacles=1)
    ctx = refload.oracle_trig() if patched else torch.no_grad()
    with ctx:
        params = copy.deepcopy(utils_mod.set_params(args)['env'])
        env = env_mod.Env(params)
        acts = []

        class Tap:          # record what the reference sampler hands out
            def __init__(self, inner): self.inner = inner
            def __call__(self):
                a = self.inner(); acts.append(a.numpy().copy()); return a
        env._sampler = Tap(env._sampler)
        T = 1000
        rec = dict(rewards=np.empty((T, 2), np.float32), terminated=np.empty((T, 2), bool),
                   truncated=np.empty((T, 2), bool), obs_sum=np.empty(T, np.uint64), obs_e0a0=np.empty((T, 8), np.float32))
        snaps = {}
        for t in range(T):
            obs, rew, term, trunc = env.step(env.sample_actions())
            o = fused(obs)
            rec['rewards'][t], rec['terminated'][t], rec['truncated'][t] = rew.numpy(), term.numpy(), trunc.numpy()
            rec['obs_sum'][t] = checksum(o); rec['obs_e0a0'][t] = o[0, 0]
            if t in (0, 1, 199, 200, 999):
                snaps[f'states_{t}'] = env.states.numpy().copy(); snaps[f'obs_{t}'] = o
        rec.update(snaps)
        rec['stats'] = np.array([env._num_trunc, env._num_col, env._num_tar], np.int64)
        rec['actions'] = np.stack(acts)
    save(name, dict(sn=sn, steps=1000), rec)


def snm1_case(name, seed, patched):
    """BASELINE.json configs[0], the literal `python -m marlnav -rc -sn -1 -se 0` setup
    (__main__.py:129-138, utils.py:217-222,237-243): triangle initialiser + ConstantSampler
    (action [0, 1] for every agent: cos = 1, sin = 0 exactly on any backend), B = 2, 1000 steps,
    through the reference's own set_params(); only the Philox reset sampler (and, for the
    patched set, acos) is injected.  The agents fly straight at the target, so this is the one
    scenario with target reaches in free running (delayed termination, Appendix B-1)."""
    args = refload.reference_args(sampler_num=-1)
    ctx = refload.oracle_trig() if patched else torch.no_grad()
    with ctx:
        params = copy.deepcopy(utils_mod.set_params(args)['env'])
        env = env_mod.Env(params)
        assert type(env._sampler).__name__ == 'ConstantSampler' and type(env._init_sampler).__name__ == 'TriangleIntitializer'
        smp = orc.PhiloxTemplateSampler(params, orc.triangle_template(params['init']), seed=seed)
        env._init_sampler = smp
        env.states, env.obstacles, env.target = smp()
        init = dict(init_states=env.states.numpy().copy(), init_obstacles=env.obstacles.numpy().copy(),
                    init_obs=fused(env.observations()))
        T, B = 1000, 2
        rec = dict(rewards=np.empty((T, B), np.float32), terminated=np.empty((T, B), bool),
                   truncated=np.empty((T, B), bool), obs_sum=np.empty(T, np.uint64), states_sum=np.empty(T, np.uint64),
                   obs_e0a0=np.empty((T, 12), np.float32), terminates=np.empty((T, B), bool),
                   step_num=np.empty((T, B), np.float32), num_tar=np.empty(T, np.int64))
        acts, snaps = [], {}
        for t in range(T):
            a = env.sample_actions()
            acts.append(a.numpy().copy())
            obs, rew, term, trunc = env.step(a)
            o = fused(obs)
            rec['rewards'][t], rec['terminated'][t], rec['truncated'][t] = rew.numpy(), term.numpy(), trunc.numpy()
            rec['obs_sum'][t], rec['states_sum'][t], rec['obs_e0a0'][t] = checksum(o), checksum(env.states.numpy()), o[0, 0]
            rec['terminates'][t], rec['step_num'][t], rec['num_tar'][t] = env._terminates.numpy(), env._step_num.numpy(), env._num_tar
            if term.any() or t in (0, 1, 999):
                snaps[f'states_{t}'] = env.states.numpy().copy(); snaps[f'obs_{t}'] = o
                snaps[f'obstacles_{t}'] = env.obstacles.numpy().copy()
        rec.update(snaps); rec.update(init)
        rec['stats'] = np.array([env._num_trunc, env._num_col, env._num_tar], np.int64)
        rec['actions'] = np.stack(acts)
        rec['reward_sum_f64'] = np.array(rec['rewards'].astype(np.float64).sum())
    save(name, dict(sn=-1, steps=T, seed=seed, B=B, A=3, O=3), rec)


def models_case(name):
    """The reference learner-side pieces SURVEY 8(f)-2/3 replace, run on fixed weights and inputs:
    Actor.forward -> dist.sample() / dist.log_prob() (models.py:14-36,113-115; the standard-normal
    draws of MultivariateNormal.rsample are injected through torch.distributions' own
    `_standard_normal` hook, everything else is the reference's code path including its Cholesky),
    Critic.forward (models.py:39-56) and MAPPO._process_rewards (models.py:131-148)."""
    import types
    import marlnav.models as models_mod
    import torch.distributions.multivariate_normal as mvn
    arrays = {}
    for tag, (S, H, N, A) in dict(a=(12, 50, 4096, 3), b=(48, 64, 512, 8)).items():
        torch.manual_seed(100 + S)
        actor = models_mod.Actor(S, H)
        critic = models_mod.Critic(A * S, H)
        with torch.no_grad():                      # spread the heads so tanh / softplus see their whole range
            actor.fc_mu.weight.mul_(1.7); actor.fc_std.weight.mul_(2.5); actor.fc_std.bias.add_(0.3)
        g = torch.Generator().manual_seed(7 + S)
        obs = (torch.rand(N // A, A, S, generator=g) * 2 - 1) * 1.2
        eps = torch.randn(N // A * A, 2, generator=g)
        saved = mvn._standard_normal
        mvn._standard_normal = lambda shape, dtype, device: eps.reshape(shape).to(dtype)
        try:
            with torch.no_grad():
                dist = actor(obs)
                actions = dist.sample()
                logp = dist.log_prob(actions)
                values = critic(obs)
        finally:
            mvn._standard_normal = saved
        arrays.update({f'{tag}_obs': obs.numpy(), f'{tag}_eps': eps.numpy(), f'{tag}_actions': actions.numpy(),
                       f'{tag}_log_probs': logp.numpy(), f'{tag}_mu': dist.loc.numpy(),
                       f'{tag}_var': torch.diagonal(dist.covariance_matrix, dim1=-2, dim2=-1).numpy().copy(),
                       f'{tag}_values': values.numpy()})
        for k, v in actor.state_dict().items(): arrays[f'{tag}_actor.{k}'] = v.numpy().copy()
        for k, v in critic.state_dict().items(): arrays[f'{tag}_critic.{k}'] = v.numpy().copy()
    # MAPPO._process_rewards on a fixed buffer, called unbound on a stand-in object (MAPPO.__init__
    # creates directories and optimisers that have nothing to do with it)
    T, B, gamma = 64, 37, 0.9
    g = torch.Generator().manual_seed(5)
    rewards = (torch.rand(T, B, generator=g) * 40 - 10)
    done = torch.rand(T, B, generator=g) < 0.06
    stub = types.SimpleNamespace(num_parallel=B, device='cpu', buffer_len=T, gamma=gamma, _mean_rew=0.,
                                 _logs={'mean_rews': []},
                                 buffer=[[None, None, None, None, rewards[i].clone(), done[i].clone()] for i in range(T)])
    raw = []
    orig_std_mean = torch.std_mean
    def tap(x, *a, **k):                            # the discounted returns before normalisation
        raw.append(x.clone()); return orig_std_mean(x, *a, **k)
    torch.std_mean = tap
    try:
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            models_mod.MAPPO._process_rewards(stub)
    finally:
        torch.std_mean = orig_std_mean
    arrays.update(ret_rewards=rewards.numpy(), ret_done=done.numpy(), ret_gamma=np.array(gamma),
                  ret_returns=raw[0].reshape(T, B).numpy(),
                  ret_normalized=torch.stack([stub.buffer[i][-2] for i in range(T)]).numpy(),
                  ret_mean=np.array(float(stub._mean_rew)))
    save(name, dict(torch=torch.__version__), arrays)


class NoisyAgentsInjector:
    """The REAL TriangleIntitializer with noisy_ags = True (utils.py:25,381-388) fed addressed draws:
    its `pos_noise.sample()` gets the Philox/Box-Muller normals through torch.distributions' own
    `_standard_normal` hook and its `torch.rand(B, 3)` the Philox uniforms, so everything it
    computes from them -- scale_tril @ eps, ags_dist * ., angle_range * (u - 0.5), the vmapped
    rotation, the cat -- is the reference's own code.  Obstacles: the addressed draw as elsewhere."""

    def __init__(self, ref_init, params, seed):
        self.ref, self.seed, self.counter = ref_init, int(seed), 0
        self.p = orc.make_params(params)
        assert ref_init.noisy_ags == 1

    def __call__(self):
        import torch.distributions.multivariate_normal as mvn
        B = self.ref.num_parallel
        normals, uniforms = orc.agent_draws(self.seed, self.counter, 0, B, 3)
        saved = (mvn._standard_normal, torch.rand)
        mvn._standard_normal = lambda shape, dtype, device: torch.from_numpy(normals).reshape(tuple(shape)).clone()
        torch.rand = lambda *a, **k: torch.from_numpy(uniforms).clone()
        try:
            states = self.ref._sample_agents()
        finally:
            mvn._standard_normal, torch.rand = saved
        obst = orc.philox_obstacles(self.p, self.seed, self.counter, 0)
        self.counter += 1
        return states, torch.from_numpy(obst), self.ref.target


def noisy_case(name, B, steps, seed, act_seed, patched):
    """SURVEY 8(f)-4's last piece: the triangle initialiser with its Gaussian position noise and
    heading rotation switched on (`noisy_ags`, hard-coded False at utils.py:25), free-running."""
    acts = actions_for(B, 3, steps, act_seed)
    ctx = refload.oracle_trig() if patched else torch.no_grad()
    with ctx:
        args = refload.reference_args(num_parallel=B, num_agents=3, num_obstacles=3, episode_len=60)
        params = copy.deepcopy(utils_mod.set_params(args)['env'])
        params['init']['noisy_ags'] = True
        env = env_mod.Env(params)
        inj = NoisyAgentsInjector(env._init_sampler, params, seed)
        env._init_sampler = inj
        env.states, env.obstacles, env.target = inj()
        init = dict(init_states=env.states.numpy().copy(), init_obstacles=env.obstacles.numpy().copy(),
                    init_obs=fused(env.observations()))
        arr = run_trace(env, acts, record_pre=not patched)
    arr.update(init)
    arr['actions'] = np.stack([a.numpy() for a in acts])
    save(name, dict(B=B, A=3, O=3, steps=steps, seed=seed, act_seed=act_seed, angle=0.2, episode_len=60, noisy=1), arr)


def factor_cases():
    """The CLI defaults multiply the risk and the distance score by 0 (`-rf 0 -df 0`, __main__.py:89-92), so
    the default traces cannot see an error in them: the same free-running traces with every reward
    factor non-zero, (3,3) and the (8,16) ring."""
    fac = dict(risk_factor=37.5, distance_factor=123.0, heading_factor=410.0, target_factor=650.0,
               soft_factor=275.0, bond_factor=7.25)
    random_case('patched_tri_3x3_factors', 40, 3, 3, 140, seed=9, act_seed=5, patched=True, episode_len=45, **fac)
    random_case('patched_ring_8x16_factors', 8, 8, 16, 50, seed=10, act_seed=6, patched=True,
                template=orc.ring_template(8), episode_len=30, **fac)
    random_case('stock_tri_3x3_factors', 40, 3, 3, 80, seed=9, act_seed=5, patched=False, episode_len=45, **fac)


def quirk_case(name):
    """delayed target termination, collision+target,
    truncation -- 4 hand-placed envs, stock reference, constant action [0, -10]."""
    B, A, O = 4, 3, 3
    with refload.oracle_trig():
        env = make_env(B, A, O, seed=0, episode_len=5)
        env.obstacles[:] = 100.0
        base = env.states.clone()
        for e, dy in ((0, 10.0), (1, 3.0)):      # agents stacked near the target at (1350,375)
            x = 1330.0 if e == 0 else 1340.0
            for i, off in enumerate((0.0, dy, -dy)):
                base[e, i] = torch.tensor([x, 375.0 + off, 1.0, 0.0, 3.0])
        env.states = base
        acts = [torch.tensor([0.0, -10.0]).repeat(B, A, 1) for _ in range(12)]
        init = dict(init_states=env.states.numpy().copy(), init_obstacles=env.obstacles.numpy().copy())
        arr = run_trace(env, acts)
        # _terminates and step_num after every step are the point of these scenarios
    arr.update(init)
    arr['actions'] = np.stack([a.numpy() for a in acts])
    save(name, dict(B=B, A=A, O=O, steps=12, seed=0, episode_len=5), arr)


if __name__ == '__main__':
    assert refload.available(), "the reference must be mounted at /root/reference"
    only = set(sys.argv[1:])                       # e.g. `make_golden.py snm1 models`: just those families
    if only:
        if 'snm1' in only:
            snm1_case('patched_rc_snm1', 0, patched=True)
            snm1_case('stock_rc_snm1', 0, patched=False)
        if 'models' in only:
            models_case('ref_models')
        if 'factors' in only:
            factor_cases()
        if 'noisy' in only:
            noisy_case('patched_noisy_3x3', 24, 200, seed=11, act_seed=77, patched=True)
            noisy_case('stock_noisy_3x3', 24, 80, seed=11, act_seed=77, patched=False)
        sys.exit(0)
    random_case('patched_tri_3x3', 48, 3, 3, 260, seed=7, act_seed=1234, patched=True)
    random_case('patched_tri_3x3_wide', 16, 3, 3, 60, seed=8, act_seed=99, patched=True, angle=4.0)
    random_case('patched_ring_8x16', 12, 8, 16, 80, seed=5, act_seed=4321, patched=True,
                template=orc.ring_template(8))
    random_case('patched_ring_4x2', 16, 4, 2, 60, seed=3, act_seed=11, patched=True,
                template=orc.ring_template(4))
    random_case('patched_ring_9x5', 6, 9, 5, 40, seed=2, act_seed=12, patched=True,
                template=orc.ring_template(9))
    random_case('stock_tri_3x3', 48, 3, 3, 120, seed=7, act_seed=1234, patched=False)
    random_case('stock_ring_8x16', 8, 8, 16, 40, seed=5, act_seed=4321, patched=False,
                template=orc.ring_template(8))
    for sn in (0, 1):
        scenario_case(f'patched_rc_sn{sn}', sn, patched=True)
        scenario_case(f'stock_rc_sn{sn}', sn, patched=False)
    quirk_case('patched_quirks')
    snm1_case('patched_rc_snm1', 0, patched=True)
    snm1_case('stock_rc_snm1', 0, patched=False)
    models_case('ref_models')
    noisy_case('patched_noisy_3x3', 24, 200, seed=11, act_seed=77, patched=True)
    noisy_case('stock_noisy_3x3', 24, 80, seed=11, act_seed=77, patched=False)
    factor_cases()
