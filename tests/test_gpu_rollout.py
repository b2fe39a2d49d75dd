"""GPU: the caller-side kernels against torch restatements of
the reference's learner code (oracle.actor_reference / oracle.discounted_returns_reference)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _actor_weights(S=12, H=50, seed=0):
    g = torch.Generator().manual_seed(seed)
    w = {'fc1.weight': torch.empty(H, S), 'fc_mu.weight': torch.empty(2, H), 'fc_std.weight': torch.empty(2, H)}
    for v in w.values():
        torch.nn.init.orthogonal_(v, generator=g)          # models.py:20-24
    w.update({'fc1.bias': torch.rand(H, generator=g) * 0.2 - 0.1, 'fc_mu.bias': torch.rand(2, generator=g) - 0.5,
              'fc_std.bias': torch.rand(2, generator=g) - 0.5})
    return w


@pytest.mark.parametrize("S,H,N", [(12, 50, 3072), (48, 50, 1000), (8, 7, 33)])
def test_fused_actor_matches_torch(oracle, S, H, N):
    import marlnav_b200 as mb
    w = _actor_weights(S, H)
    g = torch.Generator().manual_seed(1)
    obs = torch.rand(N, S, generator=g) * 2 - 1
    eps = torch.randn(N, 2, generator=g)
    fa = mb.FusedActor(w, seed=5)
    act, lp, mu, var = fa.act(obs.cuda(), eps=eps, want_moments=True)
    # the same torch ops in float64: a reference that does not depend on how the host's float32 GEMM
    # splits its sums (MKL threading made a float32 reference flaky at the 1e-5 level); the kernel's
    # float32 dot products must agree with it to 1e-5 relative.  (The float32 outputs of the reference's
    # own Actor are pinned in test_models_golden.py.)
    w64 = {k: v.double() for k, v in w.items()}
    r_act, r_lp, r_mu, r_var = oracle.actor_reference(obs.double(), w64, eps.double())
    np.testing.assert_allclose(mu.cpu().numpy(), r_mu.numpy(), rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(var.cpu().numpy(), r_var.numpy(), rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(act.cpu().numpy(), r_act.numpy(), rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), r_lp.numpy(), rtol=1e-4, atol=1e-4)


def test_fused_actor_sampling_is_standard_normal_and_addressed():
    import marlnav_b200 as mb
    w = _actor_weights()
    obs = torch.zeros(1 << 18, 12)
    fa = mb.FusedActor(w, seed=9)
    act, lp, mu, var = fa.act(obs.cuda(), want_moments=True)
    z = ((act - mu) / var.sqrt()).cpu().double()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert abs((z[:, 0] * z[:, 1]).mean()) < 0.01                      # independent components
    assert abs((z ** 4).mean() - 3) < 0.1                              # Gaussian kurtosis
    fb = mb.FusedActor(w, seed=9)
    act2, _ = fb.act(obs.cuda())
    assert torch.equal(act, act2)                                      # (seed, counter, row) addressed
    act3, _ = fb.act(obs.cuda())
    assert not torch.equal(act2, act3)                                 # next call, next counter
    want_lp = -0.5 * (z ** 2).sum(1) - 0.5 * var.cpu().double().log().sum(1) - math.log(2 * math.pi)
    np.testing.assert_allclose(lp.cpu().double().numpy(), want_lp.numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("T,B", [(1000, 257), (37, 4096)])
def test_discounted_returns_bit_exact(oracle, T, B):
    """Same float64 operations in the same order as models.py:131-139 -> identical bits."""
    import marlnav_b200 as mb
    g = torch.Generator().manual_seed(3)
    rew = (torch.rand(T, B, generator=g) * 700 - 350)
    done = torch.rand(T, B, generator=g) < 0.02
    got = mb.discounted_returns(rew.cuda(), done.cuda(), 0.9).cpu()
    want = oracle.discounted_returns_reference(rew, done, 0.9)
    assert torch.equal(got, want)
    norm = mb.discounted_returns(rew.cuda(), done.cuda(), 0.9, normalize=True).cpu()
    std, mean = torch.std_mean(want.reshape(-1))
    np.testing.assert_allclose(norm.numpy(), ((want - mean) / (std + 1e-12)).numpy(), rtol=1e-9, atol=1e-9)


def test_collect_rollout_matches_stepwise_loop(oracle):
    """collect_rollout == calling FusedActor.act / Env.step_fused by hand with the same seeds, and
    its buffers have the layouts MAPPO.get_data stores (models.py:121)."""
    import marlnav_b200 as mb
    B, A, O, T = 512, 3, 3, 40
    max_d = math.sqrt(1500.0 ** 2 + 750.0 ** 2)
    lo = [-math.pi, 0.] + O * [-math.pi] + O * [0.] + (A - 1) * [-math.pi] + (A - 1) * [0.]
    hi = [math.pi, max_d] + O * [math.pi] + O * [max_d] + (A - 1) * [math.pi] + (A - 1) * [max_d]
    norm, scal = dict(min_obs=lo, max_obs=hi), dict(min_action=[-math.pi, -0.5], max_action=[math.pi, 0.5])
    w = _actor_weights()
    envs = []
    for _ in range(2):
        p = mb.default_env_params(B, A, O, sampling_style='policy'); p['seed'] = 4
        envs.append(mb.Env(p))
    critic = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(A * 12, 50), torch.nn.ReLU(), torch.nn.Linear(50, 1)).cuda()
    buf = mb.collect_rollout(envs[0], mb.FusedActor(w, seed=11), T, critic=critic, normalizer_params=norm, scaler_params=scal)
    assert buf['obs'].shape == (T, B, A, 12) and buf['actions'].shape == (T, B * A, 2)
    assert buf['log_probs'].shape == (T, B * A) and buf['rewards'].shape == (T, B) and buf['values'].shape == (T, B, 1)
    assert buf['done'].dtype == torch.bool and bool(buf['done'].any())
    env, fa = envs[1], mb.FusedActor(w, seed=11)
    env.fuse_io(norm, scal)
    obs = (env.observations_fused() - env._io_tensors[0]) / env._io_tensors[1]
    for t in range(T):
        assert torch.equal(buf['obs'][t], obs)
        act, lp = fa.act(obs)
        assert torch.equal(buf['actions'][t], act) and torch.equal(buf['log_probs'][t], lp)
        obs, rew, term, trunc = env.step_fused(act.view(B, A, 2))
        assert torch.equal(buf['rewards'][t], rew) and torch.equal(buf['done'][t], term | trunc)
    ret = mb.discounted_returns(buf['rewards'], buf['done'], 0.9, normalize=True)
    assert ret.shape == (T, B) and ret.dtype == torch.float64 and bool(torch.isfinite(ret).all())


@pytest.mark.parametrize("B,O,H", [(515, 3, 50), (64, 1, 7), (300, 4, 64)])
def test_fused_actor_step_matches_two_launches(B, O, H):
    """marlnav_act_step_f32 (actor sampled inside the step launch) == marlnav_actor_sample_f32 followed
    by marlnav_step_f32: every buffer of the rollout and the final env state, bit for bit."""
    import marlnav_b200 as mb
    A, T = 3, 70
    S = 2 + 2 * O + 2 * (A - 1)
    max_d = math.sqrt(1500.0 ** 2 + 750.0 ** 2)
    lo = [-math.pi, 0.] + O * [-math.pi] + O * [0.] + (A - 1) * [-math.pi] + (A - 1) * [0.]
    hi = [math.pi, max_d] + O * [math.pi] + O * [max_d] + (A - 1) * [math.pi] + (A - 1) * [max_d]
    norm, scal = dict(min_obs=lo, max_obs=hi), dict(min_action=[-math.pi, -0.5], max_action=[math.pi, 0.5])
    w = _actor_weights(S, H)
    bufs, envs = [], []
    for fuse in (True, False):
        p = mb.default_env_params(B, A, O, sampling_style='policy', episode_len=30); p['seed'] = 21
        env = mb.Env(p)
        env.fuse_io(norm, scal)
        fa = mb.FusedActor(w, seed=9)
        assert env.supports_fused_actor(fa)
        bufs.append(mb.collect_rollout(env, fa, T, fuse_actor=fuse))
        envs.append(env)
    for key in ('obs', 'last_obs', 'actions', 'log_probs', 'rewards', 'done'):
        assert torch.equal(bufs[0][key], bufs[1][key]), key
    assert bool(bufs[0]['done'].any())
    assert torch.equal(envs[0].states, envs[1].states) and torch.equal(envs[0].obstacles, envs[1].obstacles)
    assert (envs[0]._num_trunc, envs[0]._num_col) == (envs[1]._num_trunc, envs[1]._num_col)


@pytest.mark.parametrize("K,H,B", [(36, 50, 1024), (384, 50, 300), (16, 7, 5), (36, 50, 5000), (36, 64, 40001), (72, 33, 2049)])
def test_fused_critic_matches_torch(K, H, B):
    """Critic.forward (models.py:39-56) as one kernel vs the torch module."""
    import marlnav_b200 as mb
    g = torch.Generator().manual_seed(2)
    fc1, fc2 = torch.nn.Linear(K, H), torch.nn.Linear(H, 1)
    torch.nn.init.orthogonal_(fc1.weight, generator=g); torch.nn.init.orthogonal_(fc2.weight, generator=g)
    sd = {'fc1.weight': fc1.weight, 'fc1.bias': fc1.bias, 'fc2.weight': fc2.weight, 'fc2.bias': fc2.bias}
    x = torch.rand(B, K, generator=g) * 2 - 1
    # float64 reference: independent of how the host's float32 GEMM splits its sums (K up to 384)
    want = (torch.relu(x.double() @ fc1.weight.double().T + fc1.bias.double()) @ fc2.weight.double().T
            + fc2.bias.double()).detach()
    got = mb.FusedCritic(sd)(x.cuda()).cpu()
    assert got.shape == (B, 1)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=2e-5, atol=1e-5)


def test_device_counter_mode_is_identical(oracle):
    """Env.use_device_counter(): same results as the host-supplied counter, bit for bit."""
    import marlnav_b200 as mb
    from helpers import action_pool
    p = mb.default_env_params(700, 3, 3, sampling_style='policy'); p['seed'] = 6
    e_host, e_dev = mb.Env(dict(p)), mb.Env(dict(p))
    e_dev.use_device_counter(True)
    pool = action_pool(700, 3, angle=0.3)
    for t in range(120):
        a = pool[t % len(pool)].cuda()
        o1, r1, t1, tr1 = e_host.step_fused(a)
        o2, r2, t2, tr2 = e_dev.step_fused(a)
        assert torch.equal(o1, o2) and torch.equal(r1, r2) and torch.equal(t1, t2) and torch.equal(tr1, tr2)
    assert torch.equal(e_host.obstacles, e_dev.obstacles) and e_host._num_col == e_dev._num_col > 0


@pytest.mark.parametrize("B", [300, 40000])
def test_cuda_graph_of_back_to_back_steps_matches_eager(B):
    """Plain step launches carry the programmatic-dependent-launch attribute; captured back to back
    (device counter in batch mode: no kernel in between) they become programmatic graph edges.
    Replays must equal the eager loop, resets included."""
    import marlnav_b200 as mb
    from helpers import action_pool
    T = 64
    p = mb.default_env_params(B, 3, 3, sampling_style='policy', episode_len=25); p['seed'] = 31
    pool = [a.cuda() for a in action_pool(B, 3, angle=0.3)]
    eager = mb.Env(dict(p))
    ref = [[t.clone() for t in eager.step_fused(pool[i % len(pool)])] for i in range(2 * T)]
    env = mb.Env(dict(p))
    env.use_device_counter(True)
    outs = [env._alloc_outputs() for _ in range(T)]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            env.batch_device_counter(True)
            for i in range(T):
                env.step_fused(pool[i % len(pool)], out=outs[i])
            env.batch_device_counter(False)
    torch.cuda.current_stream().wait_stream(side)
    for k in range(2):                       # T is a multiple of the pool length: same actions per replay
        g.replay()
        torch.cuda.synchronize()
        for i in range(T):
            o, r, te, tr = outs[i]
            ro, rr, rte, rtr = ref[k * T + i]
            assert torch.equal(o, ro) and torch.equal(r, rr), (k, i)
            assert torch.equal(te.view(torch.bool), rte) and torch.equal(tr.view(torch.bool), rtr), (k, i)
    assert torch.equal(env.states, eager.states) and torch.equal(env.obstacles, eager.obstacles)


def test_rollout_graph_replays_continue_the_eager_streams():
    """A captured rollout replayed twice == two eager collect_rollout calls (fresh reset positions
    and fresh action noise on every replay), bit for bit."""
    import marlnav_b200 as mb
    B, A, O, T = 256, 3, 3, 60
    max_d = math.sqrt(1500.0 ** 2 + 750.0 ** 2)
    lo = [-math.pi, 0.] + O * [-math.pi] + O * [0.] + (A - 1) * [-math.pi] + (A - 1) * [0.]
    hi = [math.pi, max_d] + O * [math.pi] + O * [max_d] + (A - 1) * [math.pi] + (A - 1) * [max_d]
    norm, scal = dict(min_obs=lo, max_obs=hi), dict(min_action=[-math.pi, -0.5], max_action=[math.pi, 0.5])
    w = _actor_weights()
    g = torch.Generator().manual_seed(8)
    fc1, fc2 = torch.nn.Linear(A * 12, 50), torch.nn.Linear(50, 1)
    csd = {'fc1.weight': fc1.weight, 'fc1.bias': fc1.bias, 'fc2.weight': fc2.weight, 'fc2.bias': fc2.bias}

    def make():
        p = mb.default_env_params(B, A, O, sampling_style='policy', episode_len=25); p['seed'] = 12
        env = mb.Env(p); env.fuse_io(norm, scal)
        return env, mb.FusedActor(w, seed=3), mb.FusedCritic(csd)
    env_e, act_e, cri_e = make()
    mb.collect_rollout(env_e, act_e, 2, critic=cri_e)                 # RolloutGraph's warm-up, eagerly
    eager = [mb.collect_rollout(env_e, act_e, T, critic=cri_e) for _ in range(2)]
    env_g, act_g, cri_g = make()
    rg = mb.RolloutGraph(env_g, act_g, T, critic=cri_g, warmup_steps=2)
    for k in range(2):
        buf = rg.replay()
        torch.cuda.synchronize()
        for key in ('obs', 'actions', 'log_probs', 'rewards', 'done', 'values'):
            assert torch.equal(buf[key], eager[k][key]), (k, key)
    assert not torch.equal(eager[0]['actions'], eager[1]['actions'])
    assert torch.equal(env_g.states, env_e.states) and torch.equal(env_g.obstacles, env_e.obstacles)


@pytest.mark.gpu
@pytest.mark.parametrize("B,keep", [(96, True), (20000, False)])
def test_step_graph_replays_continue_the_eager_loop(B, keep):
    """marlnav_b200.StepGraph (K captured steps, batched device counter) replayed twice == 2K eager
    steps, bit for bit; thread-per-agent (small batch) and thread-per-env (large batch) kernels."""
    import marlnav_b200 as mb
    from helpers import action_pool
    p = mb.default_env_params(B, 3, 3, sampling_style='policy', episode_len=25); p['seed'] = 77
    pool = [a.cuda() for a in action_pool(B, 3, angle=0.3)]
    K = 2 * len(pool)
    acts = [pool[i % len(pool)] for i in range(K)]
    eager = mb.Env(dict(p))
    seq = [acts[0]] + acts + acts                         # StepGraph's warm-up step runs acts[0] once
    ref = [[t.clone() for t in eager.step_fused(a)] for a in seq]
    env = mb.Env(dict(p))
    sg = mb.StepGraph(env, acts, keep_outputs=keep)
    for k in range(2):
        outs = sg.replay()
        torch.cuda.synchronize()
        steps = range(K) if keep else [K - 1]
        for i in steps:
            o, r, te, tr = outs[i if keep else 0]
            ro, rr, rte, rtr = ref[1 + k * K + i]
            assert torch.equal(o, ro) and torch.equal(r, rr), (k, i)
            assert torch.equal(te.view(torch.bool), rte) and torch.equal(tr.view(torch.bool), rtr), (k, i)
    assert torch.equal(env.states, eager.states) and torch.equal(env.obstacles, eager.obstacles)
    env.use_device_counter(False)
    assert env._reset_counter == eager._reset_counter


def test_sharded_rollout_equals_single_process():
    """The actor's sampling noise is addressed by GLOBAL (env, agent) row (marlnav_actor_spec.row_offset,
    ABI 4): two ranks' rollouts of two half batches are the single-process rollout, bit for bit --
    with the fused {actor -> step} launch and with the two-launch route."""
    import marlnav_b200 as mb
    B, A, O, T = 600, 3, 3, 60
    S = 2 + 2 * O + 2 * (A - 1)
    max_d = math.sqrt(1500.0 ** 2 + 750.0 ** 2)
    lo = [-math.pi, 0.] + O * [-math.pi] + O * [0.] + (A - 1) * [-math.pi] + (A - 1) * [0.]
    hi = [math.pi, max_d] + O * [math.pi] + O * [max_d] + (A - 1) * [math.pi] + (A - 1) * [max_d]
    norm, scal = dict(min_obs=lo, max_obs=hi), dict(min_action=[-math.pi, -0.5], max_action=[math.pi, 0.5])
    w = _actor_weights(S, 50)
    full = mb.default_env_params(B, A, O, sampling_style='policy', episode_len=25); full['seed'] = 8
    for fuse in (True, False):
        env = mb.Env(dict(full)); env.fuse_io(norm, scal)
        whole = mb.collect_rollout(env, mb.FusedActor(w, seed=13), T, fuse_actor=fuse)
        parts = []
        for r in range(2):
            e = mb.Env(mb.shard_env_params(full, r, 2)); e.fuse_io(norm, scal)
            parts.append(mb.collect_rollout(e, mb.FusedActor(w, seed=13), T, fuse_actor=fuse))
        for key, dim in (('obs', 1), ('actions', 1), ('log_probs', 1), ('rewards', 1), ('done', 1)):
            assert torch.equal(whole[key], torch.cat([p[key] for p in parts], dim=dim)), (fuse, key)
        assert bool(whole['done'].any())


@pytest.mark.parametrize("B", [1, 5, 33, 1000, 16385, 32769, 40001])
def test_fused_actor_step_ragged_batches(B):
    """Batches that are not a multiple of the tile (8 envs per warp thread-per-agent up to 32 768 envs, 32
    thread-per-env above)
    through the fused {actor -> step} launch == the two-launch route."""
    import marlnav_b200 as mb
    A, O, T = 3, 3, 24
    S = 12
    max_d = math.sqrt(1500.0 ** 2 + 750.0 ** 2)
    lo = [-math.pi, 0.] + O * [-math.pi] + O * [0.] + (A - 1) * [-math.pi] + (A - 1) * [0.]
    hi = [math.pi, max_d] + O * [math.pi] + O * [max_d] + (A - 1) * [math.pi] + (A - 1) * [max_d]
    norm, scal = dict(min_obs=lo, max_obs=hi), dict(min_action=[-math.pi, -0.5], max_action=[math.pi, 0.5])
    w = _actor_weights(S, 50)
    bufs = []
    for fuse in (True, False):
        p = mb.default_env_params(B, A, O, sampling_style='policy', episode_len=10); p['seed'] = 2
        env = mb.Env(p); env.fuse_io(norm, scal)
        bufs.append(mb.collect_rollout(env, mb.FusedActor(w, seed=4), T, fuse_actor=fuse))
    for key in ('obs', 'last_obs', 'actions', 'log_probs', 'rewards', 'done'):
        assert torch.equal(bufs[0][key], bufs[1][key]), key


def test_actor_accepts_its_largest_advertised_shape():
    """S = 48 (the (8,16) team) with H = 256 needs 54 KiB of dynamic shared memory (> the 48 KiB default)."""
    import marlnav_b200 as mb
    w = _actor_weights(48, 256)
    obs = torch.rand(100, 48) * 2 - 1
    act, lp = mb.FusedActor(w, seed=1).act(obs.cuda())
    assert bool(torch.isfinite(act).all()) and bool(torch.isfinite(lp).all())


def _mappo_like(tmp_path, B=256, T=40, use_seed=3):
    """An object with what MappoRollout needs from the reference's MAPPO (models.py:59-104), built from
    torch modules of the reference's shapes -- the reference itself does not travel to the GPU box."""
    import types
    import marlnav_b200 as mb
    A, O, S, H = 3, 3, 12, 50
    torch.manual_seed(use_seed)

    class Actor(torch.nn.Module):                      # models.py:14-36 (parameter names matter)
        def __init__(self):
            super().__init__()
            self.fc1, self.fc_mu, self.fc_std = torch.nn.Linear(S, H), torch.nn.Linear(H, 2), torch.nn.Linear(H, 2)

    class Critic(torch.nn.Module):                     # models.py:39-56
        def __init__(self):
            super().__init__()
            self.fc1, self.fc2 = torch.nn.Linear(A * S, H), torch.nn.Linear(H, 1)

        def forward(self, x):
            return self.fc2(torch.relu(self.fc1(x.flatten(1))))

    max_d = math.sqrt(1500.0 ** 2 + 750.0 ** 2)
    lo = torch.tensor([-math.pi, 0.] + O * [-math.pi] + O * [0.] + (A - 1) * [-math.pi] + (A - 1) * [0.]).cuda()
    hi = torch.tensor([math.pi, max_d] + O * [math.pi] + O * [max_d] + (A - 1) * [math.pi] + (A - 1) * [max_d]).cuda()
    a_lo, a_hi = torch.tensor([-math.pi, -0.5]).cuda(), torch.tensor([math.pi, 0.5]).cuda()
    norm = types.SimpleNamespace(mean=0.5 * (lo + hi), scale_tensor=(0.5 * (hi - lo)).repeat(1, A, 1))     # utils.py:523-528
    scal = types.SimpleNamespace(mean=0.5 * (a_lo + a_hi), scale_tensor=(0.5 * (a_hi - a_lo)).repeat(1, A, 1))
    p = mb.default_env_params(B, A, O, sampling_style='policy', episode_len=25); p['seed'] = 6
    env = mb.Env(p)
    m = types.SimpleNamespace(env=env, actor=Actor().cuda(), critic=Critic().cuda(), _normalize=norm, _scale_up=scal,
                              buffer_len=T, gamma=0.9, num_parallel=B, num_agents=A, action_size=2, buffer=[], obs=None,
                              _logs={'mean_rews': [], 'epi_stats': {'trunc': [], 'col': [], 'tar': []}},
                              _mean_rew=0., _max_rew=float('-inf'),
                              _actor_path=str(tmp_path / 'actor.pt'), _critic_path=str(tmp_path / 'critic.pt'))

    def _update_epi_stats():                           # models.py:151-158
        for k, attr in (('trunc', '_num_trunc'), ('col', '_num_col'), ('tar', '_num_tar')):
            m._logs['epi_stats'][k] += [getattr(env, attr)]; setattr(env, attr, 0)
    m._update_epi_stats = _update_epi_stats
    return m


@pytest.mark.parametrize("use_graph", [False, True])
def test_mappo_rollout_fills_the_reference_buffer(tmp_path, oracle, use_graph):
    """MappoRollout.get_data == MAPPO.get_data's contract (models.py:106-158): the list-of-lists buffer
    in the reference's layout, returns scanned and normalised like _process_rewards, episode counters
    logged and reset, weights saved; repeated rollouts continue the env and see refreshed weights."""
    import marlnav_b200 as mb
    B, T, A = 256, 40, 3
    m = _mappo_like(tmp_path, B, T)
    mb.MappoRollout(m, use_graph=use_graph, seed=1).attach()
    m.get_data()
    assert len(m.buffer) == T and [tuple(x.shape) for x in m.buffer[0]] == [(B, A, 12), (B, A, 2), (B * A,), (B, 1), (B,), (B,)]
    obs, act, lp, val, ret, done = (torch.stack([row[k] for row in m.buffer]) for k in range(6))
    assert ret.dtype == torch.float64 and done.dtype == torch.bool and bool(done.any())
    assert float(obs.abs().max()) <= 1.0 + 1e-6                              # normalised observations
    # the critic's values are the module's own outputs on those observations
    np.testing.assert_allclose(val.cpu().numpy(), m.critic(obs.reshape(T * B, A, 12)).detach().reshape(T, B, 1).cpu().numpy(),
                               rtol=2e-5, atol=2e-5)
    # returns: unnormalise with the logged mean and compare the recurrence on the raw rewards is not
    # possible from the buffer alone, so check the normalisation itself and the recurrence's zeros
    assert abs(float(ret.mean())) < 1e-9 and abs(float(ret.std()) - 1.0) < 1e-9
    assert len(m._logs['mean_rews']) == 1 and len(m._logs['epi_stats']['col']) == 1
    assert m._logs['epi_stats']['trunc'][0] + m._logs['epi_stats']['col'][0] >= int(done.sum()) - B   # (B-2 delayed terminations at most)
    assert (m.env._num_trunc, m.env._num_col, m.env._num_tar) == (0, 0, 0)
    assert os.path.exists(m._actor_path) and os.path.exists(m._critic_path)
    first_obs = obs[0].clone()
    with torch.no_grad():                                                    # "an optimiser step"
        for prm in m.actor.parameters():
            prm.add_(0.05 * torch.randn_like(prm))
    m.get_data()
    obs2 = torch.stack([row[0] for row in m.buffer])
    assert len(m._logs['mean_rews']) == 2 and not torch.equal(obs2[0], first_obs)
    assert torch.equal(m.obs, m.env.observations_fused().sub(m.env._io_tensors[0]).div(m.env._io_tensors[1]))
