"""GPU: the drop-in surface the reference's callers touch and full-size
(BASELINE.json configs[3]/[4]) checks: bit-exact against the oracle for a few steps at
1 048 576 envs, plus size-independent invariants of the step."""
import numpy as np
import pytest
import torch

from helpers import action_pool, assert_bits_equal, cpu_params

pytestmark = pytest.mark.gpu


def _env(B=64, A=3, O=3, seed=1, **kw):
    import marlnav_b200 as mb
    p = mb.default_env_params(B, A, O, **kw) if A == 3 else mb.template_env_params(B, A, O)
    p['seed'] = seed
    return mb.Env(p), p


def test_observations_namedtuple_matches_reference_shapes():
    env, _ = _env(B=10, sampling_style='policy')
    obs = env.observations()
    assert type(obs).__name__ == 'Observations'
    assert obs._fields == ('target_angle', 'target_distance', 'obstacles_angles', 'obstacles_distances',
                           'others_angles', 'others_distances')
    assert [tuple(t.shape) for t in obs] == [(10, 3, 1), (10, 3, 1), (10, 3, 3), (10, 3, 3), (10, 3, 2), (10, 3, 2)]
    fused = env.observations_fused()
    assert torch.equal(torch.cat(obs, dim=2), fused)            # what ObsNormalizer does (utils.py:531)
    # triangle reset observation, SURVEY.md Appendix C-4
    assert torch.allclose(obs.target_angle[0, :, 0].cpu(), torch.tensor([0.016504526, 0., 0.016504526]), atol=1e-7)
    assert torch.allclose(obs.target_distance[0, :, 0].cpu(), torch.tensor([1211.7120361, 1176.9060059, 1211.7120361]))
    assert torch.allclose(obs.others_angles[0].cpu(), torch.tensor([[0.52359891, 1.57079637], [2.61799383, 2.61799383],
                                                                    [1.57079637, 0.52359891]]), atol=1e-6)


def test_step_returns_fresh_tensors_and_reference_types():
    env, _ = _env(B=32, sampling_style='policy')
    act = torch.zeros(32, 3, 2, device='cuda')
    o1, r1, t1, tr1 = env.step(act)
    o2, r2, t2, tr2 = env.step(act)
    assert r1.data_ptr() != r2.data_ptr() and o1[0].data_ptr() != o2[0].data_ptr()   # models.py:121 keeps references
    assert r1.dtype == torch.float32 and t1.dtype == torch.bool and tr1.dtype == torch.bool
    assert r1.shape == (32,) and t1.shape == (32,) and tr1.shape == (32,)
    assert env.states.shape == (32, 3, 5) and env.obstacles.shape == (32, 3, 2) and env.target.shape == (32, 1, 2)
    obs, params = env.reset()                                    # environment.py:70-74
    assert params is env.params and len(obs) == 6


def test_episode_counters_are_readable_and_assignable():
    """models.py:151-158 reads env._num_* and assigns 0."""
    env, _ = _env(B=256, sampling_style='policy', episode_len=3)
    act = torch.zeros(256, 3, 2, device='cuda')
    for _ in range(3):
        env.step(act)
    assert env._num_trunc == 256 and isinstance(env._num_trunc, int)
    env._num_trunc = 0; env._num_col = 0; env._num_tar = 0
    assert (env._num_trunc, env._num_col, env._num_tar) == (0, 0, 0)
    for _ in range(3):
        env.step(act)
    assert env._num_trunc == 256


def test_accepts_cpu_and_noncontiguous_actions_and_side_streams(oracle):
    env, p = _env(B=100, sampling_style='policy')
    oe = oracle.OracleEnv(cpu_params(p), seed=1)
    g = torch.Generator().manual_seed(0)
    big = torch.rand(100, 3, 4, generator=g) - 0.5
    act_nc = big[:, :, ::2]                                      # non-contiguous CPU tensor
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        obs, rew, _, _ = env.step_fused(act_nc)
    s.synchronize()
    o_obs, o_rew, _, _ = oe.step_fused(act_nc.contiguous().numpy())
    assert_bits_equal("obs", obs.cpu().numpy(), o_obs)
    assert_bits_equal("rew", rew.cpu().numpy(), o_rew)
    with pytest.raises(Exception):
        env.step_fused(torch.zeros(99, 3, 2))


def test_sample_actions_const_sampler():
    env, _ = _env(B=4)                                           # -sn -1 -sa sampler: ConstantSampler
    a = env.sample_actions()
    assert a.shape == (4, 3, 2) and a.is_cuda and torch.equal(a[1, 2].cpu(), torch.tensor([0., 1.]))
    env2, _ = _env(B=4, sampling_style='policy')
    assert env2._sampler is None


@pytest.mark.parametrize("B,A,O,steps", [(1048576, 3, 3, 3), (262144, 8, 16, 2)])
def test_full_size_bit_exact_against_oracle(oracle, B, A, O, steps):
    """BASELINE.json configs[3] / configs[4] at their real sizes, a few steps, every bit."""
    env, p = _env(B=B, A=A, O=O, seed=2, sampling_style='policy') if A == 3 else _env(B=B, A=A, O=O, seed=2)
    oe = oracle.OracleEnv(cpu_params(p), seed=2)
    g = torch.Generator().manual_seed(7)
    # start mid-episode so that the steps include truncations (resets) for a slice of the batch
    sn = (torch.rand(B, generator=g) * 200).floor()
    env._step_num.copy_(sn.cuda()); oe.step_num[...] = sn.numpy()
    for t in range(steps):
        act = torch.stack([(torch.rand(B, A, generator=g) * 2 - 1) * 0.3, torch.rand(B, A, generator=g) - 0.5], 2)
        obs, rew, term, trunc = env.step_fused(act.cuda())
        o_obs, o_rew, o_term, o_trunc = oe.step_fused(act.numpy())
        assert_bits_equal(f"step {t} truncated", trunc.cpu().numpy(), o_trunc)
        assert_bits_equal(f"step {t} terminated", term.cpu().numpy(), o_term)
        assert_bits_equal(f"step {t} rewards", rew.cpu().numpy(), o_rew)
        assert_bits_equal(f"step {t} obs", obs.cpu().numpy(), o_obs)
    assert_bits_equal("states", env.states.cpu().numpy(), oe.states)
    assert_bits_equal("obstacles", env.obstacles.cpu().numpy(), oe.obstacles)
    assert int(o_trunc.sum()) > B // 400                       # the reset path ran at scale


def test_full_size_step_invariants():
    """Size-independent properties at 1 048 576 envs, 60 free-running steps."""
    B = 1048576
    env, p = _env(B=B, seed=3, sampling_style='policy', episode_len=40)
    tmpl = env._tmpl_states.clone()
    pool = action_pool(B, 3, n=4, angle=0.3)
    pool = [a.cuda() for a in pool]
    tot = torch.zeros(3, dtype=torch.int64, device='cuda')
    for t in range(60):
        prev_ob = env.obstacles.clone(); prev_sn = env._step_num.clone(); prev_terminates = env._terminates.clone()
        obs, rew, term, trunc = env.step_fused(pool[t % 4])
        done = term | trunc
        # truncation rule (environment.py:97) and counter reset (:83-84)
        assert torch.equal(trunc, prev_sn + 1 > p['episode_len'] - 1)
        assert torch.equal(env._step_num, torch.where(done, torch.zeros_like(prev_sn), prev_sn + 1))
        # only reset envs change obstacles; reset envs sit on the template with fresh in-box obstacles
        assert torch.equal(env.obstacles[~done], prev_ob[~done])
        if done.any():
            assert torch.equal(env.states[done], tmpl.expand(int(done.sum()), 3, 5))
            ob = env.obstacles[done]
            assert (ob[..., 0] >= 500).all() and (ob[..., 0] < 1000).all()
            assert (ob[..., 1] >= 250).all() and (ob[..., 1] < 500).all()
        # delayed target termination (environment.py:213-219): last step's _terminates terminates now
        assert bool((term | ~prev_terminates).all())
        # observations: distances non-negative, |angles| <= pi, finite rewards, unit headings
        dist_cols, ang_cols = [1, 5, 6, 7, 10, 11], [0, 2, 3, 4, 8, 9]
        assert bool((obs[:, :, dist_cols] >= 0).all()) and bool((obs[:, :, ang_cols].abs() <= 3.1415928).all())
        assert bool(torch.isfinite(rew).all())
        assert bool(((env.states[:, :, 2:4].norm(dim=2) - 1).abs() < 1e-4).all())
        tot[0] += trunc.sum()
    assert env._num_trunc == int(tot[0]) and env._num_trunc > B // 2
