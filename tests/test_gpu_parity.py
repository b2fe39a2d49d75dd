"""GPU parity: the fused CUDA step (through the C ABI and the drop-in Env) against the
CPU oracle on identical seeded inputs.  Bar: EVERYTHING bit-exact -- flags, counters,
reset indices, states, observations, rewards -- free-running over many steps, because
the kernel and oracle/marlnav_oracle.c share one float32 operation order
 and the same SLEEF-u10 trig.  (Oracle vs the stock reference
is pinned separately in test_oracle_golden.py.)"""
import numpy as np
import pytest
import torch

from helpers import action_pool, assert_bits_equal, cpu_params

pytestmark = pytest.mark.gpu


def _mk(params, seed):
    import marlnav_b200 as mb
    p = dict(params, seed=seed)
    return mb.Env(p)


def _compare_state(env, oe, tag):
    assert_bits_equal(f"{tag} states", env.states.cpu().numpy(), oe.states)
    assert_bits_equal(f"{tag} obstacles", env.obstacles.cpu().numpy(), oe.obstacles)
    assert_bits_equal(f"{tag} target", env.target.cpu().numpy().reshape(-1, 2), oe.target)
    assert_bits_equal(f"{tag} step_num", env._step_num.cpu().numpy(), oe.step_num)
    assert_bits_equal(f"{tag} terminates", env._terminates.cpu().numpy(), oe.terminates.astype(bool))


def _run_free(params, oracle, steps, seed=11, angle=0.2, check_every=1):
    B, A = params['num_parallel'], params['num_agents']
    env = _mk(params, seed)
    oe = oracle.OracleEnv(cpu_params(params), seed=seed, env_id_offset=params.get('env_id_offset', 0))
    _compare_state(env, oe, "init")
    assert_bits_equal("init obs", env.observations_fused().cpu().numpy(), oe.observations_fused())
    pool = action_pool(B, A, angle=angle)
    ndone = 0
    for t in range(steps):
        act = pool[t % len(pool)]
        obs, rew, term, trunc = env.step_fused(act.cuda())
        o_obs, o_rew, o_term, o_trunc = oe.step_fused(act.numpy())
        if t % check_every == 0 or t == steps - 1:
            tag = f"step {t}"
            assert_bits_equal(f"{tag} terminated", term.cpu().numpy(), o_term)
            assert_bits_equal(f"{tag} truncated", trunc.cpu().numpy(), o_trunc)
            assert_bits_equal(f"{tag} rewards", rew.cpu().numpy(), o_rew)
            assert_bits_equal(f"{tag} obs", obs.cpu().numpy(), o_obs)
            _compare_state(env, oe, tag)
        ndone += int((o_term | o_trunc).sum())
    assert (env._num_trunc, env._num_col, env._num_tar) == tuple(int(v) for v in oe.stats)
    return ndone


def _run_free_with_geometry(params, oracle, steps, geometry, seed=5):
    """_run_free with the reference's hard-coded geometry constants (environment.py:56-68)
    overridden on both sides: the launch then does not match the compiled-in division profile
    and takes the kernels built with run-time division modes (IEEE / proven / power-of-two)."""
    B, A = params['num_parallel'], params['num_agents']
    env = _mk(params, seed)
    oe = oracle.OracleEnv(cpu_params(params), seed=seed)
    for k, v in geometry.items():
        setattr(env._c_params, k, v)
        setattr(oe.p, k, v)
    pool = action_pool(B, A)
    ndone = 0
    for t in range(steps):
        act = pool[t % len(pool)]
        obs, rew, term, trunc = env.step_fused(act.cuda())
        o_obs, o_rew, o_term, o_trunc = oe.step_fused(act.numpy())
        tag = f"step {t}"
        assert_bits_equal(f"{tag} terminated", term.cpu().numpy(), o_term)
        assert_bits_equal(f"{tag} truncated", trunc.cpu().numpy(), o_trunc)
        assert_bits_equal(f"{tag} rewards", rew.cpu().numpy(), o_rew)
        assert_bits_equal(f"{tag} obs", obs.cpu().numpy(), o_obs)
        ndone += int((o_term | o_trunc).sum())
    _compare_state(env, oe, "final")
    return ndone


@pytest.mark.parametrize("geometry", [
    dict(init_dist=1234.5, max_at_prop_d=3.0, bond_sharpness=2.0),      # proven / proven / power of two
    dict(init_dist=1024.0, max_at_prop_d=1.0, bond_sharpness=0.7, ideal_dist=35.0, target_radius=45.0),
])
def test_custom_geometry_constants(oracle, geometry):
    import marlnav_b200 as mb
    ndone = _run_free_with_geometry(mb.default_env_params(1500, 3, 3, sampling_style='policy'), oracle, 150, geometry)
    assert ndone > 100
    _run_free_with_geometry(mb.template_env_params(300, 8, 16), oracle, 60, geometry)


def test_triangle_3x3_free_running(oracle):
    import marlnav_b200 as mb
    ndone = _run_free(mb.default_env_params(4096, 3, 3, sampling_style='policy'), oracle, steps=320)
    assert ndone > 1000          # the reset path was really exercised


def test_triangle_3x3_free_running_thread_per_env(oracle):
    """Above 32 768 envs the team of 3 runs the thread-per-env kernel (the headline path at 1M envs):
    240 free-running steps with thousands of resets, every tensor compared every 6th step."""
    import marlnav_b200 as mb
    p = mb.default_env_params(40000, 3, 3, sampling_style='policy', episode_len=100)
    assert mb.Env(dict(p, seed=1)).launch_info()[3] == 32
    ndone = _run_free(p, oracle, steps=240, check_every=6)
    assert ndone > 40000


def test_triangle_3x3_wide_turns(oracle):
    import marlnav_b200 as mb
    _run_free(mb.default_env_params(1024, 3, 3, sampling_style='policy'), oracle, steps=64, angle=4.0)


@pytest.mark.parametrize("B", [1, 2, 3, 127, 129, 130, 513])
def test_ragged_batch_sizes(oracle, B):
    import marlnav_b200 as mb
    _run_free(mb.default_env_params(B, 3, 3, sampling_style='policy'), oracle, steps=60)


def test_scaled_scene_8x16(oracle):
    import marlnav_b200 as mb
    ndone = _run_free(mb.template_env_params(1000, 8, 16), oracle, steps=120)
    assert ndone > 50


@pytest.mark.parametrize("A,O", [(2, 1), (2, 2), (4, 2), (5, 3), (3, 2), (3, 4), (3, 5), (3, 6), (3, 7), (9, 5), (16, 7), (26, 64)])
def test_generic_shapes(oracle, A, O):
    import marlnav_b200 as mb
    _run_free(mb.template_env_params(200 if A < 16 else 40, A, O), oracle, steps=40)


@pytest.mark.parametrize("sn", [0, 1])
def test_reward_check_scenarios(oracle, sn):
    """`-rc -sn 0/1`: mock initialiser (aliasing quirk B-6) + scripted sampler, 1000 steps."""
    import marlnav_b200 as mb
    params = mb.default_env_params(sampler_num=sn)
    env = _mk(params, 0)
    oe = oracle.OracleEnv(cpu_params(params), seed=0)
    for t in range(1000):
        act = env.sample_actions()
        obs, rew, term, trunc = env.step_fused(act)
        o_obs, o_rew, o_term, o_trunc = oe.step_fused(act.cpu().numpy())
        assert_bits_equal(f"sn{sn} step {t} rewards", rew.cpu().numpy(), o_rew)
        assert_bits_equal(f"sn{sn} step {t} obs", obs.cpu().numpy(), o_obs)
        assert_bits_equal(f"sn{sn} step {t} term", term.cpu().numpy(), o_term)
        assert_bits_equal(f"sn{sn} step {t} trunc", trunc.cpu().numpy(), o_trunc)
    _compare_state(env, oe, f"sn{sn} final")
    assert (env._num_trunc, env._num_col, env._num_tar) == tuple(int(v) for v in oe.stats)


def test_sharded_equals_single(oracle):
    """Two half-size Envs with env_id_offset reproduce one full-size Env bit for bit."""
    import marlnav_b200 as mb
    full = mb.default_env_params(600, 3, 3, sampling_style='policy')
    e_full = _mk(full, 5)
    shards = [_mk(mb.shard_env_params(full, r, 2), 5) for r in range(2)]
    pool = action_pool(600, 3)
    for t in range(150):
        act = pool[t % len(pool)].cuda()
        o, r, te, tr = e_full.step_fused(act)
        parts = [s.step_fused(act[i * 300:(i + 1) * 300].contiguous()) for i, s in enumerate(shards)]
        assert torch.equal(o, torch.cat([p[0] for p in parts]))
        assert torch.equal(r, torch.cat([p[1] for p in parts]))
        assert torch.equal(te, torch.cat([p[2] for p in parts]))
        assert torch.equal(tr, torch.cat([p[3] for p in parts]))
    assert torch.equal(e_full.obstacles, torch.cat([s.obstacles for s in shards]))
    tot = sum(s.episode_stats for s in shards)
    assert torch.equal(tot, e_full.episode_stats)


@pytest.mark.parametrize("B,A,O", [(777, 3, 3), (40000, 3, 3), (300, 8, 16), (200, 4, 2)])
def test_fused_normalizer_and_scaler(oracle, B, A, O):
    """fuse_io(): ObsNormalizer (utils.py:519-532) and ActionScaler (utils.py:535-547) folded
    into the kernel must equal applying them around the oracle step, bit for bit -- in the NORM builds
    of all four step kernels (thread-per-agent and thread-per-env team of 3, (8,16), generic)."""
    import math
    import marlnav_b200 as mb
    params = mb.default_env_params(B, A, O, sampling_style='policy') if A == 3 else mb.template_env_params(B, A, O)
    env = _mk(params, 9)
    oe = oracle.OracleEnv(cpu_params(params), seed=9)
    max_d = math.sqrt(1500.0 ** 2 + 750.0 ** 2)
    lo = [-math.pi, 0.] + O * [-math.pi] + O * [0.] + (A - 1) * [-math.pi] + (A - 1) * [0.]
    hi = [math.pi, max_d] + O * [math.pi] + O * [max_d] + (A - 1) * [math.pi] + (A - 1) * [max_d]
    norm = dict(min_obs=lo, max_obs=hi)
    scal = dict(min_action=[-math.pi, -0.5], max_action=[math.pi, 0.5])
    env.fuse_io(norm, scal)
    lo_t, hi_t = torch.tensor(lo), torch.tensor(hi)
    o_scale, o_mean = (0.5 * (hi_t - lo_t)).numpy(), (0.5 * (lo_t + hi_t)).numpy()
    a_lo, a_hi = torch.tensor(scal['min_action']), torch.tensor(scal['max_action'])
    a_scale, a_mean = (0.5 * (a_hi - a_lo)).numpy(), (0.5 * (a_lo + a_hi)).numpy()
    g = torch.Generator().manual_seed(3)
    for t in range(80):
        raw = (torch.rand(B, A, 2, generator=g) * 2 - 1) * torch.tensor([0.08, 1.0])
        obs, rew, term, trunc = env.step_fused(raw.cuda())
        scaled = (a_scale * raw.numpy()).astype(np.float32) + a_mean
        o_obs, o_rew, o_term, o_trunc = oe.step_fused(scaled)
        want = ((o_obs - o_mean).astype(np.float32) / o_scale).astype(np.float32)
        assert_bits_equal(f"step {t} normalised obs", obs.cpu().numpy(), want)
        assert_bits_equal(f"step {t} rewards", rew.cpu().numpy(), o_rew)
        assert_bits_equal(f"step {t} term", term.cpu().numpy(), o_term)
    assert_bits_equal("states", env.states.cpu().numpy(), oe.states)


@pytest.mark.parametrize("B,A,O", [(500, 3, 3), (40000, 3, 3), (300, 8, 16), (200, 5, 3)])
def test_observations_after_a_step_equal_the_steps_observations(oracle, B, A, O):
    """Env.observations() (environment.py:139-180, its own kernel) on the state a step left behind ==
    the (post-reset) observations that step returned, and == the oracle's."""
    import marlnav_b200 as mb
    p = mb.default_env_params(B, A, O, sampling_style='policy', episode_len=12) if A == 3 else mb.template_env_params(B, A, O, episode_len=12)
    env = _mk(p, 2)
    oe = oracle.OracleEnv(cpu_params(p), seed=2)
    pool = action_pool(B, A)
    for t in range(30):
        obs, _, term, trunc = env.step_fused(pool[t % 8].cuda())
        oe.step_fused(pool[t % 8].numpy())
        if t % 6 == 5:
            again = env.observations_fused()
            assert torch.equal(again, obs), t
            assert_bits_equal(f"step {t} observations()", again.cpu().numpy(), oe.observations_fused())
            fields = env.observations()
            assert torch.equal(torch.cat(list(fields), dim=2), again)


@pytest.mark.parametrize("B", [2048, 70001, 140001])
def test_host_stepper_matches_device_step(oracle, B):
    """marlnav_step_host_f32 (pinned host in/out) == the device-resident step; 70 001 envs go through the
    chunked pipeline (two chunks on three streams, the second with its env-id offset and a ragged tail),
    140 001 through the ramped one (chunks of 8 704, 17 536, 3 x 35 072 and a ragged 8 545 envs)."""
    import marlnav_b200 as mb
    params = mb.default_env_params(B, 3, 3, sampling_style='policy', episode_len=15)
    e1, e2 = _mk(params, 4), _mk(params, 4)
    hs = mb.HostStepper(e2)
    pool = action_pool(B, 3)
    for t in range(40):
        act = pool[t % len(pool)]
        obs, rew, term, trunc = e1.step_fused(act.cuda())
        hs.actions_host.copy_(act)
        hs.step()
        assert torch.equal(obs.cpu(), hs.obs_host)
        assert torch.equal(rew.cpu(), hs.rewards_host)
        assert torch.equal(term.cpu(), hs.terminated_host.bool())
        assert torch.equal(trunc.cpu(), hs.truncated_host.bool())
    assert torch.equal(e1.states, e2.states)


# ----------------------------------------------------------------------------- round 2 additions

@pytest.mark.parametrize("make", ["tri_small", "tri_large", "ring_8x16", "ring_4x2"])
def test_noisy_agents_free_running(oracle, make):
    """noisy_ags = True (utils.py:25,381-388): Gaussian position noise + heading
    rotation on every (re-)initialised agent, in all three step kernels and the init kernel --
    thread-per-agent team of 3 (small batch), thread-per-env (large batch), (8,16) team, generic."""
    import marlnav_b200 as mb
    if make == "tri_small":
        p = mb.default_env_params(777, 3, 3, sampling_style='policy', episode_len=40)
    elif make == "tri_large":
        p = mb.default_env_params(40000, 3, 3, sampling_style='policy', episode_len=40)
    elif make == "ring_8x16":
        p = mb.template_env_params(301, 8, 16, episode_len=40)
    else:
        p = mb.template_env_params(200, 4, 2, episode_len=40)
    p['init']['noisy_ags'] = True
    ndone = _run_free(p, oracle, steps=100, check_every=1 if make != "tri_large" else 7)
    assert ndone > p['num_parallel']                # every env was re-initialised at least once


def test_sharded_equals_single_8_shards_and_noisy(oracle):
    """BASELINE configs[3]'s split: eight slices with env_id_offset reproduce one full-size Env
    bit for bit -- also with the noisy agent reset, whose draws are addressed by global env id."""
    import marlnav_b200 as mb
    full = mb.default_env_params(1000, 3, 3, sampling_style='policy', episode_len=50)
    full['init']['noisy_ags'] = True
    e_full = _mk(full, 5)
    shards = [_mk(mb.shard_env_params(full, r, 8), 5) for r in range(8)]
    bounds = [mb.shard_bounds(1000, r, 8) for r in range(8)]
    assert torch.equal(e_full.states, torch.cat([s.states for s in shards]))
    pool = action_pool(1000, 3)
    for t in range(120):
        act = pool[t % len(pool)].cuda()
        o, r, te, tr = e_full.step_fused(act)
        parts = [s.step_fused(act[lo:lo + n].contiguous()) for s, (lo, n) in zip(shards, bounds)]
        for k, full_t in enumerate((o, r, te, tr)):
            assert torch.equal(full_t, torch.cat([p[k] for p in parts])), (t, k)
    assert torch.equal(e_full.states, torch.cat([s.states for s in shards]))
    assert torch.equal(sum(s.episode_stats for s in shards), e_full.episode_stats)
    assert int(e_full.episode_stats.sum()) > 1000


@pytest.mark.parametrize("A,O,B", [(3, 3, 4096), (3, 3, 100000), (8, 16, 1024), (4, 2, 512)])
def test_misaligned_base_pointers(oracle, A, O, B):
    """Tensors whose base pointers are only 4-byte aligned (vec_ok == 0): the kernels must take their
    plain-load staging paths instead of TMA bulk copies / float4 accesses, with the same results."""
    import marlnav_b200 as mb
    p = mb.default_env_params(B, A, O, sampling_style='policy') if A == 3 else mb.template_env_params(B, A, O)
    env = _mk(p, 7)
    oe = oracle.OracleEnv(cpu_params(p), seed=7)

    def shifted(t):          # same values, storage offset by one float
        flat = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)[1:]
        flat.copy_(t.reshape(-1))
        return flat.view(t.shape)
    env.states, env.obstacles, env.target = shifted(env.states), shifted(env.obstacles), shifted(env.target)
    assert env.states.data_ptr() % 16 == 4 and env.states.is_contiguous()
    S = env.obs_size
    obs = torch.empty(B * A * S + 1, device='cuda')[1:].view(B, A, S)
    rew = torch.empty(B, device='cuda')
    term = torch.empty(B, dtype=torch.uint8, device='cuda'); trunc = torch.empty_like(term)
    pool = action_pool(B, A)
    for t in range(40):
        act = torch.empty(B * A * 2 + 1, device='cuda')[1:].view(B, A, 2)
        act.copy_(pool[t % len(pool)])
        env.step_fused(act, out=(obs, rew, term, trunc))
        o_obs, o_rew, o_term, o_trunc = oe.step_fused(pool[t % len(pool)].numpy())
        if t % 5 == 0 or t == 39:
            assert_bits_equal(f"step {t} obs", obs.cpu().numpy(), o_obs)
            assert_bits_equal(f"step {t} rewards", rew.cpu().numpy(), o_rew)
            assert_bits_equal(f"step {t} term", term.cpu().numpy().astype(bool), o_term)
    _compare_state(env, oe, "final")
    assert_bits_equal("observe", env.observations_fused().cpu().numpy(), oe.observations_fused())


def test_step_outputs_are_fresh_until_released(oracle):
    """Env.step hands out tensors from reusable output slots: a slot is only recycled when the caller
    holds none of its tensors (models.py:121 keeps every step's rewards by reference)."""
    import marlnav_b200 as mb
    p = mb.default_env_params(64, 3, 3, sampling_style='policy')
    env = _mk(p, 3)
    oe = oracle.OracleEnv(cpu_params(p), seed=3)
    pool = action_pool(64, 3)
    kept, want = [], []
    for t in range(30):                                  # MAPPO.get_data: keeps rewards, done; drops obs
        obs, rew, term, trunc = env.step(pool[t % 8].cuda())
        o_obs, o_rew, o_term, o_trunc = oe.step_fused(pool[t % 8].numpy())
        kept.append((rew, obs.others_distances[:, 0]))   # a tensor, and a view of a view
        want.append((o_rew, o_obs[:, 0, 10:12]))
    torch.cuda.synchronize()
    for (rew, od), (o_rew, o_od) in zip(kept, want):
        assert_bits_equal("kept rewards", rew.cpu().numpy(), o_rew)
        assert_bits_equal("kept view", od.cpu().numpy(), o_od)
    assert len({r.data_ptr() for r, _ in kept}) == 30   # thirty different buffers
    n_slots = len(env._ring)
    del kept, rew, obs, term, trunc, od
    ptrs = set()
    for t in range(200):                                 # nothing kept: the ring stops growing
        obs, rew, term, trunc = env.step(pool[t % 8].cuda())
        ptrs.add(rew.data_ptr())
        del obs, rew, term, trunc
    assert len(env._ring) <= n_slots + 1 and len(ptrs) <= n_slots + 1


@pytest.mark.parametrize("A,O,B", [(3, 3, 3000), (3, 3, 40000), (8, 16, 300)])
def test_zero_clamp_bounds(oracle, A, O, B):
    """min_accel = 0 / min_speed = 0: torch.clamp's `(a < b) ? b : a` keeps a -0 against a +0 bound (a
    two-FMNMX clamp would not; measured no faster and not used)."""
    import marlnav_b200 as mb
    p = mb.default_env_params(B, A, O, sampling_style='policy') if A == 3 else mb.template_env_params(B, A, O)
    p.update(min_accel=0.0, min_speed=0.0)
    ndone = _run_free(p, oracle, steps=80, check_every=4)
    assert ndone > 0


@pytest.mark.parametrize("seed", range(8))
def test_randomised_parameters(oracle, seed):
    """Every reward factor non-zero (the defaults multiply the risk and distance scores by 0, which would
    hide an error in them), random speed / acceleration limits, episode lengths, obstacle boxes and
    geometry thresholds, across all four step kernels: thread-per-env and thread-per-agent team of 3,
    the (8,16) team kernel and the generic kernel."""
    import marlnav_b200 as mb
    rng = np.random.default_rng(100 + seed)
    A, O, B = [(3, 3, 40000), (3, 3, 900), (8, 16, 320), (3, 5, 35000), (5, 4, 300), (3, 1, 1500), (8, 16, 64), (2, 2, 257)][seed]
    p = mb.default_env_params(B, A, O, sampling_style='policy') if A == 3 else mb.template_env_params(B, A, O)
    p.update(risk_factor=float(rng.uniform(1, 300)), distance_factor=float(rng.uniform(1, 300)),
             heading_factor=float(rng.uniform(1, 800)), target_factor=float(rng.uniform(1, 800)),
             soft_factor=float(rng.uniform(1, 800)), bond_factor=float(rng.uniform(0.1, 50)),
             min_speed=float(rng.uniform(0.5, 4)), max_speed=float(rng.uniform(6, 14)),
             min_accel=float(-rng.uniform(0.1, 1)), max_accel=float(rng.uniform(0.1, 1)),
             episode_len=int(rng.integers(8, 70)))
    p['init'].update(obst_min_x=float(rng.uniform(250, 500)), obst_max_x=float(rng.uniform(700, 1200)),
                     obst_min_y=float(rng.uniform(100, 300)), obst_max_y=float(rng.uniform(450, 700)))
    geometry = dict(ob_risk_dist=float(rng.uniform(55, 90)), ag_risk_dist=float(rng.uniform(10, 35)),
                    ob_coll_dist=float(rng.uniform(30, 55)), ag_coll_dist=float(rng.uniform(2, 9)),
                    agents_min_d=float(rng.uniform(20, 35)), agents_max_d=float(rng.uniform(42, 60)),
                    target_radius=float(rng.uniform(20, 60)), ideal_dist=float(rng.uniform(30, 50)))
    if seed % 2:                                     # every other case also leaves the compiled-in division profile
        geometry.update(init_dist=float(rng.uniform(800, 1600)), max_at_prop_d=float(rng.integers(1, A)),
                        bond_sharpness=float(rng.uniform(0.5, 4)))
    ndone = _run_free_with_geometry(p, oracle, 90, geometry, seed=seed)
    assert ndone > 0


@pytest.mark.parametrize("A,O,B", [(3, 3, 600), (3, 3, 36000), (8, 16, 200), (4, 2, 300)])
def test_template_with_negative_values_and_negative_zero(oracle, A, O, B):
    """An agent template with negative coordinates and a -0.0 heading component: the "+0 wash" shortcut
    of the default reset source does not apply (MARLNAV_RESET_TMPL_NONNEG is not set), so the kernels
    must evaluate the literal blend (1-m)*old + m*new for every env -- including the sign of zeros."""
    import marlnav_b200 as mb
    tmpl = mb.ring_template(A, cx=-30.0, cy=375.0)
    for i, row in enumerate(tmpl):
        row[2], row[3] = (1.0, -0.0) if i % 2 else (1.0, 0.0)
    p = mb.template_env_params(B, A, O, agent_template=tmpl, episode_len=20)
    env = _mk(p, 3)
    assert not env._tmpl_nonneg
    ndone = _run_free(p, oracle, steps=70, seed=3, check_every=3)
    assert ndone > B
