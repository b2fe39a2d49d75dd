"""CPU: pin the oracles against golden vectors produced by the REAL reference
(tests/golden/make_golden.py; the reference itself never travels to the GPU box).

  * C oracle  == reference with {Philox sampler, oracle trig} injected   -> bit-exact, free-running
  * torch port == stock reference (Philox sampler injected)              -> bit-exact on hosts whose
                  torch CPU trig matches the generating host (checked by known answers)
  * C oracle  ~= stock reference, teacher-forced per step                -> flags exact, floats 1e-5
"""
import copy

import numpy as np
import pytest
import torch

import golden_replay as gr
from helpers import assert_bits_equal

PATCHED_RANDOM = ["patched_tri_3x3", "patched_tri_3x3_wide", "patched_ring_8x16", "patched_ring_4x2",
                  "patched_ring_9x5", "patched_noisy_3x3", "patched_tri_3x3_factors", "patched_ring_8x16_factors"]


def _oracle_backend(oracle, meta, **over):
    p = gr.params_for(meta, lambda B, A, O, **k: oracle.default_env_params(B, A, O),
                      lambda B, A, O, agent_template: _tmpl(oracle, B, A, O, agent_template), **over)
    return gr.OracleBackend(oracle, p, int(meta["seed"]))


def _tmpl(oracle, B, A, O, agent_template):
    p = oracle.default_env_params(B, A, O)
    p["init"] = dict(p["init"], init_method="template", agent_template=agent_template)
    return p


@pytest.mark.parametrize("name", PATCHED_RANDOM)
def test_c_oracle_reproduces_reference_bit_exact(oracle, name):
    meta, z = gr.load(name)
    gr.replay_bit_exact(name, _oracle_backend(oracle, meta), z)


@pytest.mark.parametrize("sn", [0, 1])
def test_c_oracle_reproduces_reward_check_scenarios(oracle, sn):
    """`python -m marlnav -rc -sn 0/1` (1000 steps, mock initialiser aliasing quirk B-6)."""
    import marlnav_b200 as mb
    meta, z = gr.load(f"patched_rc_sn{sn}")
    p = mb.default_env_params(sampler_num=sn, device="cpu")
    gr.replay_scenario(f"rc_sn{sn}", gr.OracleBackend(oracle, p, 0), z)


def test_c_oracle_reproduces_termination_quirks(oracle):
    """delayed target termination (+ double counted _num_tar),
    collision+target re-termination of the fresh episode, truncation on the episode_len-th step."""
    meta, z = gr.load("patched_quirks")
    be = _oracle_backend(oracle, meta)
    be.set_states(z["init_states"]); be.set_obstacles(z["init_obstacles"])
    gr.replay_bit_exact("quirks", be, z)
    term, trunc = z["terminated"], z["truncated"]
    assert term[:4, 0].tolist() == [False, True, False, False]      # B-1: one step late
    assert term[:4, 1].tolist() == [True, True, False, False]       # B-2: fresh episode re-terminated
    assert trunc[:, 2].nonzero()[0].tolist() == [4, 9]              # B-3: every episode_len-th step
    assert tuple(z["stats"]) == (8, 1, 3)


def _torch_trig_matches_generating_host():
    x = torch.tensor([0.1234567, -2.7182817, 3.0, 0.0165045], dtype=torch.float32)
    want_cos = np.array([0x3f7e0d33, 0xbf696764, 0xbf7d7026, 0x3f7ff713], np.uint32)
    y = torch.tensor([0.99986, 0.3, -0.7], dtype=torch.float32)
    want_acos = np.array([0x3c8915e4, 0x3fa20faf, 0x4016280a], np.uint32)
    return (np.array_equal(torch.cos(x).numpy().view(np.uint32), want_cos)
            and np.array_equal(torch.acos(y).numpy().view(np.uint32), want_acos))


def _port_env(oracle, meta):
    B, A, O = int(meta["B"]), int(meta["A"]), int(meta["O"])
    p = oracle.default_env_params(B, A, O) if "template" not in meta else _tmpl(oracle, B, A, O, meta["template"].tolist())
    return oracle.TorchPortEnv(p, seed=int(meta["seed"]), num_threads=4)


@pytest.mark.parametrize("name", ["stock_tri_3x3", "stock_ring_8x16"])
def test_torch_port_reproduces_stock_reference(oracle, name):
    """The CPU-baseline port executes the reference's torch op sequence: same bits as the stock
    reference wherever torch's CPU libm is the one the goldens were generated with."""
    meta, z = gr.load(name)
    env = _port_env(oracle, meta)
    exact = _torch_trig_matches_generating_host()
    for t, act in enumerate(z["actions"]):
        obs, rew, term, trunc = env.step(torch.from_numpy(act.copy()))
        if not exact:
            continue
        assert_bits_equal(f"{name} step {t} rewards", rew.numpy(), z["rewards"][t])
        assert_bits_equal(f"{name} step {t} terminated", term.numpy(), z["terminated"][t])
        assert gr.checksum(torch.cat(list(obs), dim=2).numpy()) == z["obs_sum"][t]
    if exact:
        assert tuple(env.stats) == tuple(int(v) for v in z["stats"])
    else:
        pytest.skip("host torch CPU trig differs from the golden-generating host; ran for crashes only")


@pytest.mark.parametrize("name", ["stock_tri_3x3", "stock_ring_8x16", "stock_noisy_3x3", "stock_tri_3x3_factors"])
def test_c_oracle_vs_stock_reference_teacher_forced(oracle, name):
    """Identical pre-step states and actions into the oracle and the STOCK reference (MKL trig):
    terminal flags and reset decisions bit-exact, states/distances/rewards within 1e-5,
    angles within 1e-5 up to acos conditioning (golden_replay.teacher_forced_vs_stock)."""
    meta, z = gr.load(name)
    be = _oracle_backend(oracle, meta)
    checked, flips = gr.teacher_forced_vs_stock(name, be, z, meta, be.e.A, be.e.O)
    assert checked >= 80 and flips <= 2


def test_c_oracle_reproduces_constant_sampler_scenario(oracle):
    """BASELINE.json configs[0] as written: `python -m marlnav -rc -sn -1 -se 0` (triangle
    initialiser + ConstantSampler, B = 2, 1000 steps; utils.py:217-243, 477-485).  The one
    free-running scenario with target reaches (delayed termination, double-counted _num_tar)."""
    import marlnav_b200 as mb
    meta, z = gr.load("patched_rc_snm1")
    p = mb.default_env_params(sampler_num=-1, device="cpu")
    gr.replay_scenario("rc_snm1", gr.OracleBackend(oracle, p, int(meta["seed"])), z)
    assert int(z["stats"][2]) > 0 and int(z["stats"][1]) > 0
    # a reach is terminated one step late and counted on both steps (Appendix B-1)
    reach = np.flatnonzero(np.diff(z["num_tar"], prepend=0))
    assert len(reach) >= 2 and (np.diff(reach)[::2] == 1).all()


def test_c_oracle_vs_stock_constant_sampler_scenario(oracle):
    """The same scenario on the STOCK reference (its own acos): free-running for 1000 steps the
    oracle keeps every terminal flag and reset draw and stays within 1e-5 on the rewards."""
    import marlnav_b200 as mb
    meta, z = gr.load("stock_rc_snm1")
    be = gr.OracleBackend(oracle, mb.default_env_params(sampler_num=-1, device="cpu"), int(meta["seed"]))
    for t, act in enumerate(z["actions"]):
        obs, rew, term, trunc = be.step(act)
        assert_bits_equal(f"step {t} terminated", term, z["terminated"][t])
        assert_bits_equal(f"step {t} truncated", trunc, z["truncated"][t])
        np.testing.assert_allclose(rew, z["rewards"][t], rtol=1e-5)
        np.testing.assert_allclose(obs[0, 0], z["obs_e0a0"][t], rtol=1e-5, atol=1e-6)
        if f"obstacles_{t}" in z.files:
            assert_bits_equal(f"step {t} obstacles", be.obstacles(), z[f"obstacles_{t}"])
    assert be.stats() == tuple(int(v) for v in z["stats"])


def test_c_oracle_move_phase_vs_stock_reference(oracle):
    """Phase A of the split protocol: post-move states (cos/sin) within 1e-5 of stock torch,
    here against torch ops directly (the reference's rotation is torch.cos/sin + 2x2 matmul)."""
    B, A = 4096, 3
    p = oracle.default_env_params(B, A, 3)
    e = oracle.OracleEnv(p, seed=1)
    g = torch.Generator().manual_seed(5)
    e.states[:, :, 2:4] = torch.nn.functional.normalize(torch.randn(B, A, 2, generator=g), dim=2).numpy()
    act = torch.stack([(torch.rand(B, A, generator=g) * 2 - 1) * 3.2, torch.rand(B, A, generator=g) - 0.5], 2)
    st = torch.from_numpy(e.states.copy())
    th = torch.clamp(act[:, :, 0], -np.pi, np.pi)
    c, s = torch.cos(th), torch.sin(th)
    dx, dy = st[:, :, 2], st[:, :, 3]
    ndx, ndy = c * dx + (-s) * dy, s * dx + c * dy
    v = torch.clamp(st[:, :, 4] + torch.clamp(act[:, :, 1], -0.5, 0.5), 3., 10.)
    want = torch.stack([st[:, :, 0] + ndx * v, st[:, :, 1] + ndy * v, ndx, ndy, v], dim=2).numpy()
    e.move_only(act.numpy())
    np.testing.assert_allclose(e.states, want, rtol=1e-5, atol=2e-7)


def test_c_oracle_observe_vs_torch_port_identical_states(oracle):
    """Phase B: identical (post-move) states into both -> distances bit-exact, angles within
    2 ulp (acos backends), no flag can differ."""
    B, A, O = 3000, 3, 3
    p = oracle.default_env_params(B, A, O)
    e = oracle.OracleEnv(p, seed=2)
    port = oracle.TorchPortEnv(copy.deepcopy(p), seed=2, num_threads=4)
    g = torch.Generator().manual_seed(6)
    st = torch.from_numpy(e.states.copy())
    st[:, :, :2] += torch.rand(B, A, 2, generator=g) * 900
    st[:, :, 2:4] = torch.nn.functional.normalize(torch.randn(B, A, 2, generator=g), dim=2)
    e.states[...] = st.numpy(); port.states = st.clone()
    port.obstacles = torch.from_numpy(e.obstacles.copy()); port.target = torch.from_numpy(e.target.copy()).unsqueeze(1)
    got = e.observations_fused()
    want = torch.cat(list(port.observations()), dim=2).numpy()
    dist_cols = [1, 5, 6, 7, 10, 11]
    assert_bits_equal("distances", got[:, :, dist_cols], np.ascontiguousarray(want[:, :, dist_cols]))
    ang_cols = [0, 2, 3, 4, 8, 9]
    ga, wa = got[:, :, ang_cols], want[:, :, ang_cols]
    ulps = np.abs(ga.view(np.int32).astype(np.int64) - np.ascontiguousarray(wa).view(np.int32).astype(np.int64))
    assert ulps.max() <= 2
