"""GPU: the CUDA step against golden traces recorded from the REAL reference
(tests/golden/patched_*.npz: reference + Philox sampler + oracle trig, free-running).
Bit-exact on every recorded tensor -- the direct CUDA <-> reference link; no oracle involved."""
import numpy as np
import pytest

import golden_replay as gr

pytestmark = pytest.mark.gpu

NAMES = ["patched_tri_3x3", "patched_tri_3x3_wide", "patched_ring_8x16", "patched_ring_4x2", "patched_ring_9x5"]


def _backend(meta, **over):
    import marlnav_b200 as mb
    p = gr.params_for(meta, mb.default_env_params,
                      lambda B, A, O, agent_template: mb.template_env_params(B, A, O, agent_template), **over)
    return gr.CudaBackend(mb, p, int(meta["seed"]))


@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_reference_trace(name):
    meta, z = gr.load(name)
    gr.replay_bit_exact(name, _backend(meta), z, tag=" [cuda]")


@pytest.mark.parametrize("sn", [0, 1])
def test_cuda_reproduces_reward_check_scenarios(sn):
    """BASELINE.json configs[0] family: `python -m marlnav -rc -sn 0/1`, 1000 steps, B=2."""
    import marlnav_b200 as mb
    meta, z = gr.load(f"patched_rc_sn{sn}")
    be = gr.CudaBackend(mb, mb.default_env_params(sampler_num=sn), 0)
    # the drop-in's own sampler must hand out the reference sampler's actions
    for t in (0, 1, 2):
        assert np.array_equal(be.e.sample_actions().cpu().numpy(), z["actions"][t])
    be = gr.CudaBackend(mb, mb.default_env_params(sampler_num=sn), 0)
    gr.replay_scenario(f"rc_sn{sn} [cuda]", be, z)


def test_cuda_reproduces_termination_quirks():
    meta, z = gr.load("patched_quirks")
    be = _backend(meta)
    be.set_states(z["init_states"]); be.set_obstacles(z["init_obstacles"])
    gr.replay_bit_exact("quirks", be, z, tag=" [cuda]")


def test_cuda_vs_stock_reference_teacher_forced():
    """Stock reference (MKL trig) per-step snapshots: flags / reset decisions bit-exact,
    states and distances within 1e-5 (north-star tolerance)."""
    import torch
    import marlnav_b200 as mb
    meta, z = gr.load("stock_tri_3x3")
    be = _backend(meta)
    e = be.e
    O, A = e.num_obstacles, e.num_agents
    snaps = {int(t): i for i, t in enumerate(z["snap_steps"])}
    for k, t in enumerate(range(0, int(meta["steps"]), 4)):
        e.states.copy_(torch.as_tensor(z["pre_states"][k])); e.obstacles.copy_(torch.as_tensor(z["pre_obstacles"][k]))
        e.target.copy_(torch.as_tensor(z["pre_target"][k])); e._step_num.copy_(torch.as_tensor(z["pre_step_num"][k]))
        e._terminates_u8.copy_(torch.as_tensor(z["pre_terminates"][k].astype(np.uint8)))
        e._reset_counter = t
        obs, rew, term, trunc = be.step(z["actions"][t])
        i = snaps[t]
        assert np.array_equal(term, z["terminated"][t]) and np.array_equal(trunc, z["truncated"][t])
        assert np.array_equal(be.obstacles(), z["snap_obstacles"][i])        # same envs reset, same draws
        assert np.array_equal(be.step_num(), z["snap_step_num"][i])
        np.testing.assert_allclose(be.states(), z["snap_states"][i], rtol=1e-5, atol=1e-6)
        dist_cols = [1] + list(range(2 + O, 2 + 2 * O)) + list(range(2 + 2 * O + A - 1, 2 + 2 * O + 2 * (A - 1)))
        np.testing.assert_allclose(obs[:, :, dist_cols], z["snap_obs"][i][:, :, dist_cols], rtol=1e-5)
