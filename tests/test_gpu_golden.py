"""GPU: the CUDA step against golden traces recorded from the REAL reference
(tests/golden/patched_*.npz: reference + Philox sampler + oracle trig, free-running).
Bit-exact on every recorded tensor -- the direct CUDA <-> reference link; no oracle involved."""
import numpy as np
import pytest

import golden_replay as gr

pytestmark = pytest.mark.gpu

NAMES = ["patched_tri_3x3", "patched_tri_3x3_wide", "patched_ring_8x16", "patched_ring_4x2", "patched_ring_9x5",
         "patched_noisy_3x3", "patched_tri_3x3_factors", "patched_ring_8x16_factors"]


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


def test_cuda_reproduces_constant_sampler_scenario():
    """BASELINE.json configs[0] as written: `python -m marlnav -rc -sn -1 -se 0` (triangle initialiser,
    ConstantSampler [0, 1], B = 2, 1000 steps), reference-generated, bit for bit -- including the
    target reaches this scenario alone produces in free running (stats[2] > 0)."""
    import marlnav_b200 as mb
    meta, z = gr.load("patched_rc_snm1")
    be = gr.CudaBackend(mb, mb.default_env_params(sampler_num=-1), int(meta["seed"]))
    for t in (0, 1, 2):      # the drop-in's ConstantSampler hands out the reference sampler's actions
        assert np.array_equal(be.e.sample_actions().cpu().numpy(), z["actions"][t])
    gr.replay_scenario("rc_snm1 [cuda]", be, z)
    assert be.stats()[2] > 0 and be.stats()[1] > 0


def test_cuda_vs_stock_constant_sampler_scenario():
    """The same scenario against the STOCK reference (its own acos), free-running for 1000 steps:
    every terminal flag, reset draw and episode counter identical, rewards within 1e-5."""
    import marlnav_b200 as mb
    meta, z = gr.load("stock_rc_snm1")
    be = gr.CudaBackend(mb, mb.default_env_params(sampler_num=-1), int(meta["seed"]))
    for t, act in enumerate(z["actions"]):
        obs, rew, term, trunc = be.step(act)
        assert np.array_equal(term, z["terminated"][t]) and np.array_equal(trunc, z["truncated"][t]), t
        np.testing.assert_allclose(rew, z["rewards"][t], rtol=1e-5)
        np.testing.assert_allclose(obs[0, 0], z["obs_e0a0"][t], rtol=1e-5, atol=1e-6)
        if f"obstacles_{t}" in z.files:
            assert np.array_equal(be.obstacles(), z[f"obstacles_{t}"])
    assert be.stats() == tuple(int(v) for v in z["stats"])


@pytest.mark.parametrize("name", ["stock_tri_3x3", "stock_ring_8x16", "stock_noisy_3x3", "stock_tri_3x3_factors"])
def test_cuda_vs_stock_reference_teacher_forced(name):
    """Stock reference (MKL trig) per-step snapshots replayed on the GPU: flags / reset decisions
    bit-exact; states, distances and rewards within 1e-5 (north-star tolerance); angles within 1e-5
    up to acos conditioning; at most 2 heading-score flips on pi/8 straddles."""
    meta, z = gr.load(name)
    be = _backend(meta)
    checked, flips = gr.teacher_forced_vs_stock(name + " [cuda]", be, z, meta, be.e.num_agents, be.e.num_obstacles)
    assert checked >= 80 and flips <= 2


def test_reference_traces_through_the_thread_per_env_kernel():
    """The golden traces are small batches, which the library would run thread-per-agent: replay them in
    a child process with MARLNAV_TEAM3_MAX_ENVS=0, i.e. through the thread-per-env kernel that steps
    the 1M-env batches (the variable is read once per process)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, MARLNAV_TEAM3_MAX_ENVS="0")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_golden.py"), "-q", "-x", "-m", "gpu",
                          "-k", "reproduces or teacher_forced or constant_sampler", "-p", "no:cacheprovider"],
                         env=env, capture_output=True, text=True, cwd=os.path.dirname(here))
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert " passed" in res.stdout and "failed" not in res.stdout
