"""Replay the committed golden traces (tests/golden/*.npz, produced by the REAL reference in
make_golden.py) through a backend and compare.  Backends: the C oracle (numpy) and the CUDA
Env (torch); both expose step_fused / observations_fused and the state tensors."""
import os

import numpy as np

from helpers import assert_bits_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    meta = {k[5:]: z[k] for k in z.files if k.startswith("meta_")}
    return meta, z


def checksum(a):
    return np.ascontiguousarray(a).view(np.uint32).astype(np.uint64).sum()


def params_for(meta, make_default, make_template, **over):
    B, A, O = int(meta["B"]), int(meta["A"]), int(meta["O"])
    if "template" in meta:
        p = make_template(B, A, O, agent_template=meta["template"].tolist())
    else:
        p = make_default(B, A, O, sampling_style="policy")
    p["episode_len"] = int(meta.get("episode_len", 200))
    for k in ("risk_factor", "distance_factor", "heading_factor", "target_factor", "soft_factor", "bond_factor"):
        if k in meta:
            p[k] = float(meta[k])                  # traces recorded with non-default reward factors
    if int(meta.get("noisy", 0)):
        p["init"]["noisy_ags"] = True            # utils.py:25 switched on
    p.update(over)
    return p


class OracleBackend:
    def __init__(self, orc, params, seed):
        import copy
        p = copy.deepcopy(params); p["device"] = "cpu"
        self.e = orc.OracleEnv(p, seed=seed)

    def step(self, act):
        return self.e.step_fused(np.asarray(act, np.float32))

    def obs(self): return self.e.observations_fused()
    def states(self): return self.e.states
    def obstacles(self): return self.e.obstacles
    def step_num(self): return self.e.step_num
    def terminates(self): return self.e.terminates.astype(bool)
    def stats(self): return tuple(int(v) for v in self.e.stats)
    def set_states(self, s): self.e.states[...] = s
    def set_obstacles(self, o): self.e.obstacles[...] = o

    def load_pre(self, z, k, t):
        e = self.e
        e.states[...] = z["pre_states"][k]; e.obstacles[...] = z["pre_obstacles"][k]
        e.target[...] = z["pre_target"][k].reshape(-1, 2); e.step_num[...] = z["pre_step_num"][k]
        e.terminates[...] = z["pre_terminates"][k].astype(np.uint8)
        e.counter = t                       # step t draws Philox counter t+1 in both


class CudaBackend:
    def __init__(self, mb, params, seed):
        import torch
        self.torch = torch
        self.e = mb.Env(dict(params, seed=seed, device="cuda"))

    def step(self, act):
        t = self.torch
        o, r, te, tr = self.e.step_fused(t.as_tensor(np.asarray(act, np.float32)).cuda())
        return o.cpu().numpy(), r.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy()

    def obs(self): return self.e.observations_fused().cpu().numpy()
    def states(self): return self.e.states.cpu().numpy()
    def obstacles(self): return self.e.obstacles.cpu().numpy()
    def step_num(self): return self.e._step_num.cpu().numpy()
    def terminates(self): return self.e._terminates.cpu().numpy()
    def stats(self): return (self.e._num_trunc, self.e._num_col, self.e._num_tar)
    def set_states(self, s): self.e.states.copy_(self.torch.as_tensor(s))
    def set_obstacles(self, o): self.e.obstacles.copy_(self.torch.as_tensor(o))

    def load_pre(self, z, k, t):
        e, as_t = self.e, self.torch.as_tensor
        e.states.copy_(as_t(z["pre_states"][k])); e.obstacles.copy_(as_t(z["pre_obstacles"][k]))
        e.target.copy_(as_t(z["pre_target"][k])); e._step_num.copy_(as_t(z["pre_step_num"][k]))
        e._terminates_u8.copy_(as_t(z["pre_terminates"][k].astype(np.uint8)))
        e._reset_counter = t


def replay_bit_exact(name, backend, z, tag=""):
    """Free-running replay; everything the golden recorded must match bit for bit."""
    if "init_states" in z.files:
        assert_bits_equal(f"{name} init states", backend.states(), z["init_states"])
        assert_bits_equal(f"{name} init obstacles", backend.obstacles(), z["init_obstacles"])
    if "init_obs" in z.files:
        assert_bits_equal(f"{name} init obs", backend.obs(), z["init_obs"])
    snaps = {int(t): i for i, t in enumerate(z["snap_steps"])} if "snap_steps" in z.files else {}
    for t, act in enumerate(z["actions"]):
        obs, rew, term, trunc = backend.step(act)
        assert_bits_equal(f"{name}{tag} step {t} terminated", term, z["terminated"][t])
        assert_bits_equal(f"{name}{tag} step {t} truncated", trunc, z["truncated"][t])
        assert_bits_equal(f"{name}{tag} step {t} rewards", rew, z["rewards"][t])
        assert checksum(obs) == z["obs_sum"][t], f"{name}{tag} step {t}: observation checksum differs"
        if "states_sum" in z.files:
            assert checksum(backend.states()) == z["states_sum"][t], f"{name}{tag} step {t}: state checksum"
        if t in snaps:
            i = snaps[t]
            assert_bits_equal(f"{name}{tag} step {t} obs", obs, z["snap_obs"][i])
            assert_bits_equal(f"{name}{tag} step {t} states", backend.states(), z["snap_states"][i])
            assert_bits_equal(f"{name}{tag} step {t} obstacles", backend.obstacles(), z["snap_obstacles"][i])
            assert_bits_equal(f"{name}{tag} step {t} step_num", backend.step_num(), z["snap_step_num"][i])
            assert_bits_equal(f"{name}{tag} step {t} _terminates", backend.terminates(), z["snap_terminates"][i])
    assert backend.stats() == tuple(int(v) for v in z["stats"]), f"{name}{tag}: episode stats"


def replay_scenario(name, backend, z):
    """`-rc -sn -1/0/1` traces (reference sampler's recorded actions)."""
    if "init_states" in z.files:
        assert_bits_equal(f"{name} init states", backend.states(), z["init_states"])
        assert_bits_equal(f"{name} init obstacles", backend.obstacles(), z["init_obstacles"])
        assert_bits_equal(f"{name} init obs", backend.obs(), z["init_obs"])
    for t, act in enumerate(z["actions"]):
        obs, rew, term, trunc = backend.step(act)
        assert_bits_equal(f"{name} step {t} rewards", rew, z["rewards"][t])
        assert_bits_equal(f"{name} step {t} terminated", term, z["terminated"][t])
        assert_bits_equal(f"{name} step {t} truncated", trunc, z["truncated"][t])
        assert checksum(obs) == z["obs_sum"][t], f"{name} step {t}: observation checksum differs"
        assert_bits_equal(f"{name} step {t} obs[0,0]", obs[0, 0], z["obs_e0a0"][t])
        if "states_sum" in z.files:
            assert checksum(backend.states()) == z["states_sum"][t], f"{name} step {t}: state checksum"
        if "terminates" in z.files and (t % 10 == 0 or z["terminated"][t].any() or z["terminates"][t].any()):
            assert_bits_equal(f"{name} step {t} _terminates", backend.terminates(), z["terminates"][t])
            assert_bits_equal(f"{name} step {t} step_num", backend.step_num(), z["step_num"][t])
        if f"states_{t}" in z.files:
            assert_bits_equal(f"{name} step {t} states", backend.states(), z[f"states_{t}"])
            assert_bits_equal(f"{name} step {t} obs", obs, z[f"obs_{t}"])
            if f"obstacles_{t}" in z.files:
                assert_bits_equal(f"{name} step {t} obstacles", backend.obstacles(), z[f"obstacles_{t}"])
    assert backend.stats() == tuple(int(v) for v in z["stats"])


def angle_tol(angle):
    """Angles come from acos(dot): a heading that moved by 1 ulp moves dot by a few 6e-8, i.e. the
    angle by that over sin(angle).  1e-5 relative wherever that conditioning allows it."""
    return 1e-5 * np.abs(angle) + 5e-7 / np.maximum(np.abs(np.sin(angle)), 1e-3)


def teacher_forced_vs_stock(name, backend, z, meta, A, O):
    """Identical pre-step states and actions into `backend` and the STOCK reference (MKL trig,
    recorded every 4th step): terminal flags and reset decisions bit-exact, states / distances /
    rewards within 1e-5, angles within 1e-5 up to acos conditioning.  A reward may miss 1e-5 only
    where a heading score (|target_angle| < pi/8, worth heading_factor/A) flipped on an angle that
    straddles the threshold within the trig difference.  Returns (rewards checked, flips)."""
    snaps = {int(t): i for i, t in enumerate(z["snap_steps"])}
    checked = flips = 0
    for k, t in enumerate(range(0, int(meta["steps"]), 4)):
        backend.load_pre(z, k, t)
        obs, rew, term, trunc = backend.step(z["actions"][t])
        i = snaps[t]
        assert_bits_equal(f"{name} step {t} truncated", trunc, z["truncated"][t])
        assert_bits_equal(f"{name} step {t} terminated", term, z["terminated"][t])
        assert_bits_equal(f"{name} step {t} obstacles", backend.obstacles(), z["snap_obstacles"][i])
        assert_bits_equal(f"{name} step {t} step_num", backend.step_num(), z["snap_step_num"][i])
        np.testing.assert_allclose(backend.states(), z["snap_states"][i], rtol=1e-5, atol=1e-6)
        want = z["snap_obs"][i]
        dist_cols = [1] + list(range(2 + O, 2 + 2 * O)) + list(range(2 + 2 * O + A - 1, 2 + 2 * O + 2 * (A - 1)))
        ang_cols = [c for c in range(want.shape[2]) if c not in dist_cols]
        np.testing.assert_allclose(obs[:, :, dist_cols], want[:, :, dist_cols], rtol=1e-5)
        da = np.abs(obs[:, :, ang_cols] - want[:, :, ang_cols])
        assert (da <= angle_tol(want[:, :, ang_cols])).all(), f"{name} step {t}: angle beyond conditioning bound"
        dr = np.abs(rew - z["rewards"][t])
        bad = dr > 1e-5 * np.maximum(np.abs(z["rewards"][t]), 1.0)
        if bad.any():
            # the returned observations are post-reset; an env that did not reset still shows the
            # target angles its reward was computed from
            keep = bad & ~(term | trunc)
            ta = np.abs(want[keep][:, :, 0])
            assert (np.abs(ta - np.float32(np.pi / 8)) < 1e-5).any(axis=1).all(), f"{name} step {t}: reward off"
            # each flip moves the reward by heading_factor / A
            unit = float(meta.get("heading_factor", 500.0)) / A
            assert (np.abs(dr[bad] / unit - np.round(dr[bad] / unit)) < 1e-3).all(), f"{name} step {t}: reward off"
            flips += int(bad.sum())
        checked += rew.size
    return checked, flips
