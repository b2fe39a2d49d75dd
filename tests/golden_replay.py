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
    """`-rc -sn 0/1` traces (reference sampler's recorded actions)."""
    for t, act in enumerate(z["actions"]):
        obs, rew, term, trunc = backend.step(act)
        assert_bits_equal(f"{name} step {t} rewards", rew, z["rewards"][t])
        assert_bits_equal(f"{name} step {t} terminated", term, z["terminated"][t])
        assert_bits_equal(f"{name} step {t} truncated", trunc, z["truncated"][t])
        assert checksum(obs) == z["obs_sum"][t], f"{name} step {t}: observation checksum differs"
        assert_bits_equal(f"{name} step {t} obs[0,0]", obs[0, 0], z["obs_e0a0"][t])
        if f"states_{t}" in z.files:
            assert_bits_equal(f"{name} step {t} states", backend.states(), z[f"states_{t}"])
            assert_bits_equal(f"{name} step {t} obs", obs, z[f"obs_{t}"])
    assert backend.stats() == tuple(int(v) for v in z["stats"])
