"""oracle/oracle.py -- TEST INFRASTRUCTURE (oracle), not product code.

Python face of the CPU oracle for MARL-nav's batched environment step
(/root/reference/marlnav/environment.py:92-107).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; ``marlnav_b200`` never does.

Three things live here:

* ``OracleEnv``      -- ctypes wrapper over ``libmarlnav_oracle.so``
                        (``marlnav_oracle.c``: plain C, marlnav_trig.h trig, Philox
                        resets).  The CUDA path must match it bit for bit.
* ``TorchPortEnv``   -- the same step restated with the reference's own torch
                        op sequence (cdist / normalize / einsum / vmap-free 2x2
                        rotation), so that it (a) reproduces stock torch-CPU
                        bits including MKL's cos/sin/acos and (b) costs what the
                        reference costs on CPU.  It is the ``"port"`` CPU
                        baseline ``bench.py`` times.
* ``PhiloxTriangleSampler`` / ``PhiloxTemplateSampler`` -- drop-in replacements
                        for the reference's ``env._init_sampler`` (SURVEY.md
                        Appendix D) so the reference, both oracles and the
                        CUDA path reset to identical states.

Pinning status: the reference ships no tests or golden vectors (SURVEY.md
section 4).  The oracles are pinned against outputs of the reference itself,
generated in the build container by ``tests/golden/make_golden.py`` and
committed under ``tests/golden/``.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmarlnav_oracle.so")

# Same field names and order as the reference's namedtuple (utils.py:13-15).
Observations = namedtuple('Observations', ['target_angle', 'target_distance',
    'obstacles_angles', 'obstacles_distances', 'others_angles', 'others_distances'])


# --------------------------------------------------------------------------- C ABI

class MoParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("num_envs", "num_agents", "num_obstacles", "episode_len")] + \
               [(n, ctypes.c_float) for n in (
                   "min_speed", "max_speed", "min_accel", "max_accel",
                   "risk_factor", "distance_factor", "heading_factor", "target_factor",
                   "soft_factor", "bond_factor",
                   "ob_risk_dist", "ag_risk_dist", "ob_coll_dist", "ag_coll_dist",
                   "agents_min_d", "agents_max_d", "max_at_prop_d", "max_angle_diff",
                   "target_radius", "cap_distance", "bond_sharpness", "ideal_dist", "init_dist",
                   "obst_x_range", "obst_x_mean", "obst_y_range", "obst_y_mean")]


class MoReset(ctypes.Structure):
    _fields_ = [("tmpl_states", ctypes.c_void_p), ("tmpl_obstacles", ctypes.c_void_p),
                ("tmpl_target", ctypes.c_void_p),
                ("states_env_stride", ctypes.c_int64), ("obstacles_env_stride", ctypes.c_int64),
                ("target_env_stride", ctypes.c_int64),
                ("alias_first_step", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("seed", ctypes.c_uint64), ("step_counter", ctypes.c_uint64),
                ("env_id_offset", ctypes.c_uint64),
                ("noisy", ctypes.c_int32), ("noise_chol", ctypes.c_float),
                ("noise_mult", ctypes.c_float), ("angle_range", ctypes.c_float)]


def build(force: bool = False) -> str:
    """Compile marlnav_oracle.c with oracle/Makefile (gcc; seconds)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f))
                for f in ("marlnav_oracle.c", "torch_cpu_math.h", "marlnav_trig.h", "Makefile"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_m:
        subprocess.run(["make", "-C", _HERE, "-B", "libmarlnav_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.mo_row_sum.restype = ctypes.c_float
        _lib.mo_abi_version.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


# ------------------------------------------------------------- params / templates

# Geometry constants hard-coded in the reference's Env.__init__ (environment.py:56-68).
GEOMETRY = dict(ob_risk_dist=60., ag_risk_dist=15., ob_coll_dist=50., ag_coll_dist=5.,
                agents_min_d=30., agents_max_d=50., max_at_prop_d=2., max_angle_diff=math.pi / 8,
                target_radius=30., cap_distance=0.1, bond_sharpness=1., ideal_dist=40.,
                init_dist=1200.)


def default_env_params(num_parallel=2, num_agents=3, num_obstacles=3, sampler_num=-1,
                       device='cpu', **over):
    """The dict ``set_env_params`` builds from the reference CLI defaults
    (__main__.py:73-103, utils.py:257-282) -- restated so tests do not need the
    reference package at run time."""
    env = dict(device=device, num_parallel=num_parallel, num_agents=num_agents,
               num_obstacles=num_obstacles, x_bound=1500.0, y_bound=750.0, max_step=1000,
               episode_len=200, min_speed=3., max_speed=10., min_accel=-0.5, max_accel=0.5,
               risk_factor=0., distance_factor=0., heading_factor=500., target_factor=500.,
               soft_factor=500., bond_factor=10., sampler=None, init=None)
    if sampler_num == -1:
        env['init'] = dict(init_method='triangle', ags_cent_x=150., ags_cent_y=375., ags_dist=40.,
                           init_speed=3., tar_pos_x=1350., tar_pos_y=375., noisy_ags=False,
                           ags_std=0.01, angle_range=math.pi / 6, obst_min_x=500., obst_max_x=1000.,
                           obst_min_y=250., obst_max_y=500., num_parallel=num_parallel,
                           num_obs=num_obstacles, device=device)
    env.update(over)
    return env


def make_params(env_params) -> MoParams:
    p = MoParams()
    p.num_envs = env_params['num_parallel']
    p.num_agents = env_params['num_agents']
    p.num_obstacles = env_params['num_obstacles']
    p.episode_len = env_params['episode_len']
    for k in ("min_speed", "max_speed", "min_accel", "max_accel", "risk_factor",
              "distance_factor", "heading_factor", "target_factor", "soft_factor", "bond_factor"):
        setattr(p, k, float(env_params[k]))
    for k, v in GEOMETRY.items():
        setattr(p, k, float(v))
    init = env_params.get('init') or {}
    if 'obst_min_x' in init:   # utils.py:344-347
        p.obst_x_range = init['obst_max_x'] - init['obst_min_x']
        p.obst_y_range = init['obst_max_y'] - init['obst_min_y']
        p.obst_x_mean = 0.5 * (init['obst_min_x'] + init['obst_max_x'])
        p.obst_y_mean = 0.5 * (init['obst_min_y'] + init['obst_max_y'])
    return p


def triangle_template(init) -> np.ndarray:
    """(3,5) float32 agent template of TriangleIntitializer (utils.py:349-368)."""
    t = np.empty((3, 5), np.float32)
    lib().mo_triangle_template(ctypes.c_float(init['ags_dist']), ctypes.c_float(init['ags_cent_x']),
                               ctypes.c_float(init['ags_cent_y']), ctypes.c_float(init['init_speed']),
                               _p(t))
    return t


def ring_template(num_agents, cx=150., cy=375., speed=3., spacing=40.) -> np.ndarray:
    """N-agent reset template used for the scaled scene:
    a ring whose neighbouring agents are ``spacing`` apart, heading (1,0)."""
    r = (spacing / 2.0) / math.sin(math.pi / num_agents)
    t = np.zeros((num_agents, 5), np.float32)
    for i in range(num_agents):
        ang = 2.0 * math.pi * i / num_agents
        t[i] = (cx + r * math.cos(ang), cy + r * math.sin(ang), 1.0, 0.0, speed)
    return t


def philox_obstacles(params: MoParams, seed, step_counter, env_id_offset=0) -> np.ndarray:
    out = np.empty((params.num_envs, params.num_obstacles, 2), np.float32)
    lib().mo_philox_obstacles(ctypes.byref(params), ctypes.c_uint64(seed),
                              ctypes.c_uint64(step_counter), ctypes.c_uint64(env_id_offset), _p(out))
    return out


def philox_obstacles_numpy(params: MoParams, seed, step_counter, env_id_offset=0) -> np.ndarray:
    """Independent numpy restatement of the addressed draw (cross-checks the C one)."""
    B, O = params.num_envs, params.num_obstacles
    npair = (O + 1) // 2
    env = (np.arange(B, dtype=np.uint64) + np.uint64(env_id_offset))
    c0 = np.repeat((env & np.uint64(0xFFFFFFFF)), npair)
    c1 = np.repeat((env >> np.uint64(32)), npair)
    c2 = np.full(B * npair, step_counter & 0xFFFFFFFF, np.uint64)
    c3 = np.tile(np.arange(npair, dtype=np.uint64), B)
    k0 = np.uint64(seed & 0xFFFFFFFF); k1 = np.uint64(((seed >> 32) ^ (step_counter >> 32)) & 0xFFFFFFFF)
    M = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & M, p1 & M, \
                         ((p0 >> np.uint64(32)) ^ c3 ^ k1) & M, p0 & M
        k0 = (k0 + np.uint64(0x9E3779B9)) & M; k1 = (k1 + np.uint64(0xBB67AE85)) & M
    r = np.stack([c0, c1, c2, c3], 1).reshape(B, npair, 2, 2)      # (B, pair, half, xy)
    u = ((r >> np.uint64(8)).astype(np.float32) * np.float32(2.0 ** -24)).reshape(B, 2 * npair, 2)[:, :O]
    x = (np.float32(params.obst_x_range) * (u[..., 0] - np.float32(0.5))).astype(np.float32) \
        + np.float32(params.obst_x_mean)
    y = (np.float32(params.obst_y_range) * (u[..., 1] - np.float32(0.5))).astype(np.float32) \
        + np.float32(params.obst_y_mean)
    return np.stack([x, y], -1).astype(np.float32)


def _obs_fields(obs, A, O):
    R = A - 1
    o = 2
    return Observations(obs[:, :, 0:1], obs[:, :, 1:2], obs[:, :, o:o + O], obs[:, :, o + O:o + 2 * O],
                        obs[:, :, o + 2 * O:o + 2 * O + R], obs[:, :, o + 2 * O + R:o + 2 * O + 2 * R])


def _resolve_init(env_params, B, A, O):
    """-> (mode, tmpl_states, tmpl_obstacles|None, tmpl_target) as float32 numpy."""
    init = env_params['init']
    method = init['init_method']
    if method == 'triangle':
        assert A == 3, "TriangleIntitializer is hard-wired to 3 agents (utils.py:350-368)"
        return 'shared', triangle_template(init), None, \
            np.array([init['tar_pos_x'], init['tar_pos_y']], np.float32)
    if method == 'template':     # extension: explicit (A,5) template, Philox obstacles
        t = np.asarray(init['agent_template'], np.float32).reshape(A, 5)
        return 'shared', t, None, np.array([init['tar_pos_x'], init['tar_pos_y']], np.float32)
    if method == 'mock_init':
        s = np.asarray(init['mock_states'], np.float32).reshape(B, A, 5)
        o = np.asarray(init['mock_obstacles'], np.float32).reshape(B, O, 2)
        t = np.asarray(init['mock_target'], np.float32).reshape(B, 2)
        return 'per_env', s, o, t
    raise ValueError(method)


def noise_constants(init):
    """(cholesky factor, multiplier, angle range) of the noisy agent reset, computed with the same
    float32 torch ops as TriangleIntitializer.__init__ (utils.py:370-373: MultivariateNormal keeps
    scale_tril = cholesky(diag(ags_std, ags_std)))."""
    import torch
    chol = torch.linalg.cholesky(torch.diag(torch.tensor([init['ags_std'], init['ags_std']])))
    return float(chol[0, 0]), float(init['ags_dist']), float(np.float32(init['angle_range']))


def agent_draws(seed, step_counter, env_id_offset, B, A):
    """The addressed draws of the noisy agent reset: standard normals (B,A,2), uniforms (B,A)."""
    normals = np.empty((B, A, 2), np.float32)
    uniforms = np.empty((B, A), np.float32)
    lib().mo_agent_draws(ctypes.c_uint64(seed), ctypes.c_uint64(step_counter), ctypes.c_uint64(env_id_offset),
                         ctypes.c_int64(B), ctypes.c_int(A), _p(normals), _p(uniforms))
    return normals, uniforms


class OracleEnv:
    """C-oracle environment with the reference's Env surface (numpy tensors)."""

    def __init__(self, env_params, seed=0, env_id_offset=0):
        self.params = env_params
        self.p = make_params(env_params)
        B, A, O = self.p.num_envs, self.p.num_agents, self.p.num_obstacles
        self.B, self.A, self.O = B, A, O
        self.S = 2 + 2 * O + 2 * (A - 1)
        self.seed, self.env_id_offset = int(seed), int(env_id_offset)
        self.mode, ts, to, tt = _resolve_init(env_params, B, A, O)
        self.tmpl_states, self.tmpl_obstacles, self.tmpl_target = ts, to, tt
        self.counter = 0
        init = env_params['init']
        self.noisy = bool(init.get('noisy_ags')) and self.mode == 'shared'
        self.noise = noise_constants(init) if self.noisy else (0.0, 0.0, 0.0)
        if self.mode == 'shared':
            self.states = np.broadcast_to(ts, (B, A, 5)).copy()
            if self.noisy:
                rs = self._reset_spec(0)
                lib().mo_noisy_agents(ctypes.byref(rs), ctypes.c_uint64(0), ctypes.c_int64(B), ctypes.c_int(A),
                                      _p(self.states))
            self.obstacles = philox_obstacles(self.p, self.seed, 0, self.env_id_offset)
            self.target = np.broadcast_to(tt, (B, 2)).copy()
        else:
            self.states, self.obstacles, self.target = ts.copy(), to.copy(), tt.copy()
        self._alias = self.mode == 'per_env'
        self.step_num = np.zeros(B, np.float32)
        self.terminates = np.zeros(B, np.uint8)
        self.stats = np.zeros(3, np.uint64)   # trunc, col, tar

    def move_only(self, actions):
        """environment.py:113-137 alone (in place on self.states) -- phase-split parity."""
        actions = np.ascontiguousarray(actions, np.float32).reshape(self.B, self.A, 2)
        lib().mo_move(ctypes.byref(self.p), _p(self.states), _p(actions))

    def observations_fused(self):
        obs = np.empty((self.B, self.A, self.S), np.float32)
        lib().mo_observe(ctypes.byref(self.p), _p(self.states), _p(self.obstacles),
                         _p(self.target), _p(obs))
        return obs

    def observations(self):
        return _obs_fields(self.observations_fused(), self.A, self.O)

    def _reset_spec(self, counter):
        rs = MoReset()
        rs.tmpl_states = _p(self.tmpl_states).value
        rs.tmpl_obstacles = _p(self.tmpl_obstacles).value if self.tmpl_obstacles is not None else None
        rs.tmpl_target = _p(self.tmpl_target).value
        per_env = self.mode == 'per_env'
        rs.states_env_stride = self.A * 5 if per_env else 0
        rs.obstacles_env_stride = self.O * 2 if per_env else 0
        rs.target_env_stride = 2 if per_env else 0
        rs.alias_first_step = 1 if getattr(self, '_alias', False) else 0
        rs.seed, rs.step_counter, rs.env_id_offset = self.seed, counter, self.env_id_offset
        rs.noisy = 1 if self.noisy else 0
        rs.noise_chol, rs.noise_mult, rs.angle_range = self.noise
        return rs

    def step_fused(self, actions, want_pre=False):
        actions = np.ascontiguousarray(actions, np.float32).reshape(self.B, self.A, 2)
        self.counter += 1
        rs = self._reset_spec(self.counter)
        obs = np.empty((self.B, self.A, self.S), np.float32)
        pre = np.empty_like(obs) if want_pre else None
        rew = np.empty(self.B, np.float32)
        term = np.empty(self.B, np.uint8); trunc = np.empty(self.B, np.uint8)
        lib().mo_step(ctypes.byref(self.p), ctypes.byref(rs), _p(self.states), _p(self.obstacles),
                      _p(self.target), _p(self.step_num), _p(self.terminates), _p(actions),
                      _p(obs), _p(rew), _p(term), _p(trunc), _p(self.stats), _p(pre))
        if self._alias:      # template freezes at "state after first move"
            self.tmpl_states = self.states.copy()
            self._alias = False
        out = (obs, rew, term.astype(bool), trunc.astype(bool))
        return out + (pre,) if want_pre else out

    def step(self, actions):
        obs, rew, term, trunc = self.step_fused(actions)
        return _obs_fields(obs, self.A, self.O), rew, term, trunc


# ------------------------------------------------------ init samplers for injection

class PhiloxTemplateSampler:
    """Callable with the reference initializers' contract -- returns a full-batch
    ``(states (B,A,5), obstacles (B,O,2), target (B,1,2))`` on every call
    (utils.py:375-379) -- whose obstacles are the addressed Philox draw of the
    current call counter.  Call 0 is the construction-time
    sample; call k is the one ``Env._reinit`` makes in step k (environment.py:78)."""

    def __init__(self, env_params, agent_template, seed=0, env_id_offset=0, first_counter=0):
        import torch
        self._torch = torch
        self.p = make_params(env_params)
        B = self.p.num_envs
        init = env_params['init']
        t = np.asarray(agent_template, np.float32)
        self.states = torch.from_numpy(np.broadcast_to(t, (B,) + t.shape).copy())
        self.target = torch.tensor([[init['tar_pos_x'], init['tar_pos_y']]],
                                   dtype=torch.float32).unsqueeze(0).repeat(B, 1, 1)
        self.seed, self.offset, self.counter = int(seed), int(env_id_offset), int(first_counter)

    def __call__(self):
        obst = philox_obstacles(self.p, self.seed, self.counter, self.offset)
        self.counter += 1
        # fresh tensors each call, like torch.cat in utils.py:386,398
        return self.states.clone(), self._torch.from_numpy(obst), self.target


def PhiloxTriangleSampler(env_params, seed=0, env_id_offset=0, first_counter=0):
    return PhiloxTemplateSampler(env_params, triangle_template(env_params['init']), seed,
                                 env_id_offset, first_counter)


# ----------------------------------------------------------------- torch-op port

class TorchPortEnv:
    """The reference step restated with torch CPU ops in the reference's call
    sequence (per-agent cdist / normalize / einsum chains, environment.py:139-286),
    so bits and cost follow stock torch.  Written from SURVEY.md section 3.2 and
    Appendix A; cross-checked against the reference in tests/golden/make_golden.py."""

    def __init__(self, env_params, seed=0, env_id_offset=0, num_threads=None):
        import torch
        self.torch = torch
        if num_threads:
            torch.set_num_threads(num_threads)
        self.params = env_params
        self.p = make_params(env_params)
        B, A, O = self.p.num_envs, self.p.num_agents, self.p.num_obstacles
        self.B, self.A, self.O = B, A, O
        mode, ts, to, tt = _resolve_init(env_params, B, A, O)
        self.mode = mode
        if mode == 'shared':
            self._sampler = PhiloxTemplateSampler(env_params, ts, seed, env_id_offset)
            self.states, self.obstacles, self.target = self._sampler()
        else:
            self._tmpl = (torch.from_numpy(ts.copy()), torch.from_numpy(to.copy()),
                          torch.from_numpy(tt.copy()).unsqueeze(1))
            self._sampler = lambda: self._tmpl
            self.states, self.obstacles, self.target = self._tmpl   # aliasing quirk kept
        self.others = torch.tensor([[i for i in range(A) if i != j] for j in range(A)])
        self.step_num = torch.zeros(B)
        self.terminates = torch.zeros(B, dtype=torch.bool)
        self.stats = [0, 0, 0]

    # environment.py:271-286
    def _dist(self, own, oth):
        return self.torch.cdist(own.unsqueeze(1), oth)

    def _angle(self, own, oth, heading):
        t = self.torch
        rel = oth - own.unsqueeze(1)
        unit = t.nn.functional.normalize(rel, dim=2)
        dots = t.clamp(t.einsum('bj,bij->bi', heading, unit), -1 + 1e-8, 1 - 1e-8)
        ortho = unit - t.einsum('bi,bj->bij', dots, heading)
        return t.where(ortho[:, :, 0] > 0, -1., 1.) * t.acos(dots)

    # environment.py:139-180
    def observations(self):
        t, A, O = self.torch, self.A, self.O
        st, ob, tg = self.states, self.obstacles, self.target
        pos = [st[:, i, :2] for i in range(A)]
        hd = [st[:, i, 2:4] for i in range(A)]
        ta = t.stack([self._angle(pos[i], tg, hd[i]) for i in range(A)], dim=1)
        td = t.cat([self._dist(pos[i], tg) for i in range(A)], dim=1)
        oa = t.cat([t.stack([self._angle(pos[i], ob[:, j:j + 1, :], hd[i]) for i in range(A)], dim=1)
                    for j in range(O)], dim=2)
        od = t.cat([t.cat([self._dist(pos[i], ob[:, j:j + 1, :]) for i in range(A)], dim=1)
                    for j in range(O)], dim=2)
        nb = [t.index_select(st, 1, self.others[i])[:, :, :2] for i in range(A)]
        aa = t.stack([self._angle(pos[i], nb[i], hd[i]) for i in range(A)], dim=1)
        ad = t.cat([self._dist(pos[i], nb[i]) for i in range(A)], dim=1)
        cap = GEOMETRY['cap_distance']
        ta = t.where(td < cap, 0., ta)
        oa = t.where(od < cap, 0., oa)
        aa = t.where(ad < cap, 0., aa)
        return Observations(ta, td, oa, od, aa, ad)

    # environment.py:113-137 (same vmap-of-2x2-matmul construct, so cost and bits follow torch)
    def _move(self, actions):
        t, pp = self.torch, self.params

        def turn(vec, ang):
            c, s = t.cos(ang), t.sin(ang)
            return t.matmul(t.stack([t.stack([c, -s]), t.stack([s, c])]), vec)
        th = t.clamp(actions[:, :, 0], min=-math.pi, max=math.pi)
        self.states[:, :, 2:4] = t.vmap(t.vmap(turn))(self.states[:, :, 2:4], th)
        acc = t.clamp(actions[:, :, -1:], min=pp['min_accel'], max=pp['max_accel'])
        v = t.clamp(self.states[:, :, 4:5] + acc, min=pp['min_speed'], max=pp['max_speed'])
        self.states[:, :, 4:5] = v
        self.states[:, :, :2] += self.states[:, :, 2:4] * v

    # environment.py:184-269
    def _reward(self, o):
        t, pp, g = self.torch, self.params, GEOMETRY

        def inside(d, r):
            return t.max(t.where(d < r, 1., 0.), dim=2)[0]
        risks = t.clamp(inside(o.obstacles_distances, g['ob_risk_dist'])
                        + inside(o.others_distances, g['ag_risk_dist']), max=1)
        colls = t.clamp(inside(o.obstacles_distances, g['ob_coll_dist'])
                        + inside(o.others_distances, g['ag_coll_dist']), max=1)
        in_t = t.where(o.target_distance < g['target_radius'], 1., 0.)
        near = t.where(g['agents_min_d'] < o.others_distances, 1., 0.) \
            * t.where(o.others_distances < g['agents_max_d'], 1., 0.)
        dscore = t.div(t.clamp(t.sum(near, dim=2), max=2), 2)
        head = t.where(t.squeeze(t.abs(o.target_angle), dim=2) < g['max_angle_diff'], 1., 0.)
        soft = -1. * t.squeeze(o.target_distance / g['init_dist'], dim=2)
        sd = (o.others_distances - g['ideal_dist']) / g['bond_sharpness']
        bond = t.mean(1. / (1. + sd ** 2), dim=2)
        coll_any = t.max(colls, dim=1)[0]
        all_in = t.min(in_t, dim=1)[0]
        self.stats[2] += int(t.sum(all_in).item())
        self.stats[1] += int(t.sum(coll_any).item())
        terminated = t.logical_or(coll_any > 0, self.terminates)
        self.terminates = t.logical_and(~self.terminates, t.squeeze(all_in) > 0)
        rew = (pp['target_factor'] * all_in.expand(size=(self.B, self.A))
               + pp['heading_factor'] * head + pp['distance_factor'] * dscore
               + pp['soft_factor'] * soft + pp['bond_factor'] * bond - pp['risk_factor'] * risks)
        return t.mean(rew, dim=1), terminated

    def step(self, actions):
        t = self.torch
        self._move(actions)
        self.step_num += t.ones(self.B)
        truncated = self.step_num > self.params['episode_len'] - 1
        self.stats[0] += t.sum(truncated.long()).item()
        rew, terminated = self._reward(self.observations())
        mask = t.where(t.logical_or(truncated, terminated), 1, 0)
        ns, no, nt = self._sampler()

        def blend(old, new):   # environment.py:86-90
            return t.einsum('b,b...->b...', (1 - mask), old) + t.einsum('b,b...->b...', mask, new)
        self.states, self.obstacles, self.target = blend(self.states, ns), blend(self.obstacles, no), \
            blend(self.target, nt)
        self.step_num = blend(self.step_num, t.zeros(self.B))
        return self.observations(), rew, terminated, truncated


# ------------------------------------------------ caller-side rows

def actor_reference(obs, weights, eps):
    """Actor.forward + sample + log_prob exactly as the reference composes them with torch
    (models.py:27-36,113-115), with the normal draws injected: sample = mu + sqrt(var) * eps,
    which is what MultivariateNormal(mu, diag(var)).rsample does with its Cholesky factor."""
    import torch
    from torch.distributions import MultivariateNormal
    x = obs.reshape(-1, obs.shape[-1])
    h = torch.nn.functional.linear(x, weights['fc1.weight'], weights['fc1.bias'])
    mu = torch.tanh(torch.nn.functional.linear(h, weights['fc_mu.weight'], weights['fc_mu.bias']))
    var = torch.nn.functional.softplus(torch.nn.functional.linear(h, weights['fc_std.weight'], weights['fc_std.bias']))
    dist = MultivariateNormal(mu, torch.vmap(torch.diag)(var))
    actions = mu + torch.sqrt(var) * eps
    return actions, dist.log_prob(actions), mu, var


def critic_reference(obs, weights):
    """Critic.forward (models.py:52-56) with the same torch ops: flatten, fc1, relu, fc2."""
    import torch
    x = obs.reshape(obs.shape[0], -1)
    h = torch.relu(torch.nn.functional.linear(x, weights['fc1.weight'], weights['fc1.bias']))
    return torch.nn.functional.linear(h, weights['fc2.weight'], weights['fc2.bias'])


def discounted_returns_reference(rewards, done, gamma):
    """models.py:131-139 restated with the same torch ops (float64 accumulator, torch.where)."""
    import torch
    T, B = rewards.shape
    curr = torch.zeros(B, dtype=float)
    out = torch.empty(T, B, dtype=torch.float64)
    for i in range(T - 1, -1, -1):
        curr = torch.where(done[i], 0., rewards[i] + gamma * curr)
        out[i] = curr
    return out
