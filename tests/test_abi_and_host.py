"""CPU tests: the C-ABI library loads and exports every declared symbol (no compute calls),
host-side parameter/sampler/sharding logic, world_size-2 gloo reduction of episode stats."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from marlnav_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "marlnav_b200.h")).read()
    declared = set(re.findall(r"\b(marlnav_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.marlnav_abi_version() == 4
    assert lib.marlnav_obs_size(3, 3) == 12 and lib.marlnav_obs_size(8, 16) == 48
    assert lib.marlnav_obs_size(1, 3) == 0 and lib.marlnav_obs_size(3, 0) == 0
    assert lib.marlnav_obs_size(27, 3) == 0


def test_struct_layouts_match_header():
    from marlnav_b200 import _lib
    assert ctypes.sizeof(_lib.EnvParams) == 4 + 4 * 4 + 27 * 4
    assert ctypes.sizeof(_lib.ResetSpec) == 8 + 3 * 8 + 3 * 8 + 4 + 3 * 4 + 3 * 8 + 8
    assert ctypes.sizeof(_lib.IoTransform) == 8 + 4 * 8
    assert ctypes.sizeof(_lib.ActorSpec) == 8 + 6 * 8 + 2 * 4 + 2 * 8 + 8 + 8
    from oracle import oracle as orc
    assert ctypes.sizeof(orc.MoParams) + 4 == ctypes.sizeof(_lib.EnvParams)
    assert [f[0] for f in orc.MoParams._fields_] == [f[0] for f in _lib.EnvParams._fields_][1:]
    # the library reports the sizes of ITS build: a binding compares them with its own layouts
    lib = _lib.load()
    for cls, fn in ((_lib.EnvParams, lib.marlnav_sizeof_env_params), (_lib.ResetSpec, lib.marlnav_sizeof_reset_spec),
                    (_lib.IoTransform, lib.marlnav_sizeof_io_transform), (_lib.ActorSpec, lib.marlnav_sizeof_actor_spec)):
        assert fn() == ctypes.sizeof(cls) == cls().struct_size


def test_short_or_stale_structs_are_rejected():
    """ABI 4: every struct carries its size; a binding built against an older layout (e.g. the
    ABI-2 reset spec without step_counter_dev) gets MARLNAV_ERR_BAD_ARG, not an out-of-bounds read."""
    from marlnav_b200 import _lib
    lib = _lib.load()
    p = _lib.EnvParams(); p.num_envs, p.num_agents, p.num_obstacles = 64, 3, 3
    rs = _lib.ResetSpec(); rs.alias_first_step = 1
    rs.struct_size -= 8                                   # a struct that stops before step_counter_dev
    dummy = ctypes.c_void_p(16)
    rc = lib.marlnav_step_f32(ctypes.byref(p), ctypes.byref(rs), *([dummy] * 11), None, None)
    assert rc == -1 and b"marlnav_reset_spec.struct_size" in lib.marlnav_last_error()
    rs = _lib.ResetSpec(); rs.alias_first_step = 1
    p.struct_size = 0
    rc = lib.marlnav_step_f32(ctypes.byref(p), ctypes.byref(rs), *([dummy] * 11), None, None)
    assert rc == -1 and b"marlnav_env_params.struct_size" in lib.marlnav_last_error()
    p = _lib.EnvParams(); p.num_envs, p.num_agents, p.num_obstacles = 64, 3, 3
    io = _lib.IoTransform(); io.struct_size = 32
    rc = lib.marlnav_step_f32(ctypes.byref(p), ctypes.byref(rs), *([dummy] * 11), ctypes.byref(io), None)
    assert rc == -1 and b"marlnav_io_transform.struct_size" in lib.marlnav_last_error()
    rs.flags = _lib.RESET_NOISY_AGENTS; rs.alias_first_step = 0; rs.tmpl_states = 16; rs.tmpl_target = 16
    rs.states_env_stride = 15
    rc = lib.marlnav_step_f32(ctypes.byref(p), ctypes.byref(rs), *([dummy] * 11), None, None)
    assert rc == -1 and b"NOISY" in lib.marlnav_last_error()


def test_argument_errors_are_reported_without_a_gpu():
    from marlnav_b200 import _lib
    lib = _lib.load()
    p = _lib.EnvParams(); p.num_envs, p.num_agents, p.num_obstacles = 4, 1, 3
    rc = lib.marlnav_observe_f32(ctypes.byref(p), None, None, None, None, None)
    assert rc == -2 and b"num_agents" in lib.marlnav_last_error()
    p.num_agents = 3
    rc = lib.marlnav_observe_f32(ctypes.byref(p), None, None, None, None, None)
    assert rc == -1 and b"NULL" in lib.marlnav_last_error()
    g, b, s, e = (ctypes.c_int() for _ in range(4))
    p.num_envs = 1048576
    assert lib.marlnav_step_launch_info(ctypes.byref(p), ctypes.byref(g), ctypes.byref(b), ctypes.byref(s),
                                        ctypes.byref(e)) == 0
    assert g.value * e.value >= 1048576 and b.value % 32 == 0 and s.value < 227 * 1024


def test_fused_actor_step_argument_errors_and_small_batch_geometry():
    """marlnav_act_step_f32 validates before it launches; a small batch of the reference's team is
    laid out thread-per-agent (8 envs per one-warp CTA), a large one thread-per-env (32)."""
    from marlnav_b200 import _lib
    lib = _lib.load()
    p = _lib.EnvParams(); p.num_envs, p.num_agents, p.num_obstacles = 64, 3, 3
    rs = _lib.ResetSpec(); rs.alias_first_step = 1
    rc = lib.marlnav_act_step_f32(ctypes.byref(p), ctypes.byref(rs), *([None] * 16))
    assert rc == -1 and b"NULL" in lib.marlnav_last_error()
    g, b, s, e = (ctypes.c_int() for _ in range(4))
    info = lambda: lib.marlnav_step_launch_info(ctypes.byref(p), ctypes.byref(g), ctypes.byref(b), ctypes.byref(s), ctypes.byref(e))
    assert info() == 0 and (g.value, b.value, e.value) == (8, 32, 8)
    p.num_envs = 1 << 20
    assert info() == 0 and (g.value, b.value, e.value) == (1 << 15, 32, 32)
    p.num_agents, p.num_obstacles, p.num_envs = 8, 16, 4096
    assert info() == 0 and (g.value, b.value, e.value) == (1024, 32, 4)


def test_env_refuses_cpu_device():
    import marlnav_b200 as mb
    with pytest.raises(mb.MarlnavError, match="no CPU fallback"):
        mb.Env(mb.default_env_params(device='cpu'))


def test_default_params_mirror_reference_cli_defaults():
    import marlnav_b200 as mb
    p = mb.default_env_params()
    assert (p['num_parallel'], p['num_agents'], p['num_obstacles'], p['episode_len']) == (2, 3, 3, 200)
    assert (p['min_speed'], p['max_speed'], p['min_accel'], p['max_accel']) == (3., 10., -0.5, 0.5)
    assert [p[k] for k in ('risk_factor', 'distance_factor', 'heading_factor', 'target_factor',
                           'soft_factor', 'bond_factor')] == [0., 0., 500., 500., 500., 10.]
    assert p['init']['init_method'] == 'triangle' and p['sampler']['sample_method'] == 'const_sampler'
    p0 = mb.default_env_params(sampler_num=0)
    assert p0['num_obstacles'] == 1 and p0['init']['mock_states'][0][2] == [950., 100., 0., 1., 5.]
    assert mb.default_env_params(sampling_style='policy')['sampler'] is None


def test_samplers_follow_reference_scripts():
    import math
    from marlnav_b200 import params as P
    from marlnav_b200.samplers import action_sampler
    c = action_sampler(dict(P.CONST_SAMPLER, num_parallel=4, num_agents=3, device='cpu'))
    assert c().shape == (4, 3, 2) and torch.equal(c()[2, 1], torch.tensor([0., 1.]))
    m0 = action_sampler(dict(P.MOCK_SAMPLER_0, max_step=3, device='cpu'))
    assert torch.equal(m0()[1, 2], torch.tensor([0., -100.]))
    m0(); m0()
    with pytest.raises(StopIteration):
        m0()
    m1 = action_sampler(dict(P.MOCK_SAMPLER_1, max_step=5, device='cpu'))
    first, second = m1(), m1()
    assert torch.allclose(first[0, 0], torch.tensor([-math.pi / 6, 0.]))
    assert torch.allclose(first[1, 0], torch.tensor([-0.5 * math.radians(1.8), 0.]))
    assert torch.allclose(second[1, 2], torch.tensor([math.radians(1.8), 0.]))
    assert torch.equal(second[0], torch.zeros(3, 2))
    assert action_sampler(None) is None


def test_shard_bounds_partition_exactly():
    import marlnav_b200 as mb
    for total, world in [(1048576, 8), (1000, 3), (7, 7), (10, 4)]:
        spans = [mb.shard_bounds(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        mb.shard_bounds(10, 4, 4)


def test_shard_env_params_offsets_and_scenarios():
    import marlnav_b200 as mb
    full = mb.default_env_params(1000, 3, 3, sampling_style='policy')
    parts = [mb.shard_env_params(full, r, 3) for r in range(3)]
    assert [p['num_parallel'] for p in parts] == [334, 333, 333]
    assert [p['env_id_offset'] for p in parts] == [0, 334, 667]
    assert full['num_parallel'] == 1000 and 'env_id_offset' not in full
    mock = mb.shard_env_params(mb.default_env_params(sampler_num=1), 1, 2)
    assert mock['num_parallel'] == 1 and len(mock['init']['mock_states']) == 1
    assert mock['init']['mock_target'] == [[[750., 475.]]]


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import marlnav_b200 as mb
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
stats = torch.tensor([10 + rank, 200 * (rank + 1), 3000 - rank], dtype=torch.int64)
tot = mb.reduce_episode_stats(stats.clone())
assert tot.tolist() == [21, 600, 5999], tot
off, cnt = mb.shard_bounds(1001, rank, 2)
spans = [None, None]
dist.all_gather_object(spans, (off, cnt))
assert spans == [(0, 501), (501, 500)], spans
dist.destroy_process_group()
print("ok", rank)
"""


def test_episode_stats_allreduce_gloo_world2(tmp_path):
    """The only collective on the path: SUM of int64[3] episode counters (gloo on CPU here,
    NCCL over NVLink on the GPU box)."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_reduce_is_noop_without_process_group():
    import marlnav_b200 as mb
    s = torch.tensor([1, 2, 3], dtype=torch.int64)
    assert mb.reduce_episode_stats(s).tolist() == [1, 2, 3]
    with pytest.raises(ValueError):
        mb.reduce_episode_stats(torch.zeros(3))


def test_output_slots_are_recycled_only_when_released():
    """marlnav_b200/slots.py on CPU tensors: a slot comes back only when the caller holds neither one of
    its tensors, nor the namedtuple, nor a view of a view, and only on the stream it was used on."""
    from marlnav_b200.env import split_observations
    from marlnav_b200.slots import OutputSlots, _storage_use_count
    if _storage_use_count is None:
        pytest.skip("this torch has no storage use-count hook: slots are never recycled")
    B, A, O = 8, 3, 3
    S = 2 + 2 * O + 2 * (A - 1)

    def make():
        buf = torch.empty(B * A * S * 4 + 64, dtype=torch.uint8)
        obs = buf[:B * A * S * 4].view(torch.float32).view(B, A, S)
        rew = buf[B * A * S * 4:B * A * S * 4 + 32].view(torch.float32)
        fields = split_observations(obs, A, O)
        return dict(obs=obs, rew=rew, fields=fields, storage=obs.untyped_storage(), objs=(obs, rew, fields) + tuple(fields))

    pool = OutputSlots(make)
    a = pool.take(stream=7)
    ident = id(a)
    del a
    assert id(pool.take(7)) == ident and len(pool) == 1            # released -> the same slot again
    held = [pool.take(7)['rew'] for _ in range(5)]                 # MAPPO.get_data keeps every step's rewards
    assert len({r.data_ptr() for r in held}) == 5 and len(pool) == 5
    view = held[0][1:2][0:1]                                       # a view of a view
    field = pool.take(7)['fields'].others_distances                # one Observations field
    del held
    busy = sum(not pool.is_free(s, 7) for s in pool._ring)
    assert busy == 2                                               # exactly the two slots still referenced
    n = len(pool)
    for _ in range(50):
        s = pool.take(7); assert s['rew'].data_ptr() not in (view.data_ptr() - 4, ) ; del s
    assert len(pool) <= n + 2                                      # busy slots are skipped, the ring stops growing
    del view, field
    assert all(pool.is_free(s, 7) for s in pool._ring)
    assert not any(pool.is_free(s, 8) for s in pool._ring)         # another stream: never recycled there


def test_c_abi_from_plain_c(tmp_path):
    """include/marlnav_b200.h compiles as C99 (-Wall -Werror), every declared entry point links against
    libmarlnav_b200.so, the struct sizes agree, and a stale struct is rejected -- from a C program,
    the way a non-Python binding (cgo, JNI, ...) would see the library."""
    import shutil
    from marlnav_b200 import build
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    lib = build.build()
    exe = tmp_path / "abi_check"
    src = os.path.join(ROOT, "tests", "c", "abi_check.c")
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-o", str(exe), src, "-L" + os.path.dirname(lib),
           "-lmarlnav_b200", "-Wl,-rpath," + os.path.dirname(lib)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    assert "abi 4 ok" in run.stdout
    # every function the header declares is referenced by the C program
    header = open(os.path.join(ROOT, "include", "marlnav_b200.h")).read()
    declared = set(re.findall(r"\b(marlnav_[a-z0-9_]+)\s*\(", header))
    assert all(name in open(src).read() for name in declared)
