"""CPU tests of the oracle's arithmetic building blocks (no GPU, no reference needed)."""
import ctypes

import numpy as np
import pytest
import torch


def _trig(orc, which, x, sleef=False):
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    fn = orc.lib().mo_trig_sleef if sleef else orc.lib().mo_trig
    fn(ctypes.c_int(which), x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p),
       ctypes.c_size_t(x.size))
    return y


def _ulp_err(got, exact64):
    exact32 = exact64.astype(np.float32)
    ulp = np.spacing(np.abs(exact32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - exact64) / ulp


def test_philox_known_answers(oracle):
    """Random123 Philox4x32-10 KATs."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
        oracle.lib().mo_philox_kat(c, k, o)
        assert tuple(o) == want


def test_philox_obstacles_c_equals_numpy(oracle):
    for B, O, seed, ctr, off in [(257, 3, 0, 0, 0), (64, 16, 12345678901234, 77, 1 << 33), (5, 1, 7, 3, 9)]:
        p = oracle.make_params(oracle.default_env_params(B, 3, O))
        a = oracle.philox_obstacles(p, seed, ctr, off)
        b = oracle.philox_obstacles_numpy(p, seed, ctr, off)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert (a[..., 0] >= 500).all() and (a[..., 0] < 1000).all()
        assert (a[..., 1] >= 250).all() and (a[..., 1] < 500).all()


def test_philox_is_addressed(oracle):
    """Draws depend only on (seed, global env id, counter): shard-invariant."""
    p_full = oracle.make_params(oracle.default_env_params(100, 3, 3))
    p_half = oracle.make_params(oracle.default_env_params(50, 3, 3))
    full = oracle.philox_obstacles(p_full, 9, 4, 0)
    lo, hi = oracle.philox_obstacles(p_half, 9, 4, 0), oracle.philox_obstacles(p_half, 9, 4, 50)
    assert np.array_equal(full, np.concatenate([lo, hi]))
    assert not np.array_equal(full, oracle.philox_obstacles(p_full, 9, 5, 0))


@pytest.mark.parametrize("n", list(range(2, 41)) + [47, 48, 63, 64, 65, 100])
def test_row_sum_matches_torch_sum(oracle, n):
    """mo_torch_row_sum reproduces torch.sum over a contiguous inner dim bit for bit
    (the order behind torch.mean in environment.py:233,269)."""
    g = torch.Generator().manual_seed(n)
    x = (torch.rand(4000, n, generator=g) * 1000 - 500)
    want = torch.sum(x, dim=1).numpy()
    xn = x.numpy()
    fn = oracle.lib().mo_row_sum
    got = np.array([fn(xn[i].ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n)) for i in range(0, 4000, 7)],
                   np.float32)
    assert np.array_equal(got.view(np.uint32), want[::7].view(np.uint32))


def test_trig_accuracy_sampled(oracle):
    """marlnav_trig.h stays within the exhaustively measured bounds (1.38/1.48/2.10 ulp;
    oracle/verify_math.c sweeps every float32) on a dense sample incl. the edges."""
    rng = np.random.default_rng(0)
    t = np.concatenate([rng.uniform(-np.pi, np.pi, 2_000_000), [0.0, -0.0, np.pi, -np.pi, 1e-30, -1e-8]])
    t = t.astype(np.float32)
    t = np.clip(t, -np.float32(np.pi), np.float32(np.pi))
    assert _ulp_err(_trig(oracle, 0, t), np.sin(t.astype(np.float64))).max() < 1.40
    assert _ulp_err(_trig(oracle, 1, t), np.cos(t.astype(np.float64))).max() < 1.50
    x = np.concatenate([rng.uniform(-1, 1, 2_000_000), 1 - np.logspace(-8, -1, 5000), [1.0, -1.0, 0.0, 0.5, -0.5]])
    x = x.astype(np.float32)
    assert _ulp_err(_trig(oracle, 2, x), np.arccos(x.astype(np.float64))).max() < 2.11
    assert _trig(oracle, 2, np.array([1.0], np.float32))[0] == 0.0
    s0 = _trig(oracle, 0, np.array([-0.0], np.float32))
    assert s0[0] == 0.0 and np.signbit(s0[0])          # sin(-0) = -0 like torch


def test_trig_close_to_torch_cpu(oracle):
    """Against torch's own CPU sin/cos/acos (MKL VML or SLEEF, depending on the build) the
    oracle's sin/cos differ by at most 2 ulp -- the same order as torch's backends differ from
    each other -- and its acos (2.10 ulp from exact) by at most 3: 3.6e-7 relative, against the 1e-5
    parity tolerance versus the stock reference."""
    rng = np.random.default_rng(1)
    t = rng.uniform(-np.pi, np.pi, 1_000_000).astype(np.float32)
    x = rng.uniform(-1, 1, 1_000_000).astype(np.float32)
    for which, arr, fn in ((0, t, torch.sin), (1, t, torch.cos), (2, x, torch.acos)):
        want = fn(torch.from_numpy(arr)).numpy()
        got = _trig(oracle, which, arr)
        ulps = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
        big = np.abs(want) > 1e-3          # ulp distance is meaningless across a zero crossing
        assert ulps[big].max() <= (3 if which == 2 else 2)


def test_sleef_restatement_known_bits(oracle):
    """torch_cpu_math.h (SLEEF u10, torch's non-MKL backend) -- spot values verified against
    the Sleef_*f16_u10 symbols inside libtorch_cpu.so when this file was written."""
    x = np.array([2.4472241, -1.1192276, 0.5047433], np.float32)
    assert np.allclose(_trig(oracle, 0, x, sleef=True), np.sin(x.astype(np.float64)), rtol=2e-7)
    assert np.allclose(_trig(oracle, 1, x, sleef=True), np.cos(x.astype(np.float64)), rtol=2e-7)
