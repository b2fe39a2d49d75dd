"""Shared helpers for the parity tests (oracle side is numpy, CUDA side torch)."""
import numpy as np
import torch


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_bits_equal(name, got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{name}: shape {got.shape} != {want.shape}"
    if want.dtype == np.float32:
        bad = bits(got.astype(np.float32, copy=False)) != bits(want)
    else:
        bad = got != want
    if bad.any():
        idx = np.argwhere(bad)[:5]
        detail = [(tuple(i), got[tuple(i)], want[tuple(i)]) for i in idx]
        raise AssertionError(f"{name}: {int(bad.sum())} of {bad.size} elements differ bitwise, e.g. {detail}")


def action_pool(B, A, n=8, seed=1234, angle=0.2, accel=0.5):
    """angle ~ U(-angle, angle), accel ~ U(-accel, accel)."""
    g = torch.Generator().manual_seed(seed)
    pool = []
    for _ in range(n):
        ang = (torch.rand(B, A, generator=g) * 2 - 1) * angle
        acc = (torch.rand(B, A, generator=g) * 2 - 1) * accel
        pool.append(torch.stack([ang, acc], dim=2).contiguous())
    return pool


def cpu_params(params):
    import copy
    p = copy.deepcopy(params)
    p['device'] = 'cpu'
    for k in ('init', 'sampler'):
        if p.get(k):
            p[k]['device'] = 'cpu'
    return p
