"""tests/golden/refload.py -- import the REAL reference (``/root/reference``) in the
build container.  Used only by make_golden.py and by container-only tests that
skip when the reference is absent (it never travels to the GPU box).

* matplotlib / PyQt5 are not installed, and marlnav/utils.py:3, models.py:5,
  animation.py:1,7 import matplotlib at module top -> stub modules.
* ``oracle_trig()`` swaps ``torch.cos/sin/acos`` for custom ops backed by the C
  oracle's own trig (oracle/marlnav_trig.h; with vmap rules, because the reference calls
  cos/sin under ``vmap(vmap(...))``, environment.py:128-135).  That is the one
  substitution behind the "bit-exact" golden set: every other torch op the
  reference executes is stock.
"""
import contextlib
import ctypes
import os
import sys
from unittest import mock

import numpy as np
import torch

REF_ROOT = os.environ.get('MARLNAV_REFERENCE', '/root/reference')
_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'marlnav'))


def load():
    """-> (marlnav.environment, marlnav.utils) of the unmodified reference."""
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.animation'):
        sys.modules.setdefault(name, mock.MagicMock())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import marlnav.environment as env_mod
    import marlnav.utils as utils_mod
    return env_mod, utils_mod


_ops = None


def _define_ops():
    global _ops
    if _ops is not None:
        return _ops
    from oracle import oracle as orc

    @torch.library.custom_op("marlnav_oracle::trig", mutates_args=())
    def trig(x: torch.Tensor, which: int) -> torch.Tensor:
        xin = np.ascontiguousarray(x.detach().numpy(), np.float32)
        out = np.empty_like(xin)
        orc.lib().mo_trig(ctypes.c_int(which), xin.ctypes.data_as(ctypes.c_void_p),
                          out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(xin.size))
        return torch.from_numpy(out).reshape(x.shape)

    @trig.register_fake
    def _(x, which):
        return torch.empty_like(x)

    def trig_vmap(info, in_dims, x, which):
        return trig(x, which), in_dims[0]
    torch.library.register_vmap(trig, trig_vmap)
    _ops = trig
    return trig


@contextlib.contextmanager
def oracle_trig():
    trig = _define_ops()
    saved = (torch.sin, torch.cos, torch.acos)
    torch.sin = lambda x: trig(x, 0)
    torch.cos = lambda x: trig(x, 1)
    torch.acos = lambda x: trig(x, 2)
    try:
        yield
    finally:
        torch.sin, torch.cos, torch.acos = saved


def reference_args(**over):
    """argparse.Namespace with the reference CLI defaults (__main__.py:49-133)."""
    import argparse
    d = dict(seed=None, max_x_value=1500.0, max_y_value=750.0, fig_size_x=10.0, fig_size_y=5.0,
             parallel_index=0, agent_index=0, interval=10, random=False, weights_file=None,
             num_parallel=2, num_agents=3, num_obstacles=3, max_step=1000, episode_len=200,
             min_speed=3., max_speed=10., min_accel=-0.5, max_accel=0.5, risk_factor=0.,
             distance_factor=0., heading_factor=500., target_factor=500., soft_factor=500.,
             bond_factor=10., hidden_size=50, learning_rate=0.001, ent_const=0.001, epsilon=0.01,
             gamma=0.9, num_total=1000000, buffer_len=1000, num_epochs=50, batch_size=1000,
             rendering=False, sampling_style='sampler', reward_check=True, sampler_num=-1)
    d.update(over)
    return argparse.Namespace(**d)
