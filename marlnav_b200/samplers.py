"""Host-side action samplers behind ``Env.sample_actions()``.

These are the reference's deterministic test drivers for ``-sn -1/0/1``
(/root/reference/marlnav/utils.py:419-497), kept on the host: they are fixtures,
not part of the accelerated path.
"""
import math

import torch


class ConstantSampler:
    """utils.py:477-485 -- every agent always gets action ``[0, 1]``."""

    def __init__(self, params):
        self.actions = torch.tensor([0., 1.], device=params['device']).repeat(
            params['num_parallel'], params['num_agents'], 1)

    def __call__(self):
        return self.actions


class MockSampler:
    """utils.py:419-451 -- scripted two-env action sequences, ``max_step`` calls long.

    sampler_num 0: the same (2,3,2) action tensor every call.
    sampler_num 1: call 0 turns env 0's outer agents by -/+ pi/6 and env 1's agents by
    half their scripted angle; afterwards the scripted actions repeat."""

    def __init__(self, params):
        self._device = params['device']
        self._left = params['max_step']
        self._num = params['sampler_num']
        self._base = params['actions']
        self._calls = 0
        if self._num not in (0, 1):
            raise NotImplementedError(self._num)

    def __call__(self):
        if self._left <= 0:
            raise StopIteration     # the reference's generator is exhausted after max_step calls
        self._left -= 1
        env0, env1 = [list(a) for a in self._base[0]], [list(a) for a in self._base[1]]
        if self._num == 1 and self._calls == 0:
            env0 = [[-math.pi / 6, 0.], env0[1], [math.pi / 6, 0.]]
            env1 = [[0.5 * a[0], 0.] for a in env1]
        self._calls += 1
        return torch.tensor([env0, env1], device=self._device)


def action_sampler(params):
    """utils.py:488-497"""
    if params is None:
        return None
    if params['sample_method'] == 'mock_sampler':
        return MockSampler(params)
    if params['sample_method'] == 'const_sampler':
        return ConstantSampler(params)
    raise NotImplementedError(params['sample_method'])
