"""Multi-GPU sharding of the env batch.

Envs are independent, so the step needs no data-path collective: rank r of N
steps the contiguous slice ``[offset_r, offset_r + count_r)`` of the global env
ids with its own ``Env``.  Re-initialisation draws are addressed by GLOBAL env id
(``params['env_id_offset']``), so an N-process run is bit-identical to the
single-process run.  The only exchange is the sum of the three episode counters
(``env._num_trunc/_num_col/_num_tar``, read once per rollout by
/root/reference/marlnav/models.py:151-158): one 24-byte ``all_reduce`` over
NCCL (NVLink 5 / NVSwitch) -- or gloo on CPU tensors in the tests.
"""
import copy

import torch
import torch.distributed as dist


def shard_bounds(total, rank, world_size):
    """(offset, count) of rank's contiguous slice; counts differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(total, world_size)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def shard_env_params(params, rank, world_size):
    """Per-rank copy of a global ``params['env']`` dict."""
    offset, count = shard_bounds(int(params['num_parallel']), rank, world_size)
    if count < 1:
        raise ValueError("more ranks than environments")
    p = copy.deepcopy(params)
    p['num_parallel'] = count
    p['env_id_offset'] = int(params.get('env_id_offset', 0)) + offset
    init = p.get('init') or {}
    if 'num_parallel' in init:
        init['num_parallel'] = count
    for key in ('mock_states', 'mock_obstacles', 'mock_target'):     # per-env scenario tables
        if key in init:
            init[key] = init[key][offset:offset + count]
    smp = p.get('sampler')
    if smp and 'num_parallel' in smp:
        smp['num_parallel'] = count
    return p


def reduce_episode_stats(stats, group=None):
    """Sum an int64[3] (trunc, col, tar) tensor over all ranks, in place, and return it.
    A no-op when torch.distributed is not initialised (single process)."""
    if stats.dtype != torch.int64 or stats.numel() != 3:
        raise ValueError("episode stats must be int64[3]")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def global_episode_stats(env, group=None):
    """(num_trunc, num_col, num_tar) summed over every rank's slice, as Python ints."""
    total = reduce_episode_stats(env.episode_stats.clone(), group)
    return tuple(int(v) for v in total.tolist())
