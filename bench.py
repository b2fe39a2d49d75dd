#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused MARL-nav environment step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3]): 1,048,576 parallel envs x 3 agents x 3
obstacles per GPU, policy-like random actions (turn angle
~U(-0.2,0.2) rad, accel ~U(-0.5,0.5), a pool of 16 pre-generated action tensors
cycled), auto-reset on every termination/truncation, episode_len 200.  A "step"
is one Env.step over the whole batch.  N>1 (torchrun, one rank per GPU): every
rank steps its own 1M-env slice (weak scaling, no data-path collective); the
three episode counters are all-reduced over NCCL once after the timed region.

One JSON line on stdout (rank 0):
  value      env-steps/s, inputs resident in HBM, CUDA events, max over ranks
  e2e        same metric through marlnav_step_host_f32: pinned HOST actions in,
             HOST observations/rewards/flags out, copies inside the timed region
  roofline   algorithmic bytes per launch / mean launch duration vs measured HBM peak
  cpu_baseline  the torch-op port of the reference step (oracle/oracle.py:TorchPortEnv)
             timed on this box's host cores on a bounded sample
--impl reference times that CPU port alone (the reference is pure Python/torch
and is not installed on the GPU box; the port reproduces its op sequence and bits).
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "env_steps_per_sec", "env-steps/s"


def algorithmic_bytes(A, O):
    """every tensor read once + written once per env-step."""
    S = 2 + 2 * O + 2 * (A - 1)
    rd = 20 * A + 8 * A + 8 * O + 8 + 4 + 1
    wr = 20 * A + 4 + 1 + 4 * A * S + 4 + 1 + 1
    return rd + wr


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(A, O, B):
    """Per-launch DRAM bytes from the committed ncu capture, if one matches this shape."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            t = json.load(f)
        return t.get(f"{B}x{A}x{O}")
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None

    def run(self):
        if self._h is None:
            return
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "samples": len(s), "reasons": sorted(self.reasons)}


def make_action_pool(B, A, n, device, seed=1234, angle=0.2):
    """turn angle ~ U(-angle, angle) -- 0.2 rad = "policy-like" (agents make
    progress, collide, reset), pi = trig stress -- and acceleration ~ U(-0.5, 0.5)."""
    g = torch.Generator().manual_seed(seed)
    pool = []
    for _ in range(n):
        ang = (torch.rand(B, A, generator=g) * 2 - 1) * angle
        acc = (torch.rand(B, A, generator=g) * 2 - 1) * 0.5
        pool.append(torch.stack([ang, acc], dim=2).contiguous().to(device))
    return pool


def env_params(B, A, O, device, offset=0):
    import marlnav_b200 as mb
    if A == 3:
        p = mb.default_env_params(B, A, O, sampling_style='policy', device=device)
    else:
        p = mb.template_env_params(B, A, O, device=device)
    p['seed'] = 0
    p['env_id_offset'] = offset
    return p


# ----------------------------------------------------------------------------- CPU arm

def time_cpu_port(A, O, budget_s, steps=None, warmup=1, sample_envs=None):
    """Times oracle.TorchPortEnv (reference op sequence on torch CPU).  Returns
    (env_steps_per_s, cores, description, steps_done, ms_per_step)."""
    from oracle import oracle as orc
    try:        # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core it may run on
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    cores = torch.get_num_threads()
    B = sample_envs or 65536
    p = orc.default_env_params(B, A, O) if A == 3 else None
    if p is None:
        p = orc.default_env_params(B, A, O)
        p['init'] = dict(p['init'], init_method='template', agent_template=orc.ring_template(A))
    env = orc.TorchPortEnv(p, seed=0)
    pool = make_action_pool(B, A, 4, 'cpu')
    t0 = time.perf_counter()
    for i in range(max(warmup, 1)):
        env.step(pool[i % 4])
    per = (time.perf_counter() - t0) / max(warmup, 1)
    if steps is None:
        steps = max(3, min(200, int(budget_s / max(per, 1e-6))))
    t0 = time.perf_counter()
    for i in range(steps):
        env.step(pool[i % 4])
    dt = time.perf_counter() - t0
    desc = (f"{B} envs x {A} agents x {O} obstacles per step, {steps} timed steps "
            f"({dt:.1f} s) of oracle.TorchPortEnv (reference op sequence, torch {torch.__version__} CPU, "
            f"{cores} threads of {os.cpu_count()} cpus)")
    return B * steps / dt, cores, desc, steps, 1e3 * dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    A, O = args.agents, args.obstacles
    # bounded sample per step so that K steps end within a few minutes
    sample = 65536 if A * O <= 9 else 8192
    if args.steps * (0.13 if A * O <= 9 else 0.2) > 150:
        sample //= 4
    v, cores, desc, steps, ms = time_cpu_port(A, O, 0, steps=args.steps, warmup=max(args.warmup, 1),
                                              sample_envs=sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.envs),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, B_local):
    return {"workload": f"{B_local} envs x {args.agents} agents x {args.obstacles} obstacles per GPU, "
                        "random policy-like actions, auto-reset, episode_len 200 "
                        "(BASELINE.json configs[3]; configs[4] with --agents 8 --obstacles 16)",
            "envs_per_gpu": B_local, "num_agents": args.agents, "num_obstacles": args.obstacles,
            "action_pool": 16, "action_angle_range": args.angle, "prewarm": f"{args.prewarm_s} s of device copies before the warm-up steps",
            "l2_policy": "working set per step (states+actions+obs) exceeds the 126 MB L2"
            if B_local * algorithmic_bytes(args.agents, args.obstacles) > 2 * 126e6 else
            "working set may fit in L2 -- not an HBM number"}


# ----------------------------------------------------------------------------- GPU arm

def run_ours(args):
    import torch.distributed as dist
    import marlnav_b200 as mb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for --impl ours)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL_DEBUG=VERSION (set on the GPU boxes) makes NCCL
        # printf its version banner there, and so does WARN; an explicit INFO/TRACE is left alone
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)

    A, O = args.agents, args.obstacles
    B = args.envs if args.scaling == "weak" else mb.shard_bounds(args.envs, rank, world)[1]
    offset = rank * B if args.scaling == "weak" else mb.shard_bounds(args.envs, rank, world)[0]
    env = mb.Env(env_params(B, A, O, f"cuda:{local_rank}", offset))
    pool = make_action_pool(B, A, 16, dev, angle=args.angle)
    out = env._alloc_outputs()          # steady-state callers reuse or recycle output tensors
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Bring the device out of any idle P-state before the W warm-up steps (a fresh process starts
    # after seconds of interpreter start-up with the GPU idle, and W short steps are only ~1.5 ms).
    # Plain device copies on scratch tensors, nothing of the workload; --prewarm-s 0 disables it.
    scratch = torch.empty(2, 64 << 20, dtype=torch.float32, device=dev)
    t_ramp = time.perf_counter()
    while time.perf_counter() - t_ramp < args.prewarm_s:
        for _ in range(8):
            scratch[1].copy_(scratch[0])
        torch.cuda.synchronize()
    del scratch
    for i in range(W):
        env.step_fused(pool[i % 16], out=out)
    barrier()
    stats0 = env.episode_stats.clone()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        env.step_fused(pool[i % 16], out=out)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    stats_delta = env.episode_stats.clone() - stats0

    # end-to-end leg: host actions -> device step -> host outputs, copies timed
    hs = mb.HostStepper(env)
    K2 = max(3, min(K, args.e2e_steps))
    host_pool = [p.cpu().pin_memory() for p in pool[:4]]     # the policy's outputs, in pinned host memory
    for i in range(2):
        hs.step(host_pool[i % 4])
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    checksum = 0.0
    for i in range(K2):
        hs.step(host_pool[i % 4], sync=True)         # H2D actions -> fused step -> D2H obs/rewards/flags
        checksum += float(hs.rewards_host[0])        # the caller reads the step's result
    f1.record()
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), 1e3 * (time.perf_counter() - t0))

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mb.reduce_episode_stats(stats_delta)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        total_envs = B * world
        value = total_envs * K / (ms * 1e-3)
        e2e_value = total_envs * K2 / (ms_e2e * 1e-3)
        peak, peak_src = measured_peak()
        bytes_launch = B * algorithmic_bytes(A, O)
        achieved = bytes_launch / (ms / K * 1e-3) / 1e9
        g, blk, smem, tile = env.launch_info()
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, cores, desc, _, _ = time_cpu_port(A, O, args.cpu_budget)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, B), grid=g, block=blk, smem_bytes=smem, envs_per_cta=tile,
                           agent_steps_per_sec=value * A,
                           episode_events_in_timed_region=dict(zip(("trunc", "col", "tar"),
                                                                   [int(x) for x in stats_delta.tolist()]))),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(A, O, B), "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": algorithmic_bytes(A, O),
                         "kernel": ("mn::step_env_kernel" if (A, O) in ((3, 3), (3, 1)) else
                                    "mn::step_team_kernel" if (A, O) == (8, 16) else "mn::step_kernel"),
                         "per": "one launch = one step of one GPU's slice"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": hs.h2d_bytes * world,
                    "d2h_bytes_per_step": hs.d2h_bytes * world, "steps": K2, "ms_per_step": ms_e2e / K2,
                    "path": "HostStepper -> marlnav_step_host_f32 (pinned host buffers)"},
            "gpu_launches": K,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--envs", type=int, default=1048576, help="envs per GPU (weak) or total (strong)")
    ap.add_argument("--agents", type=int, default=3)
    ap.add_argument("--obstacles", type=int, default=3)
    ap.add_argument("--scaling", choices=("weak", "strong"), default="weak")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU baseline work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--angle", type=float, default=0.2,
                    help="turn-angle range of the random actions in rad (SURVEY 8d: 0.2 policy-like, 3.14159 trig stress)")
    ap.add_argument("--prewarm-s", type=float, default=0.5,
                    help="seconds of device copies before the warm-up steps (P-state ramp)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
