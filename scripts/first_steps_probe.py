# scripts/first_steps_probe.py -- where do the ~40 us go that a 20-step timed region costs on top of
# 20 x the long-run step time?  (a) whole-region time for K = 5..320 after a synchronize (fit a + b K),
# (b) per-step event times of the first 24 steps after a synchronize (events between launches remove
# the programmatic-dependent-launch overlap, so (b) shows shape only).
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import marlnav_b200 as mb

dev = torch.device("cuda:0")
B = 1048576
env = mb.Env(bench.env_params(B, 3, 3, "cuda:0"))
pool = bench.make_action_pool(B, 3, 16, dev)
for i in range(30):
    env.step(pool[i % 16])
torch.cuda.synchronize()


def region(K):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        env.step(pool[i % 16])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3


fit = {}
for K in (5, 10, 20, 40, 80, 160, 320):
    ts = sorted(region(K) for _ in range(7))
    fit[K] = {"us_total_median": ts[3], "us_per_step": ts[3] / K}
ks = sorted(fit)
b = (fit[ks[-1]]["us_total_median"] - fit[ks[-2]]["us_total_median"]) / (ks[-1] - ks[-2])
a = {K: fit[K]["us_total_median"] - b * K for K in ks}

torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(25)]
ev[0].record()
for i in range(24):
    env.step(pool[i % 16])
    ev[i + 1].record()
torch.cuda.synchronize()
per = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(24)]
print(json.dumps({"region_fit": fit, "slope_us_per_step": b, "intercept_us_by_K": a, "per_step_with_events_us": per}))
