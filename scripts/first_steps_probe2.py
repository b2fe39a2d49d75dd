# scripts/first_steps_probe2.py -- one-shot 20-step regions on FRESH envs (what the driver's
# --steps 20 --warmup 5 line times), against the number of warm-up steps before the region
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import marlnav_b200 as mb

dev = torch.device("cuda:0")
B = 1048576
pool = bench.make_action_pool(B, 3, 16, dev)
scratch = torch.empty(2, 64 << 20, device=dev)
for _ in range(200):
    scratch[1].copy_(scratch[0])
torch.cuda.synchronize()
res = []
for rep in range(3):
    for W in (5, 10, 30, 100):
        env = mb.Env(bench.env_params(B, 3, 3, "cuda:0"))
        for i in range(W):
            env.step(pool[i % 16])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            env.step(pool[i % 16])
        e1.record()
        torch.cuda.synchronize()
        res.append({"rep": rep, "warmup": W, "us_per_step": e0.elapsed_time(e1) * 1e3 / 20})
        del env
print(json.dumps(res))
