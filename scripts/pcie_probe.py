# scripts/pcie_probe.py -- where do the ~0.3 ms between the host leg's 3.18 ms per step and the
# 2.81 ms its 157 MB download needs at the link's one-direction rate go?  Pinned copies shaped like
# marlnav_step_host_f32's at 1 048 576 x 3 x 3 (no kernels): one big download, the download in 8
# chunks, the chunks with their three small companions (rewards, two flag arrays), and each of those
# with the 25 MB upload running on a second stream.
import json
import time

import torch

B, A, S = 1048576, 3, 12
dev = torch.device("cuda:0")
obs_d = torch.empty(B * A * S, device=dev)
obs_h = torch.empty(B * A * S).pin_memory()
rew_d, rew_h = torch.empty(B, device=dev), torch.empty(B).pin_memory()
f1_d, f1_h = torch.empty(B, device=dev, dtype=torch.uint8), torch.empty(B, dtype=torch.uint8).pin_memory()
f2_d, f2_h = torch.empty(B, device=dev, dtype=torch.uint8), torch.empty(B, dtype=torch.uint8).pin_memory()
act_h = torch.empty(B * A * 2).pin_memory()
act_d = torch.empty(B * A * 2, device=dev)
s_out, s_in = torch.cuda.Stream(), torch.cuda.Stream()


def down(chunks, small):
    per = B // chunks
    with torch.cuda.stream(s_out):
        for c in range(chunks):
            lo, hi = c * per, (c + 1) * per
            obs_h[lo * A * S:hi * A * S].copy_(obs_d[lo * A * S:hi * A * S], non_blocking=True)
            if small == "each":
                rew_h[lo:hi].copy_(rew_d[lo:hi], non_blocking=True)
                f1_h[lo:hi].copy_(f1_d[lo:hi], non_blocking=True)
                f2_h[lo:hi].copy_(f2_d[lo:hi], non_blocking=True)
        if small == "end":
            rew_h.copy_(rew_d, non_blocking=True)
            f1_h.copy_(f1_d, non_blocking=True)
            f2_h.copy_(f2_d, non_blocking=True)


def up(chunks):
    per = B // chunks
    with torch.cuda.stream(s_in):
        for c in range(chunks):
            lo, hi = c * per, (c + 1) * per
            act_d[lo * A * 2:hi * A * 2].copy_(act_h[lo * A * 2:hi * A * 2], non_blocking=True)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
        torch.cuda.synchronize()            # one "step" at a time, like the host leg
    return (time.perf_counter() - t0) / reps * 1e3


res = {}
for chunks in (1, 8):
    for small in ("none", "each", "end"):
        res[f"down_c{chunks}_small_{small}"] = timed(lambda: down(chunks, small))
        res[f"down_c{chunks}_small_{small}+up"] = timed(lambda: (up(chunks), down(chunks, small)))
res["up_only_c1"] = timed(lambda: up(1))
res["up_only_c8"] = timed(lambda: up(8))
print(json.dumps({"probe": "ms per step-shaped copy set, 1048576 x 3 x 3, pinned, no kernels", "ms": res}))
