# scripts/host_leg_parity.py -- the e2e leg at bench size: HostStepper (marlnav_step_host_f32, ramped
# chunks over three streams) == the device-resident step, bit for bit, 1 048 576 x 3 x 3, 60 steps
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import marlnav_b200 as mb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
params = mb.default_env_params(B, 3, 3, sampling_style='policy', episode_len=25)
e1, e2 = mb.Env(dict(params, seed=4)), mb.Env(dict(params, seed=4))
hs = mb.HostStepper(e2)
g = torch.Generator().manual_seed(1)
pool = [torch.stack([0.2 * (torch.rand(B, 3, generator=g) - 0.5), 0.5 + 0.5 * torch.rand(B, 3, generator=g)], dim=2)
        for _ in range(4)]
resets = 0
for t in range(60):
    act = pool[t % 4]
    obs, rew, term, trunc = e1.step_fused(act.cuda())
    hs.actions_host.copy_(act)
    hs.step()
    assert torch.equal(obs.cpu(), hs.obs_host), t
    assert torch.equal(rew.cpu(), hs.rewards_host), t
    assert torch.equal(term.cpu(), hs.terminated_host.bool()), t
    assert torch.equal(trunc.cpu(), hs.truncated_host.bool()), t
    resets += int((term | trunc).sum())
assert torch.equal(e1.states, e2.states)
print(json.dumps({"check": "HostStepper == device step, bit for bit", "envs": B, "steps": 60, "resets": resets}))
