"""A few launches of every non-GEMM kernel at cfg3 shapes, for a quick `ncu --set full` capture:
   ncu --set full --clock-control none -k regex:"gather|featurize|act_kernel|sample|push|peer_reduce" -c 40 -o prof python tools/small_kernels.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
from dmdqn_b200 import _native as N
from dmdqn_b200.group import AgentGroup
from dmdqn_b200.parallel import PeerExchange, SharedParameterStep, grid_neighbor_table

D, A = 89, 4
cfg = {"nn_layers": [256, 256], "replay_buffer_size": 30000, "batch_size": 256, "precision": "tf32x3"}
n, b = 256, 256
grp = AgentGroup(n, cfg, D, A, seed=1)
gen = torch.Generator(device=grp.device).manual_seed(3)
grp.obs[:, :, :D] = torch.randint(-1, 20, (n, 30000, D), device=grp.device, generator=gen).float()
grp.next_obs[:, :, :D] = torch.randint(-1, 20, (n, 30000, D), device=grp.device, generator=gen).float()
grp.n_written.fill_(30000); grp.n_written_host[:] = 30000
for _ in range(3):                                              # push + sample + gather
    s = torch.randint(-1, 20, (n, D), device=grp.device).float()
    grp.push(s, torch.zeros(n, dtype=torch.int32, device=grp.device), torch.zeros(n, dtype=torch.float64, device=grp.device), s,
             torch.zeros(n, dtype=torch.uint8, device=grp.device))
    grp.sample(grp.draw_words((n, b)))
halting = torch.randint(0, 20, (n, 12), dtype=torch.int32, device=grp.device)
zi = torch.zeros(n, dtype=torch.int32, device=grp.device); zd = torch.zeros(n, dtype=torch.float64, device=grp.device)
zv = torch.zeros(n, dtype=torch.uint8, device=grp.device)
nbr = torch.as_tensor(grid_neighbor_table(16, 16)).to(grp.device)
for _ in range(3):
    grp.featurize(halting, zi, zd, zd, 0.0, zv, nbr)
del grp
torch.cuda.empty_cache()
big = AgentGroup(768, dict(cfg, replay_buffer_size=4, batch_size=4), D, A, seed=2)      # 276 MB of weights: larger than L2
obs = torch.randint(-1, 20, (768, D), device=big.device).float()
e0 = torch.zeros(768, dtype=torch.float64, device=big.device); wz = torch.zeros(768, dtype=torch.int32, device=big.device)
for _ in range(4):
    big.act(obs, e0, wz, wz)
del big
sh = AgentGroup(8, dict(cfg, nn_layers=[512, 512], replay_buffer_size=64, batch_size=128, share_parameters=True, precision="fp32"), D, A, seed=3)
for _ in range(70):
    s = torch.randint(-1, 20, (8, D), device=sh.device).float()
    sh.push(s, torch.zeros(8, dtype=torch.int32, device=sh.device), torch.zeros(8, dtype=torch.float64, device=sh.device), s,
            torch.zeros(8, dtype=torch.uint8, device=sh.device))
step = SharedParameterStep.for_group(sh, exchange=PeerExchange(sh, 0, 1))
for _ in range(3):
    step.step()
torch.cuda.synchronize()
print("ok")
