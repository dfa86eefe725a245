#!/usr/bin/env python
"""2+ GPU plumbing check of shared-parameter mode, run under torchrun:
   torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shared_step_nccl.py
Every rank holds a replica of one shared network and its own agents' rings, computes gradients on its
B/G slice, and the blocks are summed + Adam applied either by the fused peer-memory kernel (dmdqn_allreduce_adam over
CUDA-IPC mapped buffers: the product path) or by NCCL all-reduce + dmdqn_adam_apply (the baseline).  Checks for both:
replicas stay bit-identical, the loss is finite; times both.  (The arithmetic itself is checked against the
oracle by tests/test_gpu_parity.py::test_shared_parameter_step... and tests/test_parallel_gloo.py.)"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmdqn_b200.group import AgentGroup          # noqa: E402
from dmdqn_b200.parallel import SharedParameterStep  # noqa: E402


def run(rank, world, fused):
    h, agents_total, batch_global, cap = 512, 1024, 1024, 2000       # BASELINE cfg5 shape, shallow rings
    n_local, b_local = agents_total // world, batch_global // world
    cfg = {"nn_layers": [h, h], "replay_buffer_size": cap, "batch_size": b_local, "learning_rate": 5e-4,
           "share_parameters": True, "target_update_frequency": 2}
    grp = AgentGroup(n_local, cfg, seed=1234)                          # same seed: identical replicas
    gen = torch.Generator(device=grp.device).manual_seed(100 + rank)   # different data per rank
    grp.obs[:, :, :89] = torch.randint(-1, 20, (n_local, cap, 89), device=grp.device, generator=gen).float()
    grp.next_obs[:, :, :89] = torch.randint(-1, 20, (n_local, cap, 89), device=grp.device, generator=gen).float()
    grp.act_ring.copy_(torch.randint(0, 4, (n_local, cap), device=grp.device, generator=gen).int())
    grp.rew_ring.copy_(-torch.rand((n_local, cap), device=grp.device, generator=gen, dtype=torch.float64) * 100)
    grp.n_written.fill_(cap); grp.n_written_host[:] = cap
    grp._gen.manual_seed(7 + rank)
    step = SharedParameterStep.for_group(grp, fused=fused)
    losses = []
    for _ in range(5):
        losses.append(float(step.step()[0]))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        step.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    digest = grp.theta.double().sum().reshape(1)
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same = all(torch.equal(all_d[0], d) for d in all_d)
    err = int(grp.debug_views()["tc_error"][0])
    out = {"world": world, "path": "fused peer-memory reduce + Adam (dmdqn_allreduce_adam)" if fused else "NCCL all_reduce + dmdqn_adam_apply",
           "replicas_identical": bool(same), "losses": losses, "ms_per_shared_update": ms, "shared_updates_per_s": 1e3 / ms,
           "agent_updates_per_s": agents_total * 1e3 / ms, "allreduce_bytes": int(grp.theta.numel() * 4), "tc_error": err}
    assert same and np.all(np.isfinite(losses)) and err == 0, out
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = [run(rank, world, fused) for fused in (True, False)]
    if rank == 0:
        close = np.allclose(res[0]["losses"], res[1]["losses"], rtol=1e-4)
        print(json.dumps({"fused": res[0], "nccl": res[1], "losses_agree_between_paths": bool(close)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
