"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous.

Independent networks (the reference's semantics, train.py:123-126) shard by contiguous
agent ranges with NO data-path collective: GPU r owns agents [r*N/G, (r+1)*N/G) with their
replay rings, theta, theta_tgt, Adam state.  Per-agent results are
bit-identical to the 1-GPU run because every kernel works on one agent's data only.

Shared parameters (BASELINE.json cfg5, not in the reference): every rank holds
a replica of theta/theta_tgt/m/v, draws batch/G transitions from its own rings, computes
dL/dtheta with the loss mean over the GLOBAL batch, the gradient blocks are summed with one
all-reduce (NCCL over NVLink/NVSwitch; gloo in the CPU tests), and every rank applies the
same Adam step.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist


def shard_range(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous agent range of ``rank``; the first ``n_total % world`` ranks get one extra."""
    if not (0 <= rank < world) or n_total < 0:
        raise ValueError(f"bad shard request n_total={n_total} world={world} rank={rank}")
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def grid_neighbor_table(rows: int, cols: int):
    """Neighbour rows ``[rows*cols, 4]`` (int32, n,s,e,w; -1 = none) of a row-major ``J_r_c`` grid -- the table
    ``dmdqn_featurize`` takes instead of junction-id strings (order_lanes.py:399-404: n=(r-1,c), s=(r+1,c),
    e=(r,c+1), w=(r,c-1))."""
    import numpy as np
    r, c = np.divmod(np.arange(rows * cols), cols)
    out = np.full((rows * cols, 4), -1, np.int32)
    for k, (dr, dc) in enumerate(((-1, 0), (1, 0), (0, 1), (0, -1))):
        rr, cc = r + dr, c + dc
        ok = (rr >= 0) & (rr < rows) & (cc >= 0) & (cc < cols)
        out[ok, k] = (rr * cols + cc)[ok]
    return out


def shard_seed(seed: int, agent_global_index: int) -> int:
    """Per-agent init seed that does not depend on how agents are sharded."""
    return int(seed) + int(agent_global_index)


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class SharedParameterStep:
    """One data-parallel learn step of a shared network: local gradients -> all-reduce ->
    identical Adam on every replica.  ``local_grads(global_batch)`` returns the rank's
    gradient block (already divided by the global batch) and a [.., 8] metrics tensor whose
    column 0 is the rank's share of the loss; ``apply(grads)`` performs the update.  The two
    callables are the native calls on a GPU (``for_group``) and oracle stand-ins in the CPU
    (gloo) tests, so the same host logic is exercised in both.

    The learn decision is COLLECTIVE.  Whether a rank's rings hold its share of the batch
    (dqn_agent.py:333-335 applied to the rank's shard) differs between ranks while the rings
    fill -- ``shard_range`` gives the first ranks one agent more, masked pushes fill unevenly --
    and a replica that stepped alone would diverge for good (theta, Adam state, the
    target-sync counter).  So ``local_ready()`` flags are combined with a MIN all-reduce and
    either every rank steps or none does; ``step()`` returns None for a skipped step, like
    ``learn()`` while the buffer is short.  Rings only grow, so once every rank has been
    ready the exchange is dropped."""

    def __init__(self, local_grads, apply, local_batch: int, group=None, local_ready=None, flag_device="cpu"):
        self.local_grads, self.apply, self.local_batch, self.group = local_grads, apply, int(local_batch), group
        self.local_ready, self.flag_device = local_ready, flag_device
        self._all_ready = False
        self.skipped = 0

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def everyone_ready(self) -> bool:
        if self._all_ready:
            return True
        ready = True if self.local_ready is None else bool(self.local_ready())
        if self.world > 1:
            flag = torch.tensor([1 if ready else 0], dtype=torch.int32, device=self.flag_device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            ready = bool(flag.item())
        self._all_ready = ready
        return ready

    def step(self):
        if not self.everyone_ready():
            self.skipped += 1
            return None
        grads, metrics = self.local_grads(self.local_batch * self.world)
        allreduce_sum_(grads, self.group)           # 4*P bytes, latency bound (1.24 MB at H=512)
        loss = allreduce_sum_(metrics[..., 0].clone(), self.group)
        self.apply(grads)
        return loss

    @classmethod
    def for_group(cls, grp, group=None) -> "SharedParameterStep":
        """Bind to an AgentGroup built with ``share_parameters=True``."""
        from . import _native as N
        from .group import _ptr
        if not grp.shared:
            raise ValueError("SharedParameterStep needs share_parameters=True")
        grads = torch.zeros_like(grp.theta)

        def local_ready() -> bool:
            return bool(grp.active_host()[0])

        def local_grads(global_batch: int):
            if not local_ready():                   # cannot happen after everyone_ready(); never reduce a stale block
                grads.zero_()
                grp.metrics.zero_()
                return grads, grp.metrics
            d = grp.draw_words((1, grp.batch_size))
            with torch.cuda.device(grp.device):
                N.check(grp.lib.dmdqn_learn_grads(C.byref(grp.dims), C.byref(grp.hp), C.byref(grp.replay), C.byref(grp.nets),
                                                  _ptr(d), None, int(global_batch), _ptr(grads), _ptr(grp.metrics),
                                                  _ptr(grp.workspace), grp.workspace.numel(), grp._stream))
            grp.learn_step_host += 1
            return grads, grp.metrics

        def apply(g: torch.Tensor):
            with torch.cuda.device(grp.device):
                N.check(grp.lib.dmdqn_adam_apply(C.byref(grp.dims), C.byref(grp.hp), C.byref(grp.nets), _ptr(g),
                                                 _ptr(grp.workspace), grp.workspace.numel(), grp._stream))
        return cls(local_grads, apply, grp.batch_size, group, local_ready=local_ready, flag_device=grp.device)
