"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: agent sharding and the
shared-parameter gradient all-reduce, with the oracle standing in for the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dmdqn_b200.parallel import SharedParameterStep, shard_range, shard_seed


def test_shard_range_partitions_agents():
    for n, w in ((256, 2), (256, 8), (4096, 8), (9, 2), (10, 4), (3, 8)):
        spans = [shard_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(4096, 8, 3) == (1536, 2048)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
    assert shard_seed(7, 1536) == 1543


def test_grid_neighbor_table_matches_oracle():
    from dmdqn_b200.parallel import grid_neighbor_table
    from oracle.featurize import grid_neighbors
    for r, c in ((1, 1), (3, 3), (4, 7), (16, 32)):
        assert np.array_equal(grid_neighbor_table(r, c), grid_neighbors(r, c))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.dqn import StackedOracle, adam_scalars, adam_update_, loss_and_grad
        torch.set_num_threads(1)
        h, b_global = 64, 64
        b_local = b_global // world
        rng = np.random.default_rng(0)                      # same data on both ranks; each takes its slice
        S = rng.integers(-1, 20, (1, b_global, 89)).astype(np.float32)
        S2 = rng.integers(-1, 20, (1, b_global, 89)).astype(np.float32)
        A_ = rng.integers(0, 4, (1, b_global)); R = rng.standard_normal((1, b_global)).astype(np.float32)
        D = (rng.random((1, b_global)) < 0.1).astype(np.float32)
        sl = slice(rank * b_local, (rank + 1) * b_local)
        stk = StackedOracle(1, 89, [h, h], 4, seed0=3)      # replica (same seed everywhere)

        def local_grads(global_batch):
            y, _, _ = stk.td_targets(R[:, sl], S2[:, sl], D[:, sl])
            params = [p.detach().requires_grad_(True) for p in stk.online]
            q = stk.forward(params, torch.as_tensor(S[:, sl]))
            pred = torch.gather(q, 2, torch.as_tensor(A_[:, sl])[..., None])[..., 0]
            terms, _ = loss_and_grad(pred, y, "mse")
            loss = terms.sum(dim=1) / global_batch          # mean over the GLOBAL batch
            grads = torch.autograd.grad(loss.sum(), params)
            flat = torch.cat([g.reshape(-1) for g in grads])
            metrics = torch.zeros(1, 8); metrics[0, 0] = loss.detach()[0]
            return flat, metrics

        def apply(flat):
            stk.learn_step[0] += 1
            alpha, eps = adam_scalars(int(stk.learn_step[0]), stk.lr)
            off = 0
            with torch.no_grad():
                for p, m, v in zip(stk.online, stk.adam_m, stk.adam_v):
                    n = p.numel()
                    adam_update_(p, flat[off:off + n].view_as(p), m, v, alpha, eps)
                    off += n

        step = SharedParameterStep(local_grads, apply, b_local)
        assert step.world == world
        loss = step.step()
        q.put((rank, float(loss[0]), [p.clone().numpy() for p in stk.online]))
    finally:
        dist.destroy_process_group()


def test_shared_parameter_allreduce_equals_single_process_big_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle.dqn import StackedOracle
    rng = np.random.default_rng(0)
    S = rng.integers(-1, 20, (1, 64, 89)).astype(np.float32); S2 = rng.integers(-1, 20, (1, 64, 89)).astype(np.float32)
    A_ = rng.integers(0, 4, (1, 64)); R = rng.standard_normal((1, 64)).astype(np.float32)
    D = (rng.random((1, 64)) < 0.1).astype(np.float32)
    ref = StackedOracle(1, 89, [64, 64], 4, seed0=3)
    out = ref.learn_on_batch(S, A_, R, S2, D)
    for rank, loss, weights in results:
        assert abs(loss - out["loss"][0]) <= 1e-5 * abs(out["loss"][0])
        for w, r in zip(weights, ref.online):
            np.testing.assert_allclose(w, r.numpy(), rtol=1e-5, atol=1e-6)
    # replicas stay identical
    for w0, w1 in zip(results[0][2], results[1][2]):
        assert np.array_equal(w0, w1)


def _uneven_worker(rank, world, port, q):
    """9 agents on 2 ranks (5 + 4), local batch 20: rank 0 holds 20 transitions after 4 pushes per agent, rank 1 only
    after 5 -- the ranks must start learning on the SAME step (replicas diverged otherwise)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dmdqn_b200.parallel import shard_range
        lo, hi = shard_range(9, world, rank)
        n_local, b_local = hi - lo, 20
        state = {"pushed": 0, "theta": torch.zeros(4), "steps": 0, "calls": 0}

        def local_ready():
            return state["pushed"] * n_local >= b_local

        def local_grads(global_batch):
            assert local_ready(), "a rank was asked for gradients before its rings held its share of the batch"
            assert global_batch == b_local * world
            state["calls"] += 1
            return torch.full((4,), float(rank + 1)), torch.ones(1, 8)

        def apply(g):
            state["theta"] -= 0.1 * g
            state["steps"] += 1

        step = SharedParameterStep(local_grads, apply, b_local, local_ready=local_ready)
        trace = []
        for _ in range(8):
            state["pushed"] += 1
            out = step.step()
            trace.append(None if out is None else float(out[0]))
        q.put((rank, n_local, trace, state["steps"], state["theta"].tolist(), step.skipped))
    finally:
        dist.destroy_process_group()


def test_shared_parameter_learn_decision_is_collective_on_uneven_shards():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_uneven_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, n0, trace0, steps0, theta0, skipped0), (_, n1, trace1, steps1, theta1, skipped1) = results
    assert (n0, n1) == (5, 4)
    # rank 0 alone would have started at push 4 (5 * 4 = 20); together they start at push 5 (4 * 5 = 20)
    assert trace0 == trace1 == [None] * 4 + [2.0] * 4
    assert steps0 == steps1 == 4 and skipped0 == skipped1 == 4
    assert theta0 == theta1 == pytest.approx([-0.1 * 3 * 4] * 4)      # summed gradient 1 + 2, four steps
