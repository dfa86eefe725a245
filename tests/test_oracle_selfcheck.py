"""Cross-checks for the unpinned (TensorFlow/Keras) part of the oracle and for the new
index contracts.  CPU only."""
import random
from collections import deque

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import replay as R
from oracle.dqn import (OracleDQNAgent, StackedOracle, adam_scalars, adam_update_, init_params,
                        loss_and_grad, mlp_forward)


def test_manual_adam_torch_form_equals_torch_optim():
    torch.manual_seed(0)
    p = torch.randn(37, 11)
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=5e-4, betas=(0.9, 0.999), eps=1e-8)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t in range(1, 30):
        g = torch.randn_like(p) * (0.1 if t % 3 else 1e-4)
        ref.grad = g.clone()
        opt.step()
        alpha, eps = adam_scalars(t, 5e-4, "torch")
        adam_update_(p, g, m, v, alpha, eps)
        np.testing.assert_allclose(p.numpy(), ref.detach().numpy(), rtol=2e-6, atol=1e-8)


def test_keras_form_differs_from_torch_form_on_small_gradients():
    """SURVEY 'Adam form' hard part: eps placement matters when |g| ~ eps."""
    p1, p2 = torch.zeros(4), torch.zeros(4)
    g = torch.full((4,), 1e-7)
    for form, p in (("keras", p1), ("torch", p2)):
        m, v = torch.zeros(4), torch.zeros(4)
        alpha, eps = adam_scalars(1, 1e-3, form)
        adam_update_(p, g, m, v, alpha, eps)
    assert abs(float(p1[0]) / float(p2[0]) - 1) > 0.1


@pytest.mark.parametrize("kind", ["mse", "huber"])
def test_hand_derived_backward_equals_autograd(kind):
    """SURVEY App. A.8 formulas (what the CUDA kernel implements) vs torch.autograd."""
    torch.manual_seed(1)
    d, h, a, b = 89, 64, 4, 48
    params = [p.double() for p in init_params(3, d, [h, h], a)]
    params[1] += 0.1
    x = torch.randn(b, d, dtype=torch.float64)
    act = torch.randint(0, a, (b,))
    y = torch.randn(b, dtype=torch.float64) * 2
    ps = [p.clone().requires_grad_(True) for p in params]
    q, acts = mlp_forward(ps, x, keep=True)
    pred = q[torch.arange(b), act]
    terms, gpred = loss_and_grad(pred, y, kind)
    grads = torch.autograd.grad(terms.mean(), ps)
    W1, b1, W2, b2, W3, b3 = params
    h1, h2 = acts[1].detach(), acts[2].detach()
    dq = torch.zeros(b, a, dtype=torch.float64)
    dq[torch.arange(b), act] = gpred.detach()
    dW3, db3 = h2.T @ dq, dq.sum(0)
    dh2 = (dq @ W3.T) * (h2 > 0)
    dW2, db2 = h1.T @ dh2, dh2.sum(0)
    dh1 = (dh2 @ W2.T) * (h1 > 0)
    dW1, db1 = x.T @ dh1, dh1.sum(0)
    for mine, ref in zip((dW1, db1, dW2, db2, dW3, db3), grads):
        np.testing.assert_allclose(mine.numpy(), ref.numpy(), rtol=1e-10, atol=1e-12)


def test_stacked_oracle_equals_per_agent_loop():
    n, d, h, a, b = 5, 89, 64, 4, 32
    rng = np.random.default_rng(0)
    cfg = dict(learning_rate=5e-4, gamma=0.99, nn_layers=[h, h], batch_size=b, target_update_frequency=3)
    agents = [OracleDQNAgent(d, a, f"J{i}", dict(cfg, seed=100 + i)) for i in range(n)]
    stk = StackedOracle(n, d, [h, h], a, learning_rate=5e-4, target_update_frequency=3, seed0=100)
    for step in range(7):
        s = rng.integers(0, 20, (n, b, d)).astype(np.float32)
        s2 = rng.integers(0, 20, (n, b, d)).astype(np.float32)
        act = rng.integers(0, a, (n, b)).astype(np.int32)
        r = rng.standard_normal((n, b)).astype(np.float32)
        dn = (rng.random((n, b)) < 0.1).astype(np.float32)
        out = stk.learn_on_batch(s, act, r, s2, dn)
        for i, ag in enumerate(agents):
            loss = ag.learn_on_batch(s[i], act[i], r[i], s2[i], dn[i])
            assert abs(loss - out["loss"][i]) <= 2e-5 * abs(loss)
            for k in range(6):
                np.testing.assert_allclose(stk.online[k][i].numpy(), ag.online[k].numpy(), rtol=2e-5, atol=2e-6)
                np.testing.assert_allclose(stk.target[k][i].numpy(), ag.target[k].numpy(), rtol=2e-5, atol=2e-6)


def test_learn_semantics_none_then_loss_and_hard_sync_after_increment():
    cfg = dict(nn_layers=[64, 64], batch_size=8, replay_buffer_size=50, target_update_frequency=2, seed=1)
    ag = OracleDQNAgent(89, 4, "J", cfg, rng=random.Random(0))
    rng = np.random.default_rng(0)
    assert ag.learn() is None and ag.replay() == 0
    for i in range(8):
        ag.remember(rng.random((1, 89)).astype(np.float32), int(i % 4), float(-i), rng.random((1, 89)).astype(np.float32), False)
    assert ag.global_step_count == 0          # remember() does not count (dqn_agent.py:312-325)
    ag.store_experience((rng.random((1, 89)).astype(np.float32), 1, -1.0, rng.random((1, 89)).astype(np.float32), True))
    assert ag.global_step_count == 1
    l1 = ag.replay()
    assert l1 > 0 and ag.learn_step_counter == 1
    assert not torch.equal(ag.online[0], ag.target[0])
    ag.replay()
    assert ag.learn_step_counter == 2 and all(torch.equal(o, t) for o, t in zip(ag.online, ag.target))


@settings(max_examples=60, deadline=None)
@given(cap=st.integers(1, 40), n=st.integers(0, 130))
def test_ring_mapping_equals_deque(cap, n):
    dq = deque(maxlen=cap)
    ring = np.full((cap,), -1)
    for i in range(n):
        dq.append(i)
        ring[i % cap] = i
    size = min(n, cap)
    assert len(dq) == size
    if size:
        phys = R.ring_physical(n, cap, np.arange(size))
        assert list(ring[phys]) == list(dq)


@settings(max_examples=60, deadline=None)
@given(size=st.integers(1, 3000), frac=st.floats(0.01, 1.0), seed=st.integers(0, 2**31))
def test_fisher_yates_contract(size, frac, seed):
    b = max(1, int(size * frac))
    w = np.random.default_rng(seed).integers(0, 2**32, b, dtype=np.uint64).astype(np.uint32)
    idx = R.fisher_yates_indices(w, size)
    assert len(set(idx.tolist())) == b and idx.min() >= 0 and idx.max() < size
    # dense-pool restatement of CPython's pool path with randbelow := mulhi
    pool = list(range(size))
    dense = []
    for i in range(b):
        j = (int(w[i]) * (size - i)) >> 32
        dense.append(pool[j])
        pool[j] = pool[size - i - 1]
    assert dense == idx.tolist()


def test_zscore_canonical_close_to_numpy_and_handles_constant():
    rng = np.random.default_rng(0)
    for b in (1, 7, 32, 64, 100, 256, 1000):
        r = -0.3 * rng.integers(0, 200, (3, b)) - 0.7 * rng.integers(0, 5000, (3, b))
        can = R.zscore_canonical(r)
        for i in range(3):
            np.testing.assert_allclose(can[i], R.zscore(r[i]), rtol=1e-9, atol=1e-9)
    assert np.all(R.zscore_canonical(np.full((2, 64), -3.7)) == 0)


def test_explore_and_random_action_contract():
    w = np.array([0, 1, 2**31, 2**32 - 1], np.uint32)
    assert R.explore_decision(w, 1.0).all() and not R.explore_decision(w, 0.0).any()
    assert list(R.explore_decision(w, 0.5)) == [True, True, False, False]
    assert list(R.random_action(w, 4)) == [0, 0, 2, 3]
