"""The oracle against fixtures produced by the reference's own code
(oracle/make_golden.py -> tests/golden/ref_*.npz).  CPU only."""
import os
import random

import numpy as np
import pytest

from oracle import featurize as F
from oracle import replay as R
from oracle.dqn import OracleDQNAgent, epsilon_linear_step, epsilon_schedule

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("tag", ["shipped", "live"])
def test_featurize_matches_reference(tag):
    z = np.load(os.path.join(G, "ref_featurize.npz"))
    nbr = F.grid_neighbors(4, 4)
    for c in range(z[f"{tag}_halting"].shape[0]):
        own = F.own_state(z[f"{tag}_halting"][c], z[f"{tag}_phase"][c], z[f"{tag}_next_switch"][c],
                          z[f"{tag}_phase_dur"][c], z[f"{tag}_sim_time"][c], z[f"{tag}_signal_valid"])
        assert np.array_equal(own, z[f"{tag}_own"][c])
        obs = F.build_obs(own, nbr)
        assert np.array_equal(obs, z[f"{tag}_obs"][c].astype(np.float32))
        assert np.array_equal((nbr >= 0).astype(np.int32), z[f"{tag}_presence"][c])
    assert np.all(z[f"{tag}_invalid"] == -1.0) and z[f"{tag}_invalid"].shape == (89,)
    if tag == "live":  # the fixture must exercise both time_spent branches and dead phases
        assert (z["live_own"][..., 16] == -1.0).any() and (z["live_own"][..., 16] >= 0).any()
        assert (z["live_own"][..., 12:16].sum(-1) == 0).any()
        assert (z["live_halting"] == -1).any()


@pytest.mark.parametrize("tag", ["shipped", "live"])
def test_episode_trace_matches_reference_train_loop(tag):
    """Array featuriser + reward restatement reproduce what the reference's
    train_agents loop stored in its replay buffers (train.py:182-316)."""
    z = np.load(os.path.join(G, f"ref_episode_{tag}.npz"))
    rows, cols = z["grid"]
    nbr = F.grid_neighbors(int(rows), int(cols))
    t_steps = z["s"].shape[1]
    assert z["halting"].shape[0] == t_steps + 1
    for t in range(t_steps + 1):
        own = F.own_state(z["halting"][t], z["phase"][t], z["next_switch"][t], z["phase_dur"][t],
                          z["sim_time"][t], z["signal_valid"])
        assert np.array_equal(own, z["own"][t])
        obs = F.build_obs(own, nbr)
        if t < t_steps:
            assert np.array_equal(obs, z["s"][:, t])
            rew, _ = F.rewards(own)
            assert np.array_equal(rew, z["r"][:, t])        # float64, bit-exact
        if t > 0:
            assert np.array_equal(obs, z["s2"][:, t - 1])
    # done only on the last step (train.py:232-236), actions in range
    assert z["done"][:, :-1].sum() == 0 and z["done"][:, -1].all()
    assert z["a"].min() >= 0 and z["a"].max() <= 3


def _fill(buf, z, tag):
    s = z[f"{tag}_in_s"].astype(np.float32)
    s2 = z[f"{tag}_in_s2"].astype(np.float32)
    for i in range(s.shape[0]):
        buf.add((s[i][None], int(z[f"{tag}_in_a"][i]), float(z[f"{tag}_in_r"][i]), s2[i][None],
                 bool(z[f"{tag}_in_d"][i])))


@pytest.mark.parametrize("tag", ["pool", "set", "wrap", "exact", "short", "const"])
def test_replay_matches_reference(tag):
    z = np.load(os.path.join(G, "ref_replay.npz"))
    cap, n_add, batch, seed = (int(x) for x in z[f"{tag}_meta"])
    buf = R.FaithfulReplayBuffer(cap, random.Random(seed))
    _fill(buf, z, tag)
    assert len(buf) == int(z[f"{tag}_len"]) == min(cap, n_add)
    out = buf.sample(batch)
    if f"{tag}_none" in z:
        assert out is None
        return
    for name, arr in zip(("s", "a", "r", "s2", "d"), out):
        ref = z[f"{tag}_out_{name}"]
        assert arr.dtype == ref.dtype and np.array_equal(arr, ref), name

    # the device-layout ring fed with reference-exact indices gives the same batch
    ring = R.RingReplay(1, cap, 89)
    s = z[f"{tag}_in_s"].astype(np.float32)
    s2 = z[f"{tag}_in_s2"].astype(np.float32)
    for i in range(n_add):
        ring.push(s[i][None], z[f"{tag}_in_a"][i:i + 1], z[f"{tag}_in_r"][i:i + 1], s2[i][None],
                  z[f"{tag}_in_d"][i:i + 1])
    idx = R.cpython_sample_indices(random.Random(seed), min(cap, n_add), batch)
    got = ring.gather(0, idx, canonical=False)
    for name, arr in zip(("s", "a", "r", "s2", "d"), got):
        assert np.array_equal(arr, z[f"{tag}_out_{name}"]), name
    # canonical (GPU-order) z-score agrees with numpy's to fp64 round-off
    can = ring.gather(0, idx, canonical=True)[2]
    np.testing.assert_allclose(can, z[f"{tag}_out_r"], rtol=0, atol=1e-6)
    if tag == "const":
        assert np.all(z[f"{tag}_out_r"] == 0)


def test_epsilon_schedule_matches_reference():
    z = np.load(os.path.join(G, "ref_epsilon.npz"))
    eps = 1.0
    np.random.seed(3)
    for g, e_ref, explored in zip(z["steps"], z["eps"], z["explored"]):
        eps = epsilon_schedule(int(g), eps, float(z["epsilon_min"]))
        assert eps == e_ref
        u = np.random.rand()
        assert (u < eps) == bool(explored)
        if explored:
            np.random.randint(0, 4)
    assert z["eps"].min() < 0.05 and z["eps"].max() == 1.0


def test_linear_epsilon_variant_matches_reference():
    """src/experimental/agent.py:121-146 (oracle/make_golden.py golden_epsilon_linear): the oracle agent with
    epsilon_schedule='linear' consumes np.random like the variant and lands on the same epsilons / explore actions."""
    z = np.load(os.path.join(G, "ref_epsilon_linear.npz"))
    cfg = {"epsilon_start": float(z["epsilon_start"]), "epsilon_min": float(z["epsilon_min"]),
           "epsilon_decay_steps": int(z["epsilon_decay_steps"]), "epsilon_schedule": "linear", "nn_layers": [64, 64]}
    agent = OracleDQNAgent(89, 4, "J_0_0", cfg)
    assert agent.epsilon_decay_rate == float(z["decay_rate"])
    np.random.seed(11)
    greedy = 0
    for e_ref, explored, action in zip(z["eps"], z["explored"], z["action"]):
        a = agent.select_action(np.zeros((1, 89), np.float32))
        assert agent.epsilon == e_ref
        if explored:
            assert a == int(action)
        else:
            greedy += 1
            assert a == agent.select_greedy_action(np.zeros((1, 89), np.float32))
    assert greedy > 100
    assert epsilon_linear_step(0.05, 0.05, 0.1) == 0.05 and epsilon_linear_step(0.06, 0.05, 0.1) == 0.05


def test_alt_env_contract_matches_reference_fixture():
    """SumoTrafficEnvironment's 74-dim observation and queue-reduction reward (sumo_env.py:532-679), incl. a PAD
    approach, an unreadable lane and a junction without a signal."""
    from oracle import featurize as F
    z = np.load(os.path.join(G, "ref_env_alt.npz"))
    prev = None
    for t in range(z["obs"].shape[0]):
        own = F.own_state_alt(z["halting"][t], z["phase"][t], z["next_switch"][t], z["sim_time"][t], z["signal_valid"])
        obs = F.build_obs_alt(own, z["nbr_idx"])
        assert obs.dtype == np.float32 and np.array_equal(obs, z["obs"][t])
        rew = np.zeros(own.shape[0]) if prev is None else F.rewards_alt(prev, own)
        assert np.array_equal(rew, z["reward"][t])
        prev = own
    assert (z["obs"][:, 2, 3:6] == 0).all() and (z["obs"][:, 7, 10] == -1).all() and (z["obs"][:, 5, 12:14] == -1).all()
