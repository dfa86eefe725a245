"""Generate tests/golden/ref_*.npz by running the REFERENCE'S OWN PYTHON CODE.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs only in the build container
(needs /root/reference); the fixtures it writes are committed and are all the GPU box
ever sees.  Usage:  python -m oracle.make_golden

The reference (TensorFlow/Keras + traci + sumolib + wandb) cannot be installed here,
but the functions that pin this path's integer / byte / fp64 behaviour are plain
Python + numpy.  They are imported unmodified from /root/reference with the missing
third-party modules replaced by stubs:

  tensorflow      -> MagicMock, ``convert_to_tensor`` = ``np.asarray`` (dtype kept)
  traci           -> dmdqn_b200.sim.fake_traci.FakeTraci (seeded queue model)
  sumolib, wandb  -> MagicMock            log_config -> a null logger

What is executed for real, and which fixture it lands in:
  ref_featurize.npz  order_lanes.build_junction_lane_mapping / order_lanes_in_edge /
                     get_own_state / _get_neighbor_info / build_state_vector
                     (src/experimental/order_lanes.py:143-155,392-562)
  ref_episode_*.npz  the whole train.train_agents loop for one short episode
                     (src/scripts/train.py:182-316): featurise, eps-greedy (eps == 1),
                     ACTION_MAP, reward mix, remember -> replay-buffer contents
  ref_replay.npz     ReplayBuffer.add / sample incl. the reward z-score
                     (src/agents/dqn_agent.py:27-89) under ``random.seed``
  ref_epsilon.npz    the epsilon schedule and explore branch of DQNAgent.select_action
                     (src/agents/dqn_agent.py:246-265)
  ref_epsilon_linear.npz  the variant's linear epsilon decay and explore branch
                     (src/experimental/agent.py:121-146)
  ref_env_alt.npz    SumoTrafficEnvironment._get_local_observation / _get_observations /
                     _get_neighbor_presence_vector / _calculate_rewards (src/agents/sumo_env.py:532-679)
The MLP / learn arithmetic (TensorFlow ops) is NOT executed: MagicMock swallows it.
"""
from __future__ import annotations

import logging
import os
import random
import sys
import types
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
REPO = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def _install_stubs(fake):
    tf = MagicMock(name="tensorflow")
    tf.float32, tf.int32 = np.float32, np.int32
    tf.convert_to_tensor = lambda x, dtype=None: np.asarray(x, dtype=dtype)
    tf.config.list_physical_devices.return_value = []
    tf.function = lambda *a, **k: (lambda f: f)
    for name in ("tensorflow", "tensorflow.keras", "tensorflow.keras.layers",
                 "tensorflow.keras.initializers"):
        sys.modules[name] = tf if name == "tensorflow" else getattr(tf, name.split(".", 1)[1].replace(".", "_"), MagicMock())
    sys.modules["tensorflow.keras"] = tf.keras
    sys.modules["tensorflow.keras.layers"] = tf.keras.layers
    sys.modules["wandb"] = MagicMock(name="wandb")
    sumolib = MagicMock(name="sumolib")
    sys.modules["sumolib"] = sumolib
    sys.modules["sumolib.net"] = sumolib.net
    log_config = types.ModuleType("log_config")
    log_config.logger = logging.getLogger("dmdqn_null")
    log_config.logger.addHandler(logging.NullHandler())
    log_config.logger.propagate = False
    sys.modules["log_config"] = log_config
    sys.modules["traci"] = fake
    return tf, sumolib


def _lane_table(fake):
    """[N,12] lane ids (or None) in the device layout slot = dir*3 + lane."""
    from dmdqn_b200.sim.fake_traci import DIRS
    table = []
    for r in range(fake.rows):
        for c in range(fake.cols):
            row = []
            for d in DIRS:
                ids = fake._lanes[(r, c, d)]
                row += ids + [None] * (3 - len(ids))
            table.append(row)
    return table


def _readings(fake, table):
    n = len(table)
    halting = np.full((n, 12), -1, np.int32)
    for j, row in enumerate(table):
        for k, lid in enumerate(row):
            if lid is not None:
                halting[j, k] = fake.queue[lid]
    jids = fake.junction_ids
    phase = np.array([fake.phase[j] for j in jids], np.int32)
    nsw = np.array([fake.trafficlight.getNextSwitch(j) for j in jids], np.float64)
    dur = np.array([fake.trafficlight.getPhaseDuration(j) for j in jids], np.float64)
    return halting, phase, nsw, dur


def golden_featurize(order_lanes, fake_cls):
    """Random readings on a 4x4 grid with missing lanes, all 12 SUMO phases, and
    negative / positive time_spent, through the reference's own functions."""
    rng = np.random.default_rng(1234)
    out = {}
    for tag, live in (("shipped", False), ("live", True)):
        drop = {(0, 0, "n", 2), (1, 2, "e", 1), (1, 2, "e", 2), (3, 3, "w", 0), (3, 3, "w", 1), (3, 3, "w", 2)}
        fake = fake_cls(rows=4, cols=4, seed=7, drop=drop, junction_get_type=live)
        sys.modules["traci"] = fake
        order_lanes.traci = fake
        lanes = [l for l in fake.lane.getIDList() if not l.startswith(":")]
        jmap = order_lanes.order_lanes_in_edge(
            order_lanes.build_junction_lane_mapping(fake.junction_ids, lanes))
        table = _lane_table(fake)
        cases = []
        for case in range(6):
            for lid in fake.queue:
                fake.queue[lid] = int(rng.integers(0, 25))
            now = float(rng.integers(0, 2000)) + 0.5 * case
            fake.time = now
            fake.phase = {j: int(rng.integers(0, 12)) for j in fake.junction_ids}
            fake.phase_duration = {j: float(rng.integers(5, 60)) for j in fake.junction_ids}
            fake.next_switch_override = {j: now + float(rng.integers(0, 70)) - 3.0 for j in fake.junction_ids}
            halting, phase, nsw, dur = _readings(fake, table)
            own = np.array([order_lanes.get_own_state(j, jmap, 3, now) for j in fake.junction_ids], np.float64)
            gstate = {j: list(own[i]) for i, j in enumerate(fake.junction_ids)}
            obs = np.array([order_lanes.build_state_vector(j, fake.junction_ids, jmap, 3, now, gstate)
                            for j in fake.junction_ids], np.float64)
            pres = np.array([order_lanes._get_neighbor_info(j, fake.junction_ids)[0] for j in fake.junction_ids], np.int32)
            cases.append((halting, phase, nsw, dur, now, own, obs, pres))
        for k, name in enumerate(("halting", "phase", "next_switch", "phase_dur", "sim_time", "own", "obs", "presence")):
            out[f"{tag}_{name}"] = np.stack([np.asarray(c[k]) for c in cases])
        out[f"{tag}_signal_valid"] = np.full((16,), int(live), np.uint8)
        # invalid junction id -> 89 x -1.0 (order_lanes.py:519-524)
        out[f"{tag}_invalid"] = np.asarray(order_lanes.build_state_vector("X_9", fake.junction_ids, jmap, 3, 0.0, {}), np.float64)
    np.savez_compressed(os.path.join(OUT, "ref_featurize.npz"), **out)
    print("ref_featurize.npz", {k: v.shape for k, v in out.items() if k.startswith("live")})


def golden_episode(train, order_lanes, fake_cls, sumolib, live: bool, tag: str):
    """One short episode of the reference's train loop under the fake TraCI."""
    fake = fake_cls(rows=3, cols=3, seed=11, junction_get_type=live)
    sys.modules["traci"] = fake
    train.traci = fake
    order_lanes.traci = fake
    nodes = []
    for j in fake.junction_ids:
        node = MagicMock()
        node.getType.return_value = "traffic_light"
        node.getID.return_value = j
        nodes.append(node)
    sumolib.net.readNet.return_value.getNodes.return_value = nodes
    order_lanes.sumolib = sumolib
    train.EPISODES = 1
    train.MAX_SIM_TIME = 600          # 60 RL steps (reference: 2400 -> 240)
    table = _lane_table(fake)
    stash, trace = {}, {"own": [], "time": [], "halting": [], "phase": [], "nsw": [], "dur": []}
    orig_create, orig_gos = train.create_agents, train.get_own_state

    def create_agents(tl):
        agents = orig_create(tl)
        stash.update(agents)
        return agents

    def get_own_state(junction_id, structured_junction_lane_map, max_lanes_per_direction, current_sim_time):
        if junction_id == fake.junction_ids[0]:
            h, p, s, d = _readings(fake, table)
            trace["halting"].append(h); trace["phase"].append(p); trace["nsw"].append(s); trace["dur"].append(d)
            trace["time"].append(current_sim_time)
            trace["own"].append([])
        blk = orig_gos(junction_id=junction_id, structured_junction_lane_map=structured_junction_lane_map,
                       max_lanes_per_direction=max_lanes_per_direction, current_sim_time=current_sim_time)
        trace["own"][-1].append(list(blk))
        return blk

    train.create_agents, train.get_own_state = create_agents, get_own_state
    np.random.seed(5)
    cwd = os.getcwd()
    os.chdir(REF)                     # SUMO_NET_PATH is relative (train.py:51); nothing is written
    try:
        train.train_agents()
    finally:
        os.chdir(cwd)
        train.create_agents, train.get_own_state = orig_create, orig_gos
    jids = fake.junction_ids
    bufs = [list(stash[j].replay_buffer.buffer) for j in jids]
    t = len(bufs[0])
    out = {
        "s": np.array([[b[i][0] for i in range(t)] for b in bufs], np.float32),
        "a": np.array([[b[i][1] for i in range(t)] for b in bufs], np.int32),
        "r": np.array([[b[i][2] for i in range(t)] for b in bufs], np.float64),
        "s2": np.array([[b[i][3] for i in range(t)] for b in bufs], np.float32),
        "done": np.array([[int(bool(b[i][4])) for i in range(t)] for b in bufs], np.uint8),
        "halting": np.array(trace["halting"], np.int32),
        "phase": np.array(trace["phase"], np.int32),
        "next_switch": np.array(trace["nsw"], np.float64),
        "phase_dur": np.array(trace["dur"], np.float64),
        "sim_time": np.array(trace["time"], np.float64),
        "own": np.array(trace["own"], np.float64),
        "signal_valid": np.full((9,), int(live), np.uint8),
        "grid": np.array([3, 3], np.int32),
        "fake_seed": np.array(11), "np_seed": np.array(5), "max_sim_time": np.array(600),
    }
    np.savez_compressed(os.path.join(OUT, f"ref_episode_{tag}.npz"), **out)
    print(f"ref_episode_{tag}.npz", {k: v.shape for k, v in out.items()})


def golden_replay(dqn_agent):
    """ReplayBuffer.add/sample (dqn_agent.py:27-89) under random.seed: wrap-around,
    both CPython sampling paths (pool: n <= 21+4^ceil(log4 3B); set: n larger)."""
    out = {}
    cases = [("pool", 100, 150, 32, 3), ("set", 400, 600, 16, 4), ("wrap", 300, 700, 128, 5),
             ("exact", 64, 64, 64, 6), ("short", 100, 10, 32, 7), ("const", 200, 200, 16, 8)]
    for tag, cap, n_add, batch, seed in cases:
        rng = np.random.default_rng(seed)
        buf = dqn_agent.ReplayBuffer(cap)
        ss = rng.integers(0, 20, (n_add, 1, 89)).astype(np.float32)
        ss[:, 0, 0] = np.arange(n_add)                      # transition id in column 0
        s2 = rng.integers(0, 20, (n_add, 1, 89)).astype(np.float32)
        acts = rng.integers(0, 4, n_add)
        rews = -0.3 * rng.integers(0, 200, n_add) - 0.7 * rng.integers(0, 2000, n_add)
        if tag == "const":
            rews[:] = -12.5                                 # std == 0 -> r_hat == 0
        dones = rng.random(n_add) < 0.05
        for i in range(n_add):
            buf.add((ss[i], int(acts[i]), float(rews[i]), s2[i], bool(dones[i])))
        random.seed(seed)
        res = buf.sample(batch)
        out[f"{tag}_meta"] = np.array([cap, n_add, batch, seed], np.int64)
        out[f"{tag}_in_s"], out[f"{tag}_in_s2"] = ss[:, 0].astype(np.int16), s2[:, 0].astype(np.int16)  # small ints
        out[f"{tag}_in_a"], out[f"{tag}_in_r"], out[f"{tag}_in_d"] = acts.astype(np.int32), rews, dones.astype(np.uint8)
        out[f"{tag}_len"] = np.array(len(buf))
        if res is None:
            out[f"{tag}_none"] = np.array(1)
            continue
        for name, arr in zip(("s", "a", "r", "s2", "d"), res):
            out[f"{tag}_out_{name}"] = np.asarray(arr)
    np.savez_compressed(os.path.join(OUT, "ref_replay.npz"), **out)
    print("ref_replay.npz", sorted(k for k in out if k.endswith("_meta")))


def golden_epsilon(dqn_agent):
    """select_action (dqn_agent.py:246-274): epsilon after the call and whether the
    explore branch ran, over a sweep of global_step_count."""
    cfg = {"epsilon_start": 1.0, "epsilon_min": 0.05}
    agent = dqn_agent.DQNAgent(89, 4, "J_0_0", cfg)
    steps = np.concatenate([np.arange(0, 8000, 997), np.arange(7990, 8010), np.arange(8000, 120000, 1499)])
    np.random.seed(3)
    eps, explored, action = [], [], []
    for g in steps:
        agent.global_step_count = int(g)
        act = agent.select_action(np.zeros((1, 89), np.float32))
        eps.append(agent.epsilon)
        is_explore = isinstance(act, (int, np.integer))
        explored.append(is_explore)
        action.append(int(act) if is_explore else -1)
    np.random.seed(3)                                       # the uniform stream the agent consumed
    u_stream = np.random.rand(4096)
    np.savez_compressed(os.path.join(OUT, "ref_epsilon.npz"), steps=steps, eps=np.array(eps, np.float64),
                        explored=np.array(explored), action=np.array(action, np.int32),
                        epsilon_min=np.array(0.05), u_check=u_stream[:4])
    print("ref_epsilon.npz", len(steps), "steps; eps range", min(eps), max(eps))


def golden_epsilon_linear(exp_agent):
    """The variant's select_action (src/experimental/agent.py:121-146): linear epsilon decay applied after
    every action, explore branch under ``np.random``.  The greedy branch returns a MagicMock (TensorFlow is
    stubbed) and is recorded as -1."""
    cfg = {"epsilon_start": 1.0, "epsilon_min": 0.05, "epsilon_decay_steps": 400}
    agent = exp_agent.DQNAgent(89, 4, "J_0_0", cfg)
    np.random.seed(11)
    eps, explored, action = [], [], []
    for _ in range(600):
        act = agent.select_action(np.zeros((1, 89), np.float32))
        is_explore = isinstance(act, (int, np.integer))
        eps.append(agent.epsilon)
        explored.append(is_explore)
        action.append(int(act) if is_explore else -1)
    np.savez_compressed(os.path.join(OUT, "ref_epsilon_linear.npz"), eps=np.array(eps, np.float64),
                        explored=np.array(explored), action=np.array(action, np.int32),
                        epsilon_start=np.array(1.0), epsilon_min=np.array(0.05), epsilon_decay_steps=np.array(400),
                        decay_rate=np.array(agent.epsilon_decay_rate, np.float64))
    print("ref_epsilon_linear.npz", len(eps), "calls; eps range", min(eps), max(eps), "explored", int(np.sum(explored)))


def golden_env_alt(FakeTraci):
    """ref_env_alt.npz: SumoTrafficEnvironment._get_local_observation / _get_observations /
    _get_neighbor_presence_vector / _calculate_rewards (src/agents/sumo_env.py:532-679) run unmodified on an
    instance whose network-derived attributes (lane order, neighbour map: the sumolib part of __init__) are set
    by hand for a 3x4 grid with one PAD lane block, one unreadable lane and one junction without a signal."""
    import tempfile
    home = tempfile.mkdtemp()
    os.makedirs(os.path.join(home, "tools"), exist_ok=True)
    os.environ["SUMO_HOME"] = home                                    # the module exits at import without it
    sys.modules["tensorflow.keras.models"] = MagicMock(name="tensorflow.keras.models")   # imported, unused on this path
    import src.agents.sumo_env as sumo_env                            # noqa: E402  (the reference, unmodified)
    from dmdqn_b200.sim.fake_traci import DIRS
    rows, cols = 3, 4
    fake = FakeTraci(rows=rows, cols=cols, seed=5, arrival_rate=0.3)
    sumo_env.traci = fake
    fake.start(["sumo"])
    ids = fake.junction_ids
    n = len(ids)
    env = object.__new__(sumo_env.SumoTrafficEnvironment)
    env.max_lanes_per_direction, env.padding_value = 3, -1.0
    env.neighbor_info_size, env.state_vector_size = 14, 74
    env.controlled_intersection_ids = list(ids)
    no_signal = ids[5]
    env.traffic_light_ids = {j: j for j in ids if j != no_signal}
    pos = {f"J_{r}_{c}": (r, c) for r in range(rows) for c in range(cols)}
    step = {"N": (-1, 0), "E": (0, 1), "S": (1, 0), "W": (0, -1)}
    env.neighbor_map, nbr_idx = {}, np.full((n, 4), -1, np.int32)
    for a, j in enumerate(ids):
        r, c = pos[j]
        env.neighbor_map[j] = {}
        for k, d in enumerate("NESW"):
            rr, cc = r + step[d][0], c + step[d][1]
            if 0 <= rr < rows and 0 <= cc < cols:
                env.neighbor_map[j][d] = f"J_{rr}_{cc}"
                nbr_idx[a, k] = ids.index(f"J_{rr}_{cc}")
    env.observed_lanes, lane_tab = {}, []
    for a, j in enumerate(ids):
        r, c = pos[j]
        lanes = []
        for d in "nesw":
            lanes += list(fake._lanes[(r, c, d)])
        if a == 2:
            lanes[3:6] = ["PAD_E_0", "PAD_E_1", "PAD_E_2"]                # a missing approach (:545-549)
        if a == 7:
            lanes[10] = "no_such_lane"                                   # TraCIException -> keeps -1.0 (:555-557)
        env.observed_lanes[j] = lanes
        lane_tab.append(lanes)
    T = 6
    halting = np.zeros((T, n, 12), np.int32); phase = np.zeros((T, n), np.int32); nsw = np.zeros((T, n), np.float64)
    times = np.zeros(T); obs_all = np.zeros((T, n, 74), np.float32); rew_all = np.zeros((T, n), np.float64)
    prev = None
    rng = np.random.default_rng(1)
    for t in range(T):
        for _ in range(7):
            fake.simulationStep()
        for j in ids:
            if rng.random() < 0.5:
                fake.trafficlight.setPhase(j, int(rng.integers(0, 12)))
        env.current_time = fake.time
        times[t] = fake.time
        for a, j in enumerate(ids):
            for e, lid in enumerate(lane_tab[a]):
                halting[t, a, e] = -2 if lid.startswith("PAD_") else (fake.queue[lid] if lid in fake.queue else -1)
            phase[t, a] = fake.phase[j]; nsw[t, a] = fake.trafficlight.getNextSwitch(j)
        obs = env._get_observations()
        rew = env._calculate_rewards(prev, obs, None)
        for a, j in enumerate(ids):
            obs_all[t, a] = obs[j]; rew_all[t, a] = float(rew[j])
        prev = obs
    valid = np.array([j != no_signal for j in ids], np.uint8)
    np.savez_compressed(os.path.join(OUT, "ref_env_alt.npz"), halting=halting, phase=phase, next_switch=nsw, sim_time=times,
                        signal_valid=valid, nbr_idx=nbr_idx, obs=obs_all, reward=rew_all)
    print("ref_env_alt.npz", obs_all.shape, "rewards", rew_all.min(), rew_all.max())


def main():
    if not os.path.isdir(REF):
        raise SystemExit("/root/reference is not present: goldens can only be made in the build container")
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, REPO)
    from dmdqn_b200.sim.fake_traci import FakeTraci
    fake = FakeTraci()
    tf, sumolib = _install_stubs(fake)
    sys.path.insert(0, REF)
    import src.experimental.order_lanes as order_lanes   # noqa: E402  (the reference, unmodified)
    import src.agents.dqn_agent as dqn_agent              # noqa: E402
    import src.scripts.train as train                     # noqa: E402
    golden_featurize(order_lanes, FakeTraci)
    golden_episode(train, order_lanes, FakeTraci, sumolib, live=False, tag="shipped")
    golden_episode(train, order_lanes, FakeTraci, sumolib, live=True, tag="live")
    golden_replay(dqn_agent)
    golden_epsilon(dqn_agent)
    import src.experimental.agent as exp_agent            # noqa: E402  (the older variant, unmodified)
    golden_epsilon_linear(exp_agent)
    golden_env_alt(FakeTraci)


if __name__ == "__main__":
    main()
