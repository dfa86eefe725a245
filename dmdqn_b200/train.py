"""Drop-in training loop: the reference's src/scripts/train.py rebuilt on the batched
device path.

Same steps per RL step as train.py:207-310 -- observe, epsilon-greedy act, setPhase with
ACTION_MAP, STEP_DURATION simulation seconds, reward from the PRE-step state, next
observation, remember, replay -- but each is ONE native call over all intersections
instead of a Python loop over agents, and the TraCI readings are taken once per lane per
step (the reference re-reads every junction three times, 324 socket round trips per step).
``mode="per_agent"`` keeps the reference's per-agent call pattern through the DQNAgent
facade instead.  SUMO stays the environment: with ``traci`` importable it is used as is;
otherwise (this image) the seeded fake in dmdqn_b200/sim/fake_traci.py stands in.

    python -m dmdqn_b200.train --episodes 1 --max-sim-time 600
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys
import time

import numpy as np
import torch
import yaml

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
AGENT_CONFIG_PATH = os.path.join(ROOT, "config", "agent_config.yaml")
ENV_CONFIG_PATH = os.path.join(ROOT, "config", "env_config.yaml")

EPISODES = 100                      # train.py:54-58
MAX_LANES_PER_DIRECTION = 3
STEP_DURATION = 10.0
ACTION_MAP = {0: 0, 1: 3, 2: 6, 3: 9}
MAX_SIM_TIME = 2400

PRESETS = {
    "shipped_train_py": {"learning_rate": 0.001, "gamma": 0.99, "epsilon_start": 1.0, "epsilon_min": 0.01,
                         "epsilon_decay_steps": 200000, "replay_buffer_size": 10000, "batch_size": 128,
                         "target_update_frequency": 500, "nn_layers": [128, 128]},   # train.py:111-121
}


def set_seeds(seed_value: int) -> None:            # train.py:72-79
    random.seed(seed_value)
    np.random.seed(seed_value)
    torch.manual_seed(seed_value)
    os.environ["PYTHONHASHSEED"] = str(seed_value)


def load_config(agent_yaml_path: str = AGENT_CONFIG_PATH, env_yaml_path: str = ENV_CONFIG_PATH) -> dict:
    """train.py:82-96 (defined there but never called; here it is the default)."""
    for p in (agent_yaml_path, env_yaml_path):
        if not os.path.exists(p):
            raise FileNotFoundError(f"config file not found: {p}")
    with open(agent_yaml_path) as f:
        agent_config = yaml.safe_load(f)
    with open(env_yaml_path) as f:
        env_config = yaml.safe_load(f)
    return {**env_config, **agent_config}


def calculate_local_reward(current_state, next_state=None):          # train.py:159-160
    return -1.0 * sum(current_state[:12])


def calculate_global_reward(global_state: dict, next_global_state: dict | None = None):   # train.py:163-165
    return -1.0 * sum(sum(state[:12]) for state in global_state.values())


def calculate_rewards(junction_id, global_state, next_global_state, alpha, beta):          # train.py:168-179
    return alpha * calculate_local_reward(global_state[junction_id]) + beta * calculate_global_reward(global_state)


class SmoothedValue:                                                  # train.py:144-156
    def __init__(self, alpha=0.5):
        self.alpha, self.value = alpha, None

    def update(self, new_val):
        self.value = new_val if self.value is None else self.alpha * new_val + (1 - self.alpha) * self.value

    def get_value(self):
        return self.value


# ------------------------------------------------------------------------------------------
def get_traci(config: dict, seed: int = 0):
    backend = config.get("backend", "auto")
    if backend in ("auto", "traci"):
        try:
            import traci  # noqa: F401
            return traci, False
        except ImportError:
            if backend == "traci":
                raise
    from .sim.fake_traci import FakeTraci
    return FakeTraci(rows=int(config.get("grid_rows", 3)), cols=int(config.get("grid_cols", 3)),
                     seed=int(config.get("fake_traci_seed", seed)),
                     max_sim_time=float(config.get("max_sim_time", MAX_SIM_TIME))), True


def initialize_environment(traci, config: dict):
    """train.py:99-106: start SUMO, list the traffic-light junctions and their incoming
    lanes.  Returns (tl_junctions, lane_table[N][12] of lane ids or None, nbr_idx[N,4])."""
    traci.start(["sumo", "-c", config.get("sumo_cfg_path", "")])
    lanes = [l for l in traci.lane.getIDList() if not l.startswith(":")]
    per = {}
    for lane_id in lanes:                          # order_lanes.py:48-106 naming scheme
        parts = lane_id.split("_to_")
        if len(parts) != 2 or not parts[1].startswith("J_"):
            continue
        to = parts[1].split("_")
        jid, lane_no = "_".join(to[:3]), int(to[3]) if len(to) > 3 else 0
        frm = parts[0].split("_")
        if frm[0] == "END":
            d = frm[1].lower()
        else:
            fr, fc, tr, tc = int(frm[1]), int(frm[2]), int(to[1]), int(to[2])
            d = "n" if fr < tr else "s" if fr > tr else "w" if fc < tc else "e"
        per.setdefault(jid, {}).setdefault(d, []).append((lane_no, lane_id))
    tl_junctions = sorted(per, key=lambda j: tuple(int(x) for x in j.split("_")[1:]))
    table = []
    for j in tl_junctions:
        row = []
        for d in ("n", "s", "e", "w"):             # DIRECTION_ORDER, order_lanes.py:10
            ids = [lid for _, lid in sorted(per[j].get(d, []))][:MAX_LANES_PER_DIRECTION]
            row += ids + [None] * (MAX_LANES_PER_DIRECTION - len(ids))
        table.append(row)
    index = {j: i for i, j in enumerate(tl_junctions)}
    nbr = np.full((len(tl_junctions), 4), -1, np.int32)
    for j, i in index.items():
        r, c = (int(x) for x in j.split("_")[1:])
        for k, (dr, dc) in enumerate(((-1, 0), (1, 0), (0, 1), (0, -1))):   # order_lanes.py:399-404
            nbr[i, k] = index.get(f"J_{r + dr}_{c + dc}", -1)
    return tl_junctions, table, nbr


def read_traci(traci, tl_junctions, table, live_signal: bool):
    """One pass over TraCI: halting count per lane, phase / switch times per junction."""
    n = len(tl_junctions)
    halting = np.full((n, 12), -1, np.int32)
    for i, row in enumerate(table):
        for k, lid in enumerate(row):
            if lid is not None:
                try:
                    halting[i, k] = traci.lane.getLastStepHaltingNumber(lid)
                except Exception:
                    pass                            # order_lanes.py:456-462: keep the -1 padding
    phase = np.zeros(n, np.int32); nsw = np.zeros(n); dur = np.zeros(n); valid = np.zeros(n, np.uint8)
    if live_signal:
        for i, j in enumerate(tl_junctions):
            try:
                phase[i] = traci.trafficlight.getPhase(j)
                nsw[i] = traci.trafficlight.getNextSwitch(j)
                dur[i] = traci.trafficlight.getPhaseDuration(j)
                valid[i] = 1
            except Exception:
                pass
    return halting, phase, nsw, dur, valid


def train_agents(config: dict | None = None, episodes: int = EPISODES, mode: str = "batched", seed: int = 0,
                 live_signal: bool = False, log=None, learn: bool = True):
    """The loop of train.py:182-316.  Returns a list of per-step dicts (also fed to ``log``).
    ``live_signal=False`` reproduces the shipped run, whose phase/time features stay
    [0,0,0,0,-1]."""
    from .agent import create_agents
    config = dict(config or load_config())
    set_seeds(seed)
    traci, is_fake = get_traci(config, seed)
    tl_junctions, table, nbr = initialize_environment(traci, config)
    agents, group = create_agents(tl_junctions, config, seed=seed)
    n = len(tl_junctions)
    step_duration = float(config.get("step_duration", STEP_DURATION))
    max_sim_time = float(config.get("max_sim_time", MAX_SIM_TIME))
    lw, gw = float(config.get("local_reward_weight", 0.3)), float(config.get("global_reward_weight", 0.7))
    smooth_global, smooth_total = SmoothedValue(0.3), SmoothedValue(0.3)
    history = []
    for episode in range(episodes):
        traci.load(["-c", config.get("sumo_cfg_path", "")])
        current_time = traci.simulation.getTime()
        obs, _, reward, glob = group.featurize(*read_traci(traci, tl_junctions, table, live_signal)[:4], current_time,
                                               read_traci(traci, tl_junctions, table, live_signal)[4], nbr,
                                               local_weight=lw, global_weight=gw)
        done, step_count = 0, 0
        while not done:
            # act (train.py:211-222): epsilon from the per-agent host schedule (dqn_agent.py:258-261)
            if mode == "per_agent":
                actions_host = np.array([agents[j].select_action(obs[i:i + 1, :89]) for i, j in enumerate(tl_junctions)])
                actions = torch.as_tensor(actions_host, dtype=torch.int32)
            else:
                eps = np.asarray([agents[j].update_epsilon_before_action() for j in tl_junctions], np.float64)
                # the reference draws rand() (and randint() when exploring) per agent, in agent order (:263-265)
                u = np.empty(n); ra = np.zeros(n, np.int64)
                for i in range(n):
                    u[i] = np.random.rand()
                    if u[i] < eps[i]:
                        ra[i] = np.random.randint(0, group.action_size)
                # supplied-draw contract of dmdqn_act: explore iff w < eps * 2^32; action = (w_a * A) >> 32
                w_explore = np.where(u < eps, 0, 0xFFFFFFFF).astype(np.uint32)
                w_action = ((ra.astype(np.uint64) << np.uint64(32)) // np.uint64(group.action_size)
                            + np.uint64(1)).astype(np.uint32)
                eps_dev = np.where(u < eps, 1.0, 0.0)
                actions = group.act(obs, eps_dev, w_explore, w_action)
                actions_host = actions.cpu().numpy()
                for j in tl_junctions:
                    agents[j].update_epsilon_after_action()
            for j, a in zip(tl_junctions, actions_host):                        # train.py:225-226
                traci.trafficlight.setPhase(j, ACTION_MAP[int(a)])
            target_time = current_time + step_duration                          # train.py:229-236
            while current_time < target_time:
                traci.simulationStep()
                current_time = traci.simulation.getTime()
                done = traci.simulation.getMinExpectedNumber() == 0 or current_time >= max_sim_time
            # reward of this transition comes from the PRE-step readings (train.py:241): `reward`
            # and `glob` were produced together with `obs`; featurise the post-step readings now.
            rd = read_traci(traci, tl_junctions, table, live_signal)
            next_obs, _, next_reward, next_glob = group.featurize(*rd[:4], current_time, rd[4], nbr,
                                                                  local_weight=lw, global_weight=gw)
            dones = np.full(n, int(bool(done)), np.uint8)
            if mode == "per_agent":                                             # train.py:274-292
                r_host = reward.cpu().numpy()
                losses = []
                for i, j in enumerate(tl_junctions):
                    agents[j].remember(obs[i:i + 1, :89], int(actions_host[i]), float(r_host[i]), next_obs[i:i + 1, :89], bool(done))
                    losses.append(agents[j].replay() if learn else 0)
                total_loss = float(sum(float(x) for x in losses))
            else:
                group.push(obs, actions, reward, next_obs, dones)
                total_loss = float(group.learn()[:, 0].sum()) if learn else 0.0
            group.check_errors()        # the host has just synchronised on the losses: one more 4-byte read
            g_r, t_r = float(glob.item()), float(reward.sum().item())
            smooth_global.update(g_r); smooth_total.update(t_r)
            rec = {"episode": episode, "step": step_count, "total_loss": total_loss, "global_reward": g_r,
                   "total_reward": t_r, "smooth_global_reward": smooth_global.get_value(),
                   "smooth_total_reward": smooth_total.get_value()}
            history.append(rec)
            if log is not None:
                log(rec)
            obs, reward, glob = next_obs, next_reward, next_glob
            step_count += 1
    traci.close()
    return history, agents, group


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--episodes", type=int, default=1)
    ap.add_argument("--max-sim-time", type=float, default=None)
    ap.add_argument("--mode", default="batched", choices=["batched", "per_agent"])
    ap.add_argument("--preset", default=None, choices=sorted(PRESETS))
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--log", default=None, help="JSONL file for per-step metrics (wandb stays optional/offline)")
    args = ap.parse_args(argv)
    config = load_config()
    if args.preset:
        config.update(PRESETS[args.preset])
    if args.max_sim_time:
        config["max_sim_time"] = args.max_sim_time
    fh = open(args.log, "w") if args.log else None
    t0 = time.time()
    history, _, _ = train_agents(config, args.episodes, args.mode, args.seed,
                                 log=(lambda r: fh.write(json.dumps(r) + "\n")) if fh else None)
    print(json.dumps({"steps": len(history), "seconds": time.time() - t0, "last": history[-1]}))


if __name__ == "__main__":
    sys.exit(main())
