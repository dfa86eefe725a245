"""A deterministic stand-in for the ``traci`` module, for environments without SUMO.

SUMO itself is the environment and is not rebuilt (BASELINE.json); neither
``sumo`` nor ``traci``/``sumolib`` exist in this image.  This module exposes exactly
the TraCI surface the reference's training loop touches
(src/scripts/train.py:99-106,182-316 and src/experimental/order_lanes.py:430-499):

    traci.start / load / close / simulationStep / isconnected
    traci.simulation.getTime / getMinExpectedNumber
    traci.lane.getIDList / getLastStepHaltingNumber
    traci.trafficlight.getPhase / getNextSwitch / getPhaseDuration / setPhase
    traci.junction.getType            (optional -- see ``junction_get_type``)
    traci.exceptions.TraCIException / FatalTraCIError

driven by a seeded per-lane queue model on a ``rows x cols`` grid of ``J_r_c``
junctions with lane IDs in the reference's naming scheme
(order_lanes.py:48-106: ``END_N_r_c_to_J_r_c_k`` / ``J_a_b_to_J_r_c_k``).
It is a test/demo fixture: it makes the loop runnable, it is not a traffic model.
"""
from __future__ import annotations

import numpy as np

DIRS = ("n", "s", "e", "w")           # order_lanes.py:10
_DELTA = {"n": (-1, 0), "s": (1, 0), "e": (0, 1), "w": (0, -1)}


class TraCIException(Exception):
    pass


class FatalTraCIError(Exception):
    pass


class _Exceptions:
    TraCIException = TraCIException
    FatalTraCIError = FatalTraCIError


def grid_lane_ids(rows: int, cols: int, lanes_per_dir=3, drop=()):
    """Incoming lane IDs of every junction, keyed ``(r, c, dir) -> [lane ids]``.
    ``drop`` is a set of ``(r, c, dir, lane)`` that do not exist (shorter approaches)."""
    out = {}
    for r in range(rows):
        for c in range(cols):
            for d in DIRS:
                dr, dc = _DELTA[d]
                rr, cc = r + dr, c + dc
                if 0 <= rr < rows and 0 <= cc < cols:
                    edge = f"J_{rr}_{cc}_to_J_{r}_{c}"
                else:
                    edge = f"END_{d.upper()}_{r}_{c}_to_J_{r}_{c}"
                out[(r, c, d)] = [f"{edge}_{k}" for k in range(lanes_per_dir)
                                  if (r, c, d, k) not in drop]
    return out


class FakeTraci:
    """One simulated SUMO connection.  ``junction_get_type=False`` reproduces the
    pinned traci 1.22 junction domain, which has no ``getType`` (the reference's
    guard at order_lanes.py:468 then raises AttributeError -> phase/time features
    stay [0,0,0,0,-1.0]); ``True`` makes the traffic-light branch live."""

    def __init__(self, rows=3, cols=3, seed=0, lanes_per_dir=3, drop=(), max_sim_time=2400.0,
                 arrival_rate=0.12, junction_get_type=False, phase_duration=30.0):
        self.rows, self.cols, self.seed = rows, cols, seed
        self.lanes_per_dir = lanes_per_dir
        self.max_sim_time = float(max_sim_time)
        self.arrival_rate = arrival_rate
        self.phase_duration = float(phase_duration)   # or {jid: seconds}
        self.next_switch_override = None              # or {jid: absolute time}
        self._lanes = grid_lane_ids(rows, cols, lanes_per_dir, drop)
        self.junction_ids = [f"J_{r}_{c}" for r in range(rows) for c in range(cols)]
        self._lane_dir = {}
        for (r, c, d), ids in self._lanes.items():
            for lid in ids:
                self._lane_dir[lid] = (f"J_{r}_{c}", DIRS.index(d))
        self.exceptions = _Exceptions
        self.lane = _Lane(self)
        self.trafficlight = _TrafficLight(self)
        self.simulation = _Simulation(self)
        self.junction = _Junction(self, junction_get_type)
        self._connected = False
        self._reset()

    # -- module-level functions -------------------------------------------
    def _reset(self):
        self._rng = np.random.default_rng(self.seed)
        self.time = 0.0
        self.queue = {lid: 0 for lid in self._lane_dir}
        self.phase = {j: 0 for j in self.junction_ids}
        self.phase_set_at = {j: 0.0 for j in self.junction_ids}
        self.n_phase_sets = 0

    def start(self, cmd, *a, **k):
        self._connected = True
        self._reset()

    def load(self, args):
        self._reset()

    def close(self, *a, **k):
        self._connected = False

    def isconnected(self):
        return self._connected

    def simulationStep(self, *a):
        """One simulated second: Bernoulli arrivals on every lane, one departure
        per lane whose approach is served by the junction's current phase
        (phase 0/3/6/9 = ACTION_MAP of train.py:57 serve n/s/e/w)."""
        self.time += 1.0
        lids = sorted(self.queue)
        arrive = self._rng.random(len(lids)) < self.arrival_rate
        for lid, a in zip(lids, arrive):
            jid, d = self._lane_dir[lid]
            q = self.queue[lid] + int(a)
            p = self.phase[jid]
            if p % 3 == 0 and p // 3 == d and q > 0:
                q -= 1
            self.queue[lid] = q


class _Lane:
    def __init__(self, sim):
        self._s = sim

    def getIDList(self):
        internal = [f":J_{r}_{c}_0_0" for r in range(self._s.rows) for c in range(self._s.cols)]
        return sorted(self._s.queue) + internal  # train.py:103 filters the ':' ones

    def getLastStepHaltingNumber(self, lane_id):
        if lane_id not in self._s.queue:
            raise TraCIException(f"unknown lane {lane_id}")
        return self._s.queue[lane_id]


class _TrafficLight:
    def __init__(self, sim):
        self._s = sim

    def getPhase(self, jid):
        return self._s.phase[jid]

    def setPhase(self, jid, phase):
        self._s.phase[jid] = int(phase)
        self._s.phase_set_at[jid] = self._s.time
        self._s.n_phase_sets += 1

    def getPhaseDuration(self, jid):
        d = self._s.phase_duration
        return d[jid] if isinstance(d, dict) else d

    def getNextSwitch(self, jid):
        if self._s.next_switch_override is not None:
            return self._s.next_switch_override[jid]
        return self._s.phase_set_at[jid] + self.getPhaseDuration(jid)


class _Simulation:
    def __init__(self, sim):
        self._s = sim

    def getTime(self):
        return self._s.time

    def getMinExpectedNumber(self):
        return 1 if self._s.time < self._s.max_sim_time else 0


class _Junction:
    def __init__(self, sim, has_get_type):
        self._s = sim
        if has_get_type:
            self.getType = lambda jid: "traffic_light"
