"""Oracle: observation / reward featurisation on index tables.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Array restatement of

* ``get_own_state``        reference src/experimental/order_lanes.py:430-499
* ``_get_neighbor_info``   reference src/experimental/order_lanes.py:392-427
* ``build_state_vector``   reference src/experimental/order_lanes.py:502-555
* reward helpers           reference src/scripts/train.py:159-165,241,251-254
* alt 74-dim contract      reference src/agents/sumo_env.py:532-679

The reference works on junction-ID strings and TraCI calls; here every junction
is a row index and the TraCI readings are arrays, so grids of any size can be
featurised (the reference's ``junction[:5]`` mapping breaks past 9x9).
PINNED against the reference's own functions run under a fake ``traci``
(oracle/make_golden.py -> tests/golden/ref_featurize.npz).
"""
from __future__ import annotations

import numpy as np

OWN = 17            # 12 queues + 4 phase one-hot + time_spent
OBS = 89            # own + presence(4) + 4 * own
PHASE_LUT_SIZE = 16


def default_phase_lut() -> np.ndarray:
    """PHASE_ENCODING (order_lanes.py:14-19): SUMO phase index 0..3 -> one-hot slot,
    anything else -> zeros (-1)."""
    lut = np.full((PHASE_LUT_SIZE,), -1, np.int32)
    lut[:4] = np.arange(4)
    return lut


def grid_neighbors(rows: int, cols: int) -> np.ndarray:
    """``nbr_idx[N,4]`` in n,s,e,w order for a row-major ``J_r_c`` grid
    (order_lanes.py:399-404: n=(r-1,c), s=(r+1,c), e=(r,c+1), w=(r,c-1)); -1 = none."""
    idx = np.full((rows * cols, 4), -1, np.int32)
    for r in range(rows):
        for c in range(cols):
            j = r * cols + c
            for k, (dr, dc) in enumerate(((-1, 0), (1, 0), (0, 1), (0, -1))):
                rr, cc = r + dr, c + dc
                if 0 <= rr < rows and 0 <= cc < cols:
                    idx[j, k] = rr * cols + cc
    return idx


def own_state(halting, phase, next_switch, phase_dur, sim_time, signal_valid, phase_lut=None):
    """``[N,17]`` float64 own blocks (order_lanes.py:430-499).

    halting[N,12] int (slot = dir*3+lane, dir order n,s,e,w; -1 = lane absent or
    read failed -> stays -1.0, :439).  ``signal_valid[N]`` = whether the
    traffic-light branch ran (:468); when 0 the block ends [0,0,0,0,-1.0] -- the
    behaviour of the shipped run.  ``time_spent`` =
    dur - (next_switch - now); a negative value trips the reference's ``assert``
    inside a swallowed ``try`` (:481,485-486) BEFORE the ``max(0.0, .)`` runs, so it
    stays -1.0."""
    halting = np.asarray(halting)
    n = halting.shape[0]
    lut = default_phase_lut() if phase_lut is None else np.asarray(phase_lut, np.int32)
    own = np.empty((n, OWN), np.float64)
    own[:, :12] = halting.astype(np.float64)
    own[:, 12:16] = 0.0
    own[:, 16] = -1.0
    phase = np.asarray(phase, np.int64)
    valid = np.asarray(signal_valid).astype(bool)
    in_lut = valid & (phase >= 0) & (phase < lut.shape[0])
    slot = np.where(in_lut, lut[np.clip(phase, 0, lut.shape[0] - 1)], -1)
    rows = np.nonzero(slot >= 0)[0]
    own[rows, 12 + slot[rows]] = 1.0
    calc = np.asarray(phase_dur, np.float64) - (np.asarray(next_switch, np.float64) - np.float64(sim_time))
    ok = valid & (calc >= 0)
    own[ok, 16] = calc[ok]
    return own


def build_obs(own, nbr_idx, global_state=None) -> np.ndarray:
    """``[N,89]`` float32 observations (order_lanes.py:502-555): own || presence ||
    nbr_n || nbr_s || nbr_e || nbr_w, absent neighbour -> 17 x -1.0.  Neighbour
    blocks come from ``global_state`` (the caller's snapshot, :547-548); the own
    block is the live one (:526-531)."""
    own = np.asarray(own, np.float64)
    snap = own if global_state is None else np.asarray(global_state, np.float64)
    nbr_idx = np.asarray(nbr_idx, np.int64)
    n = own.shape[0]
    obs = np.empty((n, OBS), np.float64)
    obs[:, :OWN] = own
    obs[:, OWN:OWN + 4] = (nbr_idx >= 0).astype(np.float64)
    for k in range(4):
        blk = np.where((nbr_idx[:, k] >= 0)[:, None], snap[np.clip(nbr_idx[:, k], 0, n - 1)], -1.0)
        obs[:, OWN + 4 + k * OWN: OWN + 4 + (k + 1) * OWN] = blk
    return obs.astype(np.float32)  # caller's cast (train.py:220,269)


def rewards(own_pre, local_weight=0.3, global_weight=0.7):
    """train.py:159-165,241,251-254 from the PRE-step own blocks.  float64,
    two multiplies and one add (no FMA).  Returns (reward[N], global_reward)."""
    own_pre = np.asarray(own_pre, np.float64)
    local = -1.0 * own_pre[:, :12].sum(axis=1)        # integer-valued: exact in any order
    glob = -1.0 * float(own_pre[:, :12].sum())
    return np.float64(local_weight) * local + np.float64(global_weight) * np.float64(glob), glob


# ----------------------------------------------------------------------------
# Alt contract: SumoTrafficEnvironment 74-dim (sumo_env.py:532-679).
# ----------------------------------------------------------------------------
OWN_ALT = 14
OBS_ALT = 74


def own_state_alt(halting_nesw, phase, next_switch, sim_time, signal_valid=None):
    """local(14) = 12 queues in N,E,S,W order || phase index || max(0, nextSwitch - now)
    (sumo_env.py:532-580).  Queue codes: -2 = PAD lane -> 0.0 (:545-549); -1 = failed
    read -> keeps the -1.0 padding value (:555-557).  A junction without a readable signal
    (``signal_valid`` 0) keeps -1.0 in both signal slots (:560-575)."""
    h = np.asarray(halting_nesw, np.float64)
    n = h.shape[0]
    valid = np.ones(n, bool) if signal_valid is None else np.asarray(signal_valid).astype(bool)
    own = np.zeros((n, OWN_ALT), np.float64)
    own[:, :12] = np.where(h == -2, 0.0, h)
    own[:, 12] = np.where(valid, np.asarray(phase, np.float64), -1.0)
    own[:, 13] = np.where(valid, np.maximum(0.0, np.asarray(next_switch, np.float64) - np.float64(sim_time)), -1.0)
    return own


def build_obs_alt(own, nbr_idx_nesw) -> np.ndarray:
    """obs = local || presence[N,E,S,W] || nbr_N || nbr_E || nbr_S || nbr_W, pad -1.0
    (sumo_env.py:131-142,582-631)."""
    own = np.asarray(own, np.float64)
    nbr = np.asarray(nbr_idx_nesw, np.int64)
    n = own.shape[0]
    obs = np.empty((n, OBS_ALT), np.float64)
    obs[:, :OWN_ALT] = own
    obs[:, OWN_ALT:OWN_ALT + 4] = (nbr >= 0).astype(np.float64)
    for k in range(4):
        blk = np.where((nbr[:, k] >= 0)[:, None], own[np.clip(nbr[:, k], 0, n - 1)], -1.0)
        obs[:, OWN_ALT + 4 + k * OWN_ALT: OWN_ALT + 4 + (k + 1) * OWN_ALT] = blk
    return obs.astype(np.float32)


def rewards_alt(own_prev, own_curr) -> np.ndarray:
    """r_j = sum max(0, prev q) - sum max(0, curr q)  (sumo_env.py:672-677)."""
    p = np.maximum(0.0, np.asarray(own_prev, np.float64)[:, :12]).sum(axis=1)
    c = np.maximum(0.0, np.asarray(own_curr, np.float64)[:, :12]).sum(axis=1)
    return p - c
