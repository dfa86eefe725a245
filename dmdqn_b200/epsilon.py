"""Host-side epsilon schedules of the two reference agents (pure Python float64, like the reference).

``reference``: src/agents/dqn_agent.py:258-261 -- evaluated BEFORE the explore draw from ``global_step_count``:
eps = 1 while g < 8000, then ``max(0.01, exp(-(g - 8000) / 16000))`` as long as eps > eps_min (it freezes once it
has fallen to eps_min; the yaml ``epsilon_*`` keys other than ``epsilon_min`` are ignored there).
``linear``: src/experimental/agent.py:82-84,140-144 -- applied AFTER the action was chosen:
``if eps > eps_min: eps -= (eps_start - eps_min) / decay_steps`` then ``eps = max(eps_min, eps)``.
"""
from __future__ import annotations

import numpy as np

KINDS = ("reference", "linear")


def before_action(kind: str, epsilon: float, epsilon_min: float, global_step_count: int) -> float:
    if kind == "reference":
        if global_step_count < 8000:
            return 1.0
        if epsilon > epsilon_min:
            return max(0.01, 1.0 * np.exp(-(global_step_count - 8000) / 16000))
    return epsilon


def after_action(kind: str, epsilon: float, epsilon_min: float, decay_rate: float) -> float:
    if kind == "linear":
        if epsilon > epsilon_min:
            epsilon -= decay_rate
        return max(epsilon_min, epsilon)
    return epsilon
