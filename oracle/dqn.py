"""Oracle: the DQN agent (MLP, epsilon-greedy act, Double-DQN learn, Adam, target sync).

TEST INFRASTRUCTURE (see oracle/__init__.py).  CPU PyTorch fp32 restatement of
reference ``src/agents/dqn_agent.py`` (TensorFlow/Keras, not installable here):

* MLP definition / init            dqn_agent.py:153-184   (Keras kernel layout [in,out], y = xW + b)
* ``select_action`` + eps schedule dqn_agent.py:246-274
* ``store_experience``/``remember`` dqn_agent.py:306-325
* ``learn``                        dqn_agent.py:328-380
* hard / soft target sync          dqn_agent.py:382-399   (soft is dead code there; ``tau`` made real)
* ``replay``                       dqn_agent.py:428-434
* variant (Huber, vanilla DQN target, linear eps): src/experimental/agent.py:99,140-144,166-167

Keras pieces restated from the published keras 3.9.2 algorithms (un-vendored
dependency, uv.lock:283-284): ``optimizers.Adam.update_step``,
``losses.MeanSquaredError``/``Huber`` (mean over the batch), ``HeNormal`` /
``GlorotUniform`` initialisers.  PARITY UNPINNED by the reference for this file:
cross-checks live in tests/test_oracle_selfcheck.py (torch.autograd for the
gradients, torch.optim.Adam for the torch-form update).

Two forms: :class:`OracleDQNAgent` -- faithful single agent, one Python object per
intersection, sequential (this is also what bench.py times as the CPU baseline) --
and :class:`StackedOracle` -- the same math over ``[N, ...]`` stacks with bmm, used
to make golden vectors and check the grouped kernels quickly.
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch

from .replay import (FaithfulReplayBuffer, explore_decision, random_action)

ADAM_BETA1 = 0.9
ADAM_BETA2 = 0.999
KERAS_ADAM_EPS = 1e-7
TORCH_ADAM_EPS = 1e-8


# ----------------------------------------------------------------------------
# Initialisers (dqn_agent.py:166-181).
# ----------------------------------------------------------------------------
def he_normal(gen: torch.Generator, fan_in: int, fan_out: int) -> torch.Tensor:
    """Keras HeNormal = VarianceScaling(2, fan_in, truncated_normal): samples a
    normal truncated at +-2 sigma with sigma = sqrt(2/fan_in)/0.87962566103423978."""
    std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
    w = torch.empty(fan_in, fan_out, dtype=torch.float32)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
    return w


def glorot_uniform(gen: torch.Generator, fan_in: int, fan_out: int) -> torch.Tensor:
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(fan_in, fan_out, generator=gen, dtype=torch.float32) * 2 - 1) * limit


def init_params(seed: int, state_size: int, nn_layers, action_size: int) -> list[torch.Tensor]:
    """[W1, b1, W2, b2, ..., Wout, bout] in Keras ``get_weights()`` order."""
    gen = torch.Generator().manual_seed(int(seed))
    dims = [state_size] + list(nn_layers)
    params = []
    for i in range(len(nn_layers)):
        params += [he_normal(gen, dims[i], dims[i + 1]), torch.zeros(dims[i + 1])]
    params += [glorot_uniform(gen, dims[-1], action_size), torch.zeros(action_size)]
    return params


def mlp_forward(params, x: torch.Tensor, keep: bool = False):
    """Dense(relu) x len(nn_layers) -> Dense(linear)  (dqn_agent.py:160-183)."""
    acts = [x]
    h = x
    n_layers = len(params) // 2
    for i in range(n_layers):
        h = h @ params[2 * i] + params[2 * i + 1]
        if i < n_layers - 1:
            h = torch.relu(h)
        acts.append(h)
    return (h, acts) if keep else h


# ----------------------------------------------------------------------------
# Scalars.
# ----------------------------------------------------------------------------
def epsilon_schedule(global_step_count: int, epsilon: float, epsilon_min: float) -> float:
    """dqn_agent.py:258-261 (stateful: returns the new epsilon)."""
    if global_step_count < 8000:
        return 1.0
    if epsilon > epsilon_min:
        return max(0.01, 1.0 * float(np.exp(-(global_step_count - 8000) / 16000)))
    return epsilon


def epsilon_linear_step(epsilon: float, epsilon_min: float, decay_rate: float) -> float:
    """Linear decay applied AFTER an action was selected (experimental/agent.py:140-144):
    ``if eps > eps_min: eps -= rate`` then ``eps = max(eps_min, eps)``; rate =
    (epsilon_start - epsilon_min) / epsilon_decay_steps (experimental/agent.py:82-84)."""
    if epsilon > epsilon_min:
        epsilon -= decay_rate
    return max(epsilon_min, epsilon)


def adam_scalars(t: int, lr: float, form: str = "keras"):
    """(alpha_t, eps_eff) such that  theta -= alpha_t * m / (sqrt(v) + eps_eff).

    keras: alpha_t = lr*sqrt(1-b2^t)/(1-b1^t), eps = 1e-7 (keras 3.9.2 Adam.update_step).
    torch: torch.optim.Adam == same alpha_t with eps_eff = 1e-8*sqrt(1-b2^t).
    Evaluated in float64 then rounded to fp32 (Keras evaluates in fp32; the
    difference is <= 1 ulp of alpha_t, i.e. ~1e-7 relative on the *update*).
    """
    bc1 = 1.0 - ADAM_BETA1 ** t
    bc2 = 1.0 - ADAM_BETA2 ** t
    alpha = lr * math.sqrt(bc2) / bc1
    eps = KERAS_ADAM_EPS if form == "keras" else TORCH_ADAM_EPS * math.sqrt(bc2)
    return np.float32(alpha), np.float32(eps)


def adam_update_(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor,
                 alpha: float, eps: float) -> None:
    """In-place Keras-form Adam on one tensor, fp32, op order of keras 3.9.2:
    m += (g-m)(1-b1); v += (g*g-v)(1-b2); p -= (m*alpha)/(sqrt(v)+eps)."""
    m.add_((g - m) * np.float32(1.0 - ADAM_BETA1))
    v.add_((g * g - v) * np.float32(1.0 - ADAM_BETA2))
    p.sub_((m * np.float32(alpha)) / (torch.sqrt(v) + np.float32(eps)))


def loss_and_grad(pred: torch.Tensor, target: torch.Tensor, kind: str):
    """Returns (per-sample loss terms, dL/dpred) for a batch-mean loss.
    mse   (dqn_agent.py:141,352): L = mean (y-p)^2,  dL/dp = 2(p-y)/B
    huber (experimental/agent.py:99, delta=1): L = mean(|e|<=1 ? e^2/2 : |e|-1/2), dL/dp = clip(p-y,-1,1)/B
    """
    b = pred.shape[-1]
    e = pred - target
    if kind == "mse":
        return e * e, 2.0 * e / b
    if kind == "huber":
        ae = e.abs()
        return torch.where(ae <= 1.0, 0.5 * e * e, ae - 0.5), torch.clamp(e, -1.0, 1.0) / b
    raise ValueError(kind)


# ----------------------------------------------------------------------------
# Faithful single agent.
# ----------------------------------------------------------------------------
class OracleDQNAgent:
    """One intersection's agent; mirrors reference ``DQNAgent`` member for member
    (dqn_agent.py:97-151).  Extra config keys (all default to reference behaviour):
    ``loss`` ('mse'|'huber'), ``tau`` (None -> hard sync), ``normalize_rewards``,
    ``double_dqn``, ``adam_form`` ('keras'|'torch'), ``seed``."""

    def __init__(self, state_size: int, action_size: int, agent_id: str, config: dict,
                 rng: random.Random | None = None):
        self.agent_id = agent_id
        self.state_size = int(state_size)  # reference hard-codes 89/4 (dqn_agent.py:108-109)
        self.action_size = int(action_size)
        self.learning_rate = config.get("learning_rate", 0.001)
        self.gamma = config.get("gamma", 0.99)
        self.epsilon = config.get("epsilon_start", 1.0)
        self.epsilon_min = config.get("epsilon_min", 0.01)
        self.epsilon_decay_steps = config.get("epsilon_decay_steps", 100000)
        self.epsilon_decay_rate = ((self.epsilon - self.epsilon_min) / self.epsilon_decay_steps
                                   if self.epsilon_decay_steps > 0 else 0)        # experimental/agent.py:82-84
        self.epsilon_schedule_kind = config.get("epsilon_schedule", "reference")   # 'reference' | 'linear'
        self.buffer_size = config.get("replay_buffer_size", 10000)
        self.batch_size = config.get("batch_size", 128)
        self.target_update_frequency = config.get("target_update_frequency", 1000)
        self.nn_layers = list(config.get("nn_layers", [64, 64]))
        self.loss_kind = config.get("loss", "mse")
        self.tau = config.get("tau", None)
        self.normalize_rewards = config.get("normalize_rewards", True)
        self.double_dqn = config.get("double_dqn", True)
        self.adam_form = config.get("adam_form", "keras")
        seed = config.get("seed", 0)

        self.online = init_params(seed, self.state_size, self.nn_layers, self.action_size)
        self.target = [p.clone() for p in self.online]  # dqn_agent.py:135-137
        self.adam_m = [torch.zeros_like(p) for p in self.online]
        self.adam_v = [torch.zeros_like(p) for p in self.online]
        self.replay_buffer = FaithfulReplayBuffer(self.buffer_size, rng)
        self.global_step_count = 0
        self.learn_step_counter = 0
        self.last_metrics: dict = {}

    # -- act ---------------------------------------------------------------
    def q_values(self, state) -> torch.Tensor:
        with torch.no_grad():
            return mlp_forward(self.online, torch.as_tensor(np.asarray(state), dtype=torch.float32))

    def select_action(self, state, w_explore: int | None = None, w_action: int | None = None) -> int:
        """dqn_agent.py:246-274.  With words supplied the decision follows the
        "supplied draws" contract (oracle/replay.py); without, ``np.random`` like
        the reference."""
        linear = self.epsilon_schedule_kind == "linear"
        if not linear:
            self.epsilon = epsilon_schedule(self.global_step_count, self.epsilon, self.epsilon_min)
        if w_explore is None:
            explore = np.random.rand() < self.epsilon
        else:
            explore = bool(explore_decision(np.uint32(w_explore), self.epsilon))
        if explore:
            if w_action is None:
                action = int(np.random.randint(0, self.action_size))
            else:
                action = int(random_action(np.uint32(w_action), self.action_size))
        else:
            action = int(torch.argmax(self.q_values(state), dim=1)[0])  # ties -> lowest index
        if linear:      # experimental/agent.py:140-144: decay after the action was chosen
            self.epsilon = epsilon_linear_step(self.epsilon, self.epsilon_min, self.epsilon_decay_rate)
        return action

    def select_greedy_action(self, state) -> int:
        """experimental/agent.py:148-152 (used by src/scripts/test.py:88)."""
        return int(torch.argmax(self.q_values(state), dim=1)[0])

    # -- remember ----------------------------------------------------------
    def store_experience(self, experience) -> None:  # dqn_agent.py:306-310
        self.replay_buffer.add(experience)
        self.global_step_count += 1

    def remember(self, state, action, reward, next_state, done) -> None:  # dqn_agent.py:312-325
        self.replay_buffer.add((state, action, reward, next_state, done))

    # -- learn -------------------------------------------------------------
    def learn_on_batch(self, states, actions, rewards, next_states, dones):
        """dqn_agent.py:342-377 on an already sampled batch.  Returns loss (float)."""
        s = torch.as_tensor(states, dtype=torch.float32)
        a = torch.as_tensor(actions, dtype=torch.int64)
        r = torch.as_tensor(rewards, dtype=torch.float32)
        s2 = torch.as_tensor(next_states, dtype=torch.float32)
        d = torch.as_tensor(dones, dtype=torch.float32)
        rows = torch.arange(s.shape[0])
        with torch.no_grad():
            tq_all = mlp_forward(self.target, s2)
            if self.double_dqn:  # dqn_agent.py:342-345
                next_actions = torch.argmax(mlp_forward(self.online, s2), dim=1)
                tq = tq_all[rows, next_actions]
            else:  # experimental/agent.py:166-167
                tq = tq_all.max(dim=1).values
            targets = r + np.float32(self.gamma) * (1.0 - d) * tq  # dqn_agent.py:347

        params = [p.detach().requires_grad_(True) for p in self.online]
        q_all = mlp_forward(params, s)
        pred = q_all[rows, a]  # == reduce_sum(q_all * one_hot(a)) (dqn_agent.py:351)
        terms, _ = loss_and_grad(pred, targets, self.loss_kind)
        loss = terms.mean()
        grads = torch.autograd.grad(loss, params)

        self.learn_step_counter += 1  # dqn_agent.py:359 (Keras `iterations` is t-1 at apply time)
        alpha, eps = adam_scalars(self.learn_step_counter, self.learning_rate, self.adam_form)
        with torch.no_grad():
            for p, g, m, v in zip(self.online, grads, self.adam_m, self.adam_v):
                adam_update_(p, g, m, v, alpha, eps)
        qd = q_all.detach()
        self.last_metrics = {  # dqn_agent.py:361-363
            "q_values_mean": float(qd.mean()),
            "q_values_std": float(qd.std(unbiased=False)),
            "action_distribution": np.bincount(np.asarray(actions), minlength=self.action_size),
        }
        if self.tau is not None:  # dqn_agent.py:389-399 made live
            self.update_target_network_soft()
        elif self.learn_step_counter % self.target_update_frequency == 0:  # dqn_agent.py:376-377
            self.update_target_network()
        return float(loss.detach())

    def learn(self):
        if len(self.replay_buffer) < self.batch_size:  # dqn_agent.py:333-335
            return None
        batch = self.replay_buffer.sample(self.batch_size, self.normalize_rewards)
        return self.learn_on_batch(*batch)

    def replay(self):  # dqn_agent.py:428-434
        loss = self.learn()
        return 0 if loss is None else loss

    def update_target_network(self) -> None:  # dqn_agent.py:382-384
        self.target = [p.clone() for p in self.online]

    def update_target_network_soft(self) -> None:  # dqn_agent.py:389-399
        tau = np.float32(self.tau)
        with torch.no_grad():
            for t, o in zip(self.target, self.online):
                t.copy_(tau * o + (np.float32(1.0) - tau) * t)

    def get_epsilon(self) -> float:
        return self.epsilon


# ----------------------------------------------------------------------------
# Stacked form over N independent agents (or one shared network with N=1).
# ----------------------------------------------------------------------------
class StackedOracle:
    """Same arithmetic as :class:`OracleDQNAgent` over ``[N, ...]`` stacks.
    ``params``: list of ``[N, in, out]`` / ``[N, out]`` tensors (Keras order)."""

    def __init__(self, n_nets: int, state_size: int, nn_layers, action_size: int, *,
                 gamma=0.99, learning_rate=5e-4, loss="mse", tau=None,
                 target_update_frequency=1000, double_dqn=True, adam_form="keras",
                 seed0: int = 0):
        self.n, self.d, self.a = n_nets, state_size, action_size
        self.nn_layers = list(nn_layers)
        self.gamma, self.lr, self.loss_kind, self.tau = gamma, learning_rate, loss, tau
        self.freq, self.double_dqn, self.adam_form = target_update_frequency, double_dqn, adam_form
        per = [init_params(seed0 + i, state_size, nn_layers, action_size) for i in range(n_nets)]
        self.online = [torch.stack([p[k] for p in per]) for k in range(len(per[0]))]
        self.target = [p.clone() for p in self.online]
        self.adam_m = [torch.zeros_like(p) for p in self.online]
        self.adam_v = [torch.zeros_like(p) for p in self.online]
        self.learn_step = np.zeros((n_nets,), np.int64)

    @staticmethod
    def forward(params, x: torch.Tensor) -> torch.Tensor:
        h = x  # [N, B, in]
        n_layers = len(params) // 2
        for i in range(n_layers):
            h = torch.baddbmm(params[2 * i + 1][:, None, :], h, params[2 * i])
            if i < n_layers - 1:
                h = torch.relu(h)
        return h

    def q_values(self, obs) -> torch.Tensor:
        """obs [N, D] (one row per agent) -> [N, A]."""
        with torch.no_grad():
            return self.forward(self.online, torch.as_tensor(obs, dtype=torch.float32)[:, None, :])[:, 0]

    def act(self, obs, eps, w_explore, w_action):
        """Batched dqn_agent.py:263-274 with supplied words.  Returns (actions i32, q [N,A])."""
        q = self.q_values(obs)
        greedy = torch.argmax(q, dim=1).numpy().astype(np.int32)
        explore = explore_decision(w_explore, eps)
        return np.where(explore, random_action(w_action, self.a), greedy).astype(np.int32), q.numpy()

    def td_targets(self, rewards, next_states, dones):
        r = torch.as_tensor(rewards, dtype=torch.float32)
        s2 = torch.as_tensor(next_states, dtype=torch.float32)
        d = torch.as_tensor(dones, dtype=torch.float32)
        with torch.no_grad():
            tq_all = self.forward(self.target, s2)
            q_next = self.forward(self.online, s2)
            if self.double_dqn:
                na = torch.argmax(q_next, dim=2, keepdim=True)
                tq = torch.gather(tq_all, 2, na)[..., 0]
            else:
                tq = tq_all.max(dim=2).values
            y = r + np.float32(self.gamma) * (1.0 - d) * tq
        return y, q_next, tq_all

    def learn_on_batch(self, states, actions, rewards, next_states, dones, active=None):
        """All inputs ``[N, B, ...]``.  ``active[N]`` bool: agents that learn this
        step (others untouched, loss 0 like ``replay()``).  Returns dict of numpy
        arrays: loss[N], y[N,B], q_all[N,B,A], grads (list)."""
        n = self.n
        active = np.ones((n,), bool) if active is None else np.asarray(active, bool)
        s = torch.as_tensor(states, dtype=torch.float32)
        a = torch.as_tensor(actions, dtype=torch.int64)
        y, q_next, tq_all = self.td_targets(rewards, next_states, dones)
        params = [p.detach().requires_grad_(True) for p in self.online]
        q_all = self.forward(params, s)
        pred = torch.gather(q_all, 2, a[..., None])[..., 0]
        terms, _ = loss_and_grad(pred, y, self.loss_kind)
        loss = terms.mean(dim=1)  # [N]
        grads = torch.autograd.grad(loss.sum(), params)
        act_t = torch.as_tensor(active)
        with torch.no_grad():
            for i in np.nonzero(active)[0]:
                self.learn_step[i] += 1
                alpha, eps = adam_scalars(int(self.learn_step[i]), self.lr, self.adam_form)
                for p, g, m, v in zip(self.online, grads, self.adam_m, self.adam_v):
                    adam_update_(p[i], g[i], m[i], v[i], alpha, eps)
                if self.tau is not None:
                    tau = np.float32(self.tau)
                    for t, o in zip(self.target, self.online):
                        t[i].copy_(tau * o[i] + (np.float32(1.0) - tau) * t[i])
                elif self.learn_step[i] % self.freq == 0:
                    for t, o in zip(self.target, self.online):
                        t[i].copy_(o[i])
        return {
            "loss": torch.where(act_t, loss.detach(), torch.zeros_like(loss)).numpy(),
            "y": y.numpy(), "q_all": q_all.detach().numpy(), "q_next": q_next.numpy(),
            "tq_all": tq_all.numpy(), "grads": [g.numpy() for g in grads],
        }
