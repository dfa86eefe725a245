"""Host-side logic of the product package that needs no GPU: weight initialisers, epsilon schedules,
yaml defaults.  CPU only."""
import math
import os

import numpy as np
import pytest
import torch
import yaml

from dmdqn_b200 import epsilon as E
from dmdqn_b200.group import keras_init

G = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


# ---- keras_init: HeNormal hidden kernels, GlorotUniform head, zero biases (reference dqn_agent.py:166-181) ----
@pytest.mark.parametrize("d,h,a", [(89, 256, 4), (89, 512, 4), (74, 128, 3)])
def test_keras_init_statistics(d, h, a):
    """Checked against the published definitions, not against the oracle's code: Keras HeNormal =
    VarianceScaling(scale=2, fan_in, truncated_normal): a normal of sigma' = sqrt(2/fan_in)/0.87962566
    truncated at +-2 sigma' (so the sample std is sqrt(2/fan_in)); GlorotUniform = U(-l, l), l = sqrt(6/(fan_in+fan_out))."""
    from scipy import stats
    ws = [keras_init(s, d, h, a) for s in range(4)]
    for w in ws:
        assert [tuple(t.shape) for t in w] == [(d, h), (h,), (h, h), (h,), (h, a), (a,)]
        assert all(t.dtype == torch.float32 for t in w)
        assert all(float(t.abs().max()) == 0.0 for t in (w[1], w[3], w[5]))           # zero biases
    for layer, fan_in in ((0, d), (2, h)):
        x = torch.cat([w[layer].flatten() for w in ws]).double().numpy()
        target_std = math.sqrt(2.0 / fan_in)
        sigma_pre = target_std / 0.87962566103423978
        assert np.abs(x).max() <= 2 * sigma_pre * (1 + 1e-6)                            # truncation at 2 sigma'
        assert np.abs(x).max() >= 1.9 * sigma_pre                                       # ... and the tails do reach it
        n = x.size
        assert abs(x.mean()) < 5 * target_std / math.sqrt(n)
        assert abs(x.std() / target_std - 1) < 0.02                                     # sample std = sqrt(2/fan_in)
        # whole distribution: KS distance to scipy's truncated normal (independent implementation)
        ks = stats.kstest(x[:: max(1, n // 20000)], stats.truncnorm(-2, 2, loc=0, scale=sigma_pre).cdf)
        assert ks.statistic < 0.02, ks
    x = torch.cat([w[4].flatten() for w in ws]).double().numpy()
    limit = math.sqrt(6.0 / (h + a))
    assert np.abs(x).max() <= limit and np.abs(x).max() > 0.97 * limit
    assert abs(x.std() / (limit / math.sqrt(3)) - 1) < 0.05 and abs(x.mean()) < 5 * limit / math.sqrt(3 * x.size)
    # seeds: reproducible and distinct (agents must not start with identical networks)
    assert all(torch.equal(p, q) for p, q in zip(keras_init(3, d, h, a), ws[3]))
    assert not torch.equal(ws[0][0], ws[1][0])


# ---- epsilon schedules against the reference's own outputs ----
def test_reference_epsilon_schedule_matches_fixture():
    z = np.load(os.path.join(G, "ref_epsilon.npz"))
    eps = 1.0
    for g, e_ref in zip(z["steps"], z["eps"]):
        eps = E.before_action("reference", eps, float(z["epsilon_min"]), int(g))
        assert eps == e_ref
        assert E.after_action("reference", eps, float(z["epsilon_min"]), 0.123) == eps   # no decay after the action


def test_linear_epsilon_schedule_matches_fixture():
    """src/experimental/agent.py:121-146 run under np.random.seed(11) (oracle/make_golden.py golden_epsilon_linear)."""
    z = np.load(os.path.join(G, "ref_epsilon_linear.npz"))
    eps, eps_min = float(z["epsilon_start"]), float(z["epsilon_min"])
    rate = (eps - eps_min) / int(z["epsilon_decay_steps"])
    assert rate == float(z["decay_rate"])
    np.random.seed(11)
    for e_ref, explored, action in zip(z["eps"], z["explored"], z["action"]):
        eps = E.before_action("linear", eps, eps_min, 10**6)        # no-op for this variant
        u = np.random.rand()
        assert (u < eps) == bool(explored)
        if explored:
            assert np.random.randint(0, 4) == int(action)
        eps = E.after_action("linear", eps, eps_min, rate)
        assert eps == e_ref
    assert z["eps"].min() == eps_min and not z["explored"].all()


def test_yaml_defaults_select_the_tensor_core_path():
    cfg = yaml.safe_load(open(os.path.join(ROOT, "config", "agent_config.yaml")))
    ref = {"learning_rate": 0.0005, "gamma": 0.99, "epsilon_start": 1.0, "epsilon_min": 0.05, "epsilon_decay_steps": 200000,
           "replay_buffer_size": 30000, "batch_size": 128, "target_update_frequency": 1000, "nn_layers": [256, 256]}
    for k, v in ref.items():                                     # reference config/agent_config.yaml:1-9, verbatim
        assert cfg[k] == v
    assert cfg["precision"] == "auto"                            # H = 256 -> tcgen05 3xTF32 (fp32-class), else FFMA
