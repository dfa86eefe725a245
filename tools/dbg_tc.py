import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from dmdqn_b200.group import AgentGroup
from oracle import replay as R
from oracle.dqn import StackedOracle

def run(n, batch, cap, precision="tf32x3", h=256):
    rng = np.random.default_rng(5)
    cfg = {"nn_layers": [h, h], "replay_buffer_size": cap, "batch_size": batch, "learning_rate": 5e-4, "precision": precision}
    grp = AgentGroup(n, cfg)
    stk = StackedOracle(n, 89, [h, h], 4, learning_rate=5e-4, seed0=50)
    for k in (1, 3, 5):
        stk.online[k] += torch.as_tensor(rng.standard_normal(stk.online[k].shape).astype(np.float32)) * 0.05
        stk.target[k].copy_(stk.online[k])
    for i in range(n):
        grp.set_weights(i, [p[i] for p in stk.online], "online"); grp.set_weights(i, [p[i] for p in stk.target], "target")
    ring = R.RingReplay(n, cap, 89)
    for t in range(cap + 7):
        s = rng.integers(-1, 20, (n, 89)).astype(np.float32); s2 = rng.integers(-1, 20, (n, 89)).astype(np.float32)
        a = rng.integers(0, 4, n).astype(np.int32); r = -0.3 * rng.integers(0, 200, n) - 0.7 * rng.integers(0, 5000, n)
        dn = rng.random(n) < 0.1
        grp.push(s, a, r, s2, dn); ring.push(s, a, r, s2, dn)
    words = rng.integers(0, 2**32, (n, batch), dtype=np.uint64).astype(np.uint32)
    grp.learn(words, sample_mode="fisher_yates")
    dbg = {k: v.cpu().numpy() for k, v in grp.debug_views().items()}
    batches = [ring.gather(i, R.fisher_yates_indices(words[i], cap)) for i in range(n)]
    out = stk.learn_on_batch(*(np.stack([b[k] for b in batches]) for k in range(5)))
    print(f"B={batch} tc_error={dbg['tc_error']}")
    for name in ("q_next", "tq_all", "q_all", "y"):
        d = np.abs(dbg[name] - out[name]); print(f"  {name:7s} max err {d.max():.3e} (mag {np.abs(out[name]).max():.2e}) argmax {np.unravel_index(d.argmax(), d.shape)}")
    names = ["W1", "b1", "W2", "b2", "W3", "b3"]
    for i in range(n):
        gm = grp.get_weights(i, "m")
        for k in range(6):
            g_gpu = gm[k].numpy() / 0.1; g_ref = out["grads"][k][i]
            d = np.abs(g_gpu - g_ref); tol = 1e-5 * np.abs(g_ref).max()
            bad = np.argwhere(d > tol)
            if len(bad):
                print(f"  net {i} grad {names[k]}: {len(bad)} bad, max {d.max():.3e} (mag {np.abs(g_ref).max():.2e}); rows {sorted(set(bad[:,0].tolist()))[:20]} cols {sorted(set(bad[:,-1].tolist()))[:20]}")
            else:
                print(f"  net {i} grad {names[k]}: ok max {d.max():.3e} (mag {np.abs(g_ref).max():.2e})")

for b in [int(x) for x in sys.argv[1:]] or [300]:
    run(2, b, b + 100)
