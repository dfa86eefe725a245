"""Debug aid: run the sample stage once, then the K3 / K4a stages many times on unchanged inputs and report every
(network, array) whose result differs from the first run or from the FFMA (fp32) kernels on the same inputs -- a race in
the persistent tcgen05 kernels shows up as a run-to-run difference.   python tools/dbg_k3_repeat.py [n_agents batch cap reps]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import torch
from dmdqn_b200.group import AgentGroup

n, batch, cap, reps = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (150, 64, 80, 20)))


def make(precision):
    grp = AgentGroup(n, {"nn_layers": [256, 256], "replay_buffer_size": cap, "batch_size": batch, "precision": precision})
    grp.init_weights(seed=3)
    rng = np.random.default_rng(0)
    for t in range(cap + 3):
        s = rng.integers(-1, 20, (n, 89)).astype(np.float32)
        s2 = rng.integers(-1, 20, (n, 89)).astype(np.float32)
        grp.push(s, rng.integers(0, 4, n).astype(np.int32), -rng.random(n) * 100, s2, rng.random(n) < 0.1)
    torch.cuda.synchronize()
    return grp


def stages(grp, d, mask):
    stream = torch.cuda.current_stream()
    rc = grp.lib.dmdqn_learn_stages(C.byref(grp.dims), C.byref(grp.hp), C.byref(grp.replay), C.byref(grp.nets), d.data_ptr(), None,
                                    grp.metrics.data_ptr(), grp.workspace.data_ptr(), grp.workspace.numel(), mask, stream.cuda_stream)
    assert rc == 0, grp.lib.dmdqn_last_error().decode()
    torch.cuda.synchronize()
    v = grp.debug_views()
    return {k: v[k].cpu().numpy().copy() for k in ("q_next", "tq_all", "q_all", "y", "dh1", "h1") if k in v}, int(v["tc_error"][0])


ref_grp, tc_grp = make("fp32"), make("tf32x3")
d = tc_grp.draw_words((n, batch))
stages(ref_grp, d, 1); stages(tc_grp, d, 1)
ref, _ = stages(ref_grp, d, 2 | 4)
first = None
for rep in range(reps):
    out, _err = stages(tc_grp, d, 2 | 4)
    if first is None:
        first = out
    for k in ("q_next", "tq_all", "q_all", "y"):
        a, b, r = out[k].reshape(n, -1), first[k].reshape(n, -1), ref[k].reshape(n, -1)
        scale = np.abs(r).max()
        bad_ref = np.where(np.abs(a - r).max(1) > 1e-4 * scale)[0]
        bad_first = np.where((a != b).any(1))[0]
        for g in bad_ref[:3]:
            rows_bad = np.where(np.abs(out[k][g] - ref[k][g]).reshape(batch, -1).max(1) > 1e-4 * scale)[0]
            err = (out[k][g] - ref[k][g]).reshape(batch, -1)
            print(f"   net {g} {k}: {len(rows_bad)} bad rows, first {rows_bad[:12].tolist()} last {rows_bad[-4:].tolist()}; "
                  f"err of first bad rows {np.round(err[rows_bad[:3]], 3).tolist()}", flush=True)
        if len(bad_ref) or len(bad_first):
            print(f"rep {rep} {k}: vs fp32 kernels nets {bad_ref.tolist()[:20]} (worst {np.abs(a - r).max() / scale:.2e}), "
                  f"vs first run nets {bad_first.tolist()[:20]}", flush=True)
    if _err:
        print("tc_error", _err)
print("done", n, batch, cap, reps)
