"""Debug aid: run learn() at several group sizes, report tc_error and the time of each stage."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import ctypes as C
import numpy as np, torch
from dmdqn_b200 import _native as N
from dmdqn_b200.group import AgentGroup

def run(n, batch, cap=400, precision="tf32x3"):
    cfg = {"nn_layers": [256, 256], "replay_buffer_size": cap, "batch_size": batch, "learning_rate": 5e-4, "precision": precision}
    grp = AgentGroup(n, cfg, seed=1)
    gen = torch.Generator(device=grp.device).manual_seed(0)
    for ring in (grp.obs, grp.next_obs):
        ring[:, :, :89] = torch.randint(-1, 20, (n, cap, 89), device=grp.device, generator=gen).float()
    grp.act_ring.copy_(torch.randint(0, 4, (n, cap), device=grp.device, generator=gen).int())
    grp.rew_ring.copy_(-torch.randint(0, 240, (n, cap), device=grp.device, generator=gen).double())
    grp.n_written.fill_(cap + 7); grp.n_written_host[:] = cap + 7
    draws = grp.draw_words((4, n, batch))
    stream = torch.cuda.current_stream()
    for it in range(3):
        ts = []
        for s in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            N.check(grp.lib.dmdqn_learn_stages(C.byref(grp.dims), C.byref(grp.hp), C.byref(grp.replay), C.byref(grp.nets),
                                               draws[it].data_ptr(), None, grp.metrics.data_ptr(), grp.workspace.data_ptr(),
                                               grp.workspace.numel(), 1 << s, stream.cuda_stream))
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        err = int(grp.debug_views()["tc_error"][0])
        print(f"n={n} B={batch} it={it} tc_error={err} stage ms: " + " ".join(f"{t:.3f}" for t in ts), flush=True)

for n in [int(x) for x in sys.argv[1:]] or [2, 74, 148, 256]:
    run(n, 256)
