"""Where does the end-to-end step (dmdqn_step_host + synchronise + read) spend its time beyond the learn kernels?
   python tools/e2e_diag.py   (one GPU, cfg3)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
import bench
from dmdqn_b200.group import AgentGroup

w = bench.WORKLOADS["cfg3"]
grp = AgentGroup(w["agents"], bench.agent_cfg(w, "tf32x3"), bench.D, bench.A, seed=0)
bench.synth_fill(grp, 0)
n, b = grp.n_agents, grp.batch_size
sb = grp.make_step_block()
hv = sb["host"]
hv["obs"].copy_(torch.randint(0, 20, (n, bench.D)).float()); hv["next_obs"].copy_(torch.randint(0, 20, (n, bench.D)).float())
hv["draws"].copy_(torch.randint(0, 2**31, (n, b), dtype=torch.int32))
stream = torch.cuda.current_stream()
draws = grp.draw_words((n, b))
K = 200

def timeit(name, f, sync_each):
    for _ in range(10): f(); stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    for _ in range(K):
        f()
        if sync_each: stream.synchronize()
    e1.record(stream); stream.synchronize(); t1 = time.perf_counter()
    print(f"{name:58s} {e0.elapsed_time(e1) / K * 1e3:8.1f} us/step (wall {(t1 - t0) / K * 1e6:8.1f})", flush=True)

timeit("learn (device draws), no sync", lambda: grp.learn(draws), False)
timeit("learn, sync every step", lambda: grp.learn(draws), True)
timeit("step_host, no sync", lambda: grp.step_host(sb), False)
timeit("step_host, sync every step", lambda: grp.step_host(sb), True)
def full():
    m = grp.step_host(sb); stream.synchronize(); return float(m[0, 0])
timeit("step_host, sync + read loss", full, False)
replay = grp.capture_step_host(sb)
def full_graph():
    m = replay(); stream.synchronize(); return float(m[0, 0])
timeit("step_host as ONE CUDA graph, sync + read loss", full_graph, False)
dev = sb["dev_block"]; host = sb["host_block"]
timeit("H2D copy of the block alone, sync every step", lambda: dev.copy_(host, non_blocking=True), True)
t0 = time.perf_counter()
for _ in range(K): grp.step_host(sb)
t1 = time.perf_counter(); stream.synchronize()
print(f"CPU time to ISSUE one step_host: {(t1 - t0) / K * 1e6:.1f} us")
