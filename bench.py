#!/usr/bin/env python
"""bench.py -- agent-updates/s of the Double-DQN learn step (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg3]

A "step" is one learn sweep: every agent of the workload runs one ``learn()`` (sample B
transitions from its ring, target + online forward, backward, Adam, target sync).
Workloads (BASELINE.json configs; synthetic replay, observation-shaped integers, rings full):
  cfg2  16 agents,  B=64,  H=256      cfg3  256 agents, B=256, H=256  (default, per GPU)
  cfg4  512 agents, B=512, H=256 per GPU (4096 over 8 GPUs)
N>1 shards agents across ranks with no collective (independent networks): weak scaling,
every rank holds the per-GPU workload.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import random
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints "NCCL version ..." on stdout at
# init), so the real stdout is kept on a private descriptor for the result and fd 1 is pointed at stderr.
_JSON_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    os.write(_JSON_FD, (json.dumps(obj) + "\n").encode())

WORKLOADS = {
    "cfg2": dict(agents=16, batch=64, hidden=256, capacity=30000, grid="4x4"),
    "cfg3": dict(agents=256, batch=256, hidden=256, capacity=30000, grid="16x16"),
    "cfg4": dict(agents=512, batch=512, hidden=256, capacity=30000, grid="64x64 city grid / 8 GPUs"),
}
D, A = 89, 4


def workload_name(key):
    w = WORKLOADS[key]
    return (f"{key}: {w['agents']} agents/GPU ({w['grid']}), batch {w['batch']}, hidden [{w['hidden']},{w['hidden']}], obs 89, "
            f"actions 4, replay capacity {w['capacity']}, independent networks")


def agent_cfg(w, precision="fp32"):
    return {"learning_rate": 5e-4, "gamma": 0.99, "replay_buffer_size": w["capacity"], "batch_size": w["batch"],
            "target_update_frequency": 1000, "nn_layers": [w["hidden"], w["hidden"]],    # config/agent_config.yaml
            "precision": precision}


def flops_bytes(w):
    """Algorithmic work per agent-update (SURVEY.md section 8 D-roofline)."""
    b, h = w["batch"], w["hidden"]
    m = D * h + h * h + h * A
    p = m + h + h + A
    k = {"target": dict(flops=2 * 2 * b * m, bytes=b * (D + 2) * 4 + 8 * p),
         "online": dict(flops=2 * b * m + 2 * b * (h * h + h * A), bytes=b * (D + 1) * 4 + 4 * p),
         "wgrad_adam": dict(flops=2 * b * m, bytes=12 * p),
         "sample": dict(flops=0, bytes=b * 13 + b * 16)}
    return {"flops": 2 * b * (4 * m + h * h + h * A), "bytes": b * (2 * D + 3) * 4 + 24 * p, "params": p, "kernels": k}


# ------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons sampled WHILE the timed region runs: NVML every 10 ms
    (nvidia_ml_py), or nvidia-smi every 200 ms when NVML cannot be loaded."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt, self.how = index, [], threading.Event(), "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(index))
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.how = "nvml"
        except Exception:
            self.nv = None

    @staticmethod
    def _physical_index(local):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v for v in vis.split(",") if v.strip()]
            if local < len(ids) and ids[local].strip().isdigit():
                return int(ids[local])
        return local

    def _nvml_row(self):
        nv = self.nv
        bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        flags = (getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8), getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))
        return [float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), self.max_sm, nv.nvmlDeviceGetPowerUsage(self.h) / 1e3] + \
               ["Active" if bits & f else "Not Active" for f in flags]

    def run(self):
        import subprocess
        while not self._stop_evt.is_set():
            try:
                if self.nv is not None:
                    self.rows.append(self._nvml_row())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    if len(f) >= 7:
                        self.rows.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.01 if self.nv is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows]
        reasons = [n for i, n in enumerate(self.NAMES, 3) if any(str(r[i]).lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.rows[0][1]), "samples": len(sm), "how": self.how,
                "power_w_max": max(float(r[2]) for r in self.rows), "reasons": reasons}


def ncu_traffic(workload, precision, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full`
    capture of this workload (profiles/ncu_traffic.json names the capture each figure comes from); None when
    this workload / precision has not been captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return t[f"{workload}/{precision}"][kernel]["dram_bytes"]
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"], "bf16_tflops_sustained": j["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------
def synth_fill(grp, seed):
    """Full rings of featuriser-shaped data: queue counts 0..19 as floats, -1 padding,
    rewards 0.3*local + 0.7*global (float64), dones ~ 1/240 (SURVEY.md D-inputs)."""
    n, c = grp.n_agents, grp.capacity
    gen = torch.Generator(device=grp.device).manual_seed(seed)
    chunk = max(1, (1 << 28) // (c * D))
    for a0 in range(0, n, chunk):
        a1 = min(n, a0 + chunk)
        for ring in (grp.obs, grp.next_obs):
            v = torch.randint(-1, 20, (a1 - a0, c, D), device=grp.device, generator=gen).float()
            ring[a0:a1, :, :D] = v
    grp.act_ring.copy_(torch.randint(0, A, (n, c), device=grp.device, generator=gen).int())
    loc = torch.randint(0, 240, (n, c), device=grp.device, generator=gen).double()
    glob = torch.randint(0, 240 * 256, (1, c), device=grp.device, generator=gen).double()
    grp.rew_ring.copy_(-0.3 * loc - 0.7 * glob)
    grp.done_ring.copy_((torch.rand((n, c), device=grp.device, generator=gen) < 1 / 240).to(torch.uint8))
    grp.n_written.fill_(c + 777)
    grp.n_written_host[:] = c + 777


def run_ours(args):
    from dmdqn_b200 import _native as N
    from dmdqn_b200.group import AgentGroup
    import ctypes as C
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = WORKLOADS[args.workload]
    fb = flops_bytes(w)
    grp = AgentGroup(w["agents"], agent_cfg(w, args.precision), D, A, seed=1000 * rank)
    synth_fill(grp, seed=rank)
    n, b = grp.n_agents, grp.batch_size
    K, W = args.steps, args.warmup
    draws = grp.draw_words((K + W, n, b))
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: K learn sweeps, inputs already in HBM ----------------
    for i in range(W):
        grp.learn(draws[i])
    barrier()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(K):
        grp.learn(draws[W + i])
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    # ---- per-kernel durations: same sweep, one stage per call, events in between ----------
    stage_names = ["sample", "target", "online", "wgrad_adam"]
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
    lib = grp.lib
    for i in range(K):
        d = draws[W + i]
        for s in range(4):
            evs[i][s].record(stream)
            N.check(lib.dmdqn_learn_stages(C.byref(grp.dims), C.byref(grp.hp), C.byref(grp.replay), C.byref(grp.nets),
                                           d.data_ptr(), None, grp.metrics.data_ptr(), grp.workspace.data_ptr(),
                                           grp.workspace.numel(), 1 << s, stream.cuda_stream))
        evs[i][4].record(stream)
    torch.cuda.synchronize()
    grp.learn_step_host += K
    stage_ms = {nm: statistics.mean(evs[i][s].elapsed_time(evs[i][s + 1]) for i in range(K)) for s, nm in enumerate(stage_names)}

    # ---- the other kernels of the path, device-resident (actions/s is BASELINE's second metric) ----
    obs_dev = grp.obs[:, 0, :].contiguous()
    eps0 = torch.zeros(n, dtype=torch.float64, device=grp.device)
    w_zero = torch.zeros(n, dtype=torch.int32, device=grp.device)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b_.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(b_) / reps
    act_api_ms = timed(lambda: grp.act(obs_dev, eps0, w_zero, w_zero), min(4 * K, 400))   # eps = 0: every agent runs its network
    act_ms = act_api_ms
    try:        # the Python call costs about as much as the kernel: replay 16 captured launches for the device time
        side_stream = torch.cuda.Stream()
        side_stream.wait_stream(stream)
        with torch.cuda.stream(side_stream):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side_stream):
                for _ in range(16):
                    grp.act(obs_dev, eps0, w_zero, w_zero)
            for _ in range(3):
                graph.replay()
            side_stream.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record(side_stream)
            for _ in range(20):
                graph.replay()
            b_.record(side_stream)
            side_stream.synchronize()
            act_ms = a_.elapsed_time(b_) / (20 * 16)
        stream.wait_stream(side_stream)
    except Exception as exc:                                                   # noqa: BLE001
        print(f"act graph timing unavailable ({exc}); reporting the API loop", file=sys.stderr)
    halting = torch.randint(0, 20, (n, 12), dtype=torch.int32, device=grp.device)
    zi = torch.zeros(n, dtype=torch.int32, device=grp.device); zd = torch.zeros(n, dtype=torch.float64, device=grp.device)
    zv = torch.zeros(n, dtype=torch.uint8, device=grp.device)
    side = max(r for r in range(1, int(n ** 0.5) + 1) if n % r == 0)          # this GPU's shard of the grid: side x n/side
    from dmdqn_b200.parallel import grid_neighbor_table
    nbr = torch.as_tensor(grid_neighbor_table(side, n // side)).to(grp.device)
    assert nbr.shape[0] == n
    feat_ms = timed(lambda: grp.featurize(halting, zi, zd, zd, 0.0, zv, nbr), 4 * K)
    act_bytes = n * 4 * fb["params"]
    extra = {"act": {"value": n * world / (act_ms / 1e3), "unit": "actions/s", "ms": act_ms, "api_ms": act_api_ms, "eps": 0.0,
                     "timing": "16 launches captured in a CUDA graph, replayed 20x (device time); api_ms = the same call from Python",
                     "hbm_gbs": act_bytes / (act_ms / 1e3) / 1e9, "hbm_frac": act_bytes / (act_ms / 1e3) / 1e9 / peaks()["hbm_gbs"],
                     "bytes_per_action": 4 * fb["params"]},
             "featurize": {"ms": feat_ms, "agents_per_s": n * world / (feat_ms / 1e3), "bytes_per_agent": 64 + 96 * 4 + 17 * 8 + 8}}

    # ---- end to end: host transitions + draws in, losses out, every step ------------------
    # dmdqn_step_host (include/dmdqn_b200.h): the step's inputs sit in ONE pinned host block (struct of arrays);
    # the call queues one H2D copy, push, learn and the copy of the metrics back; the caller synchronises and reads.
    sb = grp.make_step_block()
    hv = sb["host"]
    hv["obs"].copy_(torch.randint(0, 20, (n, D)).float()); hv["next_obs"].copy_(torch.randint(0, 20, (n, D)).float())
    hv["act"].copy_(torch.randint(0, A, (n,), dtype=torch.int32)); hv["rew"].copy_(-torch.rand(n, dtype=torch.float64) * 100)
    hv["done"].zero_(); hv["draws"].copy_(torch.randint(0, 2**31, (n, b), dtype=torch.int32))
    h2d = sb["bytes"]
    d2h = sb["metrics_host"].numel() * sb["metrics_host"].element_size()

    def e2e_step():
        m = grp.step_host(sb)                                  # remember + replay (train.py:274-292), host buffers
        stream.synchronize()                                   # the caller reads the losses
        return float(m[0, 0])
    for _ in range(W):
        e2e_step()
    barrier()
    e0.record(stream)
    for _ in range(K):
        e2e_step()
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    tc_err = int(grp.debug_views()["tc_error"][0])
    if tc_err:
        raise RuntimeError(f"a tcgen05 kernel timed out on an mbarrier (code {tc_err}): the timings are invalid")
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=grp.device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    total_agents = n * world
    value = total_agents * K / (ms / 1e3)
    pk = peaks()
    dom = max(("target", "online", "wgrad_adam"), key=lambda k_: stage_ms[k_])
    kf = fb["kernels"][dom]
    ach_tf = n * kf["flops"] / (stage_ms[dom] / 1e3) / 1e12
    ffma_peak = 148 * 128 * 2 * (clocks["sm_mhz"] or 1965.0) * 1e6 / 1e12
    out = {
        "metric": "agent-updates/sec", "value": value, "unit": "agent-updates/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic replay (observation-shaped integers, full rings), random-init weights",
        "config": {"workload": workload_name(args.workload), "precision": args.precision,
                   "agents_total": total_agents, "sharding": "agent ranges, no collectives" if world > 1 else "single GPU",
                   "l2_policy": "working set (replay ring 5.9 GB, theta/m/v 368 MB, scratch 200 MB at cfg3) exceeds the 126 MB L2",
                   "sample_mode": "fisher_yates (device draws)"},
        "e2e": {"value": total_agents * K / (ms_e2e / 1e3), "unit": "agent-updates/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K,
                "what": "dmdqn_step_host: one pinned host block (transition of every agent + draws) -> one H2D copy -> push -> learn -> losses to pinned host memory, stream synchronised and the loss read every step"},
        "gpu_launches": 4 * K,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": dom, "achieved": ach_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": ach_tf / pk["bf16_tflops_sustained"], "traffic": ncu_traffic(args.workload, args.precision, dom),
                     "algorithmic_bytes": n * kf["bytes"], "peak_source": pk["source"],
                     "note": (f"fp32 FFMA kernel: fraction of the fp32 FFMA peak at the sampled clock = {ach_tf / ffma_peak:.3f} of "
                              f"{ffma_peak:.1f} TFLOP/s") if args.precision == "fp32" else
                             ("tcgen05 kind::tf32; achieved counts ALGORITHMIC flops (each product is issued as "
                              f"{3 if args.precision == 'tf32x3' else 1} MMA(s)); peak is the measured bf16 figure (tf32 dense peak is half of it), so the "
                              f"ceiling of this precision is frac = {1 / (6 if args.precision == 'tf32x3' else 2):.3f}")},
        "kernels": {k_: {"ms": stage_ms[k_], "tflops": n * fb["kernels"][k_]["flops"] / (stage_ms[k_] / 1e3) / 1e12,
                         "gbs": n * fb["kernels"][k_]["bytes"] / (stage_ms[k_] / 1e3) / 1e9} for k_ in stage_names},
        **extra,
        "step": {"flops_per_agent_update": fb["flops"], "bytes_per_agent_update": fb["bytes"],
                 "tflops": n * fb["flops"] / (ms / K / 1e3) / 1e12, "hbm_gbs_algorithmic": n * fb["bytes"] / (ms / K / 1e3) / 1e9,
                 "hbm_frac": n * fb["bytes"] / (ms / K / 1e3) / 1e9 / pk["hbm_gbs"]},
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(w, budget_s=args.cpu_budget)
        emit(out)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------
def build_oracle_agents(w, n_agents, seed=0):
    """Faithful per-agent oracle objects with full deque buffers (reference structure:
    one DQNAgent per intersection, train.py:109-127)."""
    from oracle.dqn import OracleDQNAgent
    rng = np.random.default_rng(seed)
    agents = []
    c = w["capacity"]
    for i in range(n_agents):
        ag = OracleDQNAgent(D, A, f"J_{i}", dict(agent_cfg(w), seed=i))
        s = rng.integers(-1, 20, (c, 1, D)).astype(np.float32); s2 = rng.integers(-1, 20, (c, 1, D)).astype(np.float32)
        a = rng.integers(0, A, c); r = -0.3 * rng.integers(0, 240, c) - 0.7 * rng.integers(0, 240 * 256, c)
        dn = rng.random(c) < 1 / 240
        for t in range(c):
            ag.replay_buffer.buffer.append((s[t, 0], int(a[t]), float(r[t]), s2[t, 0], bool(dn[t])))
        agents.append(ag)
    return agents


def cpu_baseline(w, budget_s=15.0, sample_agents=None):
    """The oracle's sequential per-agent learn loop (train.py:274-292 + dqn_agent.py:328-380)
    on this box's host cores, bounded to ~budget_s of CPU work."""
    sample_agents = sample_agents or min(w["agents"], 16)
    random.seed(0)
    agents = build_oracle_agents(w, sample_agents)
    for ag in agents[:2]:
        ag.replay()                                   # warm-up
    t0 = time.perf_counter(); sweeps = 0
    while True:
        for ag in agents:
            ag.replay()
        sweeps += 1
        el = time.perf_counter() - t0
        if el > budget_s or sweeps >= 200:
            break
    return {"value": sample_agents * sweeps / el, "unit": "agent-updates/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{sample_agents} of {w['agents']} agents (full {w['capacity']}-deep deque buffers), {sweeps} learn sweeps, "
                      f"{el:.1f} s; sequential per-agent loop of the CPU PyTorch oracle (reference is TensorFlow: not installable)",
            "host_cpus": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; oracle/_ref does not exist because
    TensorFlow/Keras cannot be installed) on the host cores, same workload/metric."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    sample_agents = min(w["agents"], 16)
    random.seed(0)
    agents = build_oracle_agents(w, sample_agents)
    K, W = args.steps, args.warmup
    for _ in range(min(W, 3)):
        for ag in agents:
            ag.replay()
    t0 = time.perf_counter()
    for _ in range(K):
        for ag in agents:
            ag.replay()
    el = time.perf_counter() - t0
    v = sample_agents * K / el
    sample = (f"each step = one learn sweep over {sample_agents} of {w['agents']} agents (full {w['capacity']}-deep deque buffers); "
              "sequential per-agent loop")
    emit({
        "impl": "reference", "metric": "agent-updates/sec", "value": v, "unit": "agent-updates/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": K, "warmup": W, "ms_per_step": el / K * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic replay",
        "config": {"workload": workload_name(args.workload), "precision": "fp32 (CPU torch oracle)"},
        "cpu_baseline": {"value": v, "unit": "agent-updates/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "agent-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="tf32x3", choices=["fp32", "tf32x3", "tf32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
