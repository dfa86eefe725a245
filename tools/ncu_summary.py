#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full_x.csv [--traffic cfg3/tf32x3]
writes the per-launch metric table the judge reads and, with --traffic, the per-kernel
dram bytes per launch into profiles/ncu_traffic.json (bench.py's roofline.traffic)."""
import csv
import json
import os
import re
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
SHORT = {"tc_target": "target", "tc_online": "online", "tc_wgrad": "wgrad_adam", "sample_kernel": "sample", "act_kernel": "act",
         "featurize_kernel": "featurize", "push_kernel": "push", "target_kernel": "target", "online_kernel": "online",
         "wgrad_adam_kernel": "wgrad_adam"}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([re.sub(r"\(.*", "", r[i]) if hdr[i] == "Kernel Name" else r[i] for i in idx])
    print(f"wrote {out}: {len(data)} launches")
    if "--traffic" in sys.argv:
        key = sys.argv[sys.argv.index("--traffic") + 1]
        path = os.path.join(os.path.dirname(out), "ncu_traffic.json")
        tj = json.load(open(path)) if os.path.exists(path) else {}
        ent = tj.setdefault(key, {})
        kn, rd, wr, du = (hdr.index(x) for x in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        for r in data:
            name = next((v for k, v in SHORT.items() if k in r[kn]), None)
            if name:
                ent[name] = {"dram_bytes": float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]],
                             "ncu_duration_us": float(r[du].replace(",", "")), "capture": os.path.basename(out)}
        json.dump(tj, open(path, "w"), indent=1)
        print(f"updated {path} [{key}]: {sorted(ent)}")


if __name__ == "__main__":
    main()
