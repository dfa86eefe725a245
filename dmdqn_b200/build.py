"""Build libdmdqn_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with
the repo snapshot to the GPU box).  ``python -m dmdqn_b200.build [--force]``."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdmdqn_b200.so")
SOURCES = ["api.cu", "featurize.cu", "act.cu", "replay.cu", "learn.cu", "learn_tc.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libdmdqn_b200.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "dmdqn_b200.h"))
    return any(os.path.getmtime(p) > built for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if os.environ.get("DMDQN_TC_TIMING"):          # phase stamps printed by a few CTAs (profiling aid only)
        cmd.insert(1, "-DTC_TIMING")
        cmd.insert(1, "-DTC_EXP=" + os.environ.get("DMDQN_TC_EXP", "0"))
    if os.environ.get("DMDQN_TC_EXP_ONLY"):        # experiment switch without the stamps (results are wrong by design)
        cmd.insert(1, "-DWG_EXP=" + os.environ["DMDQN_TC_EXP_ONLY"])
    if os.environ.get("DMDQN_NVCC_DEFINES"):       # e.g. "DMDQN_NO_ROW_PREFETCH": A/B switches for tools/exp_timing.sh
        for dname in os.environ["DMDQN_NVCC_DEFINES"].split(","):
            cmd.insert(1, "-D" + dname)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
