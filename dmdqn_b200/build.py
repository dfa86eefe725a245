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


TIMING_LIB = os.path.join(HERE, "libdmdqn_b200_timing.so")


def build(force: bool = False, verbose: bool = False, timing: bool = False) -> str:
    """``timing=True`` builds the phase-stamp variant (-DTC_TIMING: a few CTAs printf clock64 deltas; results are
    unchanged) into its OWN file, loaded only when DMDQN_PROFILING_LIB is set: the product library is never
    replaced by a profiling build."""
    out = TIMING_LIB if timing else LIB
    if not timing and not force and not is_stale():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", out, *[os.path.join(CSRC, s) for s in SOURCES]]
    if timing:
        cmd.insert(1, "-DTC_TIMING=" + os.environ.get("DMDQN_TC_TIMING", "1"))    # 1: kernel totals, 2: + per-item phase stamps
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, timing="--timing" in sys.argv))
