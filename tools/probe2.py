"""Build and run tools/umma2_probe.cu on the GPU box: one cta_group::2 tf32 GEMM (D[256][256] = A[256][K] B[K][256])
against float64, A from shared memory and A from TMEM.  python tools/probe2.py [--build-only]"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libumma2_probe.so")


def build():
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-o", SO, os.path.join(HERE, "umma2_probe.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stdout + r.stderr)


def main():
    build()
    if "--build-only" in sys.argv:
        return
    import torch
    lib = C.CDLL(SO)
    lib.umma2_probe.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    torch.manual_seed(0)
    for K in (32, 96, 256):
        for a_tmem in (0, 1):
            if a_tmem and K > 256:
                continue
            A = torch.randn(256, K, device="cuda"); B = torch.randn(K, 256, device="cuda")
            D = torch.full((256, 256), float("nan"), device="cuda")
            st = torch.zeros(1, dtype=torch.int32, device="cuda")
            rc = lib.umma2_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), K, a_tmem, st.data_ptr(), None)
            torch.cuda.synchronize()
            ref = A.double() @ B.double()
            err = (D.double() - ref).abs()
            rel = err.max().item() / ref.abs().max().item()
            quad = [[(err[r * 128:(r + 1) * 128, c * 128:(c + 1) * 128].max().item() / ref.abs().max().item()) for c in range(2)] for r in range(2)]
            print(f"K={K:3d} A from {'TMEM' if a_tmem else 'smem'}: rc={rc} status={int(st[0])} rel_err={rel:.3e} per 128x128 quadrant {quad}", flush=True)


if __name__ == "__main__":
    main()
