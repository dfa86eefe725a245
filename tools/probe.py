"""Build and run tools/umma_probe.cu on the GPU box: checks the UMMA descriptor / layout
assumptions of the TF32 path against float64 matmul.  python tools/probe.py [--build-only]"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libumma_probe.so")


def build():
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-o", SO, os.path.join(HERE, "umma_probe.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stdout + r.stderr)


def main():
    if not os.path.exists(SO) or "--build-only" in sys.argv:
        build()
    if "--build-only" in sys.argv:
        return
    import torch
    lib = C.CDLL(SO)
    lib.umma_probe.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_void_p]
    torch.manual_seed(0)
    names = {0: "K-major SW128", 1: "MN-major B32", 2: "K-major SW64"}
    anames = {0: "A K-major", 1: "A MN-major", 2: "A_lo via TMEM"}
    for K, N, b_mn, three, a_mode in [(64, 256, 0, 0, 0), (64, 256, 1, 0, 0), (64, 256, 2, 0, 0), (32, 256, 2, 0, 0),
                                      (64, 256, 0, 0, 1), (64, 256, 1, 0, 1), (64, 256, 0, 1, 0), (64, 256, 1, 1, 0),
                                      (64, 256, 2, 1, 0), (64, 256, 1, 1, 1), (64, 256, 1, 1, 2), (64, 256, 2, 1, 2),
                                      (64, 256, 1, 2, 0), (64, 256, 1, 2, 2), (64, 256, 2, 2, 2)]:
        A = torch.randn(128, K, device="cuda")
        B = torch.randn((K, N) if b_mn == 1 else (N, K), device="cuda")
        D = torch.full((128, N), float("nan"), device="cuda")
        st = torch.zeros(4, dtype=torch.int32, device="cuda")
        A_in = A.t().contiguous() if a_mode == 1 else A          # MN-major A is given as [K][128]
        rc = lib.umma_probe(A_in.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, b_mn, three, a_mode, st.data_ptr(), 1, None)
        torch.cuda.synchronize()
        ref = A.double() @ (B.double() if b_mn == 1 else B.double().T)
        err = (D.double() - ref).abs().max().item()
        rel = err / ref.abs().max().item()
        t1 = torch.zeros(4, dtype=torch.int32, device="cuda"); t2 = torch.zeros(4, dtype=torch.int32, device="cuda")
        lib.umma_probe(A_in.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, b_mn, three, a_mode, t1.data_ptr(), 8, None)
        lib.umma_probe(A_in.data_ptr(), B.data_ptr(), D.data_ptr(), K, N, b_mn, three, a_mode, t2.data_ptr(), 40, None)
        torch.cuda.synchronize()
        cyc = (int(t2[1]) - int(t1[1])) / max(1, int(t2[2]) - int(t1[2]))
        cyc2 = (int(t2[3]) - int(t1[3])) / ((40 - 8) * 4 * 8)
        print(f"K={K:3d} N={N:3d} B={names[b_mn]:14s} {anames[a_mode]:14s} {['TF32  ', '3xTF32', '3xTF32 raw-hi'][three]} rc={rc} status={int(st[0])} cyc/MMA={cyc:6.1f} back-to-back={cyc2:6.1f} "
              f"max_abs_err={err:.3e} rel={rel:.3e}", flush=True)


if __name__ == "__main__":
    main()
