"""Developer tool: correctness (all operand-major combinations, ragged sizes, split-K, 3xTF32) and timing of vnl_gemm_tf32."""
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tr = importlib.import_module("vnl-brax-imitation_b200.train_kernels")


def check(M, N, K, a_mn, b_mn, x3, splitk=1, bias=True):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    pad = lambda n: (n + 3) // 4 * 4
    A = torch.randn((K, pad(M)) if a_mn else (M, pad(K)), device="cuda", generator=g)
    B = torch.randn((K, pad(N)) if b_mn else (N, pad(K)), device="cuda", generator=g)
    bvec = torch.randn(N, device="cuda", generator=g) if bias else None
    C = torch.zeros(M, pad(N), device="cuda")
    tr.gemm(A, a_mn, B, b_mn, C, M, N, K, bias=bvec, x3=x3, splitk=splitk)
    torch.cuda.synchronize()
    Ad = (A[:, :M].T if a_mn else A[:, :K]).double()
    Bd = (B[:, :N].T if b_mn else B[:, :K]).double()
    want = Ad @ Bd.T + (bvec.double() if bias else 0)
    err = float((C[:, :N].double() - want).abs().max() / want.abs().max())
    return err


def main():
    worst1, worst3 = 0.0, 0.0
    for (M, N, K) in ((128, 128, 32), (128, 64, 64), (300, 60, 256), (5120, 256, 795), (777, 1024, 232), (256, 128, 5120), (795, 256, 640)):
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                e1 = check(M, N, K, a_mn, b_mn, False)
                e3 = check(M, N, K, a_mn, b_mn, True)
                es = check(M, N, K, a_mn, b_mn, True, splitk=3, bias=True)
                print("M %5d N %5d K %5d a_mn %d b_mn %d: tf32 %.2e  3xtf32 %.2e  3xtf32 split-K %.2e" % (M, N, K, a_mn, b_mn, e1, e3, es), flush=True)
                worst1, worst3 = max(worst1, e1), max(worst3, e3, es)
    # the 256 x 256 tile (picked when the 128 x 256 tiling needs a second wave, or with split-K when tiles x splits fill the machine)
    for (M, N, K, sk) in ((5376, 1024, 232, 1), (5120, 1024, 1024, 1), (5000, 512, 96, 1), (1024, 1024, 5120, 9), (1000, 512, 5120, 12), (1024, 1024, 5120, 2)):
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                e1 = check(M, N, K, a_mn, b_mn, False, splitk=sk)
                e3 = check(M, N, K, a_mn, b_mn, True, splitk=max(sk, (K + 255) // 256))
                print("big  M %5d N %5d K %5d splitk %2d a_mn %d b_mn %d: tf32 %.2e  3xtf32 %.2e" % (M, N, K, sk, a_mn, b_mn, e1, e3), flush=True)
                worst1, worst3 = max(worst1, e1), max(worst3, e3)
    print("worst tf32 %.2e, worst 3xtf32 %.2e" % (worst1, worst3))
    assert worst1 < 5e-3 and worst3 < 1e-4  # tensor-core accumulation truncates: the 3xTF32 error grows ~6e-8 per MMA step (use split-K to shorten chains)
    # timing
    for (M, N, K, a_mn, b_mn, name) in ((5120, 1024, 1024, 0, 1, "fwd"), (5120, 1024, 1024, 0, 0, "dgrad"), (1024, 1024, 5120, 1, 1, "wgrad"),
                                       (5376, 1024, 232, 0, 1, "v0 fwd"), (232, 1024, 5120, 1, 1, "v0 wgrad"),
                                       (16384, 1024, 1024, 0, 1, "fwd 16k"), (5120, 256, 796, 0, 1, "enc0 fwd")):
        A = torch.randn((K, M) if a_mn else (M, K), device="cuda")
        B = torch.randn((K, N) if b_mn else (N, K), device="cuda")
        C = torch.zeros(M, N, device="cuda")
        for sk in ((2, 4, 6, 9) if name == "wgrad" else ((9, 18) if name == "v0 wgrad" else (1,))):
            for _ in range(3):
                tr.gemm(A, a_mn, B, b_mn, C, M, N, K, splitk=sk)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                tr.gemm(A, a_mn, B, b_mn, C, M, N, K, splitk=sk)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            print("%-8s M %5d N %5d K %5d splitk %d: %.1f us, %.1f TFLOP/s (tf32)" % (name, M, N, K, sk, us, 2.0 * M * N * K / us / 1e6))


def majors():
    """wgrad shapes of the update (both operands MN-major, K = 5120 rows) against the split-K factor; no zeroing inside the timing"""
    K = 5120
    for (M, N, sks) in ((1024, 1024, (2, 4, 6, 9)), (232, 1024, (4, 9, 18)), (796, 256, (5, 10, 20)), (296, 128, (10, 20, 40)), (256, 128, (10, 20, 40)),
                        (128, 256, (10, 20, 40)), (256, 60, (10, 20, 40))):
        A = torch.randn(K, (M + 3) // 4 * 4, device="cuda")
        B = torch.randn(K, (N + 3) // 4 * 4, device="cuda")
        C = torch.zeros(M, (N + 3) // 4 * 4, device="cuda")
        for sk in sks:
            for _ in range(3):
                tr.gemm(A, 1, B, 1, C, M, N, K, splitk=sk, zero=False)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                tr.gemm(A, 1, B, 1, C, M, N, K, splitk=sk, zero=False)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            print("wgrad M %4d N %4d K %d splitk %2d: %.1f us, %.1f TFLOP/s" % (M, N, K, sk, us, 2.0 * M * N * K / us / 1e6), flush=True)


if __name__ == "__main__":
    if os.environ.get("MAJORS") == "1":
        majors()
        sys.exit(0)
    main()
