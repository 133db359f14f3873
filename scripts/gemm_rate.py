#!/usr/bin/env python
"""Per-launch time of tic_gemm_bf16 at small-batch shapes, as a CUDA graph of serialised launches (operands L2-resident, like
the 2nd..nth kernel of the c2 chains).  Separates the fixed cost per launch from the per-k-block rate and shows how the rate
depends on the number of CTAs pulling operands at the same time (the question behind cluster split-K).

    python scripts/gemm_rate.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import tic_b200.plan as P  # noqa: E402


def time_chain(fn, n=20, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    best = 1e9
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / n


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    kmax, nmax, M = 3072, 4608, 512
    A = torch.randn(M, kmax, device=dev, generator=gen).to(torch.bfloat16)
    A_lo = (torch.randn(M, kmax, device=dev, generator=gen) * 1e-3).to(torch.bfloat16)
    W = torch.randn(nmax, kmax, device=dev, generator=gen).to(torch.bfloat16)
    bias = torch.randn(nmax, device=dev, generator=gen)
    D = torch.empty(M, nmax, device=dev)
    print("M=%d  (us per launch; CTAs = 128x64 tiles unless the cost model picks wider)" % M)
    import ctypes
    from tic_b200 import capi

    def kc_of(m, N, K, ns=0):
        bn, ks, kc = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        capi.call("tic_gemm_plan", m, N, K, ns, 0, ctypes.byref(bn), ctypes.byref(ks), ctypes.byref(kc))
        return kc.value

    short = os.environ.get("GEMM_RATE_SHORT") == "1"
    for m in ((128,) if short else (128, 256, 512)):
        for N in ((512,) if short else (64, 256, 512, 768, 1536, 2304)):
            row = []
            for K in (256, 512, 768, 1536, 3072):
                us = time_chain(lambda: P.gemm(A, kmax, 0, W, kmax, 0, D, nmax, 0, m, N, K, bias=bias, relu=True))
                row.append("K=%d: %.2f (kc%d)" % (K, us, kc_of(m, N, K)))
            print("m=%d N=%d tiles64=%d | %s" % (m, N, (m // 128) * (N // 64), "  ".join(row)))
    # split operand (hi+lo): the dX GEMM of the fusion chain
    us = time_chain(lambda: P.gemm(A, kmax, 0, W, kmax, 1, D, nmax, 0, 512, 768, 768, A_lo=A_lo))
    print("dX-like: M=512 N=768 K=768 hi+lo, B MN-major: %.2f us" % us)
    # accumulate (split-K, capped at 64 CTAs for small problems): the dW_f GEMM
    D.zero_()
    us = time_chain(lambda: P.gemm(A, kmax, 1, W, kmax, 1, D, nmax, 0, 768, 1536, 512, accumulate=True))
    print("dW_f-like: M=768 N=1536 K=512 accumulate (A, B MN-major): %.2f us" % us)
    # an empty-ish kernel chain for the launch floor
    x = torch.zeros(32, device=dev)
    us = time_chain(lambda: x.add_(1.0))
    print("floor: serialised tiny elementwise kernels in a graph: %.2f us per launch" % us)


if __name__ == "__main__":
    main()
