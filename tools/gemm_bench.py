#!/usr/bin/env python
"""Time the learner's panel GEMM (mal_debug_linear) in isolation: CUDA events around back-to-back launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from ma_league_b200 import _native as nat

nat.build()
lib = nat.lib()
DEV = "cuda"


def bench(M, K, Nout, epi=0, w_trans=0, use_tc=1, reps=20, pipelined=1):
    A = th.randn(M, K, device=DEV)
    W = th.randn(K, Nout, device=DEV) if w_trans else th.randn(Nout, K, device=DEV)
    bias = th.randn(Nout, device=DEV) if epi != 2 else None
    aux = th.randn(M, Nout, device=DEV) if epi == 2 else None
    Y = th.empty(M, Nout, device=DEV)
    def run():
        nat.check(lib.mal_debug_linear(M, K, Nout, nat.ptr(A), A.stride(0), nat.ptr(W), W.stride(0), int(w_trans),
                                       nat.ptr(bias), epi, nat.ptr(aux), aux.stride(0) if aux is not None else 0,
                                       nat.ptr(Y), Y.stride(0), (2 if pipelined else 3) if use_tc else 0, nat.current_stream()), "dbg")
    for _ in range(3):
        run()
    th.cuda.synchronize()
    s, e = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        run()
    e.record()
    th.cuda.synchronize()
    us = s.elapsed_time(e) / reps * 1e3
    byts = 4 * M * (K + Nout)
    print("M=%6d K=%3d N=%3d epi=%d wt=%d tc=%d pipe=%d: %7.1f us  %6.0f GB/s  %5.1f TFLOP/s" % (
        M, K, Nout, epi, w_trans, use_tc, pipelined, us, byts / us / 1e3, 2.0 * M * K * Nout / us / 1e6))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "light":
    for dbg in (0, 1, 2, 4, 8, 3, 6, 7, 15):
        lib.mal_set_option(b"tc_dbg", dbg)
        print("light dbg", dbg, end="  ")
        bench(32160, 64, 192, pipelined=0)
    lib.mal_set_option(b"tc_dbg", 0)
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1:
    for dbg in (0, 1, 2, 4, 8, 3, 6, 7, 15):
        lib.mal_set_option(b"tc_dbg", dbg)
        print("dbg", dbg, end="  ")
        bench(257280, 64, 192)
    lib.mal_set_option(b"tc_dbg", 0)
    sys.exit(0)

if __name__ == "__main__":
    for pipe in (1, 0):
        bench(32160, 64, 192, pipelined=pipe)
        bench(32160, 64, 64, pipelined=pipe)
        bench(32160, 192, 64, epi=2, w_trans=1, pipelined=pipe)
        bench(257280, 64, 192, pipelined=pipe)
    bench(32160, 64, 192, use_tc=0)
