"""The learner's panel GEMM kernels (tcgen05 3xTF32 and fp32 FFMA) against an fp64 matmul."""
import numpy as np
import pytest
import torch as th

from ma_league_b200 import _native as nat
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run_linear(A, W, bias, epi, aux, w_trans, use_tc, Nout, K):
    M = A.shape[0]
    Y = th.full((M, Nout + 3), float("nan"), device=DEV)       # padded ld: out-of-range columns must stay untouched
    ldw = W.stride(0)
    nat.check(nat.lib().mal_debug_linear(M, K, Nout, nat.ptr(A), A.stride(0), nat.ptr(W), ldw, int(w_trans),
                                         nat.ptr(bias), epi, nat.ptr(aux), aux.stride(0) if aux is not None else 0,
                                         nat.ptr(Y), Y.stride(0), int(use_tc), nat.current_stream()), "mal_debug_linear")
    th.cuda.synchronize()
    assert bool(th.isnan(Y[:, Nout:]).all())
    return Y[:, :Nout]


@pytest.mark.parametrize("use_tc", [2, 3, 0])   # 2: warp-specialised pipelined tcgen05 kernel, 3: one-tile-at-a-time tcgen05 kernel, 0: FFMA
@pytest.mark.parametrize("M,K,Nout,epi,w_trans", [
    (1000, 64, 192, 0, 0), (128, 64, 64, 1, 0), (333, 59, 64, 1, 0), (700, 80, 64, 0, 0), (513, 192, 64, 2, 1),
    (260, 64, 160, 0, 0), (90, 64, 32, 0, 0), (400, 160, 64, 2, 1), (300, 64, 640, 0, 0), (257, 320, 128, 1, 0),
    (1, 8, 16, 0, 0), (130, 214, 64, 1, 0), (64, 64, 96, 0, 0), (32160, 64, 192, 0, 0),
])
def test_linear_matches_fp64(M, K, Nout, epi, w_trans, use_tc):
    g = th.Generator().manual_seed(M + K + Nout)
    A = th.randn(M, K + 5, generator=g).to(DEV)[:, :K]                       # non-trivial lda
    Wm = (th.randn(Nout, K, generator=g) / np.sqrt(K)).to(DEV)
    W = Wm.t().contiguous() if w_trans else Wm                                 # w_trans: stored [K, Nout]
    bias = th.randn(Nout, generator=g).to(DEV) if epi != 2 else None
    aux = th.randn(M, Nout, generator=g).to(DEV) if epi == 2 else None
    Y = run_linear(A, W, bias, epi, aux, w_trans, use_tc, Nout, K)
    ref = A.double() @ Wm.double().t()
    if bias is not None:
        ref = ref + bias.double()
    if epi == 1:
        ref = ref.clamp_min(0)
    if epi == 2:
        ref = ref * (aux > 0)
    from tests.helpers import rel_err, max_err
    r, m = rel_err(Y.cpu().numpy(), ref.cpu().numpy()), max_err(Y.cpu().numpy(), ref.cpu().numpy())
    print("linear M=%d K=%d N=%d tc=%d: rel %.2e max %.2e" % (M, K, Nout, use_tc, r, m))
    # 3xTF32 keeps ~21 mantissa bits per product (dropped lo*lo term, TF32 rounding of the low parts)
    assert_close(Y.cpu().numpy(), ref.cpu().numpy(), 4e-6 if use_tc else 1e-6, "Y")


def test_tc_accuracy_is_fp32_level_not_tf32():
    """3xTF32 must beat plain TF32 by orders of magnitude (the 1e-5 parity bar rules plain TF32 out)."""
    g = th.Generator().manual_seed(0)
    A = th.randn(4096, 64, generator=g).to(DEV)
    W = th.randn(192, 64, generator=g).to(DEV)
    ref = (A.double() @ W.double().t()).cpu().numpy()
    y_tc = run_linear(A, W, None, 0, None, 0, 3, 192, 64).cpu().numpy()
    y_tp = run_linear(A, W, None, 0, None, 0, 2, 192, 64).cpu().numpy()
    assert np.abs(y_tp - ref).max() / np.abs(ref).max() < 4e-6
    y_ff = run_linear(A, W, None, 0, None, 0, 0, 192, 64).cpu().numpy()
    e_tc = np.abs(y_tc - ref).max() / np.abs(ref).max()
    e_ff = np.abs(y_ff - ref).max() / np.abs(ref).max()
    print('max err tc %.2e ffma %.2e' % (e_tc, e_ff))
    assert e_tc < 4e-6 and e_ff < 1e-6, (e_tc, e_ff)
