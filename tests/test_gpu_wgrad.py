"""GPU: the token-contracted weight-gradient GEMM (include/vlpet.h vlpet_wgrad_bf16, tcgen05 with MN-major operands)
against a float64 matmul of the same bf16 values.  Products of bf16 values are exact in fp32, so the only error is the
fp32 accumulation order: tolerance 2e-6 relative (Frobenius)."""
import ctypes as C

import pytest
import torch

from tests.helpers import rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vlpet_b200._lib as L_
    return L_


def _run(L, pairs, Mtok, d, nout):
    arr = (L.WgradPair * len(pairs))(*pairs)
    L.check(L.lib.vlpet_wgrad_bf16(arr, len(pairs), Mtok, d, nout, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
            "vlpet_wgrad_bf16")
    torch.cuda.synchronize()


@pytest.mark.parametrize("Mtok,d,nout", [(64, 768, 96), (100, 768, 96), (1000, 768, 96), (5000, 768, 32), (777, 256, 64),
                                         (333, 768, 16), (20000, 768, 96), (129, 128, 48)])
def test_wgrad_pairs_match_fp64(L, Mtok, d, nout):
    g = torch.Generator(device="cuda").manual_seed(Mtok + d + nout)
    bf = torch.bfloat16
    pitch = nout + 8
    As = [torch.randn(Mtok, d, device="cuda", generator=g).to(bf) for _ in range(4)]
    Bs = []
    for i in range(4):
        b = torch.zeros(Mtok, pitch, device="cuda", dtype=bf)
        b[:, :nout] = torch.randn(Mtok, nout, device="cuda", generator=g).to(bf)
        b[:, nout] = 1.0
        Bs.append(b)
    outs = [torch.full((d, nout), 0.5, device="cuda"), torch.zeros(d, nout, device="cuda"),
            torch.zeros(nout, d, device="cuda"), torch.full((nout, d), -1.0, device="cuda")]
    biases = [torch.zeros(d, device="cuda"), torch.ones(d, device="cuda"), None, None]
    scales = [1.0, 0.3, 1.0, 2.0]
    tr = [0, 0, 1, 1]
    init = [o.clone() for o in outs]
    binit = [None if b is None else b.clone() for b in biases]
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)  # noqa: E731
    pairs = [L.WgradPair(A=p(As[i]), lda=d, B=p(Bs[i]), ldb=pitch, nb_valid=nout + 1 if biases[i] is not None else nout,
                         transposed=tr[i], out=p(outs[i]), bias=p(biases[i]), scale=scales[i]) for i in range(4)]
    _run(L, pairs, Mtok, d, nout)
    for i in range(4):
        ref = scales[i] * (As[i].double().T @ Bs[i][:, :nout].double())
        if tr[i]:
            ref = ref.T
        ref = ref + init[i].double()
        assert rel(outs[i].double().cpu().numpy(), ref.cpu().numpy()) < 2e-6, i
        if biases[i] is not None:
            bref = scales[i] * As[i].double().sum(0) + binit[i].double()
            assert rel(biases[i].double().cpu().numpy(), bref.cpu().numpy()) < 2e-6, ("bias", i)


def test_wgrad_single_pair_with_view_pitch(L):
    """A is a column slice view of a wider tensor (pitch != d)."""
    Mtok, d, nout = 300, 256, 32
    g = torch.Generator(device="cuda").manual_seed(5)
    big = torch.randn(Mtok, 512, device="cuda", generator=g).to(torch.bfloat16)
    A = big[:, 128:128 + d]
    B = torch.randn(Mtok, nout, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.zeros(d, nout, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    _run(L, [L.WgradPair(A=p(A), lda=512, B=p(B), ldb=nout, nb_valid=nout, transposed=0, out=p(out), bias=C.c_void_p(0),
                         scale=1.0)], Mtok, d, nout)
    ref = A.double().T @ B.double()
    assert rel(out.double().cpu().numpy(), ref.cpu().numpy()) < 2e-6


def test_wgrad_rejects_bad_arguments(L):
    x = torch.zeros(64, 768, device="cuda", dtype=torch.bfloat16)
    o = torch.zeros(768, 96, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    pair = L.WgradPair(A=p(x), lda=768, B=p(x), ldb=768, nb_valid=96, transposed=0, out=p(o), bias=C.c_void_p(0), scale=1.0)
    arr = (L.WgradPair * 1)(pair)
    assert L.lib.vlpet_wgrad_bf16(arr, 1, 64, 700, 96, C.c_void_p(0)) == -2          # d % 128 != 0 -> UNSUPPORTED
    assert L.lib.vlpet_wgrad_bf16(arr, 0, 64, 768, 96, C.c_void_p(0)) == -1          # no pairs -> BADARG
