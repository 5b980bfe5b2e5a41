"""tcgen05 GEMM against torch.matmul (fp32 reference of the same bf16 operands)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cbas_b200 import _lib  # noqa: E402
from tests.gpu_util import gemm, rel_err  # noqa: E402


@pytest.fixture(params=[1, 2], ids=["cta1", "cta_pair"], autouse=True)
def cta_group(request):
    """Every case runs on single-CTA tiles (128 x N) and on CTA-pair tiles (tcgen05 cta_group::2, 256 x N)."""
    _lib.check(_lib.lib().cbas_b200_debug_gemm_cta_group(request.param), "cta_group")
    yield request.param
    _lib.lib().cbas_b200_debug_gemm_cta_group(0)


def _mk(M, N, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g)).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g)
    return a, w, b


@pytest.mark.parametrize("M,N,K", [
    (128, 128, 64), (128, 256, 64), (128, 256, 128), (300, 256, 256), (657, 768, 768), (1000, 384, 384),
    (4021, 2304, 768), (2010, 3072, 768), (2010, 768, 3072), (333, 1152, 384), (129, 1024, 4096), (64, 128, 256),
])
def test_gemm_bias_f32(M, N, K):
    a, w, b = _mk(M, N, K)
    out = gemm(a, w, b, epi=4)
    want = a.float() @ w.float().T + b
    torch.cuda.synchronize()
    e = rel_err(out, want)
    assert e < 2e-5, f"M{M} N{N} K{K}: rel err {e}; first bad rows {((out-want).abs().amax(1) > 1e-3).nonzero()[:8].flatten().tolist()}"


@pytest.mark.parametrize("epi", [0, 1])
def test_gemm_bf16_epilogues(epi):
    a, w, b = _mk(777, 768, 768, seed=1)
    out = gemm(a, w, b, epi=epi).float()
    want = a.float() @ w.float().T + b
    if epi == 1:
        want = torch.nn.functional.gelu(want)
    assert rel_err(out, want) < 6e-3  # bf16 output rounding (2^-9)


def test_gemm_gelu_f32():
    a, w, b = _mk(1029, 256, 1152, seed=4)
    out = gemm(a, w, b, epi=5)
    want = torch.nn.functional.gelu(a.float() @ w.float().T + b)
    assert rel_err(out, want) < 5e-6  # fp32 epilogue GELU: |error| <= 5.1e-7 absolute (gelu_erf_fast<5>)


def test_gemm_residual_inplace():
    a, w, b = _mk(515, 768, 3072, seed=2)
    h = torch.randn(515, 768, device="cuda")
    want = h + a.float() @ w.float().T + b
    gemm(a, w, b, epi=2, out=h)
    assert rel_err(h, want) < 2e-5


def test_gemm_no_bias_and_many_tiles():
    a, w, _ = _mk(128 * 150 + 5, 256, 64, seed=3)  # more tiles than SMs: persistent loop + phase wrap
    out = gemm(a, w, None, epi=4)
    want = a.float() @ w.float().T
    assert rel_err(out, want) < 2e-5


def test_gemm_rejects_bad_shapes():
    a, w, b = _mk(128, 100, 64)
    with pytest.raises(RuntimeError):
        gemm(a, w, b, epi=4)


def test_gemm_ragged_rows_every_epilogue():
    """M on both sides of every tile boundary (1 row, 127/128/129, 255/256/257, the 4096-row switch to CTA pairs in
    auto mode) x every staged epilogue: partial tiles are clipped by the TMA store / reduction, never written past M."""
    worst = 0.0
    for M in (1, 127, 128, 129, 255, 256, 257, 4095, 4096, 4097):
        for N, K in ((256, 64), (384, 128), (768, 192)):
            a, w, b = _mk(M, N, K, seed=M + N)
            want = a.float() @ w.float().T + b
            for epi in (0, 1, 2, 4, 5):
                guard = torch.full((M + 3, N), 7.0, device="cuda",
                                   dtype=torch.bfloat16 if epi in (0, 1) else torch.float32)
                out = guard[:M]
                if epi == 2:
                    out.fill_(0.5)
                gemm(a, w, b, epi=epi, out=out)
                ref = want
                if epi in (1, 5):
                    ref = torch.nn.functional.gelu(want)
                if epi == 2:
                    ref = want + 0.5
                e = rel_err(out.float(), ref)
                worst = max(worst, e)
                assert e < (6e-3 if epi in (0, 1) else 3e-5), f"M{M} N{N} K{K} epi{epi}: rel err {e}"
                assert bool((guard[M:] == 7.0).all()), f"M{M} N{N} K{K} epi{epi}: wrote past the last row"
    print(f"[parity] ragged GEMM sweep: worst rel err {worst:.3e}")
