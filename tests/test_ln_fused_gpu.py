"""Fused LayerNorm (csrc/gemm_tcgen05.cuh): the residual GEMM that leaves a shifted bf16 copy and row statistics, and
the GEMM that normalises in its epilogue, against torch fp32 LayerNorm + matmul (HF modeling_dinov3_vit.py:433-448:
h += proj(...); up(norm2(h))).  Includes rows whose mean dwarfs their spread - the case the per-row shift exists for."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from cbas_b200 import _lib  # noqa: E402
from cbas_b200.encoder import fold_layernorm  # noqa: E402
from tests.gpu_util import gemm, gemm_ln_a, gemm_resid_ln, layernorm, ln_stats_init, rel_err  # noqa: E402


@pytest.fixture(params=[1, 2], ids=["cta1", "cta_pair"])
def cta_group(request):
    _lib.check(_lib.lib().cbas_b200_debug_gemm_cta_group(request.param), "cta_group")
    yield request.param
    _lib.lib().cbas_b200_debug_gemm_cta_group(0)


def _rows_with_big_means(M, D, g, mean_scale):
    """Unit-spread rows whose means are spread over +-mean_scale (per row), plus a few outlier columns."""
    h = torch.randn(M, D, device="cuda", generator=g)
    h[:, 7] *= 12.0  # a 'massive activation' column
    h += mean_scale * torch.randn(M, 1, device="cuda", generator=g)
    return h


def _stats_mean_var(stats, D):
    s, q, shift = stats[:, 0:16].double().sum(1), stats[:, 16:32].double().sum(1), stats[:, 32].double()
    mu_y = s / D
    return shift + mu_y, q / D - mu_y * mu_y


@pytest.mark.parametrize("D", [384, 768, 1024])
def test_ln_stats_init(D):
    g = torch.Generator(device="cuda").manual_seed(D)
    h = _rows_with_big_means(333, D, g, 40.0)
    hb, st = ln_stats_init(h)
    mean = h.double().mean(1)
    assert torch.allclose(st[:, 32].double(), mean, rtol=0, atol=1e-4)
    y = h.double() - mean[:, None]
    assert float((hb.double() - y).abs().max()) <= float(y.abs().max()) * 2.0 ** -8
    m, v = _stats_mean_var(st, D)
    assert torch.allclose(m, mean, atol=1e-4, rtol=0)
    assert torch.allclose(v, h.double().var(1, unbiased=False), rtol=1e-4)
    assert bool((st[:, 0:16] == 0).all()) and bool((st[:, 17:32] == 0).all())


@pytest.mark.parametrize("M,D,K,N2", [(515, 768, 768, 2304), (1000, 768, 3072, 3072), (4100, 768, 768, 768),
                                     (257, 384, 1536, 1152), (129, 1024, 1024, 4096), (1, 768, 64, 256)])
@pytest.mark.parametrize("mean_scale", [0.5, 60.0], ids=["small_mean", "mean_60x_spread"])
def test_residual_producer_then_normalising_consumer(M, D, K, N2, mean_scale, cta_group):
    g = torch.Generator(device="cuda").manual_seed(M + D + K)
    h0 = _rows_with_big_means(M, D, g, mean_scale)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(D, K, device="cuda", generator=g) * (0.5 / K ** 0.5)).to(torch.bfloat16)
    b = torch.randn(D, device="cuda", generator=g) * 0.3 + 0.2
    guard = torch.full((M + 2, D), 3.0, device="cuda")
    h = guard[:M]
    h.copy_(h0)
    _, st0 = ln_stats_init(h)
    want_h = h0.double() + a.double() @ w.double().T + b.double()
    hb, st1 = gemm_resid_ln(a, w, b, h, st0)
    torch.cuda.synchronize()
    # 1. the fp32 residual stream itself
    assert rel_err(h, want_h) < 2e-5
    assert bool((guard[M:] == 3.0).all()), "wrote past the last row"
    # 2. the new shift is the exact mean of the OLD rows; hb is the shifted copy; the sums describe the NEW rows
    shift = st1[:, 32].double()
    assert torch.allclose(shift, h0.double().mean(1), atol=2e-4 * max(1.0, mean_scale), rtol=0)
    y = h.double() - shift[:, None]
    assert float((hb.double() - y).abs().max()) <= float(y.abs().max()) * 2.0 ** -8
    m, v = _stats_mean_var(st1, D)
    assert torch.allclose(m, h.double().mean(1), atol=1e-3, rtol=1e-5)
    assert torch.allclose(v, h.double().var(1, unbiased=False), rtol=2e-3)
    # 3. the consumer: LN(h) W2^T + b2 (and its GELU) against fp32 torch, judged next to the unfused bf16 pipeline
    gamma = 1.0 + 0.3 * torch.randn(D, device="cuda", generator=g)
    beta = 0.2 * torch.randn(D, device="cuda", generator=g)
    w2 = torch.randn(N2, D, device="cuda", generator=g) * (1.0 / D ** 0.5)
    b2 = torch.randn(N2, device="cuda", generator=g) * 0.1
    wf, c1, c2 = fold_layernorm(w2.cpu(), b2.cpu(), gamma.cpu(), beta.cpu())
    wf, c1, c2 = wf.cuda(), c1.cuda(), c2.cuda()
    truth = F.layer_norm(h.double(), (D,), gamma.double(), beta.double(), 1e-5) @ w2.double().T + b2.double()
    xn = layernorm(h, gamma, beta)  # the standalone kernel: what the unfused pipeline fed its GEMM
    for epi in (0, 1):
        got = gemm_ln_a(hb, st1, wf, c1, c2, epi=epi).double()
        unfused = gemm(xn, w2.to(torch.bfloat16), b2, epi=epi).double()
        ref = F.gelu(truth) if epi else truth
        e_fused, e_unfused = rel_err(got, ref), rel_err(unfused, ref)
        print(f"[parity] fused LN M{M} D{D} K{K} N{N2} epi{epi} mean x{mean_scale}: fused {e_fused:.3e}  unfused {e_unfused:.3e}")
        assert e_fused < 8e-3, f"epi {epi}: rel err {e_fused}"
        assert e_fused < 2.0 * e_unfused + 1e-3


def test_chain_of_updates_keeps_the_shift_current():
    """Three residual updates in a row, statistics ping-ponging like proj / down do in a block: the shift follows the
    drifting row mean, and the result is bit-reproducible (no atomics in the statistics)."""
    M, D, K = 700, 768, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    h = _rows_with_big_means(M, D, g, 20.0)
    ref = h.double().clone()
    _, st = ln_stats_init(h)
    runs = []
    for rep in range(2):
        hh, s = h.clone(), st.clone()
        gg = torch.Generator(device="cuda").manual_seed(6)
        for step in range(3):
            a = torch.randn(M, K, device="cuda", generator=gg).to(torch.bfloat16)
            w = (torch.randn(D, K, device="cuda", generator=gg) * (1.0 / K ** 0.5)).to(torch.bfloat16)
            b = torch.full((D,), 5.0, device="cuda")  # every update moves the row mean by 5
            prev_mean = hh.double().mean(1)
            hb, s = gemm_resid_ln(a, w, b, hh, s)
            if rep == 0:
                ref = ref + a.double() @ w.double().T + b.double()
                assert torch.allclose(s[:, 32].double(), prev_mean, atol=1e-3, rtol=0)
        runs.append((hh.clone(), hb.clone(), s.clone()))
    assert rel_err(runs[0][0], ref) < 3e-5
    for x, y in zip(runs[0], runs[1]):
        assert torch.equal(x, y)


def test_rejects_bad_arguments():
    a = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(2048, 64, device="cuda", dtype=torch.bfloat16)
    h = torch.zeros(128, 2048, device="cuda")
    st = torch.zeros(128, 36, device="cuda")
    with pytest.raises(RuntimeError):  # N = 2048 needs more statistics slots than a row has
        gemm_resid_ln(a, w, None, h, st)
