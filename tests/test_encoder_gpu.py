"""End-to-end encoder parity: libcbas_b200 vs the fp32 CPU oracle (transformers DINOv3ViTModel wrapped the way
cbas.py wraps it) on identical synthetic frames and random-init weights, plus the fixture produced by the
reference's own encode_file.  Gates (BASELINE.json): cosine >= 0.999, max|d|/max|ref| <= 2e-2; also reported:
mean-centred cosine and per-layer residual-stream error."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cbas_b200.encoder import DinoEncoder  # noqa: E402
from oracle import encoder as oenc  # noqa: E402
from tests.gpu_util import cosine_rows, rel_err  # noqa: E402


@pytest.fixture(params=[2, 1], ids=["attn_tcgen05", "attn_mma_sync"])
def attention_impl(request):
    """Both attention kernels: tcgen05 and mma.sync (each rotates q and k in its prologue).  The knob is per encoder
    handle: the test applies it with `enc.set_option(_lib.OPT_ATTENTION_IMPL, attention_impl)`."""
    return request.param


def _centred_cosine(got, want):
    """Cosine after removing the frame-independent part (the mean oracle row): with random-init weights the plain
    cosine is ~1 whatever the input, this one measures the input-dependent component."""
    got, want = torch.as_tensor(got).double(), torch.as_tensor(want).double()
    return cosine_rows(got - want.mean(0), want - want.mean(0)).min().item()


def _report(tag, got, want):
    got, want = torch.as_tensor(got).double(), torch.as_tensor(want).double()
    cos = cosine_rows(got, want).min().item()
    cc = _centred_cosine(got, want)
    rel = rel_err(got, want)
    # input dependence: every GPU row must be nearest to ITS OWN oracle row, not to another frame's
    dist = torch.cdist(got, want)
    nn_ok = bool((dist.argmin(dim=1) == torch.arange(len(got))).all())
    print(f"[parity] {tag}: min cosine {cos:.6f}  mean-centred {cc:.6f}  max|d|/max|ref| {rel:.3e}  nearest-row {nn_ok}")
    return cos, nn_ok, rel


@torch.no_grad()
def _eager_autocast(model, frames, dtype, mode="reference", size=224):
    """The reference's own CUDA path (cbas.py:433-435): the same HF model on this GPU under torch.autocast, fp32 pixel
    values in, CLS row out.  dtype float16 is the reference's numerics of record (autocast's CUDA default), bfloat16
    the same path with the operand type BASELINE.json prescribes for this build."""
    m = model.to("cuda")
    try:
        x = (oenc.preprocess_reference(frames) if mode == "reference" else oenc.preprocess_processor(frames, size)).cuda()
        with torch.autocast(device_type="cuda", dtype=dtype):
            out = m(x).last_hidden_state[:, 0, :]
        return out.float().cpu()
    finally:
        model.to("cpu")


def _assert_centred(tag, got, want, model, frames, floor, mode="reference", size=224):
    """Gate on the input-dependent component of the embedding: mean-centred cosine >= `floor`, and its shortfall from 1
    at most twice (plus 1e-3) what torch-eager autocast with the SAME operand type (bf16) loses against the same fp32
    oracle - i.e. the loss is the 8-bit mantissa of the operands, not a defect of these kernels.  The fp16-autocast
    figure (the reference's own CUDA numerics, 11-bit mantissa) is printed next to it."""
    cc = _centred_cosine(got, want)
    cc_bf16 = _centred_cosine(_eager_autocast(model, frames, torch.bfloat16, mode, size), want)
    cc_fp16 = _centred_cosine(_eager_autocast(model, frames, torch.float16, mode, size), want)
    print(f"[parity] {tag}: mean-centred cosine ours {cc:.6f} | torch-eager autocast bf16 {cc_bf16:.6f}, fp16 {cc_fp16:.6f}")
    assert cc >= floor, f"{tag}: mean-centred cosine {cc}"
    assert (1.0 - cc) <= 2.0 * (1.0 - cc_bf16) + 1e-3, \
        f"{tag}: centred error {1 - cc:.3e} vs torch-eager bf16 autocast {1 - cc_bf16:.3e}"


# last column: floor of the mean-centred cosine.  With 8-bit-mantissa (bf16) operands the input-dependent part of a
# random-init embedding - a few percent of its norm - carries the rounding noise of twelve or twenty-four blocks:
# measured here 0.994 / 0.966 / 0.825 / 0.974 / 0.894 against 0.993 / 0.960 / 0.774 / 0.966 / 0.860 for torch-eager bf16
# autocast of the same model on the same GPU (and 0.9999 / 0.9994 / 0.995 / 0.9994 / 0.998 for fp16 autocast, the
# reference's own 11-bit-mantissa numerics).  The floors sit just under the measured values; the bound relative to
# torch-eager bf16 inside _assert_centred is the gate that says "operand rounding, not a kernel defect".
@pytest.mark.parametrize("arch,side,n,scale,cc_floor", [("vits16", 64, 5, 4.0, 0.99), ("vitb16", 224, 4, 3.0, 0.955),
                                                       ("vitb16", 256, 3, 1.0, 0.78), ("vitl16", 96, 3, 2.0, 0.965),
                                                       ("vitl16", 224, 2, 2.0, 0.87)])
def test_reference_mode_parity(arch, side, n, scale, cc_floor, attention_impl):
    from cbas_b200 import _lib
    if attention_impl >= 2 and not _lib.lib().cbas_b200_attention_tc_supported((side // 16) ** 2 + 5, 5, 1):
        pytest.skip("no tcgen05 attention kernel for this token count")
    model = oenc.build_hf_model(arch, seed=0, init_scale=scale)
    frames = oenc.synthetic_frames(n, side, side, seed=5)
    want = oenc.encode(model, frames, mode="reference")
    enc = DinoEncoder.from_hf_model(model, "cuda", max_frames=8)
    enc.set_option(_lib.OPT_ATTENTION_IMPL, attention_impl)
    fd = torch.from_numpy(frames).cuda()
    # per-layer taps first: the first layer that drifts is the one to look at
    hs = oenc.hidden_states(model, oenc.preprocess_reference(frames))
    for li in sorted({0, 1, 2, len(hs) - 1}):
        got = enc.debug_hidden(fd, li).cpu()
        r = rel_err(got, hs[li])
        print(f"[parity] {arch}@{side} residual stream after {li} blocks: max|d|/max|ref| {r:.3e}")
        assert r < 2e-2, f"layer tap {li}"
    got = enc.encode_u8(fd).cpu()
    cos, nn_ok, rel = _report(f"{arch}@{side} reference-mode", got, want)
    assert cos >= 0.999 and rel <= 2e-2
    assert nn_ok
    _assert_centred(f"{arch}@{side}", got, want, model, frames, cc_floor)
    # float-plane entry (DinoEncoder.__call__ contract, cbas.py:435,672)
    x = torch.from_numpy(frames[:, :, :, 1] / 255.0).float().unsqueeze(1)
    got2 = enc(x).squeeze(1).cpu()
    assert got2.shape == (n, model.config.hidden_size)
    assert rel_err(got2, got) < 1e-5


@pytest.mark.parametrize("n", [4, 24], ids=["single_cta_gemms", "cta_pair_gemms"])
def test_processor_mode_parity(n):
    """BASELINE configs[1] geometry (256-px clip -> HF processor -> ViT-B/16 at 224 px) against the fp32 oracle; 24
    frames = 4 824 token rows take the CTA-pair GEMM tiles the 512-frame production chunk runs on."""
    model = oenc.build_hf_model("vitb16", seed=1, init_scale=3.0)
    frames = oenc.synthetic_frames(n, 256, 256, seed=6)
    want = oenc.encode(model, frames, mode="processor", size=224)
    enc = DinoEncoder.from_hf_model(model, "cuda", preprocess="processor", image_size=224, max_frames=24)
    got = enc.encode_u8(torch.from_numpy(frames).cuda()).cpu()
    cos, nn_ok, rel = _report(f"vitb16 processor 256->224, {n} frames", got, want)
    assert cos >= 0.999 and rel <= 2e-2 and nn_ok
    _assert_centred(f"vitb16 processor, {n} frames", got, want, model, frames, 0.998, mode="processor", size=224)


@pytest.mark.parametrize("arch,side,n,layers,trained_px", [
    ("dinov2reg-s14", 112, 5, 12, 518), ("dinov2reg-b14", 256, 3, 12, 518), ("dinov2reg-s14", 256, 2, 4, 252)])
def test_dinov2_with_registers_parity(arch, side, n, layers, trained_px):
    """CBAS's default encoder family (cbas.py:1030-1033): 14-px patches (floor(256/14) = 18 per side, the last 4
    pixels unused), learned position embedding interpolated bicubic+antialias from the 37x37 training grid (used
    as stored when the grid already matches: trained_px 252 = 18 patches), no RoPE, biased keys - through the same DinoEncoder call as the reference makes (cbas.py:672-677)."""
    model = oenc.build_hf_dinov2_model(arch, seed=2, init_scale=3.0, num_hidden_layers=layers, image_size=trained_px)
    frames = oenc.synthetic_frames(n, side, side, seed=15)
    want = oenc.encode(model, frames, mode="reference")
    enc = DinoEncoder.from_hf_model(model, "cuda", max_frames=8)
    fd = torch.from_numpy(frames).cuda()
    hs = oenc.hidden_states(model, oenc.preprocess_reference(frames))
    for li in sorted({0, 1, len(hs) - 1}):
        got = enc.debug_hidden(fd, li).cpu()
        r = rel_err(got, hs[li])
        print(f"[parity] {arch}@{side} residual stream after {li} blocks: max|d|/max|ref| {r:.3e}")
        assert r < 2e-2, f"layer tap {li}"
    got = enc.encode_u8(fd).cpu()
    cos, nn_ok, rel = _report(f"{arch}@{side} reference-mode", got, want)
    assert cos >= 0.999 and rel <= 2e-2 and nn_ok
    x = torch.from_numpy(frames[:, :, :, 1] / 255.0).float().unsqueeze(1)
    assert rel_err(enc(x).squeeze(1).cpu(), got) < 1e-5


def test_dinov2_against_reference_encode_file_fixture(golden_dir):
    """_cls.h5 contents the reference's own encode_file produced with a DINOv2-with-registers model
    (oracle/gen_golden_dinov2.py)."""
    g = np.load(os.path.join(golden_dir, "encode_file_dinov2reg.npz"))
    side = int(g["side"])
    frames = oenc.synthetic_frames(5, side, side, seed=int(g["frames_seed"]))
    model = oenc.build_hf_dinov2_model("dinov2reg-b14", seed=int(g["model_seed"]), init_scale=float(g["init_scale"]),
                                       num_hidden_layers=int(g["layers"]))
    enc = DinoEncoder.from_hf_model(model, "cuda", max_frames=8)
    got = enc.encode_u8(torch.from_numpy(frames).cuda()).cpu()
    cos, nn_ok, rel = _report("fixture encode_file, dinov2-with-registers (reference run)", got, g["cls"].astype(np.float32))
    assert cos >= 0.999 and rel <= 2e-2 and nn_ok


def test_dinov2_with_registers_processor_mode():
    model = oenc.build_hf_dinov2_model("dinov2reg-s14", seed=3, init_scale=3.0, num_hidden_layers=6)
    frames = oenc.synthetic_frames(3, 256, 256, seed=16)
    want = oenc.encode(model, frames, mode="processor", size=224)
    enc = DinoEncoder.from_hf_model(model, "cuda", preprocess="processor", image_size=224, max_frames=8)
    got = enc.encode_u8(torch.from_numpy(frames).cuda()).cpu()
    cos, nn_ok, rel = _report("dinov2reg-s14 processor 256->224", got, want)
    assert cos >= 0.999 and rel <= 2e-2 and nn_ok


def test_against_reference_encode_file_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "encode_file_vitb.npz"))
    frames = oenc.synthetic_frames(6, 64, 64, seed=int(g["frames_seed"]))
    model = oenc.build_hf_model("vitb16", seed=int(g["model_seed"]), init_scale=float(g["init_scale"]))
    enc = DinoEncoder.from_hf_model(model, "cuda", max_frames=8)
    got = enc.encode_u8(torch.from_numpy(frames).cuda()).cpu()
    cos, nn_ok, rel = _report("fixture encode_file (reference run)", got, g["cls"].astype(np.float32))
    assert cos >= 0.999 and rel <= 2e-2 and nn_ok


@pytest.mark.parametrize("ln_fusion", [0, 1, 2], ids=["standalone_layernorm", "fused_layernorm", "norm1_fused"])
def test_rows_whose_mean_dwarfs_their_spread(ln_fusion):
    """Every token row sits at a mean of ~40 with a spread of ~1 (embedding bias and prefix tokens shifted), and every
    residual update moves the row mean again (o_proj / down_proj biases shifted).  The standalone LayerNorm kernels
    (default) take their statistics from the fp32 residual stream; with CBAS_OPT_LN_FUSION the QKV / up GEMMs read a
    bf16 copy of it, shifted by the row's previous mean precisely so that this case keeps its accuracy.  Same gates
    as the ordinary parity test, both ways."""
    from cbas_b200 import _lib
    model = oenc.build_hf_model("vitb16", seed=2, init_scale=3.0)
    with torch.no_grad():
        model.embeddings.patch_embeddings.bias += 40.0
        model.embeddings.cls_token += 40.0
        model.embeddings.register_tokens += 40.0
        for i, layer in enumerate(model.model.layer):
            layer.attention.o_proj.bias += 0.5 * (-1) ** i
            layer.mlp.down_proj.bias -= 0.8
    frames = oenc.synthetic_frames(4, 224, 224, seed=12)
    want = oenc.encode(model, frames, mode="reference")
    enc = DinoEncoder.from_hf_model(model, "cuda", max_frames=8)
    enc.set_option(_lib.OPT_LN_FUSION, ln_fusion)
    fd = torch.from_numpy(frames).cuda()
    hs = oenc.hidden_states(model, oenc.preprocess_reference(frames))
    print(f"[parity] large-mean rows: row mean {float(hs[1].mean(-1).abs().min()):.1f}.., spread {float(hs[1].std(-1).max()):.2f} max")
    assert float(hs[1].mean(-1).abs().min()) > 30.0 and float(hs[1].std(-1).max()) < 10.0  # the premise of the test
    for li in (0, 1, 6, len(hs) - 1):
        centre = hs[li].mean(-1, keepdim=True)  # judge the part LayerNorm keeps, not the offset it removes
        r = rel_err(enc.debug_hidden(fd, li).cpu() - centre, hs[li] - centre)
        print(f"[parity] large-mean rows, centred residual stream after {li} blocks: max|d|/max|ref| {r:.3e}")
        assert r < 2e-2
    cos, nn_ok, rel = _report("vitb16@224 rows with mean >> spread", enc.encode_u8(fd).cpu(), want)
    assert cos >= 0.999 and rel <= 2e-2 and nn_ok


@pytest.mark.parametrize("arch,side,n", [("vitb16", 224, 6), ("vits16", 256, 24), ("vitl16", 96, 3)])
def test_fused_layernorm_matches_standalone_layernorm(arch, side, n):
    """norm1 / norm2 inside the GEMM epilogues (CBAS_OPT_LN_FUSION 1) against the same encoder with standalone
    LayerNorm kernels between the GEMMs (the default): same weights, same function, different rounding points."""
    from cbas_b200 import _lib
    enc = DinoEncoder(f"synthetic:{arch}@4", "cuda", max_frames=24)
    frames = torch.from_numpy(oenc.synthetic_frames(n, side, side, seed=13)).cuda()
    enc.set_option(_lib.OPT_LN_FUSION, 0)
    plain = enc.encode_u8(frames)
    taps_p = enc.debug_hidden(frames, 2)
    for mode, what in ((1, "norm1 + norm2 fused"), (2, "norm1 fused (the default)")):
        enc.set_option(_lib.OPT_LN_FUSION, mode)
        fused = enc.encode_u8(frames)
        taps_f = enc.debug_hidden(frames, 2)
        print(f"[parity] {what} vs standalone LayerNorm {arch}@{side}: embeddings {rel_err(fused, plain):.3e}, "
              f"residual stream after 2 blocks {rel_err(taps_f, taps_p):.3e}")
        assert rel_err(taps_f, taps_p) < 6e-3 and rel_err(fused, plain) < 8e-3


@pytest.mark.parametrize("arch,side", [("vitb16", 224), ("vits16", 256), ("vitl16", 64)])
def test_last_layer_cls_only_matches_full_block(arch, side):
    """The production path skips, in the last block, every row the CLS pooling throws away; the kept row must
    agree with running the block on all tokens."""
    from cbas_b200 import _lib
    enc = DinoEncoder(f"synthetic:{arch}@7", "cuda", max_frames=8)
    frames = torch.from_numpy(oenc.synthetic_frames(8, side, side, seed=10)).cuda()
    pruned = enc.encode_u8(frames)
    enc.set_option(_lib.OPT_PRUNE_LAST_LAYER, 0)
    full = enc.encode_u8(frames)
    assert rel_err(pruned, full) < 6e-3  # same math; bf16 rounding points differ (fp32 vs tensor-core softmax)


def test_large_batch_uses_cta_pair_gemms():
    # n*T >= 4096 rows switches the GEMMs to CTA-pair tiles; the small-batch result is the single-CTA path
    enc = DinoEncoder("synthetic:vits16@5", "cuda", max_frames=24)
    frames = torch.from_numpy(oenc.synthetic_frames(24, 224, 224, seed=9)).cuda()
    big = enc.encode_u8(frames)            # 24*201 = 4824 rows -> pairs
    small = torch.cat([enc.encode_u8(frames[i:i + 8]) for i in range(0, 24, 8)])  # 1608 rows -> single CTA
    assert rel_err(big, small) < 1e-5


def test_serpentine_row_order_changes_nothing():
    """CBAS_OPT_SERPENTINE only changes the ORDER in which a kernel walks its row blocks (so that it starts on what its
    predecessor left in L2): every row's arithmetic is the same, so the embeddings must be bitwise equal.  24 frames of
    ViT-S/16 at 224 px = 4 824 rows: CTA-pair GEMMs, several row blocks per kernel, both attention tiles."""
    from cbas_b200 import _lib
    enc = DinoEncoder("synthetic:vits16@6", "cuda", max_frames=24)
    frames = torch.from_numpy(oenc.synthetic_frames(24, 224, 224, seed=21)).cuda()
    enc.set_option(_lib.OPT_SERPENTINE, 1)
    on = enc.encode_u8(frames).clone()
    enc.set_option(_lib.OPT_SERPENTINE, 0)
    off = enc.encode_u8(frames).clone()
    assert torch.equal(on, off)


def test_batch_independence_and_chunking():
    enc = DinoEncoder("synthetic:vits16@3", "cuda", max_frames=4)
    frames = torch.from_numpy(oenc.synthetic_frames(10, 64, 64, seed=8)).cuda()
    full = enc.encode_u8(frames)  # 4 + 4 + 2
    one = torch.cat([enc.encode_u8(frames[i:i + 1]) for i in range(10)])
    assert torch.equal(full, one)  # same kernels, same tiles per row -> bitwise equal


def test_full_chunk_properties_at_baseline_size():
    """BASELINE configs[1] chunk (512 frames of 256x256 -> ViT-B/16 at 224 px, 102 912 token rows): the oracle cannot
    run this size in test time, so check size-independent properties - a frame's embedding does not depend on what
    else is in the chunk (512 at once == 4 x 128, bitwise: same kernels, same per-row arithmetic), repeated runs
    are deterministic, outputs are finite and normalised by the final LayerNorm (row mean 0, variance 1 for the
    unit-gain synthetic weights), and a sample of frames matches their stand-alone encoding."""
    enc = DinoEncoder("synthetic:vitb16@11", "cuda", preprocess="processor", image_size=224, max_frames=512)
    g = torch.Generator(device="cuda").manual_seed(5)
    frames = torch.randint(0, 256, (512, 256, 256, 3), dtype=torch.uint8, device="cuda", generator=g)
    full = enc.encode_u8(frames)
    assert full.shape == (512, 768) and torch.isfinite(full).all()
    assert torch.equal(full, enc.encode_u8(frames))
    parts = torch.cat([enc.encode_u8(frames[i:i + 128]) for i in range(0, 512, 128)])
    assert torch.equal(full, parts)
    assert float(full.mean(dim=1).abs().max()) < 1e-4 and float((full.var(dim=1, unbiased=False) - 1).abs().max()) < 1e-3
    for i in (0, 257, 511):
        assert rel_err(enc.encode_u8(frames[i:i + 1]), full[i:i + 1]) < 1e-5  # single-CTA tiles vs CTA pairs


@pytest.mark.parametrize("preprocess", ["reference", "processor"])
def test_degenerate_frames(preprocess):
    """All-black, all-white and single-colour frames (a covered lens, a saturated sensor): finite output that matches
    the oracle - constant patches give LayerNorm rows with zero variance in the first block's input only through eps."""
    model = oenc.build_hf_model("vits16", seed=5, init_scale=3.0)
    frames = np.zeros((4, 64, 64, 3), np.uint8)
    frames[1] = 255
    frames[2, :, :, 1] = 128
    frames[3, ::2] = 255
    side = 64 if preprocess == "reference" else 48
    want = oenc.encode(model, frames, mode=preprocess, size=side)
    enc = DinoEncoder.from_hf_model(model, "cuda", preprocess=preprocess, image_size=side, max_frames=4)
    got = enc.encode_u8(torch.from_numpy(frames).cuda()).cpu()
    assert torch.isfinite(got).all()
    cos, _, rel = _report(f"degenerate frames, {preprocess}", got, want)
    assert cos >= 0.999 and rel <= 2e-2


def test_empty_batch():
    enc = DinoEncoder("synthetic:vits16", "cuda", max_frames=4)
    out = enc.encode_u8(torch.zeros(0, 64, 64, 3, dtype=torch.uint8, device="cuda"))
    assert out.shape == (0, 384)
