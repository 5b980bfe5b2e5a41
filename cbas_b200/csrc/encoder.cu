// DINOv3 ViT encoder forward on sm_100a: preprocess -> patch-embed GEMM -> L x (LN, QKV GEMM, attention with
// RoPE prologue, proj GEMM + residual, LN, up GEMM + GELU, down GEMM + residual) -> final LN on the CLS rows.
// Reference path: cbas.py:672-677 DinoEncoder.forward -> transformers DINOv3ViTModel.forward
// (modeling_dinov3_vit.py:530-555).  Residual stream is fp32; GEMM operands are bf16 with fp32 accumulation.
#include "../../include/cbas_b200.h"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "attention_tc_split.cuh"
#include "common.h"
#include "gemm_tcgen05.cuh"
#include "layernorm.cuh"
#include "preprocess.cuh"

#include <algorithm>
#include <atomic>
#include <vector>

using namespace cbas;

namespace {

// float planes [n,H,W] in [0,1] (DinoEncoder.__call__ input, cbas.py:435) -> folded-K patch matrix (x*255 as bf16)
__global__ void __launch_bounds__(256)
preprocess_plane_kernel(const float* __restrict__ planes, __nv_bfloat16* __restrict__ A, int n_frames, int H, int W) {
    const int nw = W >> 4, nh = H >> 4;
    const long long total = (long long)n_frames * H * nw;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int px = gid % nw;
    const long long t = gid / nw;
    const int y = t % H;
    const int f = t / H;
    const float4* src = reinterpret_cast<const float4*>(planes + ((long long)f * H + y) * W + px * 16);
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(src + i);
        o[2 * i] = pack_bf16(v.x * 255.0f, v.y * 255.0f);
        o[2 * i + 1] = pack_bf16(v.z * 255.0f, v.w * 255.0f);
    }
    __nv_bfloat16* dst = A + ((long long)f * nh * nw + (long long)(y >> 4) * nw + px) * 256 + (y & 15) * 16;
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

template <typename OutT>
int launch_layernorm(const float* in, long long in_row_stride, const float* g, const float* b, OutT* out, int rows,
                     int D, float eps, cudaStream_t s, int tag = PROF_LAYERNORM) {
    if (rows <= 0) return 0;
    ProfScope prof(tag, s);
    const int threads = 256, rows_per_block = threads / 32;
    const int grid = (rows + rows_per_block - 1) / rows_per_block;
    switch (D) {
        case 384: layernorm_kernel<384, OutT><<<grid, threads, 0, s>>>(in, in_row_stride, g, b, out, rows, eps); break;
        case 768: layernorm_kernel<768, OutT><<<grid, threads, 0, s>>>(in, in_row_stride, g, b, out, rows, eps); break;
        case 1024: layernorm_kernel<1024, OutT><<<grid, threads, 0, s>>>(in, in_row_stride, g, b, out, rows, eps); break;
        case 128: layernorm_kernel<128, OutT><<<grid, threads, 0, s>>>(in, in_row_stride, g, b, out, rows, eps); break;
        case 256: layernorm_kernel<256, OutT><<<grid, threads, 0, s>>>(in, in_row_stride, g, b, out, rows, eps); break;
        default: return fail("LayerNorm width " + std::to_string(D) + " not instantiated (384/768/1024)");
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "layernorm_kernel launch");
}

int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, const float* cs, const float* sn, int frames,
                     int T, int prefix, int heads, cudaStream_t s) {
    if (frames <= 0) return 0;
    ProfScope prof(PROF_ATTENTION, s);
    const int TP = (T + 15) & ~15;
    const int smem = 3 * TP * 128;
    static std::atomic<int> configured_smem{0};  // two host threads may race here: the attribute call is idempotent
    if (smem > configured_smem) {
        CBAS_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_smem = smem;
    }
    const float scale_log2 = 0.125f * 1.4426950408889634f;  // head_dim^-0.5 * log2(e)
    if (!cs || !sn) prefix = T;  // no rotary embedding (DINOv2): no token is a "patch token" for the rotation
    // warps per CTA: the one in 4..8 that wastes the fewest 16-row query-tile slots (ties: more warps hide more latency)
    const int m_tiles = TP / 16;
    int warps = 4;
    double best = 0.0;
    for (int w = 4; w <= ATT_MAX_THREADS / 32; ++w) {
        const double eff = (double)m_tiles / (double)(((m_tiles + w - 1) / w) * w);
        if (eff >= best) { best = eff; warps = w; }
    }
    attention_kernel<<<frames * heads, warps * 32, smem, s>>>(qkv, out, cs, sn, T, prefix, heads, heads * ATT_HEAD_DIM,
                                                            scale_log2);
    count_launch();
    return check_cuda(cudaGetLastError(), "attention_kernel launch");
}

bool g_prune_last_layer = true;  // false: run the last block on every token (test knob)
int g_resize_tiled = 2;  // test knob: 0 per-pixel kernel, 1 general tiled kernel, 2 column-per-thread kernel when it applies
long long* g_attention_trace = nullptr;  // device buffer [64][ATC_TRACE_SLOTS] for the stage-timing aid (tools/attn_trace.py)
int g_attention_impl = 0;  // 0 auto (tcgen05 when T <= 256), 1 mma.sync, 2 tcgen05 (both rotate q,k in their prologue)

bool attention_tc_fits(int T, int prefix) {
    const int TK = (T + 15) & ~15;
    return TK <= 256 && atc_smem_bytes(TK, T, prefix, true) <= 232448;
}
// 257..384 tokens: the key-split tcgen05 kernel (attention_tc_split.cuh)
bool attention_tc_split_fits(int T, int prefix, bool rope) {
    const int TK = (T + 15) & ~15;
    return TK > 256 && TK <= 384 && ats_smem_bytes(TK, T, prefix, rope) <= 232448;
}
bool use_attention_tc(int T, int prefix, bool rope = true) {
    if (g_attention_impl == 1) return false;
    if (attention_tc_split_fits(T, prefix, rope)) return true;
    return g_attention_impl >= 2 || attention_tc_fits(T, prefix);
}

// cs/sn: RoPE tables applied in the kernel's prologue, or null when q and k arrive rotated (EPI_QKV_ROPE_BF16)
int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, const float* cs, const float* sn, int frames,
                        int T, int prefix, int heads, cudaStream_t s) {
    if (frames <= 0) return 0;
    const int TK = (T + 15) & ~15;
    if (TK > 256) {
        // key-split kernel for 257..384 tokens
        const bool rope = cs != nullptr && sn != nullptr;
        if (TK > 384 || !attention_tc_split_fits(T, prefix, rope))
            return fail("tcgen05 attention handles at most 384 tokens per frame");
        if (rope && (prefix < 0 || prefix >= T)) return fail("attention: bad prefix token count");
        ProfScope prof(PROF_ATTENTION, s);
        const int D = heads * 64;
        const long long M = (long long)frames * T;
        const int nq = (T + 127) / 128, TK0 = ats_key_block0(TK), TK1 = TK - TK0;
        CUtensorMap tq, tk0, tk1, to, to1;
        if (int rc = make_tmap_2d(&tq, qkv, false, (int)M, 3 * D, 3 * D, 64, 128)) return rc;
        if (int rc = make_tmap_2d(&tk0, qkv, false, (int)M, 3 * D, 3 * D, 64, TK0)) return rc;
        if (int rc = make_tmap_2d(&tk1, qkv, false, (int)M, 3 * D, 3 * D, 64, TK1)) return rc;
        if (int rc = make_tmap_3d_bf16(&to, out, D, T, frames, D, 64, 128)) return rc;
        if (int rc = make_tmap_3d_bf16(&to1, out, D, T, frames, D, 64, T - 128 * (nq - 1))) return rc;
        const int smem = ats_smem_bytes(TK, T, prefix, rope);
        static std::atomic<int> configured_split{0};
        if (smem > configured_split) {
            CBAS_CHECK(cudaFuncSetAttribute(attention_tc_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            configured_split = smem;
        }
        AttnTcParams p{qkv, out, frames, heads, T, TK, D, 0.125f * 1.4426950408889634f, rope ? cs : nullptr,
                       rope ? sn : nullptr, prefix, nullptr};
        const int items = frames * heads;
        const int grid = items < sm_count() ? items : sm_count();
        attention_tc_split_kernel<<<grid, ATC_THREADS, smem, s>>>(tq, tk0, tk1, to, to1, p);
        count_launch();
        return check_cuda(cudaGetLastError(), "attention_tc_split_kernel launch");
    }
    ProfScope prof(PROF_ATTENTION, s);
    const int D = heads * 64;
    const long long M = (long long)frames * T;
    CUtensorMap tq, tkv, to, to1;
    if (int rc = make_tmap_2d(&tq, qkv, false, (int)M, 3 * D, 3 * D, 64, 128)) return rc;
    if (int rc = make_tmap_2d(&tkv, qkv, false, (int)M, 3 * D, 3 * D, 64, TK)) return rc;
    // the output as [frames][T][D]: a 128-row store box is clipped at the end of ITS frame
    if (int rc = make_tmap_3d_bf16(&to, out, D, T, frames, D, 64, T < 128 ? T : 128)) return rc;
    to1 = to;
    if (T > 128)
        if (int rc = make_tmap_3d_bf16(&to1, out, D, T, frames, D, 64, T - 128)) return rc;
    const bool rope = cs != nullptr && sn != nullptr;
    if (rope && (prefix < 0 || prefix >= T)) return fail("attention: bad prefix token count");
    const int smem = atc_smem_bytes(TK, T, prefix, rope);
    if (smem > 232448) return fail("attention: frame does not fit in shared memory");
    static std::atomic<int> configured_smem{0};  // two host threads may race here: the attribute call is idempotent
    if (smem > configured_smem) {
        CBAS_CHECK(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured_smem = smem;
    }
    AttnTcParams p{qkv, out, frames, heads, T, TK, D, 0.125f * 1.4426950408889634f, rope ? cs : nullptr,
                   rope ? sn : nullptr, prefix, g_attention_trace};
    const int items = frames * heads;
    const int grid = items < sm_count() ? items : sm_count();
    attention_tc_kernel<<<grid, ATC_THREADS, smem, s>>>(tq, tkv, to, to1, p);
    count_launch();
    return check_cuda(cudaGetLastError(), "attention_tc_kernel launch");
}

int launch_preprocess_green(const uint8_t* frames, __nv_bfloat16* A, int n, int H, int W, long long fs, int rs,
                            cudaStream_t s) {
    if (n <= 0) return 0;
    if (H % 16 || W % 16) return fail("frame size must be a multiple of the 16-pixel patch");
    ProfScope prof(PROF_PREPROCESS, s);
    const long long total = (long long)n * H * (W / 16);
    preprocess_green_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(frames, A, n, H, W, fs, rs);
    count_launch();
    return check_cuda(cudaGetLastError(), "preprocess_green_kernel launch");
}

int launch_preprocess_resize(const uint8_t* frames, __nv_bfloat16* A, int n, int H, int W, long long fs, int rs,
                             int side, const ResizeTaps& tp, cudaStream_t s) {
    if (n <= 0) return 0;
    if (side % 16) return fail("resize target must be a multiple of the 16-pixel patch");
    if (tp.taps_x > RESIZE_MAX_TAPS * 4 || tp.taps_y > RESIZE_MAX_TAPS * 4) return fail("too many resize taps");
    ProfScope prof(PROF_PREPROCESS, s);
    const long long total = (long long)n * side * (side / 8);
    // ImageNet mean/std (transformers image_utils.IMAGENET_DEFAULT_MEAN / _STD)
    const float3 mean = make_float3(0.485f, 0.456f, 0.406f);
    const float3 istd = make_float3(1.0f / 0.229f, 1.0f / 0.224f, 1.0f / 0.225f);
    // tiled shared-memory path when the geometry allows it: 16-byte aligned rows, strip fits in shared memory
    const int max_rows = (int)((16LL * H + side - 1) / side) + tp.taps_y + 1;
    const int src_pitch = (W * 3 + 15) & ~15;
    const size_t tile_smem = (size_t)max_rows * src_pitch + (size_t)max_rows * side * 3 * 4;
    const bool aligned = (reinterpret_cast<uintptr_t>(frames) & 15) == 0 && fs % 16 == 0 && rs % 16 == 0;
    if (g_resize_tiled >= 2 && aligned && side <= 256 && tp.taps_x <= RESIZE_FAST_TAPS && tp.taps_y <= RESIZE_FAST_TAPS) {
        // production path: column-per-thread kernel (bitwise identical to the general tiled kernel below)
        const int io_bytes = std::max(max_rows * src_pitch, (side / 16) * 1536);
        const size_t smem = (size_t)io_bytes + (size_t)3 * max_rows * side * 4 + 16 * 32 + 32;  // + row taps, + slack for the word-wise reads of the last staged row
        if (smem <= 110 * 1024) {
            static std::atomic<size_t> configured_fast{0};
            if (smem > configured_fast) {
                CBAS_CHECK(cudaFuncSetAttribute(preprocess_resize_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
                configured_fast = smem;
            }
            preprocess_resize_fast_kernel<<<n * (side / 16), 256, smem, s>>>(frames, A, H, W, fs, rs, side, tp, mean, istd,
                                                                            max_rows, io_bytes);
            count_launch();
            return check_cuda(cudaGetLastError(), "preprocess_resize_fast_kernel launch");
        }
    }
    if (g_resize_tiled && tile_smem <= 200 * 1024 && aligned) {
        static std::atomic<size_t> configured{0};
        if (tile_smem > configured) {
            CBAS_CHECK(cudaFuncSetAttribute(preprocess_resize_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)tile_smem));
            configured = tile_smem;
        }
        preprocess_resize_tile_kernel<<<n * (side / 16), 256, tile_smem, s>>>(frames, A, H, W, fs, rs, side, tp, mean,
                                                                             istd, max_rows);
        count_launch();
        return check_cuda(cudaGetLastError(), "preprocess_resize_tile_kernel launch");
    }
    preprocess_resize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(frames, A, n, H, W, fs, rs, side, tp, mean,
                                                                            istd);
    count_launch();
    return check_cuda(cudaGetLastError(), "preprocess_resize_kernel launch");
}

}  // namespace

struct cbas_encoder {
    cbas_encoder_cfg cfg;
    cbas_encoder_weights w;
    std::vector<cbas_layer_weights> layers;
    int T = 0, Np = 0, Kp = 0, P = 16, ns = 0;  // tokens, patches, patch-matrix pitch, patch size, patches per side
    // workspace (device)
    __nv_bfloat16* a_patch = nullptr;  // [max*Np, Kp]
    float* h = nullptr;                // [max*T, D]   residual stream
    __nv_bfloat16* xn = nullptr;       // [max*T, D]   LN output / attention output
    __nv_bfloat16* qkv = nullptr;      // [max*T, 3D]
    __nv_bfloat16* u = nullptr;        // [max*T, I]
    __nv_bfloat16* cls_q = nullptr;    // [max, D]  last block, CLS rows only: query
    __nv_bfloat16* cls_att = nullptr;  // [max, D]  attention output
    __nv_bfloat16* cls_xn = nullptr;   // [max, D]  LN2 output
};

namespace {

int encoder_embed(cbas_encoder* e, const uint8_t* frames_u8, const float* planes, int n, long long fs, int rs,
                  cudaStream_t s) {
    const cbas_encoder_cfg& c = e->cfg;
    if (e->P != 16) {
        // generic patch size (DINOv2-with-registers, 14 px): per-element kernels, the stage is < 3 % of the step
        ProfScope prof(PROF_PREPROCESS, s);
        if (planes || c.mode == CBAS_PRE_REFERENCE) {
            if (planes && c.mode != CBAS_PRE_REFERENCE)
                return fail("float-plane input is only defined for REFERENCE preprocessing");
            const long long total = (long long)n * e->ns * e->P * e->ns;
            const unsigned grid = (unsigned)((total + 255) / 256);
            if (planes)
                preprocess_green_generic_kernel<true><<<grid, 256, 0, s>>>(planes, e->a_patch, n, c.in_h, c.in_w, 0, 0,
                                                                          e->P, e->ns, e->Kp);
            else
                preprocess_green_generic_kernel<false><<<grid, 256, 0, s>>>(frames_u8, e->a_patch, n, c.in_h, c.in_w, fs,
                                                                           rs, e->P, e->ns, e->Kp);
        } else {
            ResizeTaps tp{(const int*)e->w.rs_ymin, (const float*)e->w.rs_wy, (const int*)e->w.rs_xmin,
                          (const float*)e->w.rs_wx, c.resize_taps_y, c.resize_taps_x};
            const long long used = (long long)e->ns * e->P;
            const long long total = (long long)n * used * used;
            preprocess_resize_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(
                frames_u8, e->a_patch, n, c.in_h, c.in_w, fs, rs, c.side, tp, make_float3(0.485f, 0.456f, 0.406f),
                make_float3(1.0f / 0.229f, 1.0f / 0.224f, 1.0f / 0.225f), e->P, e->ns, e->Kp);
        }
        count_launch();
        if (int rc = check_cuda(cudaGetLastError(), "generic preprocess kernel launch")) return rc;
    } else if (planes) {
        if (c.mode != CBAS_PRE_REFERENCE) return fail("float-plane input is only defined for REFERENCE preprocessing");
        const long long total = (long long)n * c.in_h * (c.in_w / 16);
        preprocess_plane_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(planes, e->a_patch, n, c.in_h, c.in_w);
        count_launch();
        if (int rc = check_cuda(cudaGetLastError(), "preprocess_plane_kernel launch")) return rc;
    } else if (c.mode == CBAS_PRE_REFERENCE) {
        if (int rc = launch_preprocess_green(frames_u8, e->a_patch, n, c.in_h, c.in_w, fs, rs, s)) return rc;
    } else {
        ResizeTaps tp{(const int*)e->w.rs_ymin, (const float*)e->w.rs_wy, (const int*)e->w.rs_xmin,
                      (const float*)e->w.rs_wx, c.resize_taps_y, c.resize_taps_x};
        if (int rc = launch_preprocess_resize(frames_u8, e->a_patch, n, c.in_h, c.in_w, fs, rs, c.side, tp, s)) return rc;
    }
    const int D = c.hidden;
    {
        const long long total = (long long)n * c.prefix_tokens * (D / 4);
        fill_prefix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(e->h, (const float*)e->w.prefix, n, e->T,
                                                                          c.prefix_tokens, D);
        count_launch();
        if (int rc = check_cuda(cudaGetLastError(), "fill_prefix_kernel launch")) return rc;
    }
    GemmParams p{};
    p.M = n * e->Np; p.N = D; p.K = e->Kp;
    p.bias = (const float*)e->w.b_patch;
    p.out = e->h; p.ldo = D;
    p.rows_in = e->Np; p.rows_out = e->T; p.prefix = c.prefix_tokens;
    if (int rc = launch_gemm(e->a_patch, e->Kp, (const __nv_bfloat16*)e->w.w_patch, e->Kp, p, EPI_PATCH_F32, s,
                             PROF_PATCH_GEMM)) return rc;
    if (e->w.pos_embed) {
        // learned absolute position embedding (Dinov2WithRegistersEmbeddings.forward: added before the registers are
        // spliced in, so it lands on the patch rows; the CLS part is already inside prefix row 0)
        ProfScope prof(PROF_PATCH_GEMM, s);
        const long long total = (long long)n * e->Np * (D / 4);
        add_pos_embed_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(e->h, (const float*)e->w.pos_embed, n, e->T,
                                                                            c.prefix_tokens, e->Np, D);
        count_launch();
        return check_cuda(cudaGetLastError(), "add_pos_embed_kernel launch");
    }
    return 0;
}

int encoder_layer(cbas_encoder* e, int li, int n, cudaStream_t s) {
    const cbas_encoder_cfg& c = e->cfg;
    const cbas_layer_weights& L = e->layers[li];
    const int D = c.hidden, I = c.intermediate, M = n * e->T;
    if (int rc = launch_layernorm<__nv_bfloat16>(e->h, 1, (const float*)L.ln1_g, (const float*)L.ln1_b, e->xn, M, D,
                                                 c.ln_eps, s)) return rc;
    GemmParams p{};
    p.M = M; p.N = 3 * D; p.K = D; p.bias = (const float*)L.b_qkv; p.out = e->qkv; p.ldo = 3 * D;
    if (use_attention_tc(e->T, c.prefix_tokens, e->w.rope_cos != nullptr)) {
        // QKV projection (V stored as f16); the tcgen05 attention kernel rotates q and k in its prologue
        p.f16_from = 2 * D;
        if (int rc = launch_gemm(e->xn, D, (const __nv_bfloat16*)L.w_qkv, D, p, EPI_BIAS_BF16_VF16, s, PROF_QKV_GEMM))
            return rc;
        if (int rc = launch_attention_tc(e->qkv, e->xn, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n,
                                         e->T, c.prefix_tokens, c.heads, s)) return rc;
    } else {
        if (int rc = launch_gemm(e->xn, D, (const __nv_bfloat16*)L.w_qkv, D, p, EPI_BIAS_BF16, s, PROF_QKV_GEMM))
            return rc;
        if (int rc = launch_attention(e->qkv, e->xn, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n,
                                      e->T, c.prefix_tokens, c.heads, s)) return rc;
    }
    p = GemmParams{};
    p.M = M; p.N = D; p.K = D; p.bias = (const float*)L.b_o; p.out = e->h; p.ldo = D;
    if (int rc = launch_gemm(e->xn, D, (const __nv_bfloat16*)L.w_o, D, p, EPI_RESID_F32, s, PROF_PROJ_GEMM)) return rc;
    if (int rc = launch_layernorm<__nv_bfloat16>(e->h, 1, (const float*)L.ln2_g, (const float*)L.ln2_b, e->xn, M, D,
                                                 c.ln_eps, s)) return rc;
    p = GemmParams{};
    p.M = M; p.N = I; p.K = D; p.bias = (const float*)L.b_up; p.out = e->u; p.ldo = I;
    if (int rc = launch_gemm(e->xn, D, (const __nv_bfloat16*)L.w_up, D, p, EPI_BIAS_GELU_BF16, s, PROF_UP_GEMM)) return rc;
    p = GemmParams{};
    p.M = M; p.N = D; p.K = I; p.bias = (const float*)L.b_down; p.out = e->h; p.ldo = D;
    return launch_gemm(e->u, I, (const __nv_bfloat16*)L.w_down, I, p, EPI_RESID_F32, s, PROF_DOWN_GEMM);
}

// Last block, production path: the final hidden state is only read at the CLS row (HF :547-548, cbas.py:677), so
// K and V are projected for every token but the query, attention output, proj, LN2 and the MLP run on the n CLS
// rows alone.  Mathematically identical to the full block for the row that is kept.
int encoder_last_layer_cls_only(cbas_encoder* e, int li, int n, cudaStream_t s) {
    const cbas_encoder_cfg& c = e->cfg;
    const cbas_layer_weights& L = e->layers[li];
    const int D = c.hidden, I = c.intermediate, T = e->T, M = n * T;
    const bool tc = use_attention_tc(T, c.prefix_tokens, e->w.rope_cos != nullptr);
    if (int rc = launch_layernorm<__nv_bfloat16>(e->h, 1, (const float*)L.ln1_g, (const float*)L.ln1_b, e->xn, M, D,
                                                 c.ln_eps, s)) return rc;
    const __nv_bfloat16* wqkv = (const __nv_bfloat16*)L.w_qkv;
    const float* bqkv = (const float*)L.b_qkv;
    GemmParams p{};
    // K | V for all tokens -> columns [D, 3D) of the qkv buffer (V in the format the attention kernels expect)
    p.M = M; p.N = 2 * D; p.K = D; p.bias = bqkv + D; p.out = e->qkv + D; p.ldo = 3 * D; p.f16_from = D;
    if (int rc = launch_gemm(e->xn, D, wqkv + (size_t)D * D, D, p, tc ? EPI_BIAS_BF16_VF16 : EPI_BIAS_BF16, s,
                             PROF_QKV_GEMM)) return rc;
    // Q for the CLS rows only (A rows are T*D apart)
    p = GemmParams{};
    p.M = n; p.N = D; p.K = D; p.bias = bqkv; p.out = e->cls_q; p.ldo = D;
    if (int rc = launch_gemm(e->xn, T * D, wqkv, D, p, EPI_BIAS_BF16, s, PROF_QKV_GEMM)) return rc;
    {
        ProfScope prof(PROF_ATTENTION, s);
        const int items = n * c.heads;
        cls_attention_kernel<<<(items * 32 + 127) / 128, 128, 0, s>>>(
            e->cls_q, e->qkv, e->cls_att, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n, T,
            e->w.rope_cos ? c.prefix_tokens : T, c.heads, D, 0.125f * 1.4426950408889634f, tc ? 1 : 0);
        count_launch();
        if (int rc = check_cuda(cudaGetLastError(), "cls_attention_kernel launch")) return rc;
    }
    // proj + residual into the CLS rows of h (output rows are T*D apart)
    p = GemmParams{};
    p.M = n; p.N = D; p.K = D; p.bias = (const float*)L.b_o; p.out = e->h; p.ldo = T * D;
    if (int rc = launch_gemm(e->cls_att, D, (const __nv_bfloat16*)L.w_o, D, p, EPI_RESID_F32, s, PROF_PROJ_GEMM)) return rc;
    if (int rc = launch_layernorm<__nv_bfloat16>(e->h, T, (const float*)L.ln2_g, (const float*)L.ln2_b, e->cls_xn, n, D,
                                                 c.ln_eps, s)) return rc;
    p = GemmParams{};
    p.M = n; p.N = I; p.K = D; p.bias = (const float*)L.b_up; p.out = e->u; p.ldo = I;
    if (int rc = launch_gemm(e->cls_xn, D, (const __nv_bfloat16*)L.w_up, D, p, EPI_BIAS_GELU_BF16, s, PROF_UP_GEMM)) return rc;
    p = GemmParams{};
    p.M = n; p.N = D; p.K = I; p.bias = (const float*)L.b_down; p.out = e->h; p.ldo = T * D;
    return launch_gemm(e->u, I, (const __nv_bfloat16*)L.w_down, I, p, EPI_RESID_F32, s, PROF_DOWN_GEMM);
}

int encoder_forward(cbas_encoder* e, const uint8_t* frames_u8, const float* planes, int n, long long fs, int rs,
                    int stop_after_layer, float* hidden_out, float* emb_out, cudaStream_t s) {
    if (!e) return fail("null encoder");
    if (n < 0 || n > e->cfg.max_frames) return fail("n exceeds the encoder's max_frames");
    if (n == 0) return 0;
    if (int rc = encoder_embed(e, frames_u8, planes, n, fs, rs, s)) return rc;
    const int L = stop_after_layer >= 0 ? stop_after_layer : e->cfg.layers;
    // only the pooled CLS embedding is wanted: the last block can skip every row that is thrown away
    const bool prune = g_prune_last_layer && stop_after_layer < 0 && !hidden_out && e->T <= CLS_ATT_MAX_T && L >= 1;
    for (int li = 0; li < L; ++li) {
        if (prune && li == L - 1) {
            if (int rc = encoder_last_layer_cls_only(e, li, n, s)) return rc;
        } else if (int rc = encoder_layer(e, li, n, s)) return rc;
    }
    const int D = e->cfg.hidden;
    if (hidden_out)
        CBAS_CHECK(cudaMemcpyAsync(hidden_out, e->h, (size_t)n * e->T * D * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (emb_out) {
        // final norm on the CLS row of every frame only (rows frame*T): modeling_dinov3_vit.py:547-548, cbas.py:677
        if (int rc = launch_layernorm<float>(e->h, e->T, (const float*)e->w.lnf_g, (const float*)e->w.lnf_b, emb_out,
                                             n, D, e->cfg.ln_eps, s, PROF_FINAL_LN)) return rc;
    }
    return 0;
}

}  // namespace

extern "C" {

int cbas_b200_encoder_create(const cbas_encoder_cfg* cfg, const cbas_encoder_weights* w, cbas_encoder** out) {
    if (!cfg || !w || !out) return fail("null argument");
    if (cfg->hidden % 128 || cfg->hidden != cfg->heads * ATT_HEAD_DIM)
        return fail("hidden must be heads*64 and a multiple of 128");
    if (cfg->hidden != 384 && cfg->hidden != 768 && cfg->hidden != 1024)
        return fail("hidden size must be 384, 768 or 1024 (DINOv3 ViT-S/B/L)");
    if (cfg->intermediate % 128) return fail("intermediate size must be a multiple of 128");
    const int P = cfg->patch ? cfg->patch : 16;
    if (P != 16 && P != 14) return fail("patch size must be 16 (DINOv3) or 14 (DINOv2-with-registers)");
    if (cfg->side <= 0 || (P == 16 && cfg->side % 16)) return fail("ViT input side must be a positive multiple of 16");
    if (cfg->side / P < 1) return fail("ViT input side is smaller than one patch");
    if (cfg->mode == CBAS_PRE_REFERENCE && (cfg->in_h != cfg->side || cfg->in_w != cfg->side))
        return fail("REFERENCE preprocessing keeps the native (square) resolution: in_h == in_w == side");
    if (cfg->mode != CBAS_PRE_REFERENCE && cfg->mode != CBAS_PRE_PROCESSOR) return fail("unknown preprocessing mode");
    if (cfg->max_frames <= 0) return fail("max_frames must be positive");
    auto* e = new cbas_encoder();
    e->cfg = *cfg;
    e->w = *w;
    e->layers.assign(w->layers, w->layers + cfg->layers);
    e->w.layers = e->layers.data();
    const int ns = cfg->side / P;  // floor: a stride-P convolution ignores the remainder of the frame
    e->P = P; e->ns = ns;
    e->Np = ns * ns;
    e->T = e->Np + cfg->prefix_tokens;
    if (3 * ((e->T + 15) & ~15) * 128 > 232448) {
        delete e;
        return fail("too many tokens per frame: the attention kernels keep a frame's K and V in shared memory (T <= 592)");
    }
    e->Kp = ((cfg->mode == CBAS_PRE_REFERENCE ? P * P : 3 * P * P) + 63) & ~63;
    const size_t mt = (size_t)cfg->max_frames * e->T, D = cfg->hidden;
    cudaError_t err = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (err == cudaSuccess) err = cudaMalloc(p, bytes); };
    alloc((void**)&e->a_patch, (size_t)cfg->max_frames * e->Np * e->Kp * 2);
    if (err == cudaSuccess && P != 16)  // the padding columns of the patch matrix stay zero for the handle's lifetime
        err = cudaMemset(e->a_patch, 0, (size_t)cfg->max_frames * e->Np * e->Kp * 2);
    alloc((void**)&e->h, mt * D * 4);
    alloc((void**)&e->xn, mt * D * 2);
    alloc((void**)&e->qkv, mt * 3 * D * 2);
    alloc((void**)&e->u, mt * (size_t)cfg->intermediate * 2);
    alloc((void**)&e->cls_q, (size_t)cfg->max_frames * D * 2);
    alloc((void**)&e->cls_att, (size_t)cfg->max_frames * D * 2);
    alloc((void**)&e->cls_xn, (size_t)cfg->max_frames * D * 2);
    if (err != cudaSuccess) {
        cbas_b200_encoder_destroy(e);
        return check_cuda(err, "encoder workspace cudaMalloc");
    }
    *out = e;
    return 0;
}

void cbas_b200_encoder_destroy(cbas_encoder* e) {
    if (!e) return;
    cudaFree(e->a_patch); cudaFree(e->h); cudaFree(e->xn); cudaFree(e->qkv); cudaFree(e->u);
    cudaFree(e->cls_q); cudaFree(e->cls_att); cudaFree(e->cls_xn);
    delete e;
}

int cbas_b200_encoder_forward_u8(cbas_encoder* enc, const uint8_t* frames_dev, int32_t n, int64_t frame_stride,
                                 int32_t row_stride, float* emb_out_dev, void* stream) {
    return encoder_forward(enc, frames_dev, nullptr, n, frame_stride, row_stride, -1, nullptr, emb_out_dev,
                           (cudaStream_t)stream);
}

int cbas_b200_encoder_forward_f32(cbas_encoder* enc, const float* planes_dev, int32_t n, float* emb_out_dev,
                                  void* stream) {
    return encoder_forward(enc, nullptr, planes_dev, n, 0, 0, -1, nullptr, emb_out_dev, (cudaStream_t)stream);
}

int cbas_b200_encoder_debug_hidden(cbas_encoder* enc, const uint8_t* frames_dev, int32_t n, int64_t frame_stride,
                                   int32_t row_stride, int32_t after_layer, float* hidden_out_dev, void* stream) {
    if (!enc) return fail("null encoder");
    if (after_layer < 0 || after_layer > enc->cfg.layers) return fail("after_layer out of range");
    return encoder_forward(enc, frames_dev, nullptr, n, frame_stride, row_stride, after_layer, hidden_out_dev, nullptr,
                           (cudaStream_t)stream);
}

int cbas_b200_gemm_bf16(const void* a_dev, const void* w_dev, const float* bias_dev, void* out_dev, int32_t M,
                        int32_t N, int32_t K, int32_t epi, void* stream) {
    if (epi == EPI_PATCH_F32) return fail("the patch epilogue is internal to the encoder");
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.bias = bias_dev; p.out = out_dev; p.ldo = N;
    return launch_gemm((const __nv_bfloat16*)a_dev, K, (const __nv_bfloat16*)w_dev, K, p, epi, (cudaStream_t)stream);
}

int cbas_b200_debug_attention_impl(int32_t impl) {
    if (impl < 0 || impl > 2) return fail("attention impl must be 0 (auto), 1 (mma.sync) or 2 (tcgen05)");
    g_attention_impl = impl;
    return 0;
}

int cbas_b200_attention_tc(const void* qkv_bf16_dev, void* out_bf16_dev, const float* rope_cos_dev,
                           const float* rope_sin_dev, int32_t frames, int32_t T, int32_t prefix, int32_t heads,
                           void* stream) {
    return launch_attention_tc((const __nv_bfloat16*)qkv_bf16_dev, (__nv_bfloat16*)out_bf16_dev, rope_cos_dev,
                               rope_sin_dev, frames, T, prefix, heads, (cudaStream_t)stream);
}

int cbas_b200_debug_attention_trace(void* trace_dev) {
    g_attention_trace = (long long*)trace_dev;
    return 0;
}

int cbas_b200_debug_prune_last_layer(int32_t on) {
    g_prune_last_layer = on != 0;
    return 0;
}

int cbas_b200_debug_resize_tiled(int32_t on) {
    g_resize_tiled = on < 0 ? 0 : (on > 2 ? 2 : on);  // 0 per-pixel, 1 general tiled, 2 column-per-thread (default)
    return 0;
}

int cbas_b200_debug_gemm_cta_group(int32_t cg) {
    if (cg < 0 || cg > 2) return fail("cta group must be 0 (auto), 1 or 2");
    set_gemm_cta_group(cg);
    return 0;
}

int cbas_b200_layernorm(const float* in_dev, const float* gamma_dev, const float* beta_dev, void* out_bf16_dev,
                        int32_t rows, int32_t D, float eps, void* stream) {
    return launch_layernorm<__nv_bfloat16>(in_dev, 1, gamma_dev, beta_dev, (__nv_bfloat16*)out_bf16_dev, rows, D, eps,
                                           (cudaStream_t)stream);
}

int cbas_b200_attention_tc_supported(int32_t T, int32_t prefix, int32_t rope) {
    return (attention_tc_fits(T, prefix) || attention_tc_split_fits(T, prefix, rope != 0)) ? 1 : 0;
}

int cbas_b200_attention(const void* qkv_bf16_dev, void* out_bf16_dev, const float* rope_cos_dev,
                        const float* rope_sin_dev, int32_t frames, int32_t T, int32_t prefix, int32_t heads,
                        void* stream) {
    if (T <= 0 || T > 1024) return fail("attention: tokens per frame must be in [1, 1024]");
    if (3 * ((T + 15) & ~15) * 128 > 227 * 1024) return fail("attention: frame does not fit in shared memory");
    return launch_attention((const __nv_bfloat16*)qkv_bf16_dev, (__nv_bfloat16*)out_bf16_dev, rope_cos_dev,
                            rope_sin_dev, frames, T, prefix, heads, (cudaStream_t)stream);
}

int cbas_b200_preprocess_green(const uint8_t* frames_dev, void* a_bf16_dev, int32_t n, int32_t H, int32_t W,
                               int64_t frame_stride, int32_t row_stride, void* stream) {
    return launch_preprocess_green(frames_dev, (__nv_bfloat16*)a_bf16_dev, n, H, W, frame_stride, row_stride,
                                   (cudaStream_t)stream);
}

int cbas_b200_preprocess_resize(const uint8_t* frames_dev, void* a_bf16_dev, int32_t n, int32_t H, int32_t W,
                                int64_t frame_stride, int32_t row_stride, int32_t side, const int32_t* ymin_dev,
                                const float* wy_dev, int32_t taps_y, const int32_t* xmin_dev, const float* wx_dev,
                                int32_t taps_x, void* stream) {
    ResizeTaps tp{ymin_dev, wy_dev, xmin_dev, wx_dev, taps_y, taps_x};
    return launch_preprocess_resize(frames_dev, (__nv_bfloat16*)a_bf16_dev, n, H, W, frame_stride, row_stride, side,
                                    tp, (cudaStream_t)stream);
}

}  // extern "C"
