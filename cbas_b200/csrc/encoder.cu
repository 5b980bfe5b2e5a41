// DINOv3 ViT encoder forward on sm_100a: preprocess -> patch-embed GEMM -> row statistics -> L x (QKV GEMM, attention
// with RoPE prologue, proj GEMM + residual, up GEMM + GELU, down GEMM + residual) -> final LN on the CLS rows.
// norm1 / norm2 of every block are fused into the GEMMs around them (gemm_tcgen05.cuh: the residual GEMMs leave a
// shifted bf16 copy of h and per-row partial sums, the QKV / up GEMMs normalise in their epilogue).
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

#ifndef CBAS_LN_FUSED_DEFAULT
#define CBAS_LN_FUSED_DEFAULT 0
#endif

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

// Direction of the next row-streaming launch of this host thread (encoder_layer_unfused alternates it kernel by
// kernel when the handle's serpentine option is on; 0 everywhere else).
thread_local int t_reverse = 0;

template <typename OutT>
int launch_layernorm(const float* in, long long in_row_stride, const float* g, const float* b, OutT* out, int rows,
                     int D, float eps, cudaStream_t s, int tag = PROF_LAYERNORM, float* stats = nullptr) {
    if (rows <= 0) return 0;
    ProfScope prof(tag, s);
    const int threads = 256, rows_per_block = threads / 32;
    const int grid = (rows + rows_per_block - 1) / rows_per_block;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = pdl_attr(&attr[0]);  // the kernel waits for its predecessor itself (pdl_wait)
    const int rev = t_reverse;
    cudaError_t err;
    switch (D) {
        case 384: err = cudaLaunchKernelEx(&cfg, layernorm_kernel<384, OutT>, in, in_row_stride, g, b, out, rows, eps, rev, stats); break;
        case 768: err = cudaLaunchKernelEx(&cfg, layernorm_kernel<768, OutT>, in, in_row_stride, g, b, out, rows, eps, rev, stats); break;
        case 1024: err = cudaLaunchKernelEx(&cfg, layernorm_kernel<1024, OutT>, in, in_row_stride, g, b, out, rows, eps, rev, stats); break;
        case 128: err = cudaLaunchKernelEx(&cfg, layernorm_kernel<128, OutT>, in, in_row_stride, g, b, out, rows, eps, rev, stats); break;
        case 256: err = cudaLaunchKernelEx(&cfg, layernorm_kernel<256, OutT>, in, in_row_stride, g, b, out, rows, eps, rev, stats); break;
        default: return fail("LayerNorm width " + std::to_string(D) + " not instantiated (384/768/1024)");
    }
    if (err != cudaSuccess) return check_cuda(err, "layernorm_kernel launch");
    count_launch();
    return check_cuda(cudaGetLastError(), "layernorm_kernel launch");
}

// first link of the fused-LayerNorm chain (layernorm.cuh::ln_stats_init_kernel)
int launch_ln_stats_init(float* h, const float* prefix_tokens, int T, int P, __nv_bfloat16* hb, float* stats, int rows,
                         int D, cudaStream_t s) {
    if (rows <= 0) return 0;
    ProfScope prof(PROF_LAYERNORM, s);
    const int threads = 256, rows_per_block = threads / 32;
    const int grid = (rows + rows_per_block - 1) / rows_per_block;
    switch (D) {
        case 384: ln_stats_init_kernel<384><<<grid, threads, 0, s>>>(h, prefix_tokens, T, P, hb, stats, rows); break;
        case 768: ln_stats_init_kernel<768><<<grid, threads, 0, s>>>(h, prefix_tokens, T, P, hb, stats, rows); break;
        case 1024: ln_stats_init_kernel<1024><<<grid, threads, 0, s>>>(h, prefix_tokens, T, P, hb, stats, rows); break;
        case 128: ln_stats_init_kernel<128><<<grid, threads, 0, s>>>(h, prefix_tokens, T, P, hb, stats, rows); break;
        case 256: ln_stats_init_kernel<256><<<grid, threads, 0, s>>>(h, prefix_tokens, T, P, hb, stats, rows); break;
        default: return fail("LayerNorm width " + std::to_string(D) + " not instantiated (128/256/384/768/1024)");
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "ln_stats_init_kernel launch");
}

int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, const float* cs, const float* sn, int frames,
                     int T, int prefix, int heads, cudaStream_t s) {
    if (frames <= 0) return 0;
    ProfScope prof(PROF_ATTENTION, s);
    const int TP = (T + 15) & ~15;
    const int smem = 3 * TP * 128;
    static DeviceSmemOptIn optin;  // per device (common.h)
    CBAS_CHECK(optin.ensure(attention_kernel, smem));
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

// Test / profiling knobs.  The encoder's own (which attention kernel, last-block pruning, which resize kernel) live in
// the handle (cbas_b200_encoder_set_option); the two below belong to the kernel-level entry points and are per host
// thread, so two threads never see each other's setting.
thread_local long long* g_attention_trace = nullptr;  // device buffer [64][ATC_TRACE_SLOTS] (tools/attn_trace.py)
thread_local int g_resize_tiled = 2;  // cbas_b200_preprocess_resize: 0 per-pixel, 1 general tiled, 2 column-per-thread

bool attention_tc_fits(int T, int prefix) {
    const int TK = (T + 15) & ~15;
    return TK <= 256 && atc_smem_bytes(TK, T, prefix, true) <= 232448;
}
// 257..384 tokens: the key-split tcgen05 kernel (attention_tc_split.cuh)
bool attention_tc_split_fits(int T, int prefix, bool rope) {
    const int TK = (T + 15) & ~15;
    return TK > 256 && TK <= 384 && ats_smem_bytes(TK, T, prefix, rope) <= 232448;
}
// impl: 0 auto (tcgen05 whenever a tcgen05 kernel covers T), 1 mma.sync, 2 tcgen05
bool use_attention_tc(int impl, int T, int prefix, bool rope = true) {
    if (impl == 1) return false;
    if (attention_tc_split_fits(T, prefix, rope)) return true;
    return impl >= 2 || attention_tc_fits(T, prefix);
}

template <int TK>
int launch_attention_tc_tk(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& to, const CUtensorMap& to1,
                           const AttnTcParams& p, int smem, int grid, cudaStream_t s) {
    static DeviceSmemOptIn optin;
    CBAS_CHECK(optin.ensure(attention_tc_kernel<TK>, smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(ATC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    cfg.numAttrs = pdl_attr(&attr[0]);  // the kernel waits for the QKV GEMM itself (pdl_wait after its prologue)
    const cudaError_t err = cudaLaunchKernelEx(&cfg, attention_tc_kernel<TK>, tq, tkv, to, to1, p);
    count_launch();
    if (err != cudaSuccess) return check_cuda(err, "attention_tc_kernel launch");
    return check_cuda(cudaGetLastError(), "attention_tc_kernel launch");
}

// does the tcgen05 attention take q and k as IEEE f16 (like v) for this geometry?  The one-tile-per-128-rows kernel
// (T <= 256) does: its RoPE prologue and logits run on f16 operands; the key-split kernel keeps bf16 q and k.
bool attention_qk_f16(int T) { return ((T + 15) & ~15) <= 256; }

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
        static DeviceSmemOptIn optin_split;
        CBAS_CHECK(optin_split.ensure(attention_tc_split_kernel, smem));
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
    AttnTcParams p{qkv, out, frames, heads, T, TK, D, 0.125f * 1.4426950408889634f, rope ? cs : nullptr,
                   rope ? sn : nullptr, prefix, g_attention_trace, t_reverse};
    const int items = frames * heads;
    const int grid = items < sm_count() ? items : sm_count();
    switch (TK) {  // the padded key count is a template parameter: every softmax / MMA loop is fully unrolled
#define ATC_CASE(N) case N: return launch_attention_tc_tk<N>(tq, tkv, to, to1, p, smem, grid, s);
        ATC_CASE(16) ATC_CASE(32) ATC_CASE(48) ATC_CASE(64) ATC_CASE(80) ATC_CASE(96) ATC_CASE(112) ATC_CASE(128)
        ATC_CASE(144) ATC_CASE(160) ATC_CASE(176) ATC_CASE(192) ATC_CASE(208) ATC_CASE(224) ATC_CASE(240) ATC_CASE(256)
#undef ATC_CASE
    }
    return fail("attention: unexpected padded key count");
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
                             int side, const ResizeTaps& tp, cudaStream_t s, int resize_tiled) {
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
    if (resize_tiled >= 2 && aligned && side <= 256 && tp.taps_x <= RESIZE_FAST_TAPS && tp.taps_y <= RESIZE_FAST_TAPS) {
        // production path: column-per-thread kernel (bitwise identical to the general tiled kernel below)
        const int io_bytes = std::max(max_rows * src_pitch, (side / 16) * 1536);
        const size_t smem = (size_t)io_bytes + (size_t)3 * max_rows * side * 4 + 16 * 32 + 32;  // + row taps, + slack for the word-wise reads of the last staged row
        if (smem <= 110 * 1024) {
            static DeviceSmemOptIn optin_fast;
            CBAS_CHECK(optin_fast.ensure(preprocess_resize_fast_kernel, (long long)smem));
            preprocess_resize_fast_kernel<<<n * (side / 16), 256, smem, s>>>(frames, A, H, W, fs, rs, side, tp, mean, istd,
                                                                            max_rows, io_bytes);
            count_launch();
            return check_cuda(cudaGetLastError(), "preprocess_resize_fast_kernel launch");
        }
    }
    if (resize_tiled && tile_smem <= 200 * 1024 && aligned) {
        static DeviceSmemOptIn optin_tile;
        CBAS_CHECK(optin_tile.ensure(preprocess_resize_tile_kernel, (long long)tile_smem));
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
    int device = 0;                    // CUDA device ordinal the handle (workspace, weights) lives on
    // per-handle test knobs (cbas_b200_encoder_set_option)
    bool prune_last_layer = true;      // false: run the last block on every token
    int attention_impl = 0;            // 0 auto, 1 mma.sync, 2 tcgen05
    int resize_tiled = 2;              // 0 per-pixel kernel, 1 general tiled kernel, 2 column-per-thread kernel
    int ln_fused = CBAS_LN_FUSED_DEFAULT;  // 0: standalone LayerNorm kernels, 1: norm1 / norm2 inside the GEMM epilogues,
                                           // 2: norm1 fused (down GEMM -> next block's QKV GEMM), norm2 standalone
    bool serpentine = true;            // standalone-LayerNorm path: consecutive kernels walk the rows in opposite directions
    int direction = 0;                 //   ... direction of the last kernel launched by the previous block
    float* ones = nullptr;             // [D] ones / zeros: unit gamma and zero beta for the standalone LayerNorm path
    float* zeros = nullptr;            //     (gamma and beta themselves are folded into the weights either way)
    // workspace (device)
    __nv_bfloat16* a_patch = nullptr;  // [max*Np, Kp]
    float* h = nullptr;                // [max*T, D]   residual stream
    __nv_bfloat16* hb = nullptr;       // [max*T, D]   bf16(h - row shift): A operand of the QKV / up GEMMs
    float* stats[2] = {nullptr, nullptr};  // [max*T, LN_STAT_FLOATS] row statistics of h, ping-pong (proj: 0 -> 1, down: 1 -> 0)
    __nv_bfloat16* cls_hb = nullptr;   // [max, D]  last block, CLS rows only: shifted copy after the proj update
    float* cls_stats = nullptr;        // [max, LN_STAT_FLOATS]  ... and its statistics
    __nv_bfloat16* xn = nullptr;       // [max*T, D]   attention output
    __nv_bfloat16* qkv = nullptr;      // [max*T, 3D]
    __nv_bfloat16* u = nullptr;        // [max*T, I]
    __nv_bfloat16* cls_q = nullptr;    // [max, D]  last block, CLS rows only: query
    __nv_bfloat16* cls_att = nullptr;  // [max, D]  attention output
};

namespace {

// frames_u8 with pix == 3: interleaved RGB frames; pix == 1: one uint8 plane per frame (the green channel; REFERENCE
// mode only).  planes: float planes in [0,1] (the DinoEncoder.__call__ contract).
int encoder_embed(cbas_encoder* e, const uint8_t* frames_u8, int pix, const float* planes, int n, long long fs, int rs,
                  cudaStream_t s) {
    const cbas_encoder_cfg& c = e->cfg;
    if (pix == 1 && c.mode != CBAS_PRE_REFERENCE)
        return fail("single-plane uint8 input is only defined for REFERENCE preprocessing (green / 255)");
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
                                                                          e->P, e->ns, e->Kp, 1);
            else
                preprocess_green_generic_kernel<false><<<grid, 256, 0, s>>>(frames_u8, e->a_patch, n, c.in_h, c.in_w, fs,
                                                                           rs, e->P, e->ns, e->Kp, pix);
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
    } else if (c.mode == CBAS_PRE_REFERENCE && pix == 1) {
        if (c.in_h % 16 || c.in_w % 16) return fail("frame size must be a multiple of the 16-pixel patch");
        ProfScope prof(PROF_PREPROCESS, s);
        const long long total = (long long)n * c.in_h * (c.in_w / 16);
        preprocess_plane_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(frames_u8, e->a_patch, n, c.in_h,
                                                                                  c.in_w, fs, rs);
        count_launch();
        if (int rc = check_cuda(cudaGetLastError(), "preprocess_plane_u8_kernel launch")) return rc;
    } else if (c.mode == CBAS_PRE_REFERENCE) {
        if (int rc = launch_preprocess_green(frames_u8, e->a_patch, n, c.in_h, c.in_w, fs, rs, s)) return rc;
    } else {
        ResizeTaps tp{(const int*)e->w.rs_ymin, (const float*)e->w.rs_wy, (const int*)e->w.rs_xmin,
                      (const float*)e->w.rs_wx, c.resize_taps_y, c.resize_taps_x};
        if (int rc = launch_preprocess_resize(frames_u8, e->a_patch, n, c.in_h, c.in_w, fs, rs, c.side, tp, s,
                                              e->resize_tiled)) return rc;
    }
    const int D = c.hidden;
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
        if (int rc = check_cuda(cudaGetLastError(), "add_pos_embed_kernel launch")) return rc;
    }
    if (!e->ln_fused) {
        // standalone LayerNorm kernels read h themselves: only the CLS / register rows are left to write
        // (HF modeling_dinov3_vit.py:85-90: cat(cls, registers, patches))
        ProfScope prof(PROF_PATCH_GEMM, s);
        const long long total = (long long)n * c.prefix_tokens * (D / 4);
        if (total > 0) {
            fill_prefix_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(e->h, (const float*)e->w.prefix, n, e->T,
                                                                              c.prefix_tokens, D);
            count_launch();
            if (int rc = check_cuda(cudaGetLastError(), "fill_prefix_kernel launch")) return rc;
        }
        return 0;
    }
    // CLS / register rows into h, then the row statistics and the shifted bf16 copy that block 0's QKV GEMM reads
    return launch_ln_stats_init(e->h, (const float*)e->w.prefix, e->T, c.prefix_tokens, e->hb, e->stats[0], n * e->T, D, s);
}

// GemmParams of a LayerNorm-consumer GEMM (QKV, up): A is the shifted copy, bias = c2, c1 = column sums of W * gamma
GemmParams ln_consumer(const cbas_encoder* e, int M, int N, const void* c1, const void* c2, void* out, int ldo,
                       const float* stats, int stat_stride) {
    GemmParams p{};
    p.M = M; p.N = N; p.K = e->cfg.hidden; p.bias = (const float*)c2; p.out = out; p.ldo = ldo;
    p.ln_in = stats; p.ln_in_stride = stat_stride; p.ln_c1 = (const float*)c1;
    p.ln_inv_dim = 1.0f / (float)e->cfg.hidden; p.ln_eps = e->cfg.ln_eps;
    return p;
}
// ... and of a LayerNorm-producer GEMM (proj, down): h += A W^T + b, new shifted copy, new statistics
GemmParams ln_producer(const cbas_encoder* e, int M, int K, const void* bias, int ldh, const float* st_in, int in_stride,
                       float* st_out, int out_stride, __nv_bfloat16* hb, int ldhb) {
    GemmParams p{};
    p.M = M; p.N = e->cfg.hidden; p.K = K; p.bias = (const float*)bias; p.out = e->h; p.ldo = ldh;
    p.ln_in = st_in; p.ln_in_stride = in_stride; p.ln_out = st_out; p.ln_out_stride = out_stride;
    p.ln_inv_dim = 1.0f / (float)e->cfg.hidden; p.ln_eps = e->cfg.ln_eps;
    p.hb = hb; p.ldhb = ldhb;
    return p;
}

// The same block with norm1 / norm2 as standalone kernels (CBAS_OPT_LN_FUSION 0).  gamma and beta are folded into the
// QKV / up weights for both paths (W' = W * gamma, c2 = W beta + b), so this path normalises with unit gamma / zero
// beta and runs the plain GEMMs on W' and c2: the same function, the LayerNorm arithmetic in its own HBM pass.
int encoder_layer_unfused(cbas_encoder* e, int li, int n, cudaStream_t s, bool cls_only) {
    const cbas_encoder_cfg& c = e->cfg;
    const cbas_layer_weights& L = e->layers[li];
    const int D = c.hidden, I = c.intermediate, T = e->T, M = n * T;
    const bool tc = use_attention_tc(e->attention_impl, T, c.prefix_tokens, e->w.rope_cos != nullptr);
    // Serpentine order: every kernel of the block walks the rows (frames) in the direction opposite to its
    // predecessor's, so it starts on the ~100 MB its predecessor touched last and finds them in the 126 MB L2
    // instead of HBM.  The direction carries over from block to block (seven kernels per block: it alternates).
    struct Direction {
        bool on;
        ~Direction() { t_reverse = 0; }
        void flip() const { if (on) t_reverse ^= 1; }
    } dir{e->serpentine && !cls_only};
    t_reverse = dir.on ? e->direction : 0;
    dir.flip();
    if (int rc = launch_layernorm<__nv_bfloat16>(e->h, 1, e->ones, e->zeros, e->hb, M, D, c.ln_eps, s)) return rc;
    const __nv_bfloat16* wqkv = (const __nv_bfloat16*)L.w_qkv;
    const float* bqkv = (const float*)L.b_qkv;
    GemmParams p{};
    if (!cls_only) {
        // V is stored as f16 for the tcgen05 kernels; the T <= 256 kernel takes q and k as f16 as well
        p.M = M; p.N = 3 * D; p.K = D; p.bias = bqkv; p.out = e->qkv; p.ldo = 3 * D;
        p.f16_from = attention_qk_f16(T) ? 0 : 2 * D;
        dir.flip(); p.reverse = t_reverse;
        if (int rc = launch_gemm(e->hb, D, wqkv, D, p, tc ? EPI_BIAS_BF16_VF16 : EPI_BIAS_BF16, s, PROF_QKV_GEMM)) return rc;
        dir.flip();
        if (tc) {
            if (int rc = launch_attention_tc(e->qkv, e->xn, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n, T,
                                             c.prefix_tokens, c.heads, s)) return rc;
        } else if (int rc = launch_attention(e->qkv, e->xn, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n, T,
                                             c.prefix_tokens, c.heads, s)) return rc;
        p = GemmParams{};
        p.M = M; p.N = D; p.K = D; p.bias = (const float*)L.b_o; p.out = e->h; p.ldo = D;
        dir.flip(); p.reverse = t_reverse;
        if (int rc = launch_gemm(e->xn, D, (const __nv_bfloat16*)L.w_o, D, p, EPI_RESID_F32, s, PROF_PROJ_GEMM)) return rc;
        dir.flip();
        if (int rc = launch_layernorm<__nv_bfloat16>(e->h, 1, e->ones, e->zeros, e->hb, M, D, c.ln_eps, s)) return rc;
        p = GemmParams{};
        p.M = M; p.N = I; p.K = D; p.bias = (const float*)L.b_up; p.out = e->u; p.ldo = I;
        dir.flip(); p.reverse = t_reverse;
        if (int rc = launch_gemm(e->hb, D, (const __nv_bfloat16*)L.w_up, D, p, EPI_BIAS_GELU_BF16, s, PROF_UP_GEMM)) return rc;
        p = GemmParams{};
        p.M = M; p.N = D; p.K = I; p.bias = (const float*)L.b_down; p.out = e->h; p.ldo = D;
        dir.flip(); p.reverse = t_reverse;
        e->direction = t_reverse;
        return launch_gemm(e->u, I, (const __nv_bfloat16*)L.w_down, I, p, EPI_RESID_F32, s, PROF_DOWN_GEMM);
    }
    // last block, CLS rows only (see encoder_last_layer_cls_only)
    p.M = M; p.N = 2 * D; p.K = D; p.bias = bqkv + D; p.out = e->qkv + D; p.ldo = 3 * D; p.f16_from = D;
    if (int rc = launch_gemm(e->hb, D, wqkv + (size_t)D * D, D, p, tc ? EPI_BIAS_BF16_VF16 : EPI_BIAS_BF16, s,
                             PROF_QKV_GEMM)) return rc;
    p = GemmParams{};
    p.M = n; p.N = D; p.K = D; p.bias = bqkv; p.out = e->cls_q; p.ldo = D;
    if (int rc = launch_gemm(e->hb, T * D, wqkv, D, p, EPI_BIAS_BF16, s, PROF_QKV_GEMM)) return rc;
    {
        ProfScope prof(PROF_ATTENTION, s);
        const int items = n * c.heads;
        cls_attention_kernel<<<(items * 32 + 127) / 128, 128, 0, s>>>(
            e->cls_q, e->qkv, e->cls_att, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n, T,
            e->w.rope_cos ? c.prefix_tokens : T, c.heads, D, 0.125f * 1.4426950408889634f, tc ? 1 : 0);
        count_launch();
        if (int rc = check_cuda(cudaGetLastError(), "cls_attention_kernel launch")) return rc;
    }
    p = GemmParams{};
    p.M = n; p.N = D; p.K = D; p.bias = (const float*)L.b_o; p.out = e->h; p.ldo = T * D;
    if (int rc = launch_gemm(e->cls_att, D, (const __nv_bfloat16*)L.w_o, D, p, EPI_RESID_F32, s, PROF_PROJ_GEMM)) return rc;
    if (int rc = launch_layernorm<__nv_bfloat16>(e->h, T, e->ones, e->zeros, e->cls_hb, n, D, c.ln_eps, s)) return rc;
    p = GemmParams{};
    p.M = n; p.N = I; p.K = D; p.bias = (const float*)L.b_up; p.out = e->u; p.ldo = I;
    if (int rc = launch_gemm(e->cls_hb, D, (const __nv_bfloat16*)L.w_up, D, p, EPI_BIAS_GELU_BF16, s, PROF_UP_GEMM)) return rc;
    p = GemmParams{};
    p.M = n; p.N = D; p.K = I; p.bias = (const float*)L.b_down; p.out = e->h; p.ldo = T * D;
    return launch_gemm(e->u, I, (const __nv_bfloat16*)L.w_down, I, p, EPI_RESID_F32, s, PROF_DOWN_GEMM);
}

// Hybrid (CBAS_OPT_LN_FUSION 2): only norm1 is fused.  The down GEMM is long enough (K = 4 D) to hide the producer
// epilogue and the QKV GEMM's consumer epilogue costs ~7 %, together less than half of the LayerNorm pass they replace;
// the short-K proj GEMM (HBM-bound, +57 % as a producer) and the GELU epilogue of the up GEMM (already the busiest part
// of that kernel) are left alone: proj reduces into h as in the standalone path and norm2 stays a kernel, which also
// writes the statistics row (shift = the row's exact mean) the down producer starts from.
int encoder_layer_hybrid(cbas_encoder* e, int li, int n, cudaStream_t s) {
    const cbas_encoder_cfg& c = e->cfg;
    const cbas_layer_weights& L = e->layers[li];
    const int D = c.hidden, I = c.intermediate, T = e->T, M = n * T;
    const bool tc = use_attention_tc(e->attention_impl, T, c.prefix_tokens, e->w.rope_cos != nullptr);
    // norm1 + QKV projection: A = shifted copy left by the previous block's down GEMM (or by the embedding stage)
    GemmParams p = ln_consumer(e, M, 3 * D, L.c1_qkv, L.b_qkv, e->qkv, 3 * D, e->stats[0], 1);
    if (tc) p.f16_from = attention_qk_f16(T) ? 0 : 2 * D;
    if (int rc = launch_gemm(e->hb, D, (const __nv_bfloat16*)L.w_qkv, D, p, tc ? EPI_BIAS_BF16_VF16 : EPI_BIAS_BF16, s,
                             PROF_QKV_GEMM)) return rc;
    if (tc) {
        if (int rc = launch_attention_tc(e->qkv, e->xn, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n, T,
                                         c.prefix_tokens, c.heads, s)) return rc;
    } else if (int rc = launch_attention(e->qkv, e->xn, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n, T,
                                         c.prefix_tokens, c.heads, s)) return rc;
    // output projection: plain TMA add-reduction into the fp32 residual stream
    p = GemmParams{};
    p.M = M; p.N = D; p.K = D; p.bias = (const float*)L.b_o; p.out = e->h; p.ldo = D;
    if (int rc = launch_gemm(e->xn, D, (const __nv_bfloat16*)L.w_o, D, p, EPI_RESID_F32, s, PROF_PROJ_GEMM)) return rc;
    // norm2 as a kernel (unit gamma / zero beta: both are folded into W_up), leaving the statistics row of h
    if (int rc = launch_layernorm<__nv_bfloat16>(e->h, 1, e->ones, e->zeros, e->hb, M, D, c.ln_eps, s, PROF_LAYERNORM,
                                                 e->stats[1])) return rc;
    p = GemmParams{};
    p.M = M; p.N = I; p.K = D; p.bias = (const float*)L.b_up; p.out = e->u; p.ldo = I;
    if (int rc = launch_gemm(e->hb, D, (const __nv_bfloat16*)L.w_up, D, p, EPI_BIAS_GELU_BF16, s, PROF_UP_GEMM)) return rc;
    // down projection + residual; leaves hb / statistics for the next block's norm1
    p = ln_producer(e, M, I, L.b_down, D, e->stats[1], 1, e->stats[0], 1, e->hb, D);
    return launch_gemm(e->u, I, (const __nv_bfloat16*)L.w_down, I, p, EPI_RESID_LN_F32, s, PROF_DOWN_GEMM);
}

int encoder_layer(cbas_encoder* e, int li, int n, cudaStream_t s) {
    if (!e->ln_fused) return encoder_layer_unfused(e, li, n, s, false);
    if (e->ln_fused == 2) return encoder_layer_hybrid(e, li, n, s);
    const cbas_encoder_cfg& c = e->cfg;
    const cbas_layer_weights& L = e->layers[li];
    const int D = c.hidden, I = c.intermediate, M = n * e->T;
    // norm1 + QKV projection (HF :433-434, :305-311)
    GemmParams p = ln_consumer(e, M, 3 * D, L.c1_qkv, L.b_qkv, e->qkv, 3 * D, e->stats[0], 1);
    if (use_attention_tc(e->attention_impl, e->T, c.prefix_tokens, e->w.rope_cos != nullptr)) {
        // V stored as f16 (q and k too for frames of at most 256 tokens); the attention kernel rotates q and k in
        // its prologue
        p.f16_from = attention_qk_f16(e->T) ? 0 : 2 * D;
        if (int rc = launch_gemm(e->hb, D, (const __nv_bfloat16*)L.w_qkv, D, p, EPI_BIAS_BF16_VF16, s, PROF_QKV_GEMM))
            return rc;
        if (int rc = launch_attention_tc(e->qkv, e->xn, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n,
                                         e->T, c.prefix_tokens, c.heads, s)) return rc;
    } else {
        if (int rc = launch_gemm(e->hb, D, (const __nv_bfloat16*)L.w_qkv, D, p, EPI_BIAS_BF16, s, PROF_QKV_GEMM))
            return rc;
        if (int rc = launch_attention(e->qkv, e->xn, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n,
                                      e->T, c.prefix_tokens, c.heads, s)) return rc;
    }
    // output projection + LayerScale + residual (HF :440-441); leaves hb / statistics for norm2
    p = ln_producer(e, M, D, L.b_o, D, e->stats[0], 1, e->stats[1], 1, e->hb, D);
    if (int rc = launch_gemm(e->xn, D, (const __nv_bfloat16*)L.w_o, D, p, EPI_RESID_LN_F32, s, PROF_PROJ_GEMM)) return rc;
    // norm2 + up projection + GELU (HF :445-446, :385-386)
    p = ln_consumer(e, M, I, L.c1_up, L.b_up, e->u, I, e->stats[1], 1);
    if (int rc = launch_gemm(e->hb, D, (const __nv_bfloat16*)L.w_up, D, p, EPI_BIAS_GELU_BF16, s, PROF_UP_GEMM)) return rc;
    // down projection + LayerScale + residual (HF :447-448); leaves hb / statistics for the next block's norm1
    p = ln_producer(e, M, I, L.b_down, D, e->stats[1], 1, e->stats[0], 1, e->hb, D);
    return launch_gemm(e->u, I, (const __nv_bfloat16*)L.w_down, I, p, EPI_RESID_LN_F32, s, PROF_DOWN_GEMM);
}

// Last block, production path: the final hidden state is only read at the CLS row (HF :547-548, cbas.py:677), so
// K and V are projected for every token but the query, attention output, proj, norm2 and the MLP run on the n CLS
// rows alone.  Mathematically identical to the full block for the row that is kept.
int encoder_last_layer_cls_only(cbas_encoder* e, int li, int n, cudaStream_t s) {
    if (!e->ln_fused) return encoder_layer_unfused(e, li, n, s, true);
    const cbas_encoder_cfg& c = e->cfg;
    const cbas_layer_weights& L = e->layers[li];
    const int D = c.hidden, I = c.intermediate, T = e->T, M = n * T;
    const bool tc = use_attention_tc(e->attention_impl, T, c.prefix_tokens, e->w.rope_cos != nullptr);
    const __nv_bfloat16* wqkv = (const __nv_bfloat16*)L.w_qkv;
    const float* c1 = (const float*)L.c1_qkv;
    const float* c2 = (const float*)L.b_qkv;
    // K | V for all tokens -> columns [D, 3D) of the qkv buffer (V in the format the attention kernels expect)
    GemmParams p = ln_consumer(e, M, 2 * D, c1 + D, c2 + D, e->qkv + D, 3 * D, e->stats[0], 1);
    p.f16_from = D;
    if (int rc = launch_gemm(e->hb, D, wqkv + (size_t)D * D, D, p, tc ? EPI_BIAS_BF16_VF16 : EPI_BIAS_BF16, s,
                             PROF_QKV_GEMM)) return rc;
    // Q for the CLS rows only (A rows and statistics rows are T apart)
    p = ln_consumer(e, n, D, c1, c2, e->cls_q, D, e->stats[0], T);
    if (int rc = launch_gemm(e->hb, T * D, wqkv, D, p, EPI_BIAS_BF16, s, PROF_QKV_GEMM)) return rc;
    {
        ProfScope prof(PROF_ATTENTION, s);
        const int items = n * c.heads;
        cls_attention_kernel<<<(items * 32 + 127) / 128, 128, 0, s>>>(
            e->cls_q, e->qkv, e->cls_att, (const float*)e->w.rope_cos, (const float*)e->w.rope_sin, n, T,
            e->w.rope_cos ? c.prefix_tokens : T, c.heads, D, 0.125f * 1.4426950408889634f, tc ? 1 : 0);
        count_launch();
        if (int rc = check_cuda(cudaGetLastError(), "cls_attention_kernel launch")) return rc;
    }
    // proj + residual into the CLS rows of h (rows T*D apart); compact shifted copy and statistics for norm2
    p = ln_producer(e, n, D, L.b_o, T * D, e->stats[0], T, e->cls_stats, 1, e->cls_hb, D);
    if (int rc = launch_gemm(e->cls_att, D, (const __nv_bfloat16*)L.w_o, D, p, EPI_RESID_LN_F32, s, PROF_PROJ_GEMM)) return rc;
    p = ln_consumer(e, n, I, L.c1_up, L.b_up, e->u, I, e->cls_stats, 1);
    if (int rc = launch_gemm(e->cls_hb, D, (const __nv_bfloat16*)L.w_up, D, p, EPI_BIAS_GELU_BF16, s, PROF_UP_GEMM)) return rc;
    // nothing normalises this output inside the block (the final norm reads h itself): plain residual update
    p = GemmParams{};
    p.M = n; p.N = D; p.K = I; p.bias = (const float*)L.b_down; p.out = e->h; p.ldo = T * D;
    return launch_gemm(e->u, I, (const __nv_bfloat16*)L.w_down, I, p, EPI_RESID_F32, s, PROF_DOWN_GEMM);
}

int encoder_forward(cbas_encoder* e, const uint8_t* frames_u8, int pix, const float* planes, int n, long long fs, int rs,
                    int stop_after_layer, float* hidden_out, float* emb_out, cudaStream_t s) {
    if (!e) return fail("null encoder");
    DeviceGuard guard(e->device);  // the handle's device, whatever the calling thread's current device is
    if (!guard.ok()) return fail("cudaSetDevice to the encoder's device failed");
    if (n < 0 || n > e->cfg.max_frames) return fail("n exceeds the encoder's max_frames");
    if (n == 0) return 0;
    if (int rc = encoder_embed(e, frames_u8, pix, planes, n, fs, rs, s)) return rc;
    const int L = stop_after_layer >= 0 ? stop_after_layer : e->cfg.layers;
    // only the pooled CLS embedding is wanted: the last block can skip every row that is thrown away
    const bool prune = e->prune_last_layer && stop_after_layer < 0 && !hidden_out && e->T <= CLS_ATT_MAX_T && L >= 1;
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
    CBAS_CHECK(cudaGetDevice(&e->device));  // the handle belongs to the device that is current at create time
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
    alloc((void**)&e->hb, mt * D * 2);
    for (int i = 0; i < 2; ++i) {
        // slots no GEMM of this geometry writes must read as zero for the handle's lifetime
        alloc((void**)&e->stats[i], mt * LN_STAT_FLOATS * 4);
        if (err == cudaSuccess) err = cudaMemset(e->stats[i], 0, mt * LN_STAT_FLOATS * 4);
    }
    alloc((void**)&e->ones, D * 4);
    alloc((void**)&e->zeros, D * 4);
    if (err == cudaSuccess) {
        std::vector<float> one(D, 1.0f);
        err = cudaMemcpy(e->ones, one.data(), D * 4, cudaMemcpyHostToDevice);
        if (err == cudaSuccess) err = cudaMemset(e->zeros, 0, D * 4);
    }
    alloc((void**)&e->cls_hb, (size_t)cfg->max_frames * D * 2);
    alloc((void**)&e->cls_stats, (size_t)cfg->max_frames * LN_STAT_FLOATS * 4);
    if (err == cudaSuccess) err = cudaMemset(e->cls_stats, 0, (size_t)cfg->max_frames * LN_STAT_FLOATS * 4);
    alloc((void**)&e->xn, mt * D * 2);
    alloc((void**)&e->qkv, mt * 3 * D * 2);
    alloc((void**)&e->u, mt * (size_t)cfg->intermediate * 2);
    alloc((void**)&e->cls_q, (size_t)cfg->max_frames * D * 2);
    alloc((void**)&e->cls_att, (size_t)cfg->max_frames * D * 2);
    if (err != cudaSuccess) {
        cbas_b200_encoder_destroy(e);
        return check_cuda(err, "encoder workspace cudaMalloc");
    }
    *out = e;
    return 0;
}

void cbas_b200_encoder_destroy(cbas_encoder* e) {
    if (!e) return;
    DeviceGuard guard(e->device);
    cudaFree(e->a_patch); cudaFree(e->h); cudaFree(e->xn); cudaFree(e->qkv); cudaFree(e->u);
    cudaFree(e->hb); cudaFree(e->stats[0]); cudaFree(e->stats[1]); cudaFree(e->cls_hb); cudaFree(e->cls_stats);
    cudaFree(e->cls_q); cudaFree(e->cls_att); cudaFree(e->ones); cudaFree(e->zeros);
    delete e;
}

int cbas_b200_encoder_forward_u8(cbas_encoder* enc, const uint8_t* frames_dev, int32_t n, int64_t frame_stride,
                                 int32_t row_stride, float* emb_out_dev, void* stream) {
    return encoder_forward(enc, frames_dev, 3, nullptr, n, frame_stride, row_stride, -1, nullptr, emb_out_dev,
                           (cudaStream_t)stream);
}

int cbas_b200_encoder_forward_u8_plane(cbas_encoder* enc, const uint8_t* planes_dev, int32_t n, int64_t frame_stride,
                                       int32_t row_stride, float* emb_out_dev, void* stream) {
    return encoder_forward(enc, planes_dev, 1, nullptr, n, frame_stride, row_stride, -1, nullptr, emb_out_dev,
                           (cudaStream_t)stream);
}

int cbas_b200_encoder_forward_f32(cbas_encoder* enc, const float* planes_dev, int32_t n, float* emb_out_dev,
                                  void* stream) {
    return encoder_forward(enc, nullptr, 3, planes_dev, n, 0, 0, -1, nullptr, emb_out_dev, (cudaStream_t)stream);
}

int cbas_b200_encoder_debug_hidden(cbas_encoder* enc, const uint8_t* frames_dev, int32_t n, int64_t frame_stride,
                                   int32_t row_stride, int32_t after_layer, float* hidden_out_dev, void* stream) {
    if (!enc) return fail("null encoder");
    if (after_layer < 0 || after_layer > enc->cfg.layers) return fail("after_layer out of range");
    return encoder_forward(enc, frames_dev, 3, nullptr, n, frame_stride, row_stride, after_layer, hidden_out_dev, nullptr,
                           (cudaStream_t)stream);
}

int cbas_b200_gemm_bf16(const void* a_dev, const void* w_dev, const float* bias_dev, void* out_dev, int32_t M,
                        int32_t N, int32_t K, int32_t epi, void* stream) {
    if (epi == EPI_PATCH_F32) return fail("the patch epilogue is internal to the encoder");
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.bias = bias_dev; p.out = out_dev; p.ldo = N;
    return launch_gemm((const __nv_bfloat16*)a_dev, K, (const __nv_bfloat16*)w_dev, K, p, epi, (cudaStream_t)stream);
}

int cbas_b200_encoder_set_option(cbas_encoder* enc, int32_t option, int32_t value) {
    if (!enc) return fail("null encoder");
    switch (option) {
        case CBAS_OPT_ATTENTION_IMPL:
            if (value < 0 || value > 2) return fail("attention impl must be 0 (auto), 1 (mma.sync) or 2 (tcgen05)");
            enc->attention_impl = value;
            return 0;
        case CBAS_OPT_PRUNE_LAST_LAYER: enc->prune_last_layer = value != 0; return 0;
        case CBAS_OPT_RESIZE_KERNEL: enc->resize_tiled = value < 0 ? 0 : (value > 2 ? 2 : value); return 0;
        case CBAS_OPT_LN_FUSION:
            if (value < 0 || value > 2) return fail("LayerNorm fusion must be 0 (standalone), 1 (both norms fused) or 2 (norm1 fused)");
            enc->ln_fused = value;
            return 0;
        case CBAS_OPT_SERPENTINE: enc->serpentine = value != 0; return 0;
    }
    return fail("unknown encoder option " + std::to_string(option));
}

int cbas_b200_attention_tc(const void* qkv_bf16_dev, void* out_bf16_dev, const float* rope_cos_dev,
                           const float* rope_sin_dev, int32_t frames, int32_t T, int32_t prefix, int32_t heads,
                           void* stream) {
    return launch_attention_tc((const __nv_bfloat16*)qkv_bf16_dev, (__nv_bfloat16*)out_bf16_dev, rope_cos_dev,
                               rope_sin_dev, frames, T, prefix, heads, (cudaStream_t)stream);
}

int cbas_b200_attention_tc_qk_f16(int32_t T) { return attention_qk_f16(T) ? 1 : 0; }

int cbas_b200_debug_attention_trace(void* trace_dev) {
    g_attention_trace = (long long*)trace_dev;
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

int cbas_b200_ln_stats_init(float* h_dev, void* hb_out_dev, float* stats_out_dev, int32_t rows, int32_t D,
                            void* stream) {
    return launch_ln_stats_init(h_dev, nullptr, 1, 0, (__nv_bfloat16*)hb_out_dev, stats_out_dev, rows, D,
                                (cudaStream_t)stream);
}

int cbas_b200_gemm_resid_ln(const void* a_dev, const void* w_dev, const float* bias_dev, float* h_dev, void* hb_out_dev,
                            const float* stats_in_dev, float* stats_out_dev, int32_t M, int32_t N, int32_t K,
                            void* stream) {
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.bias = bias_dev; p.out = h_dev; p.ldo = N;
    p.ln_in = stats_in_dev; p.ln_in_stride = 1; p.ln_out = stats_out_dev; p.ln_out_stride = 1;
    p.ln_inv_dim = 1.0f / (float)N; p.hb = hb_out_dev; p.ldhb = N;
    return launch_gemm((const __nv_bfloat16*)a_dev, K, (const __nv_bfloat16*)w_dev, K, p, EPI_RESID_LN_F32,
                       (cudaStream_t)stream);
}

int cbas_b200_gemm_ln_a(const void* hb_dev, const float* stats_dev, const void* w_folded_dev, const float* c1_dev,
                        const float* c2_dev, void* out_bf16_dev, int32_t M, int32_t N, int32_t K, int32_t epi,
                        float eps, void* stream) {
    if (epi != EPI_BIAS_BF16 && epi != EPI_BIAS_GELU_BF16) return fail("LayerNorm-consumer GEMM: epi must be 0 or 1");
    if (!stats_dev) return fail("LayerNorm-consumer GEMM needs the row statistics");
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.bias = c2_dev; p.out = out_bf16_dev; p.ldo = N;
    p.ln_in = stats_dev; p.ln_in_stride = 1; p.ln_c1 = c1_dev; p.ln_inv_dim = 1.0f / (float)K; p.ln_eps = eps;
    return launch_gemm((const __nv_bfloat16*)hb_dev, K, (const __nv_bfloat16*)w_folded_dev, K, p, epi,
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
                                    tp, (cudaStream_t)stream, g_resize_tiled);
}

}  // extern "C"
