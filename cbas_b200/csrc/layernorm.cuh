// Row LayerNorm over the fp32 residual stream (HF modeling_dinov3_vit.py:411,416,433,445,520,547;
// nn.LayerNorm(D, eps=1e-5)).  One warp per row, the row held in registers, two-pass fp32 statistics
// (mean, then variance of the centred values - same arithmetic as ATen), 16-byte loads, bf16 or fp32 out.
// HBM-bound: reads D*4 B and writes D*2 B per row.
#pragma once
#include "ptx.cuh"

namespace cbas {

// in:  fp32 rows, row r at in + (r * in_row_stride) * D      (in_row_stride lets the final LN visit CLS rows only)
// out: OutT rows, contiguous [rows, D]
template <int D, typename OutT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ in, long long in_row_stride, const float* __restrict__ gamma,
                 const float* __restrict__ beta, OutT* __restrict__ out, int rows, float eps, int reverse,
                 float* __restrict__ stats) {
    static_assert(D % 128 == 0, "D must be a multiple of 128");
    constexpr int V = D / 128;  // float4 per lane
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();  // programmatic dependent launch: the residual stream is complete from here on
    if (warp >= rows) return;
    if (reverse) warp = rows - 1 - warp;  // last rows first: the ones the preceding kernel left in L2
    const float4* src = reinterpret_cast<const float4*>(in + (long long)warp * in_row_stride * D);
    float4 x[V];
#pragma unroll
    for (int i = 0; i < V; ++i) x[i] = src[lane + 32 * i];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
        q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q * (1.0f / D) + eps);
    if (stats) {
        // hybrid LayerNorm fusion (encoder.cu): the statistics row a LayerNorm-producer GEMM expects of the stream it
        // is about to update - shift = this row's exact mean, sum y = 0, sum y^2 = q  (LN_STAT_FLOATS = 36 floats)
        float* srow = stats + (long long)warp * 36;
        srow[lane] = lane == 16 ? q : 0.f;
        if (lane < 4) srow[32 + lane] = lane == 0 ? mean : 0.f;
    }
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
        float4 y;
        y.x = x[i].x * rstd * g.x + b.x;
        y.y = x[i].y * rstd * g.y + b.y;
        y.z = x[i].z * rstd * g.z + b.z;
        y.w = x[i].w * rstd * g.w + b.w;
        if constexpr (sizeof(OutT) == 2) {
            uint2 o;
            o.x = pack_bf16(y.x, y.y);
            o.y = pack_bf16(y.z, y.w);
            reinterpret_cast<uint2*>(out + (long long)warp * D)[lane + 32 * i] = o;
        } else {
            reinterpret_cast<float4*>(out + (long long)warp * D)[lane + 32 * i] = y;
        }
    }
}

// Start of the fused-LayerNorm chain (gemm_tcgen05.cuh header): exact two-pass statistics of every row of the freshly
// embedded residual stream, the shifted bf16 copy hb = bf16(h - mean) the first QKV GEMM reads as its A operand, and
// the statistics row {sum y = 0, sum y^2 = sum (x - mean)^2, shift = mean}.  Rows of the CLS / register prefix are
// taken from the token table and written into h here (HF modeling_dinov3_vit.py:85-90: cat(cls, registers, patches)),
// so the embedding stage needs no separate fill kernel.  One launch per forward pass; HBM-bound like layernorm_kernel.
//   h: [rows, D] fp32, row = frame * T + token;  prefix_tokens: [P, D] or null (then every row is read from h)
template <int D>
__global__ void __launch_bounds__(256)
ln_stats_init_kernel(float* __restrict__ h, const float* __restrict__ prefix_tokens, int T, int P,
                     __nv_bfloat16* __restrict__ hb, float* __restrict__ stats, int rows) {
    static_assert(D % 128 == 0, "D must be a multiple of 128");
    constexpr int V = D / 128;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float4* hrow = reinterpret_cast<float4*>(h + (long long)row * D);
    const int token = prefix_tokens ? row % T : P;
    float4 x[V];
    if (token < P) {
        const float4* src = reinterpret_cast<const float4*>(prefix_tokens + (long long)token * D);
#pragma unroll
        for (int i = 0; i < V; ++i) { x[i] = __ldg(src + lane + 32 * i); hrow[lane + 32 * i] = x[i]; }
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) x[i] = hrow[lane + 32 * i];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (x[i].x + x[i].y) + (x[i].z + x[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        x[i].x -= mean; x[i].y -= mean; x[i].z -= mean; x[i].w -= mean;
        q += (x[i].x * x[i].x + x[i].y * x[i].y) + (x[i].z * x[i].z + x[i].w * x[i].w);
        uint2 o;
        o.x = pack_bf16(x[i].x, x[i].y);
        o.y = pack_bf16(x[i].z, x[i].w);
        reinterpret_cast<uint2*>(hb + (long long)row * D)[lane + 32 * i] = o;
    }
    q = warp_sum(q);
    // LN_STAT_FLOATS = 36 floats per row (gemm_tcgen05.cuh): sums[16] = 0, squares[16] = {q, 0...}, shift, padding
    float* srow = stats + (long long)row * 36;
    srow[lane] = lane == 16 ? q : 0.f;
    if (lane < 4) srow[32 + lane] = lane == 0 ? mean : 0.f;
}

}  // namespace cbas
