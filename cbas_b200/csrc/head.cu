// CBAS temporal-delta BiLSTM head (classifier_head.py:57-172) driven the way infer_file drives it
// (cbas.py:497-551: one stride-1 window per frame, replicate padding at the ends of the video,
// softmax(logits / max(1e-3, T))), plus the actogram binning of cbas.py:969-999.
//
// The reference builds every window on the host ([512,31,768] fp32 per batch, each embedding row shipped 31x)
// and runs ~40 small launches per batch.  Here the per-frame work is hoisted out of the windows and the window stages
// run over chunks of up to 37 888 windows (two waves of 128-window tiles on 148 SMs; 7 launches per chunk):
//
//   H1 head_split_embed   f16 embeddings -> exact bf16 hi/lo split [n, 2F]
//   H2 tcgen05 GEMM       P[n,512] = x (Wc|Wd|Wa|W1|0)^T   - EMA, deltas and lin1 are linear, so the three 768->128
//                         bottleneck projections are taken ONCE per frame and the EMA / delta / acceleration
//                         recurrences run in the 128-d projected space (verified identity)
//   H3 head_features      per window: window-local EMA (alpha), reflect-padded delta / delta-delta, +bias, GELU,
//                         LayerNorm(128) x3 -> a_t[384] as bf16 hi/lo; linear branch = mean_t EMA(q) + b1
//   H4 tcgen05 GEMM       Z = GELU(a W0^T + b0)                      [rows, 256] fp32
//   H5 head_center_split  z - mean_t z  -> bf16 hi/lo
//   H6 tcgen05 GEMMs      G_dir = z Wih_dir^T + (b_ih + b_hh), one GEMM per direction over the steps that direction
//                         runs.  Rows of the window stages are ordered t-MAJOR (row = t * windows + window), so the
//                         forward direction's steps t < r and the reverse direction's t >= l (21 of 31 each for the
//                         default head) are contiguous row ranges: [steps * windows, 256] fp32 per direction
//   H7 head_lstm_tc       the recurrence (lstm_hidden_size 64): persistent tcgen05 kernel, one CTA per SM and
//                         direction, two 128-window tiles in flight, W_hh (bf16 hi / lo) resident in shared memory, G
//                         by TMA, h fed back through shared memory, c in registers, gates in TMEM (head_lstm_tc.cuh).  lstm_hidden_size 128: head_lstm_dir, fp32 FMA, 4 windows per
//                         warp, the 256 KB W_hh streamed through L1/L2.
//                         lstm_layers = 2: layer 0 runs all steps, its [fwd|rev] outputs are split to bf16 hi/lo
//                         and go through one more input-gate GEMM (H6') and recurrence (H7').
//   use_acceleration = False is the same pipeline with a zero acceleration stream (zero bottleneck / LayerNorm
//                         parameters and zero lin0 columns give exactly the two-stream concat of
//                         classifier_head.py:165-167).
//   H8 head_pool          attention pooling over the centre frames (warp-shuffle reductions), lin2, sigmoid-gate
//                         lerp with the linear branch, temperature softmax
//
// GEMM precision: the reference head is fp32.  Operands are split into bf16 hi + lo parts and the product is
// taken as hi*hi + lo*hi + hi*lo (K tripled: activations are stored [hi | lo] and the GEMM's third K block re-reads
// hi, GemmParams::a_wrap; weights are [hi | hi | lo]; fp32 accumulation in TMEM), which keeps ~16 mantissa bits per
// factor - three orders of magnitude inside the 1e-3 probability gate - while staying on the tensor cores.
#include "../../include/cbas_b200.h"
#include "common.h"
#include "gemm_tcgen05.cuh"
#include "head_lstm_tc.cuh"

#include <cmath>
#include <cstring>
#include <vector>

using namespace cbas;

namespace {

constexpr int HEAD_BN = 128;    // bottleneck width
constexpr int HEAD_PW = 512;    // per-frame projection row: cls | delta | acc bottlenecks (384), lin1 logits (C <= 32), zero padding
constexpr int HEAD_LIN0 = 256;  // lin0 width = LSTM input width
constexpr int HEAD_MAX_HS = 128; // LSTM hidden size: 64 (default) or 128 (the reference's sweep, sweep_runner.py:106)
constexpr int HEAD_MAX_C = 32;

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) {
    // 1 - 2/(e^{2x}+1): exact limits at +-inf, ~1e-7 relative elsewhere
    const float e = __expf(2.0f * x);
    return 1.0f - 2.0f / (e + 1.0f);
}

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// ---------------------------------------------------------------------------------------------- H1
// x' = [hi | lo] (2F bf16; the GEMM re-reads hi for its third K block, GemmParams::a_wrap).  Eight elements per thread:
// 16-byte loads and stores, HBM-bound (2 or 4 bytes in, 4 bytes out per element).
template <typename InT>  // __half: the stored `cls` rows (split exact); float: forward(x) windows (16-bit split)
__global__ void __launch_bounds__(256)
head_split_embed_kernel(const InT* __restrict__ emb, long long n, int F, __nv_bfloat16* __restrict__ xs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one group of 8 elements
    const int g8 = F >> 3;
    if (i >= n * g8) return;
    const long long f = i / g8;
    const int k = (int)(i - f * g8) << 3;
    float v[8];
    if constexpr (sizeof(InT) == 2) {
        const uint4 q = __ldcs(reinterpret_cast<const uint4*>(emb + f * F + k));
        const __half2* h2 = reinterpret_cast<const __half2*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(h2[j]); v[2 * j] = t.x; v[2 * j + 1] = t.y; }
    } else {
        const float4 a = __ldcs(reinterpret_cast<const float4*>(emb + f * F + k));
        const float4 b = __ldcs(reinterpret_cast<const float4*>(emb + f * F + k + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(v[2 * j], h0, l0);
        split_bf16(v[2 * j + 1], h1, l1);
        hi[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lo[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    __nv_bfloat16* o = xs + f * 2 * F + k;
    *reinterpret_cast<uint4*>(o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(o + F) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// ---------------------------------------------------------------------------------------------- H3
struct FeatParams {
    const float* P;      // [n, HEAD_PW] per-frame projections (cls | delta | acc | lin1 logits | zeros), no bias
    long long n;         // frames in the video
    long long w0;        // centre frame of the first window of this chunk
    long long stride;    // centre-frame distance between consecutive windows: 1 (infer_file) or T (forward(x))
    int windows;         // windows in this chunk
    int T, hsl, l, r;    // seq_len, seq_len/2, centre window [l, r)
    float alpha;
    const float* bias;   // [384] cls_b | delta_b | acc_b
    const float* ln_g;   // [384]
    const float* ln_b;   // [384]
    const float* lin1_b; // [C]
    int C;
    __nv_bfloat16* A;    // [T*windows, 768] = [hi(384) | lo(384)], row = t * windows + window
    float* lin_logits;   // [windows, C]
};

__device__ __forceinline__ void feat_emit(const float (&v)[4], int stream, int lane, const FeatParams& p,
                                          __nv_bfloat16* row) {
    // v = 4 channels (lane*4..) of one stream before bias; GELU, LayerNorm over the 128 channels of the stream
    const int ch = stream * HEAD_BN + lane * 4;
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + ch));
    float y[4] = {gelu_erf_fast<5>(v[0] + b.x), gelu_erf_fast<5>(v[1] + b.y), gelu_erf_fast<5>(v[2] + b.z),
                  gelu_erf_fast<5>(v[3] + b.w)};
    const float mean = warp_sum((y[0] + y[1]) + (y[2] + y[3])) * (1.0f / HEAD_BN);
    float d[4] = {y[0] - mean, y[1] - mean, y[2] - mean, y[3] - mean};
    const float var = warp_sum((d[0] * d[0] + d[1] * d[1]) + (d[2] * d[2] + d[3] * d[3])) * (1.0f / HEAD_BN);
    const float rstd = rsqrtf(var + 1e-5f);
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_g + ch));
    const float4 bb = __ldg(reinterpret_cast<const float4*>(p.ln_b + ch));
    const float o[4] = {d[0] * rstd * g.x + bb.x, d[1] * rstd * g.y + bb.y, d[2] * rstd * g.z + bb.z,
                        d[3] * rstd * g.w + bb.w};
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16(o[i], hi[i], lo[i]);
    uint2 uh, ul;
    uh.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
    uh.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16);
    ul.x = (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16);
    ul.y = (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16);
    *reinterpret_cast<uint2*>(row + ch) = uh;
    *reinterpret_cast<uint2*>(row + 384 + ch) = ul;
}

// one warp per window; lane owns channels lane*4..+3 of each of the three streams and (lane < C) one q channel
__global__ void __launch_bounds__(256)
head_features_kernel(const FeatParams p) {
    const int wl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wl >= p.windows) return;
    const long long f = p.w0 + wl * p.stride;
    __nv_bfloat16* rows = p.A + (long long)wl * 768;  // row of (window, t) = t * windows + window
    const long long tstride = (long long)p.windows * 768;
    const float a = p.alpha;

    auto frame_of = [&](int t) {  // replicate padding == clamped frame index (cbas.py:512-525)
        long long g = f - p.hsl + t;
        return g < 0 ? 0 : (g >= p.n ? p.n - 1 : g);
    };
    auto load3 = [&](int t, float (&c)[4], float (&d)[4], float (&e)[4], float& qq) {
        const long long g = frame_of(t);
        const float* pr = p.P + g * HEAD_PW + lane * 4;
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(pr));
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(pr + 128));
        const float4 x2 = __ldg(reinterpret_cast<const float4*>(pr + 256));
        c[0] = x0.x; c[1] = x0.y; c[2] = x0.z; c[3] = x0.w;
        d[0] = x1.x; d[1] = x1.y; d[2] = x1.z; d[3] = x1.w;
        e[0] = x2.x; e[1] = x2.y; e[2] = x2.z; e[3] = x2.w;
        qq = lane < p.C ? __ldg(p.P + g * HEAD_PW + 3 * HEAD_BN + lane) : 0.f;  // lin1 logits ride in the same row
    };

    // EMA states of the three projected streams at t = 0, 1, 2 (the reflect padding needs s1, s2 for t = 0)
    float sc[4], sd0[4], sa0[4], sq;
    load3(0, sc, sd0, sa0, sq);
    float xc[4], xd[4], xa[4], xq;
    float sc1[4], sd1[4], sa1[4], sq1;
    load3(1, xc, xd, xa, xq);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sc1[i] = sc[i] + a * (xc[i] - sc[i]);
        sd1[i] = sd0[i] + a * (xd[i] - sd0[i]);
        sa1[i] = sa0[i] + a * (xa[i] - sa0[i]);
    }
    sq1 = sq + a * (xq - sq);
    float sc2[4], sd2[4], sa2[4], sq2;
    load3(2, xc, xd, xa, xq);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sc2[i] = sc1[i] + a * (xc[i] - sc1[i]);
        sd2[i] = sd1[i] + a * (xd[i] - sd1[i]);
        sa2[i] = sa1[i] + a * (xa[i] - sa1[i]);
    }
    sq2 = sq1 + a * (xq - sq1);

    float qsum = 0.f;
    float v[4];
    // t = 0: cls = s0 ; delta = s0 - s1 ; acc = s0 - 2 s1 + s2
    feat_emit(sc, 0, lane, p, rows);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = sd0[i] - sd1[i];
    feat_emit(v, 1, lane, p, rows);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = sa0[i] - 2.0f * sa1[i] + sa2[i];
    feat_emit(v, 2, lane, p, rows);
    if (0 >= p.l && 0 < p.r) qsum += sq;
    // t = 1: cls = s1 ; delta = s1 - s0 ; acc = 2 (s1 - s0)
    feat_emit(sc1, 0, lane, p, rows + tstride);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = sd1[i] - sd0[i];
    feat_emit(v, 1, lane, p, rows + tstride);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = 2.0f * (sa1[i] - sa0[i]);
    feat_emit(v, 2, lane, p, rows + tstride);
    if (1 >= p.l && 1 < p.r) qsum += sq1;
    // t = 2: cls = s2 ; delta = s2 - s1 ; acc = s2 - 2 s1 + s0
    feat_emit(sc2, 0, lane, p, rows + 2 * tstride);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = sd2[i] - sd1[i];
    feat_emit(v, 1, lane, p, rows + 2 * tstride);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = sa2[i] - 2.0f * sa1[i] + sa0[i];
    feat_emit(v, 2, lane, p, rows + 2 * tstride);
    if (2 >= p.l && 2 < p.r) qsum += sq2;

    // t >= 3: roll the states (cur = s_{t-1}, prev = s_{t-2})
    float ccur[4], dcur[4], dprev[4], acur[4], aprev[4], qcur = sq2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ccur[i] = sc2[i];
        dcur[i] = sd2[i];
        acur[i] = sa2[i]; aprev[i] = sa1[i];
    }
    for (int t = 3; t < p.T; ++t) {
        load3(t, xc, xd, xa, xq);
        float cn[4], dn[4], an[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cn[i] = ccur[i] + a * (xc[i] - ccur[i]);
            dn[i] = dcur[i] + a * (xd[i] - dcur[i]);
            an[i] = acur[i] + a * (xa[i] - acur[i]);
        }
        qcur = qcur + a * (xq - qcur);
        __nv_bfloat16* row = rows + t * tstride;
        feat_emit(cn, 0, lane, p, row);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = dn[i] - dcur[i];
        feat_emit(v, 1, lane, p, row);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = an[i] - 2.0f * acur[i] + aprev[i];
        feat_emit(v, 2, lane, p, row);
        if (t >= p.l && t < p.r) qsum += qcur;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ccur[i] = cn[i];
            dprev[i] = dcur[i]; dcur[i] = dn[i];
            aprev[i] = acur[i]; acur[i] = an[i];
        }
    }
    (void)dprev;
    if (lane < p.C) p.lin_logits[(long long)wl * p.C + lane] = qsum / (float)(p.r - p.l) + __ldg(p.lin1_b + lane);
}

// ---------------------------------------------------------------------------------------------- H5
// Subtract the mean over the T rows of a window (fp32 mean, classifier_head.py:166-167) and split to bf16 [hi | lo].
// Two warps per window, 128 columns each: a lane's 4 columns of all T <= CENTER_MAX_T rows stay in registers between
// the mean and the output pass, so Z is read from HBM once (the rows of a window are a whole chunk apart).
constexpr int CENTER_MAX_T = 32;
template <bool RESIDENT>
__global__ void __launch_bounds__(256)
head_center_split_kernel(const float* __restrict__ Z, int windows, int T, __nv_bfloat16* __restrict__ Zs) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int wl = gw >> 1, col = (gw & 1) * 128 + lane * 4;
    if (wl >= windows) return;
    const float* z = Z + (long long)wl * HEAD_LIN0 + col;  // rows are t-major: row = t * windows + window
    const long long zs = (long long)windows * HEAD_LIN0;
    float4 v[RESIDENT ? CENTER_MAX_T : 1];
    float m[4] = {0, 0, 0, 0};
    if constexpr (RESIDENT) {
#pragma unroll
        for (int t = 0; t < CENTER_MAX_T; ++t)
            if (t < T) v[t] = __ldcs(reinterpret_cast<const float4*>(z + t * zs));
#pragma unroll
        for (int t = 0; t < CENTER_MAX_T; ++t)
            if (t < T) { m[0] += v[t].x; m[1] += v[t].y; m[2] += v[t].z; m[3] += v[t].w; }
    } else {
        for (int t = 0; t < T; ++t) {
            const float4 a = *reinterpret_cast<const float4*>(z + t * zs);
            m[0] += a.x; m[1] += a.y; m[2] += a.z; m[3] += a.w;
        }
    }
    const float inv = 1.0f / (float)T;
#pragma unroll
    for (int i = 0; i < 4; ++i) m[i] *= inv;
    __nv_bfloat16* o = Zs + (long long)wl * 512 + col;
    auto emit = [&](int t, const float4& a) {
        const float x[4] = {a.x - m[0], a.y - m[1], a.z - m[2], a.w - m[3]};
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16(x[i], h[i], l[i]);
        __nv_bfloat16* r = o + (long long)t * windows * 512;
        *reinterpret_cast<uint2*>(r) = *reinterpret_cast<uint2*>(h);
        *reinterpret_cast<uint2*>(r + 256) = *reinterpret_cast<uint2*>(l);
    };
    if constexpr (RESIDENT) {
#pragma unroll
        for (int t = 0; t < CENTER_MAX_T; ++t)
            if (t < T) emit(t, v[t]);
    } else {
        for (int t = 0; t < T; ++t) emit(t, *reinterpret_cast<const float4*>(z + t * zs));
    }
}

// ---------------------------------------------------------------------------------------------- H7
// grid = (window tiles, 2 directions); 8 warps x 4 windows.  Shared memory: W_hh^T of this direction as
// Wt[k][unit] float4 (i,f,g,o) = 64 KB, plus per-warp h[k][4 windows].
constexpr int LSTM_WARPS = 8;
constexpr int LSTM_WPW = 4;  // windows per warp
constexpr int lstm_smem_bytes(int HS, bool resident) {
    return (resident ? HS * HS * 16 : 0) + LSTM_WARPS * HS * LSTM_WPW * 4;
}

// RESIDENT: W_hh of this direction lives in shared memory; otherwise it is read through the read-only path every
// step (HS = 128: 256 KB per direction does not fit next to anything else; all warps of the CTA walk k in step, so
// the reads are L1 hits after the first warp).
template <int HS, bool RESIDENT>
__global__ void __launch_bounds__(LSTM_WARPS * 32)
head_lstm_dir_kernel(const float* __restrict__ Gf,       // [r * windows, 4*HS] input gates of the forward direction, row = t * windows + window
                     const float* __restrict__ Gr,       // [(T - l) * windows, 4*HS] ... of the reverse direction, row = (t - l) * windows + window
                     const float* __restrict__ whh_t,    // [2][HS k][HS unit][4 gate]
                     int windows, int T, int l, int r, int hout_t_major,
                     float* __restrict__ Hout) {         // [windows, r-l, 2*HS] (fwd | rev)
    constexpr int UPL = HS / 32;  // hidden units per lane
    extern __shared__ __align__(16) uint8_t lstm_smem[];
    float4* Wt = reinterpret_cast<float4*>(lstm_smem);                       // [HS][HS] when RESIDENT
    float* hb_all = reinterpret_cast<float*>(lstm_smem + (RESIDENT ? HS * HS * 16 : 0));
    const int dir = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4* wsrc = reinterpret_cast<const float4*>(whh_t) + (long long)dir * HS * HS;
    if (RESIDENT) {
        for (int i = threadIdx.x; i < HS * HS; i += blockDim.x) Wt[i] = __ldg(wsrc + i);
    }
    float4* hb = reinterpret_cast<float4*>(hb_all + warp * HS * LSTM_WPW);  // hb[k] = h of the 4 windows at unit k
    for (int k = lane; k < HS; k += 32) hb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    const int w_base = (blockIdx.x * LSTM_WARPS + warp) * LSTM_WPW;
    if (w_base >= windows) return;
    const int steps = dir == 0 ? r : T - l;
    const int n_keep = r - l;
    float c[UPL][LSTM_WPW];
#pragma unroll
    for (int u = 0; u < UPL; ++u)
#pragma unroll
        for (int w = 0; w < LSTM_WPW; ++w) c[u][w] = 0.f;

    for (int s = 0; s < steps; ++s) {
        const int t = dir == 0 ? s : T - 1 - s;
        // input half of the gates (prefetched while the recurrent half is accumulated)
        float acc[UPL][4][LSTM_WPW];
#pragma unroll
        for (int w = 0; w < LSTM_WPW; ++w) {
            const int win = min(w_base + w, windows - 1);
            const float* g = (dir ? Gr : Gf) + ((long long)(dir ? t - l : t) * windows + win) * (4 * HS);
#pragma unroll
            for (int u = 0; u < UPL; ++u)
#pragma unroll
                for (int gt = 0; gt < 4; ++gt) acc[u][gt][w] = __ldg(g + gt * HS + lane + 32 * u);
        }
#pragma unroll 4
        for (int k = 0; k < HS; ++k) {
            const float4 h = hb[k];  // h_k of the 4 windows (broadcast)
            const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int u = 0; u < UPL; ++u) {
                // unit lane + 32u: i,f,g,o weights for h_k
                const float4 wv = RESIDENT ? Wt[k * HS + lane + 32 * u] : __ldg(wsrc + k * HS + lane + 32 * u);
#pragma unroll
                for (int w = 0; w < LSTM_WPW; ++w) {
                    acc[u][0][w] = fmaf(wv.x, hv[w], acc[u][0][w]);
                    acc[u][1][w] = fmaf(wv.y, hv[w], acc[u][1][w]);
                    acc[u][2][w] = fmaf(wv.z, hv[w], acc[u][2][w]);
                    acc[u][3][w] = fmaf(wv.w, hv[w], acc[u][3][w]);
                }
            }
        }
        __syncwarp();  // every lane has read the old h
        float hn[UPL][LSTM_WPW];
#pragma unroll
        for (int u = 0; u < UPL; ++u)
#pragma unroll
            for (int w = 0; w < LSTM_WPW; ++w) {
                const float ig = sigmoid_f(acc[u][0][w]), fg = sigmoid_f(acc[u][1][w]);
                const float gg = tanh_f(acc[u][2][w]), og = sigmoid_f(acc[u][3][w]);
                c[u][w] = fg * c[u][w] + ig * gg;
                hn[u][w] = og * tanh_f(c[u][w]);
            }
#pragma unroll
        for (int u = 0; u < UPL; ++u) hb[lane + 32 * u] = make_float4(hn[u][0], hn[u][1], hn[u][2], hn[u][3]);
        if (t >= l && t < r) {
#pragma unroll
            for (int w = 0; w < LSTM_WPW; ++w) {
                if (w_base + w < windows) {
                    float* o = Hout + (hout_t_major ? (long long)(t - l) * windows + (w_base + w)
                                                    : (long long)(w_base + w) * n_keep + (t - l)) * (2 * HS) + dir * HS;
#pragma unroll
                    for (int u = 0; u < UPL; ++u) o[lane + 32 * u] = hn[u][w];
                }
            }
        }
        __syncwarp();  // new h visible before the next step reads it
    }
}

// fp32 rows -> bf16 [hi | lo] (the A-operand layout of the split GEMMs); K multiple of 4
__global__ void __launch_bounds__(256)
head_split_rows_kernel(const float* __restrict__ X, long long rows, int K, __nv_bfloat16* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 per thread
    const int k4 = K / 4;
    if (i >= rows * k4) return;
    const long long row = i / k4;
    const int k = (int)(i - row * k4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(X + row * K + k);
    const float x[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(x[j], hi[j], lo[j]);
    __nv_bfloat16* o = out + row * 2 * K + k;
    *reinterpret_cast<uint2*>(o) = *reinterpret_cast<uint2*>(hi);
    *reinterpret_cast<uint2*>(o + K) = *reinterpret_cast<uint2*>(lo);
}

// ---------------------------------------------------------------------------------------------- H8
struct PoolParams {
    const float* H;           // [windows, n_keep, 2*HS]
    const float* lin_logits;  // [windows, C]
    int windows, n_keep, C;
    const float* att_w; float att_b; float inv_att_temp;  // scores / (softplus(attention_temp) + 1e-3)
    const float* lin2_w; const float* lin2_b;             // [C, 2*HS], [C]
    float gate_sig;                                       // sigmoid(gate)
    float inv_temperature;                                // 1 / max(1e-3, T)
    float* probs;   // [windows, C] or null
    float* logits;  // [windows, C] or null
    float* rawm;    // [windows, 2*HS] or null (attended latent, classifier_head.py:145)
};

// one warp per window; lane owns CPL = 2*HS/32 consecutive latent channels
template <int HS>
__global__ void __launch_bounds__(256)
head_pool_kernel(const PoolParams p) {
    constexpr int W2 = 2 * HS, CPL = W2 / 32;
    const int wl = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wl >= p.windows) return;
    const float* h = p.H + (long long)wl * p.n_keep * W2 + lane * CPL;
    float aw[CPL], rr[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) { aw[j] = __ldg(p.att_w + lane * CPL + j); rr[j] = 0.f; }
    // online softmax over the kept steps (classifier_head.py:141-145)
    float mx = -INFINITY, den = 0.f;
    for (int t = 0; t < p.n_keep; ++t) {
        float v[CPL];
#pragma unroll
        for (int j = 0; j < CPL; j += 4) {
            const float4 q = *reinterpret_cast<const float4*>(h + (long long)t * W2 + j);
            v[j] = q.x; v[j + 1] = q.y; v[j + 2] = q.z; v[j + 3] = q.w;
        }
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) part = fmaf(v[j], aw[j], part);
        float sc = warp_sum(part);
        sc = (sc + p.att_b) * p.inv_att_temp;
        const float mn = fmaxf(mx, sc);
        const float corr = __expf(mx - mn), e = __expf(sc - mn);
        den = den * corr + e;
#pragma unroll
        for (int j = 0; j < CPL; ++j) rr[j] = rr[j] * corr + e * v[j];
        mx = mn;
    }
    const float inv = 1.0f / den;
#pragma unroll
    for (int j = 0; j < CPL; ++j) rr[j] *= inv;
    if (p.rawm) {
#pragma unroll
        for (int j = 0; j < CPL; j += 4)
            *reinterpret_cast<float4*>(p.rawm + (long long)wl * W2 + lane * CPL + j) =
                make_float4(rr[j], rr[j + 1], rr[j + 2], rr[j + 3]);
    }
    // lin2, gate lerp (classifier_head.py:147,171), temperature softmax (cbas.py:545-546)
    float mine = -INFINITY;  // lane c keeps final logit c
    for (int c = 0; c < p.C; ++c) {
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < CPL; ++j) part = fmaf(rr[j], __ldg(p.lin2_w + c * W2 + lane * CPL + j), part);
        const float lstm = warp_sum(part) + __ldg(p.lin2_b + c);
        const float lin = p.lin_logits[(long long)wl * p.C + c];
        const float fin = lin + p.gate_sig * (lstm - lin);
        if (lane == c) mine = fin;
    }
    const float scaled = lane < p.C ? mine * p.inv_temperature : -INFINITY;
    const float m = warp_max(scaled);
    const float e = lane < p.C ? __expf(scaled - m) : 0.f;
    const float s = warp_sum(e);
    if (lane < p.C) {
        if (p.probs) p.probs[(long long)wl * p.C + lane] = e / s;
        if (p.logits) p.logits[(long long)wl * p.C + lane] = mine;
    }
}

// ---------------------------------------------------------------------------------------------- actogram
// bins[k] = #{ f in bin k : p_b(f) * [max_{b'!=b} p_b'(f) < p_b(f)] >= threshold }   (cbas.py:989-999)
// Real = float for probabilities that never left the GPU, double for tables parsed from the CSV files (the reference
// compares the float64 values pandas parsed against a float64 threshold, cbas.py:991-993).
template <typename Real>
__global__ void __launch_bounds__(256)
actogram_bins_kernel(const Real* __restrict__ probs, long long n, int C, int behavior, Real thr, long long bin_frames,
                     int* __restrict__ bins) {
    const long long bin = blockIdx.x;
    const long long f0 = bin * bin_frames;
    const long long f1 = f0 + bin_frames < n ? f0 + bin_frames : n;
    int count = 0;
    for (long long f = f0 + threadIdx.x; f < f1; f += blockDim.x) {
        const Real* r = probs + f * C;
        const Real pb = r[behavior];
        Real others = -INFINITY;
        bool any = false;
        for (int c = 0; c < C; ++c)
            if (c != behavior) {
                any = true;
                others = fmax(others, r[c]);  // like pandas max(axis=1), fmax skips NaN
            }
        // with no other column the row maximum is NaN and `NaN < p` is False
        const bool is_max = any && (others < pb);
        const Real v = is_max ? pb : pb * Real(0);
        count += (v >= thr) ? 1 : 0;
    }
    __shared__ int sh[8];
    count = (int)warp_sum((float)count);  // counts per warp <= 2^24: exact in fp32
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = count;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += sh[i];
        bins[bin] = tot;
    }
}

template <typename Real>
static int actogram_bins_impl(const Real* probs_dev, int64_t n, int32_t C, int32_t behavior, Real threshold,
                              int64_t bin_frames, int32_t* bins_out_dev, void* stream) {
    if (n < 0 || C < 1 || behavior < 0 || behavior >= C) return fail("actogram: bad shape or behaviour index");
    if (bin_frames <= 0) return fail("actogram: bin size must be positive");
    if (n == 0) return 0;
    if (!probs_dev || !bins_out_dev) return fail("null argument");
    const long long nb = (n + bin_frames - 1) / bin_frames;
    if (nb > 0x7fffffffLL) return fail("actogram: too many bins");
    // no handle here: run on the device that owns the input, whatever the calling thread's current device is
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, probs_dev) != cudaSuccess || attr.type != cudaMemoryTypeDevice)
        return fail("actogram: probs_dev is not device memory");
    DeviceGuard guard(attr.device);
    if (!guard.ok()) return fail("cudaSetDevice to the input's device failed");
    ProfScope prof(PROF_ACTOGRAM, (cudaStream_t)stream);
    actogram_bins_kernel<Real><<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(probs_dev, n, C, behavior, threshold,
                                                                              bin_frames, bins_out_dev);
    count_launch();
    return check_cuda(cudaGetLastError(), "actogram_bins_kernel launch");
}


// host-side helpers ------------------------------------------------------------------------------
uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// W [N,K] fp32 (host) -> device bf16 [N, 3K] = [hi | hi | lo] matching activations laid out [hi | lo | hi]
int upload_split_weight(const std::vector<float>& W, int N, int K, __nv_bfloat16** out) {
    std::vector<uint16_t> buf((size_t)N * 3 * K);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
            const float w = W[(size_t)n * K + k];
            const uint16_t hi = f2bf(w);
            const uint16_t lo = f2bf(w - bf2f(hi));
            uint16_t* r = buf.data() + (size_t)n * 3 * K;
            r[k] = hi; r[K + k] = hi; r[2 * K + k] = lo;
        }
    CBAS_CHECK(cudaMalloc((void**)out, buf.size() * 2));
    return check_cuda(cudaMemcpy(*out, buf.data(), buf.size() * 2, cudaMemcpyHostToDevice), "split weight upload");
}

int download(const float* dev, size_t n, std::vector<float>& host) {
    host.resize(n);
    if (!dev) return fail("null head weight pointer");
    return check_cuda(cudaMemcpy(host.data(), dev, n * 4, cudaMemcpyDeviceToHost), "head weight download");
}

int upload_f32(const std::vector<float>& v, float** out) {
    CBAS_CHECK(cudaMalloc((void**)out, v.size() * 4));
    return check_cuda(cudaMemcpy(*out, v.data(), v.size() * 4, cudaMemcpyHostToDevice), "head weight upload");
}

}  // namespace

struct cbas_head {
    cbas_head_cfg cfg;
    int device = 0;  // the CUDA device that was current at create time: every entry point runs there
    int l = 0, r = 0;
    // derived device weights
    __nv_bfloat16* wp = nullptr;    // [HEAD_PW, 3F]: cls | delta | acc bottleneck rows, lin1 rows, zero rows
    __nv_bfloat16* w0 = nullptr;    // [256, 1152]
    // per LSTM layer: input-gate weights of both directions [8Hs, 3*K_in] (K_in = 256, then 2Hs), b_ih + b_hh [8Hs],
    // recurrent weights [2][Hs k][Hs unit][4 gate]
    __nv_bfloat16* wih[2] = {nullptr, nullptr};
    float* bg[2] = {nullptr, nullptr};
    float* whh_t[2] = {nullptr, nullptr};
    // lstm_hidden 64 (tcgen05 recurrence): the gate columns of wih / bg are ordered 4 * unit + gate inside a direction
    // and the recurrent weights are [2 dir][hi, lo][256 columns][64 k] bf16 (head_lstm_tc.cuh)
    __nv_bfloat16* whh_tc[2] = {nullptr, nullptr};
    float *b3 = nullptr, *ln_g = nullptr, *ln_b = nullptr, *b0 = nullptr;
    float *lin1_b = nullptr, *lin2_w = nullptr, *lin2_b = nullptr, *att_w = nullptr;  // (lin1_w lives in wp)
    float att_b = 0.f, inv_att_temp = 1.f, gate_sig = 0.5f;
    // workspace
    long long cap_frames = 0;
    int max_chunk_windows = 2 * 148 * 128;  // windows per pass of the window stages: two waves of 128-window tiles
    int chunk_windows = 0;                  // ... the chunk workspace currently holds (grown on demand)
    __nv_bfloat16* xs = nullptr;  // [cap, 3F]
    float* P = nullptr;           // [cap, HEAD_PW]
    __nv_bfloat16* A = nullptr;   // [chunk*T, 1152]  (reused as Z' [chunk*T, 768])
    float* Z = nullptr;           // [chunk*T, 256]
    float* G = nullptr;           // [chunk*T, 8Hs]
    float* H0 = nullptr;          // [chunk*T, 2Hs]   layer-0 outputs of a two-layer LSTM
    float* H = nullptr;           // [chunk, r-l, 2Hs]
    float* lin = nullptr;         // [chunk, C]
};

namespace {
void head_free_ws(cbas_head* h) {
    cudaFree(h->xs); cudaFree(h->P);
    h->xs = nullptr; h->P = nullptr; h->cap_frames = 0;
}
int head_ensure_ws(cbas_head* h, long long n) {
    if (n <= h->cap_frames) return 0;
    head_free_ws(h);
    const int F = h->cfg.in_features, C = h->cfg.out_features;
    CBAS_CHECK(cudaMalloc((void**)&h->xs, (size_t)n * 2 * F * 2));
    CBAS_CHECK(cudaMalloc((void**)&h->P, (size_t)n * HEAD_PW * 4));
    h->cap_frames = n;
    return 0;
}
// workspace of the per-window stages for chunks of up to `windows` windows (grows, never shrinks)
int head_ensure_chunk_ws(cbas_head* h, int windows) {
    if (windows <= h->chunk_windows) return 0;
    cudaFree(h->A); cudaFree(h->Z); cudaFree(h->G); cudaFree(h->H0); cudaFree(h->H); cudaFree(h->lin);
    h->A = nullptr; h->Z = nullptr; h->G = nullptr; h->H0 = nullptr; h->H = nullptr; h->lin = nullptr;
    h->chunk_windows = 0;
    const int T = h->cfg.seq_len, HS = h->cfg.lstm_hidden, C = h->cfg.out_features;
    const size_t rows = (size_t)windows * T;
    CBAS_CHECK(cudaMalloc((void**)&h->A, rows * 768 * 2));
    CBAS_CHECK(cudaMalloc((void**)&h->Z, rows * HEAD_LIN0 * 4));
    CBAS_CHECK(cudaMalloc((void**)&h->G, rows * 8 * HS * 4));
    if (h->cfg.lstm_layers == 2) CBAS_CHECK(cudaMalloc((void**)&h->H0, rows * 2 * HS * 4));
    CBAS_CHECK(cudaMalloc((void**)&h->H, (size_t)windows * (h->r - h->l) * 2 * HS * 4));
    CBAS_CHECK(cudaMalloc((void**)&h->lin, (size_t)windows * C * 4));
    h->chunk_windows = windows;
    return 0;
}
}  // namespace

extern "C" {

int cbas_b200_head_create(const cbas_head_cfg* cfg, const cbas_head_weights* w, cbas_head** out) {
    if (!cfg || !w || !out) return fail("null argument");
    if (cfg->bottleneck != HEAD_BN) return fail("head: bottleneck_dim must be 128");
    if (cfg->lstm_hidden != 64 && cfg->lstm_hidden != 128) return fail("head: lstm_hidden_size must be 64 or 128");
    if (cfg->lstm_layers != 1 && cfg->lstm_layers != 2) return fail("head: lstm_layers must be 1 or 2");
    if (cfg->in_features % 64 || cfg->in_features <= 0) return fail("head: in_features must be a multiple of 64");
    if (cfg->out_features < 1 || cfg->out_features > HEAD_MAX_C) return fail("head: 1..32 behaviours supported");
    if (cfg->seq_len < 3 || cfg->seq_len % 2 == 0 || cfg->seq_len > 255) return fail("head: seq_len must be odd, 3..255");
    const int F = cfg->in_features, C = cfg->out_features, T = cfg->seq_len, HS = cfg->lstm_hidden;
    const bool acc_stream = cfg->use_acceleration != 0;
    auto* h = new cbas_head();
    h->cfg = *cfg;
    if (cudaGetDevice(&h->device) != cudaSuccess) { delete h; return fail("head: no current CUDA device"); }
    const int hsl = T / 2, sw = cfg->center_window;
    h->l = hsl - sw > 0 ? hsl - sw : 0;
    h->r = hsl + sw + 1 < T ? hsl + sw + 1 : T;
    if (h->l >= h->r) { delete h; return fail("head: empty centre window"); }

    int rc = 0;
    std::vector<float> a, b, c3, tmp;
    auto cat3 = [&](const float* x, const float* y, const float* z, size_t n, std::vector<float>& o) -> int {
        std::vector<float> t;
        o.clear();
        for (const float* p : {x, y, z}) {
            if (p == z && !acc_stream) {
                // use_acceleration=False: a zero third stream (LayerNorm of a constant row with gamma = beta = 0 is 0)
                o.insert(o.end(), n, 0.f);
                continue;
            }
            if (int e = download(p, n, t)) return e;
            o.insert(o.end(), t.begin(), t.end());
        }
        return 0;
    };
    std::vector<float> W;
    if (!rc) rc = cat3(w->cls_w, w->delta_w, w->acc_w, (size_t)HEAD_BN * F, W);
    if (!rc) {
        // rows 384 .. 384 + C - 1: lin1 (the linear branch's per-frame logits come out of the same GEMM), zero rows after
        std::vector<float> l1;
        rc = download(w->lin1_w, (size_t)C * F, l1);
        if (!rc) {
            W.insert(W.end(), l1.begin(), l1.end());
            W.resize((size_t)HEAD_PW * F, 0.f);
        }
    }
    if (!rc) rc = upload_split_weight(W, HEAD_PW, F, &h->wp);
    if (!rc) rc = cat3(w->cls_b, w->delta_b, w->acc_b, HEAD_BN, W);
    if (!rc) rc = upload_f32(W, &h->b3);
    if (!rc) rc = cat3(w->cls_ln_g, w->delta_ln_g, w->acc_ln_g, HEAD_BN, W);
    if (!rc) rc = upload_f32(W, &h->ln_g);
    if (!rc) rc = cat3(w->cls_ln_b, w->delta_ln_b, w->acc_ln_b, HEAD_BN, W);
    if (!rc) rc = upload_f32(W, &h->ln_b);
    if (!rc) rc = download(w->lin0_w, (size_t)HEAD_LIN0 * (acc_stream ? 3 : 2) * HEAD_BN, W);
    if (!rc && !acc_stream) {  // [256, 256] -> [256, 384] with zero columns for the absent stream
        std::vector<float> P((size_t)HEAD_LIN0 * 3 * HEAD_BN, 0.f);
        for (int n = 0; n < HEAD_LIN0; ++n)
            memcpy(&P[(size_t)n * 3 * HEAD_BN], &W[(size_t)n * 2 * HEAD_BN], 2 * HEAD_BN * sizeof(float));
        W.swap(P);
    }
    if (!rc) rc = upload_split_weight(W, HEAD_LIN0, 3 * HEAD_BN, &h->w0);
    if (!rc) rc = download(w->lin0_b, HEAD_LIN0, W);
    if (!rc) rc = upload_f32(W, &h->b0);
    // per layer: input-gate weights of both directions stacked, bias = b_ih + b_hh, recurrent weights re-laid-out
    for (int layer = 0; layer < cfg->lstm_layers && !rc; ++layer) {
        const int Kin = layer == 0 ? HEAD_LIN0 : 2 * HS;
        const float* wif = layer ? w->w_ih_f1 : w->w_ih_f; const float* wir = layer ? w->w_ih_r1 : w->w_ih_r;
        const float* whf = layer ? w->w_hh_f1 : w->w_hh_f; const float* whr = layer ? w->w_hh_r1 : w->w_hh_r;
        const float* pbif = layer ? w->b_ih_f1 : w->b_ih_f; const float* pbhf = layer ? w->b_hh_f1 : w->b_hh_f;
        const float* pbir = layer ? w->b_ih_r1 : w->b_ih_r; const float* pbhr = layer ? w->b_hh_r1 : w->b_hh_r;
        std::vector<float> wf, wr, bif, bhf, bir, bhr;
        if (!rc) rc = download(wif, (size_t)4 * HS * Kin, wf);
        if (!rc) rc = download(wir, (size_t)4 * HS * Kin, wr);
        // lstm_hidden 64: gate columns in the order the tcgen05 recurrence wants them (4 * unit + gate per direction)
        const bool tc_order = HS == 64;
        auto gate_col = [&](int d, int g, int u) { return tc_order ? d * 4 * HS + u * 4 + g : d * 4 * HS + g * HS + u; };
        if (!rc) {
            W.assign((size_t)8 * HS * Kin, 0.f);
            for (int d = 0; d < 2; ++d)
                for (int g = 0; g < 4; ++g)
                    for (int u = 0; u < HS; ++u)
                        memcpy(&W[(size_t)gate_col(d, g, u) * Kin], &(d ? wr : wf)[(size_t)(g * HS + u) * Kin],
                               (size_t)Kin * sizeof(float));
            rc = upload_split_weight(W, 8 * HS, Kin, &h->wih[layer]);
        }
        if (!rc) rc = download(pbif, 4 * HS, bif);
        if (!rc) rc = download(pbhf, 4 * HS, bhf);
        if (!rc) rc = download(pbir, 4 * HS, bir);
        if (!rc) rc = download(pbhr, 4 * HS, bhr);
        if (!rc) {
            W.assign(8 * HS, 0.f);
            for (int g = 0; g < 4; ++g)
                for (int u = 0; u < HS; ++u) {
                    W[gate_col(0, g, u)] = bif[g * HS + u] + bhf[g * HS + u];
                    W[gate_col(1, g, u)] = bir[g * HS + u] + bhr[g * HS + u];
                }
            rc = upload_f32(W, &h->bg[layer]);
        }
        // recurrent weights: Wt[dir][k][unit][gate] = W_hh[dir][gate*Hs + unit][k]
        std::vector<float> hf, hr;
        if (!rc) rc = download(whf, (size_t)4 * HS * HS, hf);
        if (!rc) rc = download(whr, (size_t)4 * HS * HS, hr);
        if (!rc) {
            W.assign((size_t)2 * HS * HS * 4, 0.f);
            for (int d = 0; d < 2; ++d) {
                const std::vector<float>& S = d ? hr : hf;
                for (int k = 0; k < HS; ++k)
                    for (int u = 0; u < HS; ++u)
                        for (int g = 0; g < 4; ++g)
                            W[(((size_t)d * HS + k) * HS + u) * 4 + g] = S[(size_t)(g * HS + u) * HS + k];
            }
            rc = upload_f32(W, &h->whh_t[layer]);
        }
        if (!rc && tc_order) {
            // [dir][hi, lo][column 4 * unit + gate][k] bf16 for head_lstm_tc_kernel
            std::vector<uint16_t> B((size_t)2 * 2 * 4 * HS * HS);
            for (int d = 0; d < 2; ++d) {
                const std::vector<float>& S = d ? hr : hf;
                for (int g = 0; g < 4; ++g)
                    for (int u = 0; u < HS; ++u)
                        for (int k = 0; k < HS; ++k) {
                            const float w = S[(size_t)(g * HS + u) * HS + k];
                            const uint16_t hi = f2bf(w), lo = f2bf(w - bf2f(hi));
                            const size_t col = (size_t)u * 4 + g;
                            B[(((size_t)d * 2 + 0) * 4 * HS + col) * HS + k] = hi;
                            B[(((size_t)d * 2 + 1) * 4 * HS + col) * HS + k] = lo;
                        }
            }
            rc = check_cuda(cudaMalloc((void**)&h->whh_tc[layer], B.size() * 2), "head weights");
            if (!rc) rc = check_cuda(cudaMemcpy(h->whh_tc[layer], B.data(), B.size() * 2, cudaMemcpyHostToDevice),
                                     "recurrent weight upload");
        }
    }
    if (!rc) rc = download(w->lin1_b, C, W);
    if (!rc) rc = upload_f32(W, &h->lin1_b);
    if (!rc) rc = download(w->lin2_w, (size_t)C * 2 * HS, W);
    if (!rc) rc = upload_f32(W, &h->lin2_w);
    if (!rc) rc = download(w->lin2_b, C, W);
    if (!rc) rc = upload_f32(W, &h->lin2_b);
    if (!rc) rc = download(w->att_w, 2 * HS, W);
    if (!rc) rc = upload_f32(W, &h->att_w);
    if (!rc) rc = download(w->att_b, 1, W);
    if (!rc) {
        h->att_b = W[0];
        const double sp = std::log1p(std::exp((double)w->attention_temp));  // F.softplus
        h->inv_att_temp = (float)(1.0 / (sp + 1e-3));
        h->gate_sig = (float)(1.0 / (1.0 + std::exp(-(double)w->gate)));
    }
    // (the chunk workspace is allocated by head_ensure_chunk_ws on the first run, sized to the work)
    if (!rc && HS != 64)
        rc = check_cuda(cudaFuncSetAttribute(head_lstm_dir_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             lstm_smem_bytes(128, false)), "lstm smem attribute");
    if (rc) { cbas_b200_head_destroy(h); return rc; }
    *out = h;
    return 0;
}

void cbas_b200_head_destroy(cbas_head* h) {
    if (!h) return;
    DeviceGuard guard(h->device);
    head_free_ws(h);
    cudaFree(h->wp); cudaFree(h->w0); cudaFree(h->b3); cudaFree(h->ln_g); cudaFree(h->ln_b); cudaFree(h->b0);
    for (int i = 0; i < 2; ++i) { cudaFree(h->wih[i]); cudaFree(h->bg[i]); cudaFree(h->whh_t[i]); cudaFree(h->whh_tc[i]); }
    cudaFree(h->H0); cudaFree(h->lin1_b);
    cudaFree(h->lin2_w); cudaFree(h->lin2_b); cudaFree(h->att_w);
    cudaFree(h->A); cudaFree(h->Z); cudaFree(h->G); cudaFree(h->H); cudaFree(h->lin);
    delete h;
}

static int head_run(cbas_head* h, const void* x_dev, bool x_is_f16, long long n_frames, long long n_windows,
                    long long first_center, long long stride, float temperature, float* probs_out, float* logits_out,
                    float* rawm_out, cudaStream_t s) {
    const int F = h->cfg.in_features, C = h->cfg.out_features, T = h->cfg.seq_len;
    if (int rc = head_ensure_ws(h, n_frames)) return rc;
    // H1 + H2: per-frame work, once for the whole sequence
    {
        ProfScope prof(PROF_HEAD_SPLIT, s);
        const long long groups = n_frames * (F / 8);
        const unsigned grid = (unsigned)((groups + 255) / 256);
        if (x_is_f16) head_split_embed_kernel<__half><<<grid, 256, 0, s>>>((const __half*)x_dev, n_frames, F, h->xs);
        else head_split_embed_kernel<float><<<grid, 256, 0, s>>>((const float*)x_dev, n_frames, F, h->xs);
        count_launch();
        if (int rc = check_cuda(cudaGetLastError(), "head_split_embed_kernel launch")) return rc;
    }
    for (long long f0 = 0; f0 < n_frames; f0 += (1 << 20)) {  // the GEMM takes int M
        const int m = (int)((n_frames - f0) < (1 << 20) ? (n_frames - f0) : (1 << 20));
        GemmParams p{};
        p.M = m; p.N = HEAD_PW; p.K = 3 * F; p.a_wrap = 2 * F; p.bias = nullptr; p.out = h->P + f0 * HEAD_PW; p.ldo = HEAD_PW;
        if (int rc = launch_gemm(h->xs + f0 * 2 * F, 2 * F, h->wp, 3 * F, p, EPI_BIAS_F32, s, PROF_HEAD_PROJ_GEMM))
            return rc;
    }
    const float inv_temp = 1.0f / (temperature > 1e-3f ? temperature : 1e-3f);
    const int n_keep = h->r - h->l;
    const int chunk = (int)(n_windows < h->max_chunk_windows ? n_windows : h->max_chunk_windows);
    if (int rc = head_ensure_chunk_ws(h, chunk)) return rc;
    for (long long w0 = 0; w0 < n_windows; w0 += chunk) {
        const int nw = (int)((n_windows - w0) < chunk ? (n_windows - w0) : chunk);
        const int rows = nw * T;
        {
            ProfScope prof(PROF_HEAD_FEATURES, s);
            FeatParams fp{h->P, n_frames, first_center + w0 * stride, stride, nw, T, T / 2, h->l, h->r,
                          h->cfg.ema_alpha, h->b3, h->ln_g, h->ln_b, h->lin1_b, C, h->A, h->lin};
            head_features_kernel<<<(nw * 32 + 255) / 256, 256, 0, s>>>(fp);
            count_launch();
            if (int rc = check_cuda(cudaGetLastError(), "head_features_kernel launch")) return rc;
        }
        {
            GemmParams p{};
            p.M = rows; p.N = HEAD_LIN0; p.K = 1152; p.a_wrap = 768; p.bias = h->b0; p.out = h->Z; p.ldo = HEAD_LIN0;
            if (int rc = launch_gemm(h->A, 768, h->w0, 1152, p, EPI_BIAS_GELU_F32, s, PROF_HEAD_LIN0_GEMM)) return rc;
        }
        __nv_bfloat16* Zs = h->A;  // A is dead after the lin0 GEMM: reuse it for the split, centred z
        {
            ProfScope prof(PROF_HEAD_CENTER, s);
            const unsigned cgrid = (unsigned)(((long long)nw * 64 + 255) / 256);
            if (T <= CENTER_MAX_T) head_center_split_kernel<true><<<cgrid, 256, 0, s>>>(h->Z, nw, T, Zs);
            else head_center_split_kernel<false><<<cgrid, 256, 0, s>>>(h->Z, nw, T, Zs);
            count_launch();
            if (int rc = check_cuda(cudaGetLastError(), "head_center_split_kernel launch")) return rc;
        }
        const int HS = h->cfg.lstm_hidden, layers = h->cfg.lstm_layers;
        // One LSTM layer: the input half of the gates for the steps each direction actually runs (forward: t < keep_r,
        // reverse: t >= keep_l - rows are t-major, so both are contiguous row ranges of X), then the recurrence.
        //   X: [T * nw, a_wrap] bf16 [hi | lo];  Gf = G, Gr = G + keep_r * nw * 4HS
        auto run_lstm = [&](int layer, const __nv_bfloat16* X, int a_wrap, int keep_l, int keep_r, bool hout_t_major,
                            float* Hout) -> int {
            const int K3 = a_wrap / 2 * 3;
            float* Gf = h->G;
            float* Gr = h->G + (size_t)keep_r * nw * 4 * HS;
            for (int d = 0; d < 2; ++d) {
                GemmParams p{};
                p.M = (d ? T - keep_l : keep_r) * nw; p.N = 4 * HS; p.K = K3; p.a_wrap = a_wrap;
                p.bias = h->bg[layer] + d * 4 * HS; p.out = d ? Gr : Gf; p.ldo = 4 * HS;
                const __nv_bfloat16* A = X + (d ? (size_t)keep_l * nw * a_wrap : 0);
                if (int rc = launch_gemm(A, a_wrap, h->wih[layer] + (size_t)d * 4 * HS * K3, K3, p, EPI_BIAS_F32, s,
                                         PROF_HEAD_IH_GEMM)) return rc;
            }
            ProfScope prof(PROF_HEAD_LSTM, s);
            if (HS == 64) {
                // persistent tcgen05 recurrence: one CTA per SM and direction, two 128-window tiles in flight per CTA
                static DeviceSmemOptIn optin;
                CBAS_CHECK(optin.ensure(head_lstm_tc_kernel, HLT_SMEM_BYTES));
                CUtensorMap tgf, tgr;  // G of a direction as [steps][windows][256], box {32, 128, 1}
                if (int rc = make_tmap_3d_f32(&tgf, Gf, 256, nw, keep_r, 1024, (long long)nw * 1024, 32, 128, 1)) return rc;
                if (int rc = make_tmap_3d_f32(&tgr, Gr, 256, nw, T - keep_l, 1024, (long long)nw * 1024, 32, 128, 1)) return rc;
                const int pairs = ((nw + 127) / 128 + 1) / 2, per_dir = sm_count() / 2 > 0 ? sm_count() / 2 : 1;
                dim3 grid(pairs < per_dir ? pairs : per_dir, 2);
                head_lstm_tc_kernel<<<grid, HLT_THREADS, HLT_SMEM_BYTES, s>>>(tgf, tgr, h->whh_tc[layer], nw, T, keep_l, keep_r,
                                                                            hout_t_major ? 1 : 0, Hout);
                count_launch();
                return check_cuda(cudaGetLastError(), "head_lstm_tc_kernel launch");
            }
            const int per_cta = LSTM_WARPS * LSTM_WPW;
            dim3 grid((nw + per_cta - 1) / per_cta, 2);
            head_lstm_dir_kernel<128, false><<<grid, LSTM_WARPS * 32, lstm_smem_bytes(128, false), s>>>(
                Gf, Gr, h->whh_t[layer], nw, T, keep_l, keep_r, hout_t_major ? 1 : 0, Hout);
            count_launch();
            return check_cuda(cudaGetLastError(), "head_lstm_dir_kernel launch");
        };
        const __nv_bfloat16* X = Zs;
        int x_wrap = 2 * HEAD_LIN0;
        if (layers == 2) {
            // layer 0 over every step (outputs t-major like its inputs), then its [fwd | rev] outputs are the next layer's
            // inputs (nn.LSTM stacking)
            if (int rc = run_lstm(0, X, x_wrap, 0, T, true, h->H0)) return rc;
            {
                ProfScope prof(PROF_HEAD_CENTER, s);
                const long long quads = (long long)rows * (2 * HS / 4);
                head_split_rows_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, s>>>(h->H0, rows, 2 * HS, h->A);
                count_launch();
                if (int rc = check_cuda(cudaGetLastError(), "head_split_rows_kernel launch")) return rc;
            }
            X = h->A;
            x_wrap = 4 * HS;
        }
        if (int rc = run_lstm(layers - 1, X, x_wrap, h->l, h->r, false, h->H)) return rc;
        {
            ProfScope prof(PROF_HEAD_LSTM, s);
            PoolParams pp{h->H, h->lin, nw, n_keep, C, h->att_w, h->att_b, h->inv_att_temp, h->lin2_w, h->lin2_b,
                          h->gate_sig, inv_temp, probs_out ? probs_out + w0 * C : nullptr,
                          logits_out ? logits_out + w0 * C : nullptr,
                          rawm_out ? rawm_out + w0 * 2 * HS : nullptr};
            if (HS == 64) head_pool_kernel<64><<<(nw * 32 + 255) / 256, 256, 0, s>>>(pp);
            else head_pool_kernel<128><<<(nw * 32 + 255) / 256, 256, 0, s>>>(pp);
            count_launch();
            if (int rc = check_cuda(cudaGetLastError(), "head_pool_kernel launch")) return rc;
        }
    }
    return 0;
}

int cbas_b200_head_infer(cbas_head* h, const void* emb_f16_dev, int64_t n_frames, float temperature,
                         float* probs_out_dev, float* logits_out_dev, void* stream) {
    if (!h) return fail("null head");
    if (n_frames < 0) return fail("negative frame count");
    if (n_frames == 0) return 0;
    if (!emb_f16_dev || !probs_out_dev) return fail("null argument");
    DeviceGuard guard(h->device);  // the handle's device, whatever the calling thread's current device is
    if (!guard.ok()) return fail("cudaSetDevice to the head's device failed");
    // one window per frame, centred on it, replicate-padded at the ends (cbas.py:497-551)
    return head_run(h, emb_f16_dev, true, n_frames, n_frames, 0, 1, temperature, probs_out_dev, logits_out_dev, nullptr,
                    (cudaStream_t)stream);
}

int cbas_b200_head_forward_windows(cbas_head* h, const float* x_f32_dev, int64_t n_windows, float* logits_out_dev,
                                   float* rawm_out_dev, void* stream) {
    if (!h) return fail("null head");
    if (n_windows < 0) return fail("negative window count");
    if (n_windows == 0) return 0;
    if (!x_f32_dev || !logits_out_dev) return fail("null argument");
    DeviceGuard guard(h->device);
    if (!guard.ok()) return fail("cudaSetDevice to the head's device failed");
    // B independent windows laid end to end: window b is centred on frame b*T + T/2 and never leaves its block
    const int T = h->cfg.seq_len;
    return head_run(h, x_f32_dev, false, n_windows * T, n_windows, T / 2, T, 1.0f, nullptr, logits_out_dev, rawm_out_dev,
                    (cudaStream_t)stream);
}

int cbas_b200_actogram_bins(const float* probs_dev, int64_t n, int32_t C, int32_t behavior, float threshold,
                            int64_t bin_frames, int32_t* bins_out_dev, void* stream) {
    return actogram_bins_impl<float>(probs_dev, n, C, behavior, threshold, bin_frames, bins_out_dev, stream);
}

int cbas_b200_actogram_bins_f64(const double* probs_dev, int64_t n, int32_t C, int32_t behavior, double threshold,
                                int64_t bin_frames, int32_t* bins_out_dev, void* stream) {
    return actogram_bins_impl<double>(probs_dev, n, C, behavior, threshold, bin_frames, bins_out_dev, stream);
}

}  // extern "C"
