// LSTM recurrence of the CBAS head (classifier_head.py:105 nn.LSTM(256, 64, bidirectional), :124 forward_lstm) on
// tcgen05, hidden size 64: a persistent kernel whose recurrent weights stay in shared memory for its whole life.
//
// One CTA per SM works on ONE direction (blockIdx.y) and advances TWO independent 128-window tiles ("chains") side
// by side, so one chain's MMA / barrier round trip hides under the other chain's activations.  Per chain and step
//     gates[128 windows, 256] = h_{t-1}[128, 64] . W_hh^T[64, 256]          tcgen05.mma, fp32 accumulator in TMEM
//                             + G_t[128 windows, 256]                         input half of the gates, from HBM by TMA
// with h and W_hh as bf16 hi + lo pairs (hi*hi + lo*hi + hi*lo: the accuracy of the head's other GEMMs, see head.cu).
// W_hh (hi, lo: 2 x 32 KB, K-major, 128-byte swizzle) is written to shared memory once per CTA and shared by both
// chains; h_t is written by the activation warps straight into the swizzled K-major A tile the next step's MMA reads
// (no global round trip); c_t lives in registers.  Only the gate pre-activations G (the W_ih GEMM's output) stream in:
// the gate columns are ordered unit-major (column = 4 * unit + gate; the host permutes the rows of W_ih, W_hh and the
// biases), so a float4 is one cell's (i, f, g, o), and a TMA box {32 columns, 1 step, 128 windows} = 16 KB is the
// input of 8 units of all 128 windows.  A row-per-thread global read of the same data (the TMEM lane = window
// mapping) touches 32 different cache lines per warp instruction and ran 7x slower; TMA gathers the 128-byte row pieces
// (rows are 1 KB apart) into a 3-stage ring per chain and the threads read them conflict-free from shared memory.
//
// Warp roles: warps 0-15 activation (chain = warp >> 3, TMEM lane quarter = warp & 3, half = (warp >> 2) & 1: a
// thread owns one window and, of every 8-unit piece, the 4 units of its half); warps 16 / 17 MMA issue for chain
// 0 / 1 (+ TMEM allocation); warps 18 / 19 TMA producers.  MMA and TMA warps run warp-uniform with one elected lane.
// Per cell 5 ex2 + 3 rcp on the MUFU: sigmoid(a) tanh(b) = (e^2b - 1) / ((1 + e^-a)(e^2b + 1)) shares one reciprocal.
// Bounds: the MUFU (16 / clk / SM: 8 per cell) and the G stream (1 KB per window, direction and step).
#pragma once
#include "ptx.cuh"

namespace cbas {

constexpr int HLT_CHAINS = 2;
constexpr int HLT_ACT_WARPS = 8 * HLT_CHAINS;
constexpr int HLT_THREADS = 32 * (HLT_ACT_WARPS + 2 * HLT_CHAINS);
constexpr int HLT_W_BYTES = 256 * 128;   // [256 gate columns][64 k] bf16, one swizzled operand tile
constexpr int HLT_A_BYTES = 128 * 128;   // [128 windows][64 k] bf16
constexpr int HLT_G_BYTES = 128 * 128;   // [128 windows][32 fp32] one piece of G_t
constexpr int HLT_STAGES = 3;
constexpr int HLT_BARS = HLT_CHAINS * (2 + 2 * HLT_STAGES);
constexpr int HLT_SMEM_BYTES =
    2 * HLT_W_BYTES + HLT_CHAINS * (2 * HLT_A_BYTES + HLT_STAGES * HLT_G_BYTES) + 8 * HLT_BARS + 64 + 1024;

// one LSTM cell: pre-activations (i, f, g, o), cell state c (updated), returns h
__device__ __forceinline__ float hlt_cell(float ai, float af, float ag, float ao, float& c) {
    constexpr float kL2E = 1.4426950408889634f;
    const float ei = ex2_approx(-kL2E * ai);                          // e^-i
    const float ef = ex2_approx(-kL2E * af);
    const float eg = ex2_approx(fminf(2.f * kL2E * ag, 60.f));        // e^2g, kept finite: (eg - 1) * rcp(inf) = 0
    const float eo = ex2_approx(-kL2E * ao);
    const float ig = (eg - 1.f) * rcp_approx((1.f + ei) * (eg + 1.f)); // sigmoid(i) tanh(g)
    c = fmaf(rcp_approx(1.f + ef), c, ig);                             // sigmoid(f) c + ...
    const float ec = ex2_approx(fminf(2.f * kL2E * c, 60.f));
    return (ec - 1.f) * rcp_approx((1.f + eo) * (ec + 1.f));          // sigmoid(o) tanh(c)
}

// tmap_gf / tmap_gr: G of the forward / reverse direction as fp32 [steps][windows][256] (column = 4 * unit + gate; the
//         forward array holds t = 0 .. r - 1, the reverse one t = l .. T - 1), box {32, 128, 1}, SWIZZLE_128B, windows
//         past the end read as zero
// whh:    [2 dir][hi, lo][256 columns (4 * unit + gate)][64 k] bf16, row-major
// Hout:   [windows, r - l, 128] fp32 = (fwd 64 | rev 64) of the steps t in [l, r)  ([r - l, windows, 128] if hout_t_major)
__global__ void __launch_bounds__(HLT_THREADS, 1)
head_lstm_tc_kernel(const __grid_constant__ CUtensorMap tmap_gf, const __grid_constant__ CUtensorMap tmap_gr,
                    const __nv_bfloat16* __restrict__ whh, int windows, int T, int l, int r, int hout_t_major,
                    float* __restrict__ Hout) {
    extern __shared__ uint8_t hlt_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(hlt_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + HLT_W_BYTES;
    uint8_t* a_base = smem + 2 * HLT_W_BYTES;                       // [chain][hi, lo] A tiles
    uint8_t* g_base = a_base + HLT_CHAINS * 2 * HLT_A_BYTES;        // [chain][stage] pieces of G
    uint64_t* bars = reinterpret_cast<uint64_t*>(g_base + HLT_CHAINS * HLT_STAGES * HLT_G_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + HLT_BARS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int tiles = (windows + 127) >> 7;
    const int steps = dir == 0 ? r : T - l;  // the steps that can reach the kept frames [l, r)
    const int n_keep = r - l;

    {   // recurrent weights of this direction -> shared memory, swizzled the way a SWIZZLE_128B TMA box would land
        const uint4* src = reinterpret_cast<const uint4*>(whh + (size_t)dir * 2 * 256 * 64);
        for (int i = threadIdx.x; i < 2 * 256 * 8; i += HLT_THREADS) {
            const int part = i >> 11, n = (i >> 3) & 255, c = i & 7;
            *reinterpret_cast<uint4*>(smem + part * HLT_W_BYTES + n * 128 + ((c ^ (n & 7)) << 4)) = __ldg(src + i);
        }
    }
    if (threadIdx.x == 0) {
        for (int ch = 0; ch < HLT_CHAINS; ++ch) {
            uint64_t* b = bars + ch * (2 + 2 * HLT_STAGES);
            mbar_init(&b[0], 8 * 32);  // h written: every activation thread of the chain
            mbar_init(&b[1], 1);       // gates ready: tcgen05.commit
            for (int st = 0; st < HLT_STAGES; ++st) {
                mbar_init(&b[2 + st], 1);               // piece landed (TMA transaction bytes)
                mbar_init(&b[2 + HLT_STAGES + st], 8);  // piece consumed: one arrival per activation warp
            }
        }
        fence_mbar_init();
        tma_prefetch_desc(&tmap_gf);
        tma_prefetch_desc(&tmap_gr);
    }
    if (warp == HLT_ACT_WARPS) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    fence_proxy_async();  // the weight tiles were written through the generic proxy; the MMAs read through the async one
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // every role of a chain walks the same tiles: pair index pi -> tile 2 * pi + chain
    const int chain = warp < HLT_ACT_WARPS ? warp >> 3 : (warp - HLT_ACT_WARPS) & 1;
    uint64_t* cb = bars + chain * (2 + 2 * HLT_STAGES);
    uint64_t* h_ready = &cb[0];
    uint64_t* gates_ready = &cb[1];
    uint64_t* g_full = &cb[2];
    uint64_t* g_empty = &cb[2 + HLT_STAGES];
    uint8_t* a_hi = a_base + chain * 2 * HLT_A_BYTES;
    uint8_t* a_lo = a_hi + HLT_A_BYTES;
    uint8_t* g_ring = g_base + chain * HLT_STAGES * HLT_G_BYTES;

    if (warp < HLT_ACT_WARPS) {
        // ------------------------------------------------------------------------------------ activation warps
        const int q = warp & 3, uh = (warp >> 2) & 1;
        const int row = q * 32 + lane;
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + 256 * chain + 16 * uh;
        uint8_t* a_row_hi = a_hi + row * 128;
        uint8_t* a_row_lo = a_lo + row * 128;
        const int sw = row & 7;
        uint32_t gphase = 0, stage = 0, fphase = 0;
        for (int pi = blockIdx.x; 2 * pi + chain < tiles; pi += gridDim.x) {
            const int win = (2 * pi + chain) * 128 + row;
            const bool valid = win < windows;
            float c[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) c[i] = 0.f;
            for (int s = 0; s < steps; ++s) {
                const int t = dir == 0 ? s : T - 1 - s;
                if (s > 0) {
                    mbar_wait(gates_ready, gphase);
                    gphase ^= 1;
                    tc_fence_after();
                }
                const bool keep = valid && t >= l && t < r;
                float* ho = Hout + (hout_t_major ? (size_t)(t - l) * windows + win : (size_t)win * n_keep + (t - l)) * 128 +
                            dir * 64 + uh * 4;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    uint32_t acc[16];
                    if (s > 0) tmem_ld_32x16(t_row + ch * 32, acc);
                    mbar_wait(&g_full[stage], fphase);
                    const uint8_t* grow = g_ring + stage * HLT_G_BYTES + row * 128;
                    float4 g[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) g[j] = *reinterpret_cast<const float4*>(grow + (((4 * uh + j) ^ sw) << 4));
                    if (s > 0) {
                        tmem_ld_wait();
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] = 0u;  // h_{-1} = 0
                    }
                    float hv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        hv[j] = hlt_cell(g[j].x + __uint_as_float(acc[4 * j]), g[j].y + __uint_as_float(acc[4 * j + 1]),
                                         g[j].z + __uint_as_float(acc[4 * j + 2]), g[j].w + __uint_as_float(acc[4 * j + 3]),
                                         c[ch * 4 + j]);
                    // the piece is in registers (its values have been used): hand the ring slot back
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&g_empty[stage]);
                    if (++stage == HLT_STAGES) { stage = 0; fphase ^= 1; }
                    if (s + 1 < steps) {
                        // h as bf16 hi / lo into the K-major A tiles: units 8 ch + 4 uh .. + 3 = 8 bytes of 16-byte chunk ch
                        const uint32_t hi01 = pack_bf16(hv[0], hv[1]), hi23 = pack_bf16(hv[2], hv[3]);
                        const float2 f01 = unpack_bf16(hi01), f23 = unpack_bf16(hi23);
                        const uint32_t lo01 = pack_bf16(hv[0] - f01.x, hv[1] - f01.y);
                        const uint32_t lo23 = pack_bf16(hv[2] - f23.x, hv[3] - f23.y);
                        const int off = ((ch ^ sw) << 4) + (uh << 3);
                        *reinterpret_cast<uint2*>(a_row_hi + off) = make_uint2(hi01, hi23);
                        *reinterpret_cast<uint2*>(a_row_lo + off) = make_uint2(lo01, lo23);
                    }
                    if (keep) *reinterpret_cast<float4*>(ho + ch * 8) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                }
                if (s + 1 < steps) {
                    tc_fence_before();    // this thread's TMEM reads are done before the next MMA overwrites the gates
                    fence_proxy_async();  // ... and its h is visible to the tensor core's shared-memory reads
                    mbar_arrive(h_ready);
                }
            }
        }
    } else if (warp < HLT_ACT_WARPS + HLT_CHAINS) {
        // ------------------------------------------------------------------------------------ MMA issuer of the chain
        constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
        const uint64_t d_ahi = umma_desc_sw128(smem_u32(a_hi)), d_alo = umma_desc_sw128(smem_u32(a_lo));
        const uint64_t d_whi = umma_desc_sw128(smem_u32(w_hi)), d_wlo = umma_desc_sw128(smem_u32(w_lo));
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0) + 256 * chain;
        uint32_t hphase = 0;
        for (int pi = blockIdx.x; 2 * pi + chain < tiles; pi += gridDim.x) {
            for (int s = 1; s < steps; ++s) {
                mbar_wait(h_ready, hphase);
                hphase ^= 1;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_u, d_ahi + 2 * k, d_whi + 2 * k, idesc, k != 0);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_u, d_alo + 2 * k, d_whi + 2 * k, idesc, 1);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_u, d_ahi + 2 * k, d_wlo + 2 * k, idesc, 1);
                    umma_commit(gates_ready);
                }
                __syncwarp();
            }
        }
    } else {
        // ------------------------------------------------------------------------------------ TMA producer of the chain
        uint32_t stage = 0, ephase = 1;  // a fresh barrier passes a wait on parity 1
        const CUtensorMap* tm = dir ? &tmap_gr : &tmap_gf;
        const int t0 = dir ? l : 0;      // first step held by this direction's array
        for (int pi = blockIdx.x; 2 * pi + chain < tiles; pi += gridDim.x) {
            const int w0 = (2 * pi + chain) * 128;
            for (int s = 0; s < steps; ++s) {
                const int t = dir == 0 ? s : T - 1 - s;
                for (int ch = 0; ch < 8; ++ch) {
                    mbar_wait(&g_empty[stage], ephase);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&g_full[stage], HLT_G_BYTES);
                        tma_load_3d(g_ring + stage * HLT_G_BYTES, tm, &g_full[stage], ch * 32, w0, t - t0);
                    }
                    __syncwarp();
                    if (++stage == HLT_STAGES) { stage = 0; ephase ^= 1; }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == HLT_ACT_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace cbas
