// Fused ViT self-attention on tcgen05 for frames whose padded token count fits one MMA tile (T <= 256):
// one persistent CTA per SM walks (frame, head) work items; per item
//
//   TMA      : Q (two 128-row tiles), K and V ([TK,64] each) of the head straight out of the fused QKV
//              activation into a double-buffered shared-memory set (SWIZZLE_128B), one item ahead.
//   tcgen05  : S_mt = Q_mt K^T  (128 x TK x 64, bf16 operands, fp32 in TMEM) for both query tiles;
//              O_mt = P_mt V    (128 x 64 x TK, f16 operands) with P read FROM TMEM (f16 pairs written over S
//              by the softmax warps) and V consumed MN-major exactly as TMA laid it down - no transposes, no P
//              in shared memory.  V arrives as IEEE f16 (the QKV GEMM stores that third of its output as f16,
//              EPI_BIAS_BF16_VF16) so the probabilities can be f16 (11 mantissa bits) rather than bf16.
//   softmax  : four warpgroups, two per query tile; a thread owns one query row and HALF of its keys (TMEM lane
//              access is tied to warp_id % 4, so two warps share each lane quarter).  Pass 1 reads S for the
//              partial row max (exchanged with the partner thread through shared memory), pass 2 re-reads S,
//              exponentiates (exp2, scale folded), packs f16 pairs and stores them over consumed columns of S.
//   epilogue : each of the two threads of a row scales 32 of the 64 output columns by 1 / (sum_a + sum_b)
//              -> bf16 -> a swizzled shared-memory staging row; one 3-D TMA store per query tile writes
//              [rows, 64] at column head*64, clipped at the end of the frame.
//
//   RoPE     : attention prologue, in shared memory, by FOUR DEDICATED WARPS: as soon as Q and K of an item have
//              landed they rotate the patch-token rows in place (rotate-half form, fp32 math, cos/sin held as
//              half2 pairs in shared memory, built once per CTA from the fp32 tables), fence them to the async
//              proxy and release the MMA warps through an mbarrier.  Q/K buffers are recycled as soon as the S
//              MMAs that read them retire (V only after PV), so loads and rotation run a full item ahead of the
//              tensor core.  Rows of prefix tokens and rows past the frame are left alone.  (Pass null tables to
//              skip RoPE.)
//   overlap  : each query tile has its OWN MMA-issuing warp and its own barrier chain
//              (S -> softmax -> P V -> epilogue -> next S); tile 1 is started half an item late, so while one
//              tile's warps are in their softmax the other tile's MMAs, TMEM drain and stores run.  The MMA and TMA
//              warps run warp-uniform code with one elected lane (descriptors in uniform registers).
//
// TMEM map (512 columns): query tile mt owns columns [256*mt, 256*mt+256): S at +0..TK; P (f16 pairs) of the
// first key half at +0..CA/2 and of the second half at +CA..+CA+(TK-CA)/2; O at +192..+256 (written only after
// the softmax has consumed S, read back by the same warps).
// Reference semantics: HF modeling_dinov3_vit.py:316-329 (SDPA, scale 1/8, no mask, non-causal).
#pragma once
#include "ptx.cuh"

namespace cbas {

// warps 0-7 softmax/epilogue of query tile 0 (0-3 first half of the keys, 4-7 second half), 8-15 of tile 1; then
// warp 16 TMA producer (+ TMEM allocation), 17 / 18 MMA issuers of query tile 0 / 1, 19-22 RoPE rotation
constexpr int ATC_THREADS = 736;
constexpr int ATC_SOFTMAX_WARPS = 16;
constexpr int ATC_PRODUCER_WARP = 16, ATC_MMA_WARP0 = 17, ATC_ROT_WARP0 = 19, ATC_ROT_WARPS = 4;
constexpr int ATC_TRACE_SLOTS = 24;
#ifndef ATC_POLY_PAIRS
#define ATC_POLY_PAIRS 2  // of the 8 probability pairs per 16 columns, how many are exponentiated on the FMA pipe
#endif
#ifndef ATC_DIRECT_STORE
#define ATC_DIRECT_STORE 0  // 1: the epilogue writes O straight to global memory (64 B per thread) instead of staging a TMA store
#endif
constexpr int ATC_XCHG_BYTES = 2 * 2 * 2 * 128 * 4;  // [max|sum][tile][half][row] floats
constexpr int ATC_O_COL = 192;

struct AttnTcParams {
    const __nv_bfloat16* qkv;  // only for documentation; loads go through the tensor maps
    __nv_bfloat16* out;        // [frames*T, D]
    int frames, heads, T, TK, D;
    float scale_log2;          // head_dim^-0.5 * log2(e)
    const float* rope_cos;     // [T - prefix, 32] fp32 or null (no RoPE in this kernel)
    const float* rope_sin;
    int prefix;
    long long* trace;          // optional [64][ATC_TRACE_SLOTS] clock64 stamps of CTA 0 (profiling aid; null in production)
    int reverse;               // walk the frames from the last to the first (attention_tc_kernel only; see GemmParams::reverse)
};

__host__ __device__ inline int atc_set_bytes(int TK) { return 2 * 128 * 128 + 2 * TK * 128; }
__host__ __device__ inline int atc_rope_bytes(int T, int prefix) { return ((T - prefix) * 32 * 4 + 127) & ~127; }
__host__ __device__ inline int atc_stage_bytes(int T) { return (T * 128 + 1023) & ~1023; }  // O rows, 128 B each
__host__ __device__ inline int atc_smem_bytes(int TK, int T, int prefix, bool rope) {
    return 2 * atc_set_bytes(TK) + atc_stage_bytes(T) + (rope ? atc_rope_bytes(T, prefix) : 0) + ATC_XCHG_BYTES + 1024 +
           256;
}

// Rotate the patch-token rows of one [rows,64] bf16 tile in place.  `u` enumerates (row, chunk pair): a warp
// covers 8 consecutive rows x 4 chunk pairs, so the 16-byte shared-memory accesses are conflict-free under
// the 128-byte swizzle.  cs2 holds (cos, sin) half2 pairs, 32 per patch token.
__device__ __forceinline__ void atc_rope_unit(uint8_t* tile, int row, int c, const __half2* cs2) {
    uint8_t* r = tile + row * 128;
    const int sw = row & 7;
    uint4* plo = reinterpret_cast<uint4*>(r + ((c ^ sw) << 4));
    uint4* phi = reinterpret_cast<uint4*>(r + (((c + 4) ^ sw) << 4));
    uint4 lo = *plo, hi = *phi;
    const uint4 t0 = *reinterpret_cast<const uint4*>(cs2 + 8 * c);      // 4 (cos,sin) pairs
    const uint4 t1 = *reinterpret_cast<const uint4*>(cs2 + 8 * c + 4);  // 4 more
    const uint32_t tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
    uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = unpack_bf16(l[i]), b = unpack_bf16(h[i]);
        const float2 cs0 = __half22float2(*reinterpret_cast<const __half2*>(&tw[2 * i]));
        const float2 cs1 = __half22float2(*reinterpret_cast<const __half2*>(&tw[2 * i + 1]));
        l[i] = pack_bf16(a.x * cs0.x - b.x * cs0.y, a.y * cs1.x - b.y * cs1.y);
        h[i] = pack_bf16(b.x * cs0.x + a.x * cs0.y, b.y * cs1.x + a.y * cs1.y);
    }
    *plo = lo;
    *phi = hi;
}


// The same rotation for an f16 [rows,64] tile (the T <= 256 kernel, whose q and k arrive as f16): packed HFMA2 on f16
// pairs, cos / sin as f16 from a shared table laid out [pos][chunk pair][8 cos | 8 sin].  `n` (row, chunk pair) units
// starting at u0 with stride `step` are in flight together so their shared-memory round trips overlap.
template <int N>
__device__ __forceinline__ void atc_rope_units_f16(uint8_t* q_tile, uint8_t* k_tile, int u0, int step, int units, int nrot,
                                                   int prefix, const __half* tab) {
    uint4 lo[N], hi[N], cs[N], sn[N];
    uint4* plo[N];
    uint4* phi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int u = u0 + i * step;
        const int ridx = u < units ? u >> 2 : 0, c4 = u & 3;
        const int isk = ridx >= nrot;
        const int tok = prefix + ridx - (isk ? nrot : 0);
        uint8_t* r = (isk ? k_tile : q_tile) + tok * 128;
        const int sw = tok & 7;
        plo[i] = reinterpret_cast<uint4*>(r + ((c4 ^ sw) << 4));
        phi[i] = reinterpret_cast<uint4*>(r + (((c4 + 4) ^ sw) << 4));
        const __half* tp = tab + (tok - prefix) * 64 + c4 * 16;
        lo[i] = *plo[i];
        hi[i] = *phi[i];
        cs[i] = *reinterpret_cast<const uint4*>(tp);
        sn[i] = *reinterpret_cast<const uint4*>(tp + 8);
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        __half2* l = reinterpret_cast<__half2*>(&lo[i]);
        __half2* h = reinterpret_cast<__half2*>(&hi[i]);
        const __half2* cc = reinterpret_cast<const __half2*>(&cs[i]);
        const __half2* ss = reinterpret_cast<const __half2*>(&sn[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __half2 a = l[j], b = h[j];
            l[j] = __hfma2(a, cc[j], __hneg2(__hmul2(b, ss[j])));  // a cos - b sin
            h[j] = __hfma2(b, cc[j], __hmul2(a, ss[j]));           // b cos + a sin
        }
        if (u0 + i * step < units) {
            *plo[i] = lo[i];
            *phi[i] = hi[i];
        }
    }
}

// packed fp32 pairs (FFMA2 / FADD2: one issue slot for two lanes of arithmetic)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tmov.b64 rc, {%6,%7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2,%3};\n\tmov.b64 rb, {%4,%5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
// 2^x for a pair, x <= 0, on the FMA pipe (ex2_poly3 of ptx.cuh with packed arithmetic)
__device__ __forceinline__ float2 ex2_poly3_pair(float2 x) {
    x.x = fmaxf(x.x, -30.f);
    x.y = fmaxf(x.y, -30.f);
    const float2 t = fadd2(x, make_float2(12582912.f, 12582912.f));
    const float2 r = fadd2(t, make_float2(-12582912.f, -12582912.f));
    const float2 f = fadd2(x, make_float2(-r.x, -r.y));
    float2 p = ffma2(f, make_float2(0.05517089366912842f, 0.05517089366912842f),
                     make_float2(0.24261131882667542f, 0.24261131882667542f));
    p = ffma2(p, f, make_float2(0.6932610273361206f, 0.6932610273361206f));
    p = ffma2(p, f, make_float2(0.9999280571937561f, 0.9999280571937561f));
    return make_float2(__uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23)),
                       __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23)));
}

// Softmax of one thread's share [CB, CE) of its query row (a multiple of 16 columns; the row's other share belongs to
// the partner thread of the same TMEM lane): exact row max (pass 1, exchanged with the partner through shared memory),
// then p = exp2(s*c - max*c) as f16 pairs written over columns of S this thread has already consumed (pass 2: pair j
// of 16-column chunk k goes to column CB + 8 k + j).  Both passes keep the TMEM load of the next chunk in flight
// while the current one is worked on.  Returns the partial row sum.
// NP threads share a row; thread `part` publishes its partial max at xmax0 + 512 * part (shared-space address of a
// [NP][128] float array, already offset to this row) and the NP warps of the lane quarter meet at named barrier bar_id.
// `turn` (may be null): mbarrier to pass, at parity `turn_parity`, between the max pass and the exponential pass - the
// two query tiles of an item take turns on the MUFU-heavy pass only; the max pass (TMEM loads and FMNMX) of one tile runs
// under the other tile's exponentials.
template <int TK, int CB, int CE, int NP>
__device__ __forceinline__ float atc_softmax_range(uint32_t t_row, int T, float c, uint32_t xmax0, int part, int bar_id,
                                                   uint64_t* turn, uint32_t turn_parity) {
    constexpr int W = CE - CB;
    constexpr bool kEdge = CE == TK;  // the last 16 columns of this share may lie past the frame
    float mx = -INFINITY;
    uint32_t cur[16];
    if constexpr (W > 0) {
        // pass 1: 16 columns per TMEM load, the next load in flight while the current chunk is reduced
        constexpr int NC1 = W / 16;
        uint32_t buf[2][16];
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
        tmem_ld_32x16(t_row + CB, buf[0]);
#pragma unroll
        for (int k = 0; k < NC1; ++k) {
            tmem_ld_wait();
            if (k + 1 < NC1) tmem_ld_32x16(t_row + CB + 16 * (k + 1), buf[(k + 1) & 1]);
            uint32_t* v = buf[k & 1];
            if (kEdge && k == NC1 - 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (CB + 16 * k + j >= T) v[j] = 0xff800000u;  // -inf
            }
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(v[4]), __uint_as_float(v[5])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7])));
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[8]), __uint_as_float(v[9])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[10]), __uint_as_float(v[11])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(v[12]), __uint_as_float(v[13])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(v[14]), __uint_as_float(v[15])));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        tmem_ld_32x16(t_row + CB, cur);  // pass 2's first chunk flies during the exchange
    }
    st_shared_f32(xmax0 + 512 * part, mx);
    named_bar_sync(bar_id, 32 * NP);  // only the partner warps: same rows, the other shares of the keys
#pragma unroll
    for (int q = 0; q < NP; ++q) mx = fmaxf(mx, ld_shared_f32(xmax0 + 512 * q));
    if (turn) mbar_wait(turn, turn_parity);
    float sum = 0.f;
    if constexpr (W > 0) {
        constexpr int NC = W / 16;
        const float mc = mx * c;
        const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
        float2 s0 = make_float2(0.f, 0.f), s1 = s0;
        uint32_t nxt[16];
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            tmem_ld_wait();
            uint32_t* v = (k & 1) ? nxt : cur;
            if (k + 1 < NC) tmem_ld_32x16(t_row + CB + 16 * (k + 1), (k & 1) ? cur : nxt);
            const bool edge = kEdge && k == NC - 1;
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
                if (edge) {
                    if (CB + 16 * k + 2 * j >= T) a = -INFINITY;
                    if (CB + 16 * k + 2 * j + 1 >= T) b = -INFINITY;
                }
                const float2 x = ffma2(make_float2(a, b), c2, nmc2);
                float2 e;
                // the MUFU (16 ex2/clk/SM) is the busiest unit of this kernel: ATC_POLY_PAIRS pairs of every 8 go
                // through the FMA-pipe polynomial; masked keys stay on the MUFU, where ex2(-inf) is an exact zero
                if (j < 8 - ATC_POLY_PAIRS || edge) e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
                else e = ex2_poly3_pair(x);
                if (j & 1) s1 = fadd2(s1, e); else s0 = fadd2(s0, e);
                pk[j] = pack_f16(e.x, e.y);
            }
            tmem_st_32x8(t_row + CB + 8 * k, pk);
        }
        const float2 st = fadd2(s0, s1);
        sum = st.x + st.y;
        tmem_st_wait();
    }
    return sum;
}


// The same for shares of at most 64 columns (four threads per row): the whole share is read from TMEM ONCE and stays
// in registers between the max and the exponentials - no second TMEM pass, no load latency inside the exponential
// loop, and every pair is independent work for the scheduler.
template <int TK, int CB, int CE, int NP>
__device__ __forceinline__ float atc_softmax_range_resident(uint32_t t_row, int T, float c, uint32_t xmax0, int part,
                                                            int bar_id) {
    constexpr int W = CE - CB;
    static_assert(W <= 64 && W % 16 == 0, "register-resident share");
    constexpr bool kEdge = CE == TK;  // the last 16 columns of this share may lie past the frame
    float mx = -INFINITY;
    uint32_t v[W > 0 ? W : 16];
    if constexpr (W > 0) {
#pragma unroll
        for (int k = 0; k < W / 16; ++k) tmem_ld_32x16(t_row + CB + 16 * k, *reinterpret_cast<uint32_t(*)[16]>(&v[16 * k]));
        tmem_ld_wait();
        if (kEdge) {
#pragma unroll
            for (int j = W - 16; j < W; ++j)
                if (CB + j >= T) v[j] = 0xff800000u;  // -inf: exp2 gives 0 on the MUFU, 2^-30 -> f16 zero on the FMA pipe
        }
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
        for (int j = 0; j < W; j += 8) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    }
    st_shared_f32(xmax0 + 512 * part, mx);
    named_bar_sync(bar_id, 32 * NP);  // only the partner warps: same rows, the other shares of the keys
#pragma unroll
    for (int q = 0; q < NP; ++q) mx = fmaxf(mx, ld_shared_f32(xmax0 + 512 * q));
    float sum = 0.f;
    if constexpr (W > 0) {
        const float mc = mx * c;
        const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
        float2 s0 = make_float2(0.f, 0.f), s1 = s0;
#pragma unroll
        for (int k = 0; k < W / 16; ++k) {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 x = ffma2(make_float2(__uint_as_float(v[16 * k + 2 * j]), __uint_as_float(v[16 * k + 2 * j + 1])), c2, nmc2);
                float2 e;
                if (j < 8 - ATC_POLY_PAIRS) e = make_float2(ex2_approx(x.x), ex2_approx(x.y));
                else e = ex2_poly3_pair(x);
                if (j & 1) s1 = fadd2(s1, e); else s0 = fadd2(s0, e);
                pk[j] = pack_f16(e.x, e.y);
            }
            tmem_st_32x8(t_row + CB + 8 * k, pk);
        }
        const float2 st = fadd2(s0, s1);
        sum = st.x + st.y;
        tmem_st_wait();
    }
    return sum;
}

#define ATC_STAMP(slot)                                                                       \
    do {                                                                                      \
        if (p.trace && blockIdx.x == 0 && it < 64) p.trace[it * ATC_TRACE_SLOTS + (slot)] = clock64();     \
    } while (0)

template <int TK>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,   // box {64, 128} over qkv [M, 3D]
                    const __grid_constant__ CUtensorMap tmap_kv,  // box {64, TK}
                    const __grid_constant__ CUtensorMap tmap_o,   // box {64, min(T,128), 1} over out [frames][T][D]
                    const __grid_constant__ CUtensorMap tmap_o1,  // box {64, T-128, 1}: the rows of query tile 1
                    const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    static_assert(TK % 16 == 0 && TK >= 16 && TK <= 256, "padded key count");
    const int T = p.T;
    constexpr int set_bytes = 2 * 128 * 128 + 2 * TK * 128;
    const bool rope = p.rope_cos != nullptr;
    uint8_t* ostage = smem + 2 * set_bytes;  // [T][128 B] output rows of the current item, swizzled per 128-row tile
    __half* rope_tab = reinterpret_cast<__half*>(ostage + atc_stage_bytes(T));  // [T - prefix][4 chunk pairs][8 cos | 8 sin]
    float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(rope_tab) + (rope ? atc_rope_bytes(T, p.prefix) : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xchg) + ATC_XCHG_BYTES);
    uint64_t* qk_full = bars;        // [2] TMA -> rotation warps (or MMA when there is no RoPE): Q tiles + K landed
    uint64_t* qk_empty = bars + 2;   // [2] both MMA warps -> TMA: the S MMAs that read this Q/K set retired
    uint64_t* v_full = bars + 4;     // [2] TMA -> MMA
    uint64_t* v_empty = bars + 6;    // [2] both MMA warps -> TMA: the PV MMAs retired
    uint64_t* s_full = bars + 8;     // [2] per query tile: MMA -> softmax
    uint64_t* p_full = bars + 10;    // [2] softmax -> MMA
    uint64_t* o_full = bars + 12;    // [2] MMA -> epilogue
    uint64_t* o_empty = bars + 14;   // [2] epilogue -> MMA (TMEM half free again)
    uint64_t* qk_ready = bars + 16;  // [2] rotation warps -> MMA: Q and K of this set are rotated
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_items = p.frames * p.heads;

    if (warp == ATC_PRODUCER_WARP && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv);
        tma_prefetch_desc(&tmap_o);
        tma_prefetch_desc(&tmap_o1);
    }
    if (warp == ATC_MMA_WARP0 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&qk_empty[i], 2);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 2);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 8);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 8);
            mbar_init(&qk_ready[i], ATC_ROT_WARPS);
        }
        fence_mbar_init();
    }
    if (rope) {
        const int n = (T - p.prefix) * 32;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int pos = i >> 5, d = i & 31;
            __half* t = rope_tab + pos * 64 + (d >> 3) * 16 + (d & 7);
            t[0] = __float2half_rn(__ldg(p.rope_cos + i));
            t[8] = __float2half_rn(__ldg(p.rope_sin + i));
        }
    }
    if (warp == ATC_PRODUCER_WARP) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();     // the QKV GEMM's output is read from here on (the prologue above overlapped its tail)
    pdl_trigger();

    if (warp == ATC_PRODUCER_WARP) {
        // -------------------------------------------------------------------- TMA producer (one elected lane issues)
        int it = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
            const int b = it & 1;
            const int f = p.reverse ? p.frames - 1 - w / p.heads : w / p.heads, h = w % p.heads;
            uint8_t* set = smem + b * set_bytes;
            const int row0 = f * T;
            mbar_wait(&qk_empty[b], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&qk_full[b], 32768 + TK * 128);
                tma_load_2d(set, &tmap_q, &qk_full[b], h * 64, row0);
                tma_load_2d(set + 16384, &tmap_q, &qk_full[b], h * 64, row0 + 128);
                tma_load_2d(set + 32768, &tmap_kv, &qk_full[b], p.D + h * 64, row0);
            }
            __syncwarp();
            mbar_wait(&v_empty[b], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&v_full[b], TK * 128);
                tma_load_2d(set + 32768 + TK * 128, &tmap_kv, &v_full[b], 2 * p.D + h * 64, row0);
            }
            __syncwarp();
        }
    } else if (warp == ATC_MMA_WARP0 || warp == ATC_MMA_WARP0 + 1) {
        // -------------------------------------------------------------------- MMA issuer of query tile mt
        // The whole warp walks the loop (uniform control flow, operands in uniform registers); one elected lane
        // issues.  tcgen05.commit tracks the MMAs of the issuing thread, so the same lane must issue both: elect.sync
        // picks the same lane every time for a full mask.
        const int mt = warp - ATC_MMA_WARP0;
        constexpr uint32_t idesc_s = umma_idesc_f16(128, TK);  // q and k arrive as f16
        constexpr uint32_t idesc_o = umma_idesc_f16_bmn(128, 64);
        const uint32_t t_tile = __shfl_sync(0xffffffffu, tmem_base, 0) + 256 * mt;
        constexpr int nk = TK >> 4;
        constexpr int ka = (nk + 1) / 2;  // k-steps whose keys belong to the first half of the row
        const uint32_t smem_base = smem_u32(smem);
        int it = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
            const int b = it & 1;
            const uint32_t set = smem_base + b * set_bytes;
            if (lane == 0) ATC_STAMP(3 * mt);
            mbar_wait(rope ? &qk_ready[b] : &qk_full[b], (it >> 1) & 1);
            mbar_wait(&o_empty[mt], (it & 1) ^ 1);  // previous item's O (inside this S region) was drained
            tc_fence_after();
            if (lane == 0 && mt == 0) ATC_STAMP(21);
            const uint64_t dk = umma_desc_sw128(set + 32768);
            const uint64_t dq = umma_desc_sw128(set + mt * 16384);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ss(t_tile, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[mt]);
                umma_commit(&qk_empty[b]);
            }
            __syncwarp();
            if (lane == 0) ATC_STAMP(3 * mt + 1);
            const uint64_t dv = umma_desc_sw128_mn(set + 32768 + TK * 128);
            mbar_wait(&v_full[b], (it >> 1) & 1);
            mbar_wait(&p_full[mt], it & 1);
            tc_fence_after();
            if (lane == 0 && mt == 0) ATC_STAMP(22);
            if (elect_one()) {
                // 16 keys per MMA: 8 TMEM columns of P, 2048 B of V.  P of keys [0,CA) sits at columns [0,CA/2); P of
                // keys [CA,TK) at [CA, CA+(TK-CA)/2): each softmax thread overwrites only columns of S that it has
                // itself already consumed
#pragma unroll
                for (int k = 0; k < nk; ++k) {
                    const uint32_t pcol = 8 * k + (k >= ka ? 8 * ka : 0);
                    umma_bf16_ts(t_tile + ATC_O_COL, t_tile + pcol, dv + 128 * k, idesc_o, k != 0);
                }
                umma_commit(&o_full[mt]);
                umma_commit(&v_empty[b]);
            }
            __syncwarp();
            if (lane == 0) ATC_STAMP(3 * mt + 2);
        }
    } else if (warp >= ATC_ROT_WARP0) {
        // ------------------------------------------------------------------- RoPE rotation warps
        if (rope) {
            const int rtid = threadIdx.x - ATC_ROT_WARP0 * 32;
            const int nrot = T - p.prefix;     // rotated rows of Q, and again of K
            const int units = 2 * nrot * 4;    // (row, chunk pair)
            int it = 0;
            for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
                const int b = it & 1;
                uint8_t* set = smem + b * set_bytes;
                mbar_wait(&qk_full[b], (it >> 1) & 1);
                if (rtid == 0) ATC_STAMP(16);
                // query tile 1 follows tile 0 in the set: a query row's offset is tok * 128 in both tiles
                for (int u0 = rtid; u0 < units; u0 += 4 * ATC_ROT_WARPS * 32)
                    atc_rope_units_f16<4>(set, set + 32768, u0, ATC_ROT_WARPS * 32, units, nrot, p.prefix, rope_tab);
                fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&qk_ready[b]);
                if (rtid == 0) ATC_STAMP(17);
            }
        }
    } else {
        // ------------------------------------------------------------------- softmax + epilogue warpgroups
        const int mt = warp >> 3;          // query tile
        const int half = (warp >> 2) & 1;  // which half of the keys (and of the output columns)
        const int quarter = warp & 3;
        const int rit = quarter * 32 + lane;      // row inside the tile
        const int row = mt * 128 + rit;           // query token inside the frame
        const bool warp_has_rows = (mt * 128 + quarter * 32) < T;
        const uint32_t t_row = tmem_base + 256 * mt + (uint32_t(quarter * 32) << 16);
        const float c = p.scale_log2;
        constexpr int CA = ((TK >> 4) + 1) / 2 * 16;  // keys [0,CA) for half 0, [CA,TK) for half 1 (multiples of 16)
        float* my_sum = xchg + ((1 * 2 + mt) * 2 + half) * 128 + rit;
        float* peer_sum = xchg + ((1 * 2 + mt) * 2 + (half ^ 1)) * 128 + rit;
        const bool stamper = (threadIdx.x & 255) == 0;  // first thread of each query tile
        const int sbase = 6 + 5 * mt;
        int it = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
            const int f = p.reverse ? p.frames - 1 - w / p.heads : w / p.heads, h = w % p.heads;
            // The two query tiles take turns in the softmax: tile 1 starts item i only when tile 0 has handed over its
            // P of item i, tile 0 starts item i+1 when tile 1 has handed over item i.  One tile's S / P V MMAs, TMEM drain
            // and stores then always fall into the other tile's exponentials instead of both tiles queueing for the
            // MUFU at the same time and idling together afterwards.  (No phase can be skipped: a tile cannot complete
            // its next P hand-over before every warp of the other tile has seen this one.)
            // (The turn is taken between the two passes of the softmax: the max pass needs no MUFU.)
            uint64_t* turn = mt == 1 ? &p_full[0] : (it > 0 ? &p_full[1] : nullptr);
            const uint32_t turn_parity = mt == 1 ? (it & 1) : ((it - 1) & 1);
            mbar_wait(&s_full[mt], it & 1);
            tc_fence_after();
            if (stamper) ATC_STAMP(sbase);
            float sum = 0.f;
            if (warp_has_rows) {  // the partner warp (same rows) takes the same branch
                const int bar_id = 1 + mt * 4 + quarter;
                const uint32_t xmax0 = smem_u32(xchg + ((0 * 2 + mt) * 2 + 0) * 128 + rit);
                if (half == 0) sum = atc_softmax_range<TK, 0, CA, 2>(t_row, T, c, xmax0, 0, bar_id, turn, turn_parity);
                else sum = atc_softmax_range<TK, CA, TK, 2>(t_row, T, c, xmax0, 1, bar_id, turn, turn_parity);
            } else if (turn) {
                mbar_wait(turn, turn_parity);
            }
            *my_sum = sum;  // read by the partner thread after o_full (ordered through the mbarrier chain)
            // The previous item's TMA store must have finished reading the staging rows before ANY warp of this tile
            // rewrites them; every warp's next staging follows o_full(it), hence P V(it), hence this hand-over - so the
            // storing thread checks here, thousands of cycles after it issued the store, instead of stalling its warp
            // (and with it the whole tile's next softmax) right after the issue.
            if ((warp & 7) == 0 && it > 0) {
                if (elect_one()) tma_wait_group_read<0>();
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[mt]);
            if (stamper) ATC_STAMP(sbase + 2);

            mbar_wait(&o_full[mt], it & 1);
            tc_fence_after();
            if (stamper) ATC_STAMP(sbase + 3);
            if (warp_has_rows) {
                uint32_t v0[32];
                tmem_ld_32x32(t_row + ATC_O_COL + 32 * half, v0);
                tmem_ld_wait();
                if (threadIdx.x == 0) ATC_STAMP(18);
                // read the partner's partial sum BEFORE releasing the tile: once o_empty completes the partner may
                // run ahead into the next item and overwrite it
                const float inv_sum = 1.0f / (sum + *peer_sum);
                // O is in registers: hand the TMEM half back before the stores
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[mt]);
                // stage this thread's 64 bytes, 128-byte swizzled like TMA wants it
                uint8_t* srow = ostage + row * 128;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 q;
                    q.x = pack_bf16(__uint_as_float(v0[j]) * inv_sum, __uint_as_float(v0[j + 1]) * inv_sum);
                    q.y = pack_bf16(__uint_as_float(v0[j + 2]) * inv_sum, __uint_as_float(v0[j + 3]) * inv_sum);
                    q.z = pack_bf16(__uint_as_float(v0[j + 4]) * inv_sum, __uint_as_float(v0[j + 5]) * inv_sum);
                    q.w = pack_bf16(__uint_as_float(v0[j + 6]) * inv_sum, __uint_as_float(v0[j + 7]) * inv_sum);
#if ATC_DIRECT_STORE
                    if (row < T)
                        reinterpret_cast<uint4*>(p.out + ((long long)f * T + row) * p.D + h * 64 + half * 32)[j >> 3] = q;
#else
                    if (row < T) *reinterpret_cast<uint4*>(srow + (((half * 4 + (j >> 3)) ^ (rit & 7)) << 4)) = q;
#endif
                }
            } else {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[mt]);
            }
#if !ATC_DIRECT_STORE
            // one coalesced TMA store per query tile
            fence_proxy_async();
            if (threadIdx.x == 0) ATC_STAMP(19);
            named_bar_sync(9 + mt, 256);
            if (threadIdx.x == 0) ATC_STAMP(20);
            if ((warp & 7) == 0 && mt * 128 < T) {  // first warp of the tile, uniform operands, elected lane
                if (elect_one()) {
                    if (mt) tma_store_3d(&tmap_o1, ostage + 16384, h * 64, 128, f);
                    else tma_store_3d(&tmap_o, ostage, h * 64, 0, f);
                    tma_commit_group();  // read-completion is checked before the next hand-over (see above)
                }
                __syncwarp();
            }
#endif
            if (stamper) ATC_STAMP(sbase + 4);
        }
        if ((warp & 7) == 0 && elect_one()) tma_wait_group<0>();  // this CTA's output stores have landed
    }

    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == ATC_PRODUCER_WARP) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace cbas
