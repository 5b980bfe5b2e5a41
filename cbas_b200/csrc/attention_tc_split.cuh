// tcgen05 self-attention for frames of 257..384 tokens (256-px frames: 261 tokens with 16-px patches, 329 with the
// 14-px patches of DINOv2-with-registers).  Same building blocks as attention_tc.cuh - TMA-fed Q/K/V, S = QK^T and
// O = PV on tcgen05 with P kept in TMEM, RoPE by dedicated warps, softmax with two threads per row, staged TMA
// store - but a 128-row score tile of such a frame no longer fits the 256 TMEM columns one pipeline owns, so the
// KEYS are split instead of the queries:
//
//   * the frame's keys are cut into two blocks (TK0 + TK1 = TK, each <= 192, multiples of 16);
//   * pipeline p (its own MMA-issuing warp, eight softmax warps, TMEM columns [256p, 256p+256)) computes, for the
//     current 128-row query tile, S_p = Q K_p^T, a softmax over ITS key block only (block max m_p, block sum l_p,
//     probabilities relative to m_p) and O_p = P_p V_p;
//   * the partial results are merged exactly:  w_p = 2^(c (m_p - max(m_0, m_1))),
//     O = (w_0 O_0 + w_1 O_1) / (w_0 l_0 + w_1 l_1)  - the standard split-K (flash-decoding) identity, so there
//     is no running rescale of O inside the loop.  Pipeline 1 runs half a tile behind pipeline 0 and ITS warps do
//     the merge, staging and store; pipeline 0 never waits for pipeline 1 (its O_0 simply stays in TMEM until the
//     merge has read it), so one pipeline's MMAs and drain fall into the other's softmax;
//   * the query tiles of a frame run one after the other against the resident K and V;
//   * LEFTOVER ROWS: when the last query tile holds at most 16 real rows (256-px frames with 16-px patches: 261 tokens
//     = 2 x 128 + 5) it does not go through the tensor-core pipeline at all - a full tile pass (S MMA, two softmax
//     passes, exchange, P V, merge) for five rows cost a third of the item.  The rotation warps, idle once Q and K
//     are rotated, run those rows through the warp-level mma.sync flash routine of attention.cuh (att_key_block: same
//     swizzled shared-memory rows, fp32 online softmax, f16 P and V) while the two full tiles are on tcgen05, and
//     stores them straight to global memory (the key blocks are dealt to the four rotation warps and their partial
//     (m, l, O) merged lane by lane through shared memory).  Q / K / V are handed back to the TMA producer only when it is done too.
//
// Shared memory holds ONE frame-head at a time (Q tiles + K + V, up to 134 KB) next to the output staging rows and
// the RoPE table, so the next item's loads start when the last S / PV MMAs of the current one retire; that bubble
// (about a fifth of an item) is the price of the larger frame.
// Reference semantics: HF modeling_dinov3_vit.py:316-329 / modeling_dinov2_with_registers.py eager_attention_forward
// (scale 1/8, no mask, non-causal).
#pragma once
#include "attention.cuh"
#include "attention_tc.cuh"

namespace cbas {

constexpr int ATS_LEFT_BYTES = ATC_ROT_WARPS * 36 * 32 * 4;  // leftover rows: per rotation warp [36 values][32 lanes] fp32
constexpr int ATS_XCHG_BYTES = (2 * 2 * 128 + 2 * 2 * 2 * 128 + 2 * 2 * 128) * 4;  // max [pipe][half][row]; per tile parity: sums [pipe][half][row], block max [pipe][row]

__host__ __device__ inline int ats_key_block0(int TK) { return ((TK + 31) / 32) * 16; }
__host__ __device__ inline int ats_smem_bytes(int TK, int T, int prefix, bool rope) {
    const int nq = (T + 127) / 128;
    const bool leftover = T - 128 * (nq - 1) <= 16;  // the last query tile's rows go through the rotation warps
    return nq * 16384 + 2 * TK * 128 + atc_stage_bytes(T) + (rope ? atc_rope_bytes(T, prefix) : 0) + ATS_XCHG_BYTES + 1024 +
           256 + (leftover ? ATS_LEFT_BYTES : 0);
}

__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_split_kernel(const __grid_constant__ CUtensorMap tmap_q,    // box {64, 128} over qkv [M, 3D]
                          const __grid_constant__ CUtensorMap tmap_kv0,  // box {64, TK0}
                          const __grid_constant__ CUtensorMap tmap_kv1,  // box {64, TK1}
                          const __grid_constant__ CUtensorMap tmap_o,    // box {64, 128, 1} over out [frames][T][D]
                          const __grid_constant__ CUtensorMap tmap_o1,   // box {64, T - 128 (nq-1), 1}: the last query tile
                          const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int TK = p.TK, T = p.T;
    const int nq = (T + 127) >> 7;
    const int TK0 = ats_key_block0(TK), TK1 = TK - TK0;
    const bool rope = p.rope_cos != nullptr;
    const int left_rows = T - 128 * (nq - 1);          // real rows of the last query tile
    const bool leftover = left_rows <= 16;              // ... few enough for one mma.sync m16 tile on a rotation warp
    const int nq_tc = leftover ? nq - 1 : nq;           // query tiles that go through the tcgen05 pipelines
    uint8_t* q_s = smem;                          // nq tiles of [128][128 B]
    uint8_t* k_s = smem + nq * 16384;             // [TK][128 B]
    uint8_t* v_s = k_s + TK * 128;                // [TK][128 B]
    uint8_t* ostage = v_s + TK * 128;             // [T][128 B] output rows of the current item
    __half2* rope_tab = reinterpret_cast<__half2*>(ostage + atc_stage_bytes(T));
    float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(rope_tab) + (rope ? atc_rope_bytes(T, p.prefix) : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xchg) + ATS_XCHG_BYTES);
    float* left_x = reinterpret_cast<float*>(bars + 32);  // [rotation warp][36][32], present in leftover mode only
    uint64_t* qk_full = bars;        // TMA -> rotation warps (or MMA): Q tiles + K landed
    uint64_t* qk_empty = bars + 1;   // both MMA warps -> TMA: the item's last S MMAs retired
    uint64_t* v_full = bars + 2;     // TMA -> MMA
    uint64_t* v_empty = bars + 3;    // both MMA warps -> TMA: the item's last PV MMAs retired
    uint64_t* qk_ready = bars + 4;   // rotation warps -> MMA
    uint64_t* s_full = bars + 5;     // [2] per pipeline: MMA -> softmax
    uint64_t* p_full = bars + 7;     // [2] softmax -> MMA
    uint64_t* o_full = bars + 9;     // [2] MMA -> epilogue
    uint64_t* o_empty = bars + 11;   // [2] merge (the 8 warps of pipeline 1 read both partial outputs) -> MMA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_items = p.frames * p.heads;

    if (warp == ATC_PRODUCER_WARP && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_kv0);
        tma_prefetch_desc(&tmap_kv1);
        tma_prefetch_desc(&tmap_o);
        tma_prefetch_desc(&tmap_o1);
    }
    if (warp == ATC_MMA_WARP0 && lane == 0) {
        mbar_init(qk_full, 1);
        mbar_init(qk_empty, leftover ? 3 : 2);  // both MMA warps (+ the leftover-rows warp)
        mbar_init(v_full, 1);
        mbar_init(v_empty, leftover ? 3 : 2);
        mbar_init(qk_ready, ATC_ROT_WARPS);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 8);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 8);
        }
        fence_mbar_init();
    }
    if (rope) {
        const int n = (T - p.prefix) * 32;
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            rope_tab[i] = __floats2half2_rn(__ldg(p.rope_cos + i), __ldg(p.rope_sin + i));
    }
    if (warp == ATC_PRODUCER_WARP) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == ATC_PRODUCER_WARP) {
        // -------------------------------------------------------------------- TMA producer (one elected lane issues)
        int it = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
            const int f = w / p.heads, h = w % p.heads;
            const int row0 = f * T;
            mbar_wait(qk_empty, (it & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(qk_full, nq * 16384 + TK * 128);
                for (int qt = 0; qt < nq; ++qt) tma_load_2d(q_s + qt * 16384, &tmap_q, qk_full, h * 64, row0 + qt * 128);
                tma_load_2d(k_s, &tmap_kv0, qk_full, p.D + h * 64, row0);
                tma_load_2d(k_s + TK0 * 128, &tmap_kv1, qk_full, p.D + h * 64, row0 + TK0);
            }
            __syncwarp();
            mbar_wait(v_empty, (it & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(v_full, TK * 128);
                tma_load_2d(v_s, &tmap_kv0, v_full, 2 * p.D + h * 64, row0);
                tma_load_2d(v_s + TK0 * 128, &tmap_kv1, v_full, 2 * p.D + h * 64, row0 + TK0);
            }
            __syncwarp();
        }
    } else if (warp == ATC_MMA_WARP0 || warp == ATC_MMA_WARP0 + 1) {
        // -------------------------------------------------------------------- MMA issuer of key block pp
        const int pp = warp - ATC_MMA_WARP0;
        const int TKp = pp ? TK1 : TK0;
        const uint32_t idesc_s = umma_idesc_bf16(128, TKp);
        const uint32_t idesc_o = umma_idesc_f16_bmn(128, 64);
        const uint32_t t_pipe = __shfl_sync(0xffffffffu, tmem_base, 0) + 256 * pp;
        const int nk = TKp >> 4;
        const int ka = (nk + 1) / 2;  // k-steps whose keys belong to the first half of the block
        const uint32_t q_u = smem_u32(q_s);
        const uint64_t dk = umma_desc_sw128(smem_u32(k_s) + pp * TK0 * 128);
        const uint64_t dv = umma_desc_sw128_mn(smem_u32(v_s) + pp * TK0 * 128);
        if (pp == 1) mbar_wait(&p_full[0], 0);  // half a tile late: the two pipelines' MMA and softmax phases interleave
        int it = 0, g = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
            mbar_wait(rope ? qk_ready : qk_full, it & 1);
            for (int qt = 0; qt < nq_tc; ++qt, ++g) {
                // S overwrites the previous tile's P (same columns): that tile's PV must have retired.  O sits at
                // +192..+256, beyond any S of this kernel (TKp <= 192), so S does not wait for the merge.
                if (g > 0) mbar_wait(&o_full[pp], (g - 1) & 1);
                tc_fence_after();
                const uint64_t dq = umma_desc_sw128(q_u + qt * 16384);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_ss(t_pipe, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                    umma_commit(&s_full[pp]);
                    if (qt == nq_tc - 1) umma_commit(qk_empty);
                }
                __syncwarp();
                if (qt == 0) mbar_wait(v_full, it & 1);
                mbar_wait(&p_full[pp], g & 1);
                mbar_wait(&o_empty[pp], (g & 1) ^ 1);  // the merge has read the previous tile's partial O
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 12; ++k) {  // TKp <= 192
                        if (k < nk) {
                            const uint32_t pcol = 8 * k + (k >= ka ? 8 * ka : 0);
                            umma_bf16_ts(t_pipe + ATC_O_COL, t_pipe + pcol, dv + 128 * k, idesc_o, k != 0);
                        }
                    }
                    umma_commit(&o_full[pp]);
                    if (qt == nq_tc - 1) umma_commit(v_empty);
                }
                __syncwarp();
            }
        }
    } else if (warp >= ATC_ROT_WARP0) {
        // ------------------------------------------------------------------- RoPE rotation warps (+ leftover rows)
        if (rope || leftover) {
            const int rtid = threadIdx.x - ATC_ROT_WARP0 * 32;
            const int nrot = T - p.prefix;
            const int units = 2 * nrot * 4;
            int it = 0;
            for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
                mbar_wait(qk_full, it & 1);
                if (rope) {
                    for (int u = rtid; u < units; u += ATC_ROT_WARPS * 32) {
                        const int ridx = u >> 2, cpair = u & 3;
                        const int isk = ridx >= nrot;
                        const int tok = p.prefix + ridx - (isk ? nrot : 0);
                        atc_rope_unit(isk ? k_s : q_s, tok, cpair, rope_tab + (tok - p.prefix) * 32);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(qk_ready);
                }
                if (leftover) {
                    // the last query tile's few rows: one m16 tile against all keys.  The 16-key blocks are dealt to the
                    // four rotation warps (flash-style online softmax inside a warp), the partial (m, l, O) - identical
                    // fragment layout in every warp, so the merge is lane by lane - meet in shared memory.
                    if (rope) mbar_wait(qk_ready, it & 1);  // every rotation warp has finished its share of Q and K
                    mbar_wait(v_full, it & 1);
                    const int rw = warp - ATC_ROT_WARP0;
                    const uint32_t sQ = smem_u32(q_s) + (nq - 1) * 16384, sK = smem_u32(k_s), sV = smem_u32(v_s);
                    uint32_t qf[4][4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int row = (lane & 7) + (((lane >> 3) & 1) << 3);
                        const int chunk = 2 * kk + (lane >> 4);
                        ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], sQ + att_swz(row, chunk));
                    }
                    float o[8][4];
#pragma unroll
                    for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
                    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
                    const int nb = TK >> 4;
                    for (int b = rw * nb / ATC_ROT_WARPS; b < (rw + 1) * nb / ATC_ROT_WARPS; ++b)
                        att_key_block<2, true>(qf, sK, sV, b * 16, T, p.scale_log2, o, m, l, lane);
                    float* mine = left_x + rw * 36 * 32 + lane;
                    if (rw != 0) {
#pragma unroll
                        for (int d = 0; d < 8; ++d)
#pragma unroll
                            for (int e = 0; e < 4; ++e) mine[(4 * d + e) * 32] = o[d][e];
                        mine[32 * 32] = m[0]; mine[33 * 32] = m[1]; mine[34 * 32] = l[0]; mine[35 * 32] = l[1];
                    }
                    named_bar_sync(10, 32 * ATC_ROT_WARPS);  // partial results written; nobody reads Q / K / V any more
                    if (rw == 0) {
#pragma unroll 1
                        for (int ow = 1; ow < ATC_ROT_WARPS; ++ow) {
                            const float* other = left_x + ow * 36 * 32 + lane;
                            float sc_mine[2], sc_other[2];
#pragma unroll
                            for (int r = 0; r < 2; ++r) {
                                const float mo = other[(32 + r) * 32];
                                const float mn = fmaxf(m[r], mo);
                                sc_mine[r] = m[r] == -INFINITY ? 0.f : ex2_approx(m[r] - mn);
                                sc_other[r] = mo == -INFINITY ? 0.f : ex2_approx(mo - mn);
                                m[r] = mn;
                                l[r] = l[r] * sc_mine[r] + other[(34 + r) * 32] * sc_other[r];
                            }
#pragma unroll
                            for (int d = 0; d < 8; ++d)
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    o[d][e] = o[d][e] * sc_mine[e >> 1] + other[(4 * d + e) * 32] * sc_other[e >> 1];
                        }
#pragma unroll
                        for (int r = 0; r < 2; ++r) {
                            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
                            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
                            l[r] = 1.0f / l[r];
                        }
                        const int f = w / p.heads, h = w % p.heads;
                        const int r0 = lane >> 2, r1 = r0 + 8;
                        __nv_bfloat16* ob = p.out + ((long long)f * T + 128 * (nq - 1)) * p.D + h * 64 + 2 * (lane & 3);
#pragma unroll
                        for (int d = 0; d < 8; ++d) {
                            if (r0 < left_rows)
                                *reinterpret_cast<uint32_t*>(ob + (long long)r0 * p.D + d * 8) = pack_bf16(o[d][0] * l[0], o[d][1] * l[0]);
                            if (r1 < left_rows)
                                *reinterpret_cast<uint32_t*>(ob + (long long)r1 * p.D + d * 8) = pack_bf16(o[d][2] * l[1], o[d][3] * l[1]);
                        }
                        // Q, K and V of this item are no longer read by the rotation warps
                        if (lane == 0) {
                            mbar_arrive(qk_empty);
                            mbar_arrive(v_empty);
                        }
                    }
                    named_bar_sync(11, 32 * ATC_ROT_WARPS);  // the scratch rows may be rewritten
                }
            }
        }
    } else {
        // ------------------------------------------------------------------- softmax + merge warpgroups
        const int pp = warp >> 3;          // pipeline = key block
        const int half = (warp >> 2) & 1;  // which half of the block's keys
        const int quarter = warp & 3;
        const int rit = quarter * 32 + lane;
        const int TKp = pp ? TK1 : TK0;
        const int key_off = pp ? TK0 : 0;  // first key of this block inside the frame
        const uint32_t t_lane = tmem_base + (uint32_t(quarter * 32) << 16);
        const uint32_t t_row = t_lane + 256 * pp;
        const float c = p.scale_log2;
        const int CA = ((TKp >> 4) + 1) / 2 * 16;
        const int c_begin = half ? CA : 0, c_end = half ? TKp : CA;
        float* my_max = xchg + ((0 * 2 + pp) * 2 + half) * 128 + rit;
        float* peer_max = xchg + ((0 * 2 + pp) * 2 + (half ^ 1)) * 128 + rit;
        float* sums_all = xchg + 2 * 2 * 128;                  // [tile parity][pipe][half][row]
        float* bmax_all = xchg + 2 * 2 * 128 + 2 * 2 * 2 * 128;  // [tile parity][pipe][row]
        int it = 0, g = 0;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x, ++it) {
            const int f = w / p.heads, h = w % p.heads;
            for (int qt = 0; qt < nq_tc; ++qt, ++g) {
                const int row = qt * 128 + rit;
                const bool warp_has_rows = (qt * 128 + quarter * 32) < T;
                float* sums = sums_all + (g & 1) * 512;
                float* blockmax = bmax_all + (g & 1) * 256;
                mbar_wait(&s_full[pp], g & 1);
                tc_fence_after();
                float sum = 0.f;
                float mx = -INFINITY;
                if (warp_has_rows) {
                    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld_32x16(t_row + c0, v);
                        tmem_ld_wait();
                        if (key_off + c0 + 16 <= T) {
#pragma unroll
                            for (int j = 0; j < 16; j += 2)
                                mx = fmaxf(mx, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (key_off + c0 + j < T) mx = fmaxf(mx, __uint_as_float(v[j]));
                        }
                    }
                }
                *my_max = mx;
                named_bar_sync(1 + pp * 4 + quarter, 64);  // only the partner warp: same rows, other half of the block
                if (warp_has_rows) {
                    mx = fmaxf(mx, *peer_max);  // block max; finite: each half of each block holds keys < T
                    if (half == 0) blockmax[pp * 128 + rit] = mx;
                    const float mc = mx * c;
                    for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld_32x16(t_row + c0, v);
                        tmem_ld_wait();
                        uint32_t pk[8];
                        if (key_off + c0 + 16 <= T) {
#pragma unroll
                            for (int j = 0; j < 16; j += 2) {
                                const float x0 = fmaf(__uint_as_float(v[j]), c, -mc);
                                const float x1 = fmaf(__uint_as_float(v[j + 1]), c, -mc);
                                if ((j >> 1) < 8 - ATC_POLY_PAIRS) pk[j >> 1] = pack_f16(ex2_approx(x0), ex2_approx(x1));  // fp32 MUFU, one pack
                                else pk[j >> 1] = pack_f16(ex2_poly3(x0), ex2_poly3(x1));
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; j += 2) {
                                const float x0 = (key_off + c0 + j < T) ? fmaf(__uint_as_float(v[j]), c, -mc) : -INFINITY;
                                const float x1 = (key_off + c0 + j + 1 < T) ? fmaf(__uint_as_float(v[j + 1]), c, -mc) : -INFINITY;
                                pk[j >> 1] = ex2_approx_f16x2(pack_f16(x0, x1));
                            }
                        }
                        __half2 a0 = __hadd2(*reinterpret_cast<__half2*>(&pk[0]), *reinterpret_cast<__half2*>(&pk[1]));
                        __half2 a1 = __hadd2(*reinterpret_cast<__half2*>(&pk[2]), *reinterpret_cast<__half2*>(&pk[3]));
                        __half2 a2 = __hadd2(*reinterpret_cast<__half2*>(&pk[4]), *reinterpret_cast<__half2*>(&pk[5]));
                        __half2 a3 = __hadd2(*reinterpret_cast<__half2*>(&pk[6]), *reinterpret_cast<__half2*>(&pk[7]));
                        const float2 sf = __half22float2(__hadd2(__hadd2(a0, a1), __hadd2(a2, a3)));
                        sum += sf.x + sf.y;
                        tmem_st_32x8(t_row + c_begin + ((c0 - c_begin) >> 1), pk);
                    }
                    tmem_st_wait();
                }
                sums[(pp * 2 + half) * 128 + rit] = sum;  // read by the merging threads after o_full (mbarrier chain)
                // the previous tile's TMA store must have read its staging rows before they are rewritten (one item
                // later); every rewrite follows a later o_full[1], hence this hand-over: the storing thread checks here,
                // long after the issue, instead of stalling its warp behind the store (attention_tc.cuh does the same)
                if (warp == 8 && g > 0) {
                    if (elect_one()) tma_wait_group_read<0>();
                    __syncwarp();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[pp]);

                if (pp == 0) continue;  // pipeline 0 goes straight on to the next tile

                // ---- merge (pipeline 1 only): this thread owns 32 output columns of its row, 16 at a time
                mbar_wait(&o_full[0], g & 1);
                mbar_wait(&o_full[1], g & 1);
                tc_fence_after();
                float w0 = 0.f, w1 = 0.f;
                uint32_t packed[16];
                if (warp_has_rows) {
                    const float m0 = blockmax[rit], m1 = blockmax[128 + rit];
                    const float mm = fmaxf(m0, m1);
                    w0 = ex2_approx((m0 - mm) * c);
                    w1 = ex2_approx((m1 - mm) * c);
                    const float l0 = sums[rit] + sums[128 + rit], l1 = sums[256 + rit] + sums[384 + rit];
                    const float inv = 1.0f / fmaf(w0, l0, w1 * l1);
                    w0 *= inv;
                    w1 *= inv;
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        uint32_t o0[16], o1[16];
                        tmem_ld_32x16(t_lane + ATC_O_COL + 32 * half + 16 * part, o0);
                        tmem_ld_32x16(t_lane + 256 + ATC_O_COL + 32 * half + 16 * part, o1);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; j += 2)
                            packed[8 * part + (j >> 1)] =
                                pack_bf16(fmaf(__uint_as_float(o0[j]), w0, __uint_as_float(o1[j]) * w1),
                                          fmaf(__uint_as_float(o0[j + 1]), w0, __uint_as_float(o1[j + 1]) * w1));
                    }
                }
                // both partial outputs and the exchange rows of this tile are consumed: the next PV of either
                // pipeline may overwrite its O columns
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&o_empty[0]);
                    mbar_arrive(&o_empty[1]);
                }
                if (warp_has_rows && row < T) {
                    uint8_t* srow = ostage + row * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(srow + (((4 * half + j) ^ (rit & 7)) << 4)) =
                            make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                }
                // one TMA store per query tile; the staging rows of tile qt are rewritten one item later
                fence_proxy_async();
                named_bar_sync(9, 256);
                if (warp == 8) {
                    if (elect_one()) {
                        if (qt == nq - 1) tma_store_3d(&tmap_o1, ostage + qt * 16384, h * 64, qt * 128, f);
                        else tma_store_3d(&tmap_o, ostage + qt * 16384, h * 64, qt * 128, f);
                        tma_commit_group();
                    }
                    __syncwarp();
                }
            }
        }
        if (warp == 8 && elect_one()) tma_wait_group<0>();  // this CTA's output stores have landed
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
