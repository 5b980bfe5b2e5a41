// Fused multi-head self-attention for one ViT frame-head per CTA, head_dim 64, non-causal, no mask
// (reference: HF modeling_dinov3_vit.py:294-334 DINOv3ViTAttention.forward with the SDPA backend,
//  :238-268 apply_rotary_pos_emb, :203-207 rotate_half).
//
//   prologue : Q,K,V head slices [T,64] are pulled from the fused QKV activation into shared memory with
//              16-byte loads; RoPE (rotate-half form, fp32 math, table from rope.cu) is applied to the Q
//              and K rows of patch tokens (token >= prefix) on the way in; rows >= T are zero-filled.
//   main     : flash-style pass over key blocks of 64 with an online softmax; S = QK^T and O += PV run on
//              warp-level mma.sync (m16n8k16, bf16 in, fp32 accumulate); probabilities never leave registers.
//   epilogue : O / rowsum -> bf16 -> [M, D] at column head*64.
//
// Shared-memory rows are 128 B (64 bf16) with the 16-byte chunk index XOR-swizzled by (row & 7), so both the
// prologue stores and every ldmatrix are bank-conflict free.
#pragma once
#include "ptx.cuh"

namespace cbas {

constexpr int ATT_HEAD_DIM = 64;
constexpr int ATT_MAX_THREADS = 256;  // the launcher picks 4..8 warps so that the 16-row query tiles divide evenly
constexpr int ATT_KEY_BLOCK = 64;

__device__ __forceinline__ uint32_t att_swz(int row, int chunk) {
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void rope_pair(uint4& lo, uint4& hi, const float* __restrict__ cs,
                                          const float* __restrict__ sn) {
    // lo = x[8c..8c+8), hi = x[32+8c..32+8c+8); cos/sin are the first-half table entries (tile(2) layout:
    // cos[i+32] == cos[i]).  out_lo = lo*cos - hi*sin ; out_hi = hi*cos + lo*sin
    uint32_t* l = reinterpret_cast<uint32_t*>(&lo);
    uint32_t* h = reinterpret_cast<uint32_t*>(&hi);
    // 8 consecutive table entries each (32-byte aligned: the row pitch is 128 B and the offset 32*c B)
    const float4 ca = __ldg(reinterpret_cast<const float4*>(cs)), cb = __ldg(reinterpret_cast<const float4*>(cs) + 1);
    const float4 sa = __ldg(reinterpret_cast<const float4*>(sn)), sb = __ldg(reinterpret_cast<const float4*>(sn) + 1);
    const float cc[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
    const float ss[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = unpack_bf16(l[i]), b = unpack_bf16(h[i]);
        const float c0 = cc[2 * i], c1 = cc[2 * i + 1];
        const float s0 = ss[2 * i], s1 = ss[2 * i + 1];
        l[i] = pack_bf16(a.x * c0 - b.x * s0, a.y * c1 - b.y * s1);
        h[i] = pack_bf16(b.x * c0 + a.x * s0, b.y * c1 + a.y * s1);
    }
}

// One key block: NT n-tiles of 8 keys starting at key0.  VF16: V (and therefore P) as IEEE f16 - the layout the
// tcgen05 kernels' QKV epilogue leaves (attention_tc_split.cuh runs a frame's few leftover query rows through here).
template <int NT, bool VF16 = false>
__device__ __forceinline__ void att_key_block(const uint32_t (&qf)[4][4], uint32_t sK, uint32_t sV, int key0, int T,
                                              float scale_log2, float (&o)[8][4], float (&m)[2], float (&l)[2],
                                              int lane) {
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    // S = Q K^T
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int j = 0; j < NT; j += 2) {
            uint32_t b0, b1, b2, b3;
            const int key = key0 + j * 8 + (lane & 7) + ((lane >> 4) << 3);
            const int chunk = 2 * kk + ((lane >> 3) & 1);
            ldmatrix_x4(b0, b1, b2, b3, sK + att_swz(key, chunk));
            mma_bf16_16816(s[j], qf[kk], b0, b1);
            mma_bf16_16816(s[j + 1], qf[kk], b2, b3);
        }
    }
    // scale, mask the padded keys, block row-max
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < NT; ++j) {
        const int col = key0 + j * 8 + 2 * (lane & 3);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool ok = (col + (e & 1)) < T;
            s[j][e] = ok ? s[j][e] * scale_log2 : -INFINITY;
            mx[e >> 1] = fmaxf(mx[e >> 1], s[j][e]);
        }
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float mn = fmaxf(m[r], mx[r]);
        alpha[r] = ex2_approx(m[r] - mn);  // m = -inf on the first block -> 0
        m[r] = mn;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            s[j][e] = ex2_approx(s[j][e] - m[e >> 1]);  // one MUFU; masked keys are -inf -> 0
            rs[e >> 1] += s[j][e];
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l[r] = l[r] * alpha[r] + rs[r];
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        o[d][0] *= alpha[0]; o[d][1] *= alpha[0];
        o[d][2] *= alpha[1]; o[d][3] *= alpha[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
        uint32_t pa[4];
        if constexpr (VF16) {
            pa[0] = pack_f16(s[2 * kk][0], s[2 * kk][1]);
            pa[1] = pack_f16(s[2 * kk][2], s[2 * kk][3]);
            pa[2] = pack_f16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            pa[3] = pack_f16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        } else {
            pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
            pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
            pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        }
#pragma unroll
        for (int d = 0; d < 8; d += 2) {
            uint32_t b0, b1, b2, b3;
            const int key = key0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
            const int chunk = d + (lane >> 4);
            ldmatrix_x4_trans(b0, b1, b2, b3, sV + att_swz(key, chunk));
            if constexpr (VF16) {
                mma_f16_16816(o[d], pa, b0, b1);
                mma_f16_16816(o[d + 1], pa, b2, b3);
            } else {
                mma_bf16_16816(o[d], pa, b0, b1);
                mma_bf16_16816(o[d + 1], pa, b2, b3);
            }
        }
    }
}

// qkv: [frames*T, 3*D] bf16 (q | k | v, head h at column h*64 of each third);  out: [frames*T, D] bf16.
// rope_cos / rope_sin: [T - prefix, 32] fp32.   grid = frames * heads, block = 128..256, dyn smem = 3*TP*128 B.
__global__ void __launch_bounds__(ATT_MAX_THREADS)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                 const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int T, int prefix,
                 int heads, int D, float scale_log2) {
    extern __shared__ __align__(128) uint8_t att_smem[];
    const int TP = (T + 15) & ~15;
    uint8_t* q_s = att_smem;
    uint8_t* k_s = q_s + TP * 128;
    uint8_t* v_s = k_s + TP * 128;
    const int frame = blockIdx.x / heads, head = blockIdx.x % heads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long ld = 3ll * D;
    const __nv_bfloat16* base = qkv + (long long)frame * T * ld + head * ATT_HEAD_DIM;

    // ---- prologue: gather + RoPE
    for (int idx = tid; idx < TP * 4; idx += blockDim.x) {
        const int row = idx >> 2, c = idx & 3;
        uint4 qlo = make_uint4(0, 0, 0, 0), qhi = qlo, klo = qlo, khi = qlo, vlo = qlo, vhi = qlo;
        if (row < T) {
            const __nv_bfloat16* r = base + row * ld;
            qlo = *reinterpret_cast<const uint4*>(r + 8 * c);
            qhi = *reinterpret_cast<const uint4*>(r + 32 + 8 * c);
            klo = *reinterpret_cast<const uint4*>(r + D + 8 * c);
            khi = *reinterpret_cast<const uint4*>(r + D + 32 + 8 * c);
            vlo = *reinterpret_cast<const uint4*>(r + 2 * D + 8 * c);
            vhi = *reinterpret_cast<const uint4*>(r + 2 * D + 32 + 8 * c);
            if (row >= prefix) {
                const float* cs = rope_cos + (row - prefix) * 32 + 8 * c;
                const float* sn = rope_sin + (row - prefix) * 32 + 8 * c;
                rope_pair(qlo, qhi, cs, sn);
                rope_pair(klo, khi, cs, sn);
            }
        }
        *reinterpret_cast<uint4*>(q_s + att_swz(row, c)) = qlo;
        *reinterpret_cast<uint4*>(q_s + att_swz(row, c + 4)) = qhi;
        *reinterpret_cast<uint4*>(k_s + att_swz(row, c)) = klo;
        *reinterpret_cast<uint4*>(k_s + att_swz(row, c + 4)) = khi;
        *reinterpret_cast<uint4*>(v_s + att_swz(row, c)) = vlo;
        *reinterpret_cast<uint4*>(v_s + att_swz(row, c + 4)) = vhi;
    }
    __syncthreads();

    const uint32_t sQ = smem_u32(q_s), sK = smem_u32(k_s), sV = smem_u32(v_s);
    const int m_tiles = TP >> 4;
    const int full_blocks = TP / ATT_KEY_BLOCK;
    const int tail_tiles = (TP % ATT_KEY_BLOCK) >> 3;

    for (int mt = warp; mt < m_tiles; mt += (blockDim.x >> 5)) {
        uint32_t qf[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int row = mt * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
            const int chunk = 2 * kk + (lane >> 4);
            ldmatrix_x4(qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], sQ + att_swz(row, chunk));
        }
        float o[8][4];
#pragma unroll
        for (int d = 0; d < 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
        float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};

        for (int kb = 0; kb < full_blocks; ++kb)
            att_key_block<8>(qf, sK, sV, kb * ATT_KEY_BLOCK, T, scale_log2, o, m, l, lane);
        const int key0 = full_blocks * ATT_KEY_BLOCK;
        if (tail_tiles == 2) att_key_block<2>(qf, sK, sV, key0, T, scale_log2, o, m, l, lane);
        else if (tail_tiles == 4) att_key_block<4>(qf, sK, sV, key0, T, scale_log2, o, m, l, lane);
        else if (tail_tiles == 6) att_key_block<6>(qf, sK, sV, key0, T, scale_log2, o, m, l, lane);

        // rowsum across the quad, normalise, store
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
            l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
            l[r] = 1.0f / l[r];
        }
        const int r0 = mt * 16 + (lane >> 2), r1 = r0 + 8;
        __nv_bfloat16* ob = out + (long long)frame * T * D + head * ATT_HEAD_DIM + 2 * (lane & 3);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            if (r0 < T)
                *reinterpret_cast<uint32_t*>(ob + (long long)r0 * D + d * 8) = pack_bf16(o[d][0] * l[0], o[d][1] * l[0]);
            if (r1 < T)
                *reinterpret_cast<uint32_t*>(ob + (long long)r1 * D + d * 8) = pack_bf16(o[d][2] * l[1], o[d][3] * l[1]);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// CLS-query attention for the LAST block.  Only the CLS row of the final hidden state is ever read
// (modeling_dinov3_vit.py:547-548 + cbas.py:677), so in the last block only the CLS query has to attend.
// One warp per (frame, head); the warp walks the keys four at a time, EIGHT LANES PER KEY ROW (16 bytes each), so
// every load instruction touches four full 128-byte lines instead of 32 different ones.  RoPE is applied to the
// patch-token keys on the fly (the partner half of the row sits four lanes away), the rotated key is rounded to
// bf16 exactly as the dense kernels do, scores go to a warp-private shared-memory row, then softmax and P V run
// with the same lane layout and the four key groups are summed at the end.
// q_cls: [frames, D] bf16 (un-rotated: the CLS token is a prefix token); k and v live in the fused QKV buffer.
constexpr int CLS_ATT_MAX_T = 1024;  // 16 KB of static shared memory for the four warps' score rows
__global__ void __launch_bounds__(128)
cls_attention_kernel(const __nv_bfloat16* __restrict__ q_cls, const __nv_bfloat16* __restrict__ qkv,
                     __nv_bfloat16* __restrict__ out, const float* __restrict__ rope_cos,
                     const float* __restrict__ rope_sin, int frames, int T, int prefix, int heads, int D,
                     float scale_log2, int v_is_f16) {
    __shared__ float s_sc[4][CLS_ATT_MAX_T];
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= frames * heads) return;
    float* sc = s_sc[threadIdx.x >> 5];
    const int frame = item / heads, head = item % heads;
    const long long ld = 3ll * D;
    const __nv_bfloat16* kbase = qkv + (long long)frame * T * ld + D + head * ATT_HEAD_DIM;
    const __nv_bfloat16* vbase = kbase + D;
    const int g = lane >> 3, j = lane & 7;  // key group, 16-byte chunk of the row (elements 8j .. 8j+7)
    float q[8];
    {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(q_cls + (long long)frame * D + head * ATT_HEAD_DIM) + j);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_bf16(w[e]);
            q[2 * e] = f.x * scale_log2;
            q[2 * e + 1] = f.y * scale_log2;
        }
    }
    const float sgn = j < 4 ? -1.f : 1.f;  // rotate-half: lo' = lo c - hi s, hi' = hi c + lo s
    const int iters = (T + 3) >> 2;
    float mx = -INFINITY;
#pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        const int t = 4 * it + g;
        const bool live = t < T;
        const bool rot = live && t >= prefix;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (live) u = __ldg(reinterpret_cast<const uint4*>(kbase + (long long)t * ld) + j);
        float4 c0, c1, s0, s1;
        if (rot) {
            const float4* cs = reinterpret_cast<const float4*>(rope_cos + (t - prefix) * 32) + 2 * (j & 3);
            const float4* sn = reinterpret_cast<const float4*>(rope_sin + (t - prefix) * 32) + 2 * (j & 3);
            c0 = __ldg(cs); c1 = __ldg(cs + 1); s0 = __ldg(sn); s1 = __ldg(sn + 1);
        } else {
            c0 = c1 = make_float4(1.f, 1.f, 1.f, 1.f);
            s0 = s1 = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const uint32_t own[4] = {u.x, u.y, u.z, u.w};
        uint32_t oth[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) oth[e] = __shfl_xor_sync(0xffffffffu, own[e], 4);
        const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
        const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        float acc = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = unpack_bf16(own[e]), b = unpack_bf16(oth[e]);
            const float r0 = __bfloat162float(__float2bfloat16_rn(fmaf(sgn * b.x, ss[2 * e], a.x * cc[2 * e])));
            const float r1 = __bfloat162float(__float2bfloat16_rn(fmaf(sgn * b.y, ss[2 * e + 1], a.y * cc[2 * e + 1])));
            acc = fmaf(q[2 * e], r0, acc);
            acc = fmaf(q[2 * e + 1], r1, acc);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (live) {
            if (j == 0) sc[t] = acc;
            mx = fmaxf(mx, acc);
        }
    }
    mx = warp_max(mx);
    __syncwarp();
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) {
        const float p = exp2f(sc[t] - mx);
        sc[t] = p;
        sum += p;
    }
    sum = warp_sum(sum);
    __syncwarp();
    // O = P V, lane (g, j) accumulates dims 8j..8j+7 over the keys of its group
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll 8
    for (int it = 0; it < iters; ++it) {
        const int t = 4 * it + g;
        if (t < T) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(vbase + (long long)t * ld) + j);
            const float p = sc[t];
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 v;
                if (v_is_f16) v = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                else v = unpack_bf16(w[e]);
                o[2 * e] = fmaf(p, v.x, o[2 * e]);
                o[2 * e + 1] = fmaf(p, v.y, o[2 * e + 1]);
            }
        }
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        o[e] += __shfl_xor_sync(0xffffffffu, o[e], 8);
        o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
        o[e] *= inv;
    }
    if (g == 0) {
        uint4 r;
        r.x = pack_bf16(o[0], o[1]); r.y = pack_bf16(o[2], o[3]); r.z = pack_bf16(o[4], o[5]); r.w = pack_bf16(o[6], o[7]);
        reinterpret_cast<uint4*>(out + (long long)frame * D + head * ATT_HEAD_DIM)[j] = r;
    }
}

}  // namespace cbas
