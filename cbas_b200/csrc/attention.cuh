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
constexpr int ATT_THREADS = 128;
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
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = unpack_bf16(l[i]), b = unpack_bf16(h[i]);
        const float c0 = __ldg(cs + 2 * i), c1 = __ldg(cs + 2 * i + 1);
        const float s0 = __ldg(sn + 2 * i), s1 = __ldg(sn + 2 * i + 1);
        l[i] = pack_bf16(a.x * c0 - b.x * s0, a.y * c1 - b.y * s1);
        h[i] = pack_bf16(b.x * c0 + a.x * s0, b.y * c1 + a.y * s1);
    }
}

// One key block: NT n-tiles of 8 keys starting at key0.
template <int NT>
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
        alpha[r] = exp2f(m[r] - mn);  // m = -inf on the first block -> 0
        m[r] = mn;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            s[j][e] = exp2f(s[j][e] - m[e >> 1]);
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
        pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int d = 0; d < 8; d += 2) {
            uint32_t b0, b1, b2, b3;
            const int key = key0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
            const int chunk = d + (lane >> 4);
            ldmatrix_x4_trans(b0, b1, b2, b3, sV + att_swz(key, chunk));
            mma_bf16_16816(o[d], pa, b0, b1);
            mma_bf16_16816(o[d + 1], pa, b2, b3);
        }
    }
}

// qkv: [frames*T, 3*D] bf16 (q | k | v, head h at column h*64 of each third);  out: [frames*T, D] bf16.
// rope_cos / rope_sin: [T - prefix, 32] fp32.   grid = frames * heads, block = 128, dyn smem = 3*TP*128 B.
__global__ void __launch_bounds__(ATT_THREADS)
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
    for (int idx = tid; idx < TP * 4; idx += ATT_THREADS) {
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

    for (int mt = warp; mt < m_tiles; mt += ATT_THREADS / 32) {
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
// (modeling_dinov3_vit.py:547-548 + cbas.py:677), so in the last block only the CLS query has to attend:
// one warp per (frame, head) scores the CLS query against all T keys (RoPE applied to the patch-token keys on the
// fly, fp32 math), soft-maxes over the warp and accumulates P V with the lanes striding the 64 output dims.
// q_cls: [frames, D] bf16 (un-rotated: the CLS token is a prefix token); k and v live in the fused QKV buffer.
__global__ void __launch_bounds__(128)
cls_attention_kernel(const __nv_bfloat16* __restrict__ q_cls, const __nv_bfloat16* __restrict__ qkv,
                     __nv_bfloat16* __restrict__ out, const float* __restrict__ rope_cos,
                     const float* __restrict__ rope_sin, int frames, int T, int prefix, int heads, int D,
                     float scale_log2, int v_is_f16) {
    const int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (item >= frames * heads) return;
    const int frame = item / heads, head = item % heads;
    const long long ld = 3ll * D;
    const __nv_bfloat16* kbase = qkv + (long long)frame * T * ld + D + head * ATT_HEAD_DIM;
    const __nv_bfloat16* vbase = kbase + D;
    // every lane keeps the whole 64-d query (fp32)
    float q[64];
    {
        const uint4* qp = reinterpret_cast<const uint4*>(q_cls + (long long)frame * D + head * ATT_HEAD_DIM);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 u = __ldg(qp + i);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16(w[j]);
                q[8 * i + 2 * j] = f.x;
                q[8 * i + 2 * j + 1] = f.y;
            }
        }
    }
    // scores: lane owns keys lane, lane+32, ...
    constexpr int MAX_SLOTS = 9;  // T <= 288
    float sc[MAX_SLOTS];
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < MAX_SLOTS; ++s) {
        const int t = s * 32 + lane;
        float acc = -INFINITY;
        if (t < T) {
            const uint4* kp = reinterpret_cast<const uint4*>(kbase + (long long)t * ld);
            float k[64];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 u = __ldg(kp + i);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16(w[j]);
                    k[8 * i + 2 * j] = f.x;
                    k[8 * i + 2 * j + 1] = f.y;
                }
            }
            acc = 0.f;
            if (t >= prefix) {
                const float4* cs = reinterpret_cast<const float4*>(rope_cos + (t - prefix) * 32);
                const float4* sn = reinterpret_cast<const float4*>(rope_sin + (t - prefix) * 32);
#pragma unroll
                for (int i4 = 0; i4 < 8; ++i4) {
                    const float4 c4 = __ldg(cs + i4), s4 = __ldg(sn + i4);
                    const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = 4 * i4 + e;
                        // the kernels that feed the tensor cores round the rotated key to bf16; do the same here
                        const float klo = __bfloat162float(__float2bfloat16_rn(k[i] * cc[e] - k[i + 32] * ss[e]));
                        const float khi = __bfloat162float(__float2bfloat16_rn(k[i + 32] * cc[e] + k[i] * ss[e]));
                        acc = fmaf(q[i], klo, acc);
                        acc = fmaf(q[i + 32], khi, acc);
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i) acc = fmaf(q[i], k[i], acc);
            }
            acc *= scale_log2;
        }
        sc[s] = acc;
        mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < MAX_SLOTS; ++s) {
        sc[s] = (s * 32 + lane < T) ? exp2f(sc[s] - mx) : 0.f;
        sum += sc[s];
    }
    sum = warp_sum(sum);
    // O = P V: lane owns output dims 2*lane, 2*lane+1
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int s = 0; s < MAX_SLOTS; ++s) {
        const int tmax = min(32, T - s * 32);
        for (int j = 0; j < tmax; ++j) {
            const float pj = __shfl_sync(0xffffffffu, sc[s], j);
            const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(vbase + (long long)(s * 32 + j) * ld) + lane);
            float2 v;
            if (v_is_f16) v = __half22float2(*reinterpret_cast<const __half2*>(&raw));
            else v = unpack_bf16(raw);
            o0 = fmaf(pj, v.x, o0);
            o1 = fmaf(pj, v.y, o1);
        }
    }
    const float inv = 1.0f / sum;
    reinterpret_cast<uint32_t*>(out + (long long)frame * D + head * ATT_HEAD_DIM)[lane] = pack_bf16(o0 * inv, o1 * inv);
}

}  // namespace cbas
