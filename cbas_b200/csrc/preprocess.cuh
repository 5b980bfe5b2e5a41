// uint8 frame -> bf16 patch matrix (the A operand of the patch-embedding GEMM), fused with the per-pixel
// preprocessing, in the two modes the reference stack has:
//
//  REFERENCE mode (cbas.py:431 `frames_np[:, :, :, 1] / 255.0`, cbas.py:672-675 replicate to 3 channels):
//      green plane only, native resolution.  Because the three input channels are identical, the patch
//      embedding is folded on the host to K = 256 (W'[d,ky,kx] = sum_c W[d,c,ky,kx] / 255), so the kernel
//      emits the raw green byte as bf16 - exact, no rounding at all on the activation side.
//      A[frame*Np + py*nw + px][ky*16 + kx] = G[frame][16*py + ky][16*px + kx]
//
//  PROCESSOR mode (HF image_processing_dinov3_vit.py:45-86: rescale 1/255 -> antialiased bilinear resize to
//      S x S -> (x - mean) / std):  separable triangle filter with host-precomputed taps (rope/resize tables in
//      tables.cu restate ATen's _upsample_bilinear2d_aa weights), all 3 channels, K = 768 in Conv2d weight
//      order.  A[frame*Np + py*nw + px][c*256 + ky*16 + kx]
//
// Both are HBM-bound byte shuffles: 16-byte loads where the layout allows, 16/32-byte stores.
#pragma once
#include "ptx.cuh"

namespace cbas {

// One thread per (frame, patch row py, ky, px): 48 contiguous source bytes (16 RGB pixels) -> 16 bf16.
// Requires W % 16 == 0, H % 16 == 0 (the ViT itself requires it) and a 16-byte aligned frame base / row pitch.
__global__ void __launch_bounds__(256)
preprocess_green_kernel(const uint8_t* __restrict__ frames, __nv_bfloat16* __restrict__ A, int n_frames, int H,
                        int W, long long frame_stride, int row_stride) {
    const int nw = W >> 4, nh = H >> 4;
    const long long total = (long long)n_frames * H * nw;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int px = gid % nw;
    const long long t = gid / nw;
    const int y = t % H;
    const int f = t / H;
    const uint8_t* src = frames + f * frame_stride + (long long)y * row_stride + px * 48;
    uint32_t w[12];
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i)
            w[i] = src[4 * i] | (src[4 * i + 1] << 8) | (src[4 * i + 2] << 16) | (uint32_t(src[4 * i + 3]) << 24);
    }
    // green of pixel i is byte 3*i+1
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const int b0 = 3 * i + 1, b1 = 3 * i + 4;
        const float g0 = float((w[b0 >> 2] >> ((b0 & 3) * 8)) & 0xff);
        const float g1 = float((w[b1 >> 2] >> ((b1 & 3) * 8)) & 0xff);
        o[i >> 1] = pack_bf16(g0, g1);
    }
    const int py = y >> 4, ky = y & 15;
    __nv_bfloat16* dst = A + ((long long)f * nh * nw + (long long)py * nw + px) * 256 + ky * 16;
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

constexpr int RESIZE_MAX_TAPS = 8;

struct ResizeTaps {
    // per output coordinate: first source index and up to RESIZE_MAX_TAPS normalised weights
    const int* ymin; const float* wy;  // [S], [S][taps]
    const int* xmin; const float* wx;  // [S], [S][taps]
    int taps_y, taps_x;
};

// One thread per (frame, output row y, 8 consecutive output x) for all three channels.
__global__ void __launch_bounds__(256)
preprocess_resize_kernel(const uint8_t* __restrict__ frames, __nv_bfloat16* __restrict__ A, int n_frames, int H,
                         int W, long long frame_stride, int row_stride, int S, ResizeTaps tp, float3 mean,
                         float3 inv_std) {
    const int xg = S >> 3;  // groups of 8 output pixels (S % 16 == 0)
    const long long total = (long long)n_frames * S * xg;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int gx = gid % xg;
    const long long t = gid / xg;
    const int y = t % S;
    const int f = t / S;
    const uint8_t* img = frames + f * frame_stride;
    const int y0 = __ldg(tp.ymin + y);
    float acc[8][3];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 0.f;
    for (int j = 0; j < tp.taps_y; ++j) {
        const float wyj = __ldg(tp.wy + y * tp.taps_y + j);
        if (wyj == 0.f) continue;
        const uint8_t* row = img + (long long)min(y0 + j, H - 1) * row_stride;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int x = gx * 8 + i;
            const int x0 = __ldg(tp.xmin + x);
            float r = 0.f, g = 0.f, b = 0.f;
            for (int k = 0; k < tp.taps_x; ++k) {
                const float wxk = __ldg(tp.wx + x * tp.taps_x + k);
                const uint8_t* px = row + 3 * min(x0 + k, W - 1);
                r += wxk * float(px[0]);
                g += wxk * float(px[1]);
                b += wxk * float(px[2]);
            }
            acc[i][0] += wyj * r; acc[i][1] += wyj * g; acc[i][2] += wyj * b;
        }
    }
    const int ns = S >> 4;
    const int py = y >> 4, ky = y & 15;
    const int px = gx >> 1, kx0 = (gx & 1) * 8;
    __nv_bfloat16* dst = A + ((long long)f * ns * ns + (long long)py * ns + px) * 768 + ky * 16 + kx0;
    const float mu[3] = {mean.x, mean.y, mean.z}, is[3] = {inv_std.x, inv_std.y, inv_std.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float v0 = (acc[i][c] * (1.0f / 255.0f) - mu[c]) * is[c];
            const float v1 = (acc[i + 1][c] * (1.0f / 255.0f) - mu[c]) * is[c];
            o[i >> 1] = pack_bf16(v0, v1);
        }
        *reinterpret_cast<uint4*>(dst + c * 256) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// CLS + register rows of the residual stream (HF modeling_dinov3_vit.py:86-90): h[frame*T + j] = prefix[j], j < P
__global__ void __launch_bounds__(256)
fill_prefix_kernel(float* __restrict__ h, const float* __restrict__ prefix_tokens, int n_frames, int T, int P, int D) {
    const long long total = (long long)n_frames * P * (D / 4);
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int d4 = gid % (D / 4);
    const long long t = gid / (D / 4);
    const int j = t % P;
    const int f = t / P;
    reinterpret_cast<float4*>(h + ((long long)f * T + j) * D)[d4] =
        __ldg(reinterpret_cast<const float4*>(prefix_tokens + (long long)j * D) + d4);
}

}  // namespace cbas
