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

// The same for a frame that arrives as ONE uint8 plane [H, W] (the decode workers ship only the green plane in
// REFERENCE mode: a third of the host-to-device bytes): 16 contiguous bytes -> 16 bf16.
__global__ void __launch_bounds__(256)
preprocess_plane_u8_kernel(const uint8_t* __restrict__ planes, __nv_bfloat16* __restrict__ A, int n_frames, int H,
                           int W, long long frame_stride, int row_stride) {
    const int nw = W >> 4, nh = H >> 4;
    const long long total = (long long)n_frames * H * nw;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int px = gid % nw;
    const long long t = gid / nw;
    const int y = t % H;
    const int f = t / H;
    const uint8_t* src = planes + f * frame_stride + (long long)y * row_stride + px * 16;
    uint32_t w[4];
    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            w[i] = src[4 * i] | (src[4 * i + 1] << 8) | (src[4 * i + 2] << 16) | (uint32_t(src[4 * i + 3]) << 24);
    }
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 16; i += 2)
        o[i >> 1] = pack_bf16(float((w[i >> 2] >> ((i & 3) * 8)) & 0xff), float((w[(i + 1) >> 2] >> (((i + 1) & 3) * 8)) & 0xff));
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


// Shared-memory version of the resize path for the common case (16-byte aligned rows): one CTA per (frame, patch
// row).  The source rows the 16 output rows need are staged with 16-byte loads, resampled horizontally into an
// fp32 strip, then vertically, normalised and written as 16-byte pieces of the patch matrix.  Every source byte
// is read from HBM once (the per-pixel version above re-reads each one ~25 times through L1).
struct ResizeTileGeom {
    int max_rows;  // largest source-row span any patch row needs (host-computed from the taps)
};

__global__ void __launch_bounds__(256)
preprocess_resize_tile_kernel(const uint8_t* __restrict__ frames, __nv_bfloat16* __restrict__ A, int H, int W,
                              long long frame_stride, int row_stride, int S, ResizeTaps tp, float3 mean,
                              float3 inv_std, int max_rows) {
    extern __shared__ __align__(16) uint8_t rs_smem[];
    const int ns = S >> 4;
    const int f = blockIdx.x / ns, py = blockIdx.x % ns;
    const int y_first = py * 16;
    const int r0 = __ldg(tp.ymin + y_first);
    int r1 = __ldg(tp.ymin + y_first + 15) + tp.taps_y;
    r1 = r1 < H ? r1 : H;
    const int nrows = r1 - r0;
    const int row_bytes = W * 3;
    const int src_pitch = (row_bytes + 15) & ~15;
    uint8_t* src = rs_smem;                                                     // [max_rows][src_pitch]
    float* hbuf = reinterpret_cast<float*>(rs_smem + max_rows * src_pitch);     // [max_rows][S*3]
    const uint8_t* img = frames + f * frame_stride + (long long)r0 * row_stride;
    // A: stage the source rows
    const int vec_per_row = src_pitch >> 4;
    for (int i = threadIdx.x; i < nrows * vec_per_row; i += blockDim.x) {
        const int r = i / vec_per_row, v = i - r * vec_per_row;
        const uint8_t* g = img + (long long)r * row_stride + v * 16;
        uint4 q;
        if (v * 16 + 16 <= row_bytes) q = __ldg(reinterpret_cast<const uint4*>(g));
        else {
            uint8_t tmp[16];
#pragma unroll
            for (int b = 0; b < 16; ++b) tmp[b] = (v * 16 + b < row_bytes) ? g[b] : 0;
            q = *reinterpret_cast<uint4*>(tmp);
        }
        *reinterpret_cast<uint4*>(src + r * src_pitch + v * 16) = q;
    }
    __syncthreads();
    // B: horizontal pass (fp32), all three channels of one (row, x) per thread iteration
    for (int i = threadIdx.x; i < nrows * S; i += blockDim.x) {
        const int r = i / S, x = i - r * S;
        const int x0 = __ldg(tp.xmin + x);
        const uint8_t* row = src + r * src_pitch;
        float cr = 0.f, cg = 0.f, cb = 0.f;
        for (int k = 0; k < tp.taps_x; ++k) {
            const float w = __ldg(tp.wx + x * tp.taps_x + k);
            const uint8_t* px = row + 3 * min(x0 + k, W - 1);
            cr += w * float(px[0]);
            cg += w * float(px[1]);
            cb += w * float(px[2]);
        }
        float* o = hbuf + (r * S + x) * 3;
        o[0] = cr; o[1] = cg; o[2] = cb;
    }
    __syncthreads();
    // C: vertical pass, normalise, store 8 consecutive kx (16 bytes) per task
    const float mu[3] = {mean.x, mean.y, mean.z}, is[3] = {inv_std.x, inv_std.y, inv_std.z};
    const int tasks = 16 * (S >> 3) * 3;
    for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
        const int c = t % 3;
        const int gx = (t / 3) % (S >> 3);
        const int ky = t / (3 * (S >> 3));
        const int y = y_first + ky;
        const int yr = __ldg(tp.ymin + y) - r0;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < tp.taps_y; ++j) {
            const float w = __ldg(tp.wy + y * tp.taps_y + j);
            const int r = min(yr + j, nrows - 1);
            const float* hrow = hbuf + (r * S + gx * 8) * 3 + c;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] += w * hrow[3 * i];
        }
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float v0 = (acc[i] * (1.0f / 255.0f) - mu[c]) * is[c];
            const float v1 = (acc[i + 1] * (1.0f / 255.0f) - mu[c]) * is[c];
            o[i >> 1] = pack_bf16(v0, v1);
        }
        const int px = gx >> 1, kx0 = (gx & 1) * 8;
        __nv_bfloat16* dst = A + ((long long)f * ns * ns + (long long)py * ns + px) * 768 + c * 256 + ky * 16 + kx0;
        *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// Production version of the tiled resize for the benchmark geometry (S <= 256 output columns, <= 5 taps per axis:
// 256 -> 224 has 5).  Same arithmetic, in the same order, as preprocess_resize_tile_kernel - the two are bitwise
// identical - but shaped for instruction-level parallelism instead of generality: one thread per output COLUMN keeps
// its horizontal taps and byte offsets in registers and walks the staged rows (no index division, no tap loads in the
// loop); the row taps of the 16 output rows sit in shared memory; the finished patch-matrix rows of this patch row
// (ns x 768 bf16, contiguous in global memory) are assembled in shared memory and leave with 16-byte coalesced stores.
constexpr int RESIZE_FAST_TAPS = 5;

__global__ void __launch_bounds__(256)
preprocess_resize_fast_kernel(const uint8_t* __restrict__ frames, __nv_bfloat16* __restrict__ A, int H, int W,
                              long long frame_stride, int row_stride, int S, ResizeTaps tp, float3 mean,
                              float3 inv_std, int max_rows, int io_bytes) {
    extern __shared__ __align__(16) uint8_t rs_smem[];
    const int ns = S >> 4;
    const int f = blockIdx.x / ns, py = blockIdx.x % ns;
    const int y_first = py * 16;
    const int r0 = __ldg(tp.ymin + y_first);
    int r1 = __ldg(tp.ymin + y_first + 15) + tp.taps_y;
    r1 = r1 < H ? r1 : H;
    const int nrows = r1 - r0;
    const int row_bytes = W * 3;
    const int src_pitch = (row_bytes + 15) & ~15;
    uint8_t* src = rs_smem;                                                   // [max_rows][src_pitch], later out_s
    float* hbuf = reinterpret_cast<float*>(rs_smem + io_bytes);               // [3][max_rows][S]
    float4* ytab = reinterpret_cast<float4*>(hbuf + 3 * max_rows * S);        // [16][2]: w0..w3 | w4, first row, -, -
    const int t = threadIdx.x;
    const uint8_t* img = frames + f * frame_stride + (long long)r0 * row_stride;
    // A: stage the source rows; the row taps of the 16 output rows
    const int vec_per_row = src_pitch >> 4;
    for (int i = t; i < nrows * vec_per_row; i += 256) {
        const int r = i / vec_per_row, v = i - r * vec_per_row;
        const uint8_t* g = img + (long long)r * row_stride + v * 16;
        uint4 q;
        if (v * 16 + 16 <= row_bytes) q = __ldg(reinterpret_cast<const uint4*>(g));
        else {
            uint8_t tmp[16];
#pragma unroll
            for (int b = 0; b < 16; ++b) tmp[b] = (v * 16 + b < row_bytes) ? g[b] : 0;
            q = *reinterpret_cast<uint4*>(tmp);
        }
        *reinterpret_cast<uint4*>(src + r * src_pitch + v * 16) = q;
    }
    if (t < 16) {
        float w[RESIZE_FAST_TAPS];
#pragma unroll
        for (int j = 0; j < RESIZE_FAST_TAPS; ++j)
            w[j] = j < tp.taps_y ? __ldg(tp.wy + (y_first + t) * tp.taps_y + j) : 0.f;
        ytab[2 * t] = make_float4(w[0], w[1], w[2], w[3]);
        ytab[2 * t + 1] = make_float4(w[4], __int_as_float(__ldg(tp.ymin + y_first + t) - r0), 0.f, 0.f);
    }
    __syncthreads();
    // B: horizontal pass, one output column per thread
    if (t < S) {
        float w[RESIZE_FAST_TAPS];
        const int x0 = __ldg(tp.xmin + t);
#pragma unroll
        for (int k = 0; k < RESIZE_FAST_TAPS; ++k) w[k] = k < tp.taps_x ? __ldg(tp.wx + t * tp.taps_x + k) : 0.f;
        // the 5 taps x 3 channels are 15 consecutive bytes from 3*x0 (taps past the right edge have weight 0, so what
        // lies there - the row padding or the next staged row - does not matter): five aligned words, funnel-shifted
        // into place, instead of fifteen byte loads (the kernel is bound by shared-memory instructions)
        const int b0 = 3 * x0;
        const int sh = (b0 & 3) * 8;
        const uint8_t* base = src + (b0 & ~3);
#pragma unroll 4
        for (int r = 0; r < nrows; ++r) {
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(base + r * src_pitch);
            const uint32_t q0 = wp[0], q1 = wp[1], q2 = wp[2], q3 = wp[3], q4 = wp[4];
            const uint32_t a0 = __funnelshift_r(q0, q1, sh), a1 = __funnelshift_r(q1, q2, sh);
            const uint32_t a2 = __funnelshift_r(q2, q3, sh), a3 = __funnelshift_r(q3, q4, sh);
            // byte i of the 15-byte run = pixel i/3, channel i%3
            float cr = 0.f, cg = 0.f, cb = 0.f;
            cr += w[0] * float(a0 & 0xff);         cg += w[0] * float((a0 >> 8) & 0xff);  cb += w[0] * float((a0 >> 16) & 0xff);
            cr += w[1] * float(a0 >> 24);          cg += w[1] * float(a1 & 0xff);         cb += w[1] * float((a1 >> 8) & 0xff);
            cr += w[2] * float((a1 >> 16) & 0xff); cg += w[2] * float(a1 >> 24);          cb += w[2] * float(a2 & 0xff);
            cr += w[3] * float((a2 >> 8) & 0xff);  cg += w[3] * float((a2 >> 16) & 0xff); cb += w[3] * float(a2 >> 24);
            cr += w[4] * float(a3 & 0xff);         cg += w[4] * float((a3 >> 8) & 0xff);  cb += w[4] * float((a3 >> 16) & 0xff);
            hbuf[(0 * max_rows + r) * S + t] = cr;
            hbuf[(1 * max_rows + r) * S + t] = cg;
            hbuf[(2 * max_rows + r) * S + t] = cb;
        }
    }
    __syncthreads();
    // C: vertical pass, normalise, assemble the patch-matrix rows in shared memory (over the dead source rows)
    __nv_bfloat16* out_s = reinterpret_cast<__nv_bfloat16*>(src);  // [ns][768]
    if (t < S) {
        const float mu[3] = {mean.x, mean.y, mean.z}, is[3] = {inv_std.x, inv_std.y, inv_std.z};
        __nv_bfloat16* o = out_s + (t >> 4) * 768 + (t & 15);
        const float* hx = hbuf + t;
        const int plane = max_rows * S;
#pragma unroll 2
        for (int ky = 0; ky < 16; ++ky) {
            const float4 wa = ytab[2 * ky], wb = ytab[2 * ky + 1];
            const int yo = __float_as_int(wb.y);
            const float wy[RESIZE_FAST_TAPS] = {wa.x, wa.y, wa.z, wa.w, wb.x};
            int rr[RESIZE_FAST_TAPS];
#pragma unroll
            for (int j = 0; j < RESIZE_FAST_TAPS; ++j) rr[j] = min(yo + j, nrows - 1) * S;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < RESIZE_FAST_TAPS; ++j) acc += wy[j] * hx[c * plane + rr[j]];
                o[c * 256 + ky * 16] = __float2bfloat16_rn((acc * (1.0f / 255.0f) - mu[c]) * is[c]);
            }
        }
    }
    __syncthreads();
    // D: ns consecutive rows of A are one contiguous block
    uint4* dst = reinterpret_cast<uint4*>(A + ((long long)f * ns * ns + (long long)py * ns) * 768);
    const uint4* so = reinterpret_cast<const uint4*>(out_s);
    for (int i = t; i < ns * 96; i += 256) dst[i] = so[i];
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

// ---------------------------------------------------------------------------------------------------------------
// Generic patch sizes (DINOv2-with-registers uses 14-pixel patches: transformers Dinov2WithRegistersPatchEmbeddings,
// a stride-P Conv2d, so the grid is floor(side / P) and the right / bottom remainder of the frame is not read).
// The patch-matrix row pitch Kp is P*P (or 3*P*P) rounded up to 64; the padding columns are zeroed once at create.
// One thread per (frame, patch row, ky, patch column): P pixels in, P bf16 out.
template <bool PLANE>
__global__ void __launch_bounds__(256)
preprocess_green_generic_kernel(const void* __restrict__ src, __nv_bfloat16* __restrict__ A, int n_frames, int H, int W,
                                long long frame_stride, int row_stride, int P, int ns, int Kp, int pix_stride) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n_frames * ns * P * ns;
    if (idx >= total) return;
    const int px = (int)(idx % ns);
    long long r = idx / ns;
    const int ky = (int)(r % P); r /= P;
    const int py = (int)(r % ns);
    const int f = (int)(r / ns);
    const int y = py * P + ky, x0 = px * P;
    __nv_bfloat16* dst = A + ((long long)f * ns * ns + (long long)py * ns + px) * Kp + ky * P;
    if (PLANE) {
        const float* g = reinterpret_cast<const float*>(src) + ((long long)f * H + y) * W + x0;
        for (int kx = 0; kx < P; ++kx) dst[kx] = __float2bfloat16_rn(g[kx] * 255.0f);
    } else {
        // pix_stride 3: interleaved RGB, take green; 1: the frame IS the green plane
        const uint8_t* g = reinterpret_cast<const uint8_t*>(src) + (long long)f * frame_stride + (long long)y * row_stride +
                           x0 * pix_stride + (pix_stride == 3 ? 1 : 0);
        for (int kx = 0; kx < P; ++kx) dst[kx] = __float2bfloat16_rn((float)g[pix_stride * kx]);
    }
}

// PROCESSOR mode for any patch size: one thread per output pixel (all three channels), separable antialiased
// bilinear taps exactly as preprocess_resize_kernel applies them, ImageNet normalisation, scatter into the patch
// matrix A[frame*Np + py*ns + px][c*P*P + ky*P + kx].
__global__ void __launch_bounds__(256)
preprocess_resize_generic_kernel(const uint8_t* __restrict__ frames, __nv_bfloat16* __restrict__ A, int n_frames, int H,
                                 int W, long long frame_stride, int row_stride, int S, ResizeTaps tp, float3 mean,
                                 float3 istd, int P, int ns, int Kp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int used = ns * P;  // pixels of the resized image that fall inside the patch grid
    const long long total = (long long)n_frames * used * used;
    if (idx >= total) return;
    const int x = (int)(idx % used);
    const int y = (int)((idx / used) % used);
    const int f = (int)(idx / ((long long)used * used));
    const uint8_t* img = frames + (long long)f * frame_stride;
    const int y0 = tp.ymin[y], x0 = tp.xmin[x];
    float acc[3] = {0.f, 0.f, 0.f};
    for (int j = 0; j < tp.taps_y; ++j) {
        const float wy = tp.wy[y * tp.taps_y + j];
        if (wy == 0.f) continue;
        const int sy = min(y0 + j, H - 1);
        const uint8_t* row = img + (long long)sy * row_stride;
        float h[3] = {0.f, 0.f, 0.f};
        for (int i = 0; i < tp.taps_x; ++i) {
            const float wx = tp.wx[x * tp.taps_x + i];
            const int sx = min(x0 + i, W - 1);
            h[0] = fmaf(wx, (float)row[3 * sx], h[0]);
            h[1] = fmaf(wx, (float)row[3 * sx + 1], h[1]);
            h[2] = fmaf(wx, (float)row[3 * sx + 2], h[2]);
        }
        acc[0] = fmaf(wy, h[0], acc[0]); acc[1] = fmaf(wy, h[1], acc[1]); acc[2] = fmaf(wy, h[2], acc[2]);
    }
    const float mu[3] = {mean.x, mean.y, mean.z}, is[3] = {istd.x, istd.y, istd.z};
    __nv_bfloat16* dst = A + ((long long)f * ns * ns + (long long)(y / P) * ns + x / P) * Kp + (y % P) * P + (x % P);
    for (int c = 0; c < 3; ++c) dst[c * P * P] = __float2bfloat16_rn((acc[c] * (1.0f / 255.0f) - mu[c]) * is[c]);
}

// h[(f*T + prefix + p), :] += pos[p, :]   (learned absolute position embedding of the patch tokens)
__global__ void __launch_bounds__(256)
add_pos_embed_kernel(float* __restrict__ h, const float* __restrict__ pos, int n_frames, int T, int prefix, int Np,
                     int D) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int d4 = D / 4;
    if (idx >= (long long)n_frames * Np * d4) return;
    const int c = (int)(idx % d4);
    const int pidx = (int)((idx / d4) % Np);
    const int f = (int)(idx / ((long long)d4 * Np));
    float4* dst = reinterpret_cast<float4*>(h + ((long long)f * T + prefix + pidx) * D) + c;
    const float4 a = *dst, b = __ldg(reinterpret_cast<const float4*>(pos + (long long)pidx * D) + c);
    *dst = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

}  // namespace cbas
