/* cbas_b200 - C ABI of the B200-native CBAS hot path (libcbas_b200.so).
 *
 * The reference (jones-lab-tamu/CBAS) has no FFI: its seam is Python.  Each entry point below names the
 * reference Python symbol (file:line under the reference checkout) whose GPU work it replaces; the Python
 * mirror in cbas_b200/ keeps those symbols' signatures and binds this library with ctypes.
 *
 * Conventions: plain pointers and sizes only; every *_dev pointer is a CUDA device pointer in the calling
 * process' primary context; `stream` is a cudaStream_t passed as void* (PyTorch: torch.cuda.current_stream()
 * .cuda_stream); calls are asynchronous on that stream unless stated.  Every function returns 0 on success,
 * non-zero on failure; cbas_b200_last_error() returns a thread-local message for the last failure.
 * The library is re-entrant across host threads as long as two threads do not share one Encoder/Head handle
 * (workthreads.py:272,370 - one EncodeThread and one ClassificationThread, each on its own CUDA stream).
 * A handle belongs to the CUDA device that was current when it was created; every call on the handle makes that
 * device current for its duration and restores the caller's, so one process can drive several GPUs from several
 * threads (workthreads.py:267-272,1245-1273).  The kernel-level entry points (no handle) run on the calling thread's
 * current device.  Test knobs are either per handle (cbas_b200_encoder_set_option) or per host thread.
 */
#ifndef CBAS_B200_H
#define CBAS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBAS_B200_ABI_VERSION 3

/* preprocessing modes (SURVEY.md 8a rows P / P') */
#define CBAS_PRE_REFERENCE 0 /* cbas.py:431 + cbas.py:672-675: green/255 replicated x3, native resolution   */
#define CBAS_PRE_PROCESSOR 1 /* HF image_processing_dinov3_vit.py:45-86: rescale, AA-bilinear resize, normalise */

const char* cbas_b200_last_error(void);
int cbas_b200_abi_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
unsigned long long cbas_b200_launch_count(void);

/* Per-kernel device timing for bench.py's roofline leg: while enabled, every launch is bracketed by CUDA events
 * on its own stream (tags: 0 preprocess, 1 patch GEMM, 2 LayerNorm, 3 QKV GEMM, 4 attention, 5 proj GEMM,
 * 6 up GEMM(+GELU), 7 down GEMM, 8 final LN, 9.. head stages, 16 actogram, 17 other).  enable(1) also clears.
 * Tag 2 counts the one statistics pass after the embedding: norm1 / norm2 themselves are fused into the GEMMs.
 * profile_read synchronises on the recorded events and returns the summed duration and launch count of a tag. */
int cbas_b200_profile_enable(int on);
int cbas_b200_profile_read(int tag, double* total_ms, long long* launches);

/* ------------------------------------------------------------------------------------------------ encoder
 * Replaces DinoEncoder.forward (cbas.py:672-677) = preprocessing + transformers DINOv3ViTModel.forward
 * (modeling_dinov3_vit.py:530-555) + CLS pooling.  Weights are device pointers prepared once by the host
 * mirror (cbas_b200/encoder.py) from the HF state_dict; the library keeps the pointers, not copies. */
typedef struct {
    int32_t hidden;        /* D: 384 / 768 / 1024 (multiple of 128)                                   */
    int32_t layers;        /* L                                                                       */
    int32_t heads;         /* D / 64 (head_dim must be 64)                                            */
    int32_t intermediate;  /* I                                                                       */
    int32_t prefix_tokens; /* 1 CLS + num_register_tokens                                             */
    int32_t mode;          /* CBAS_PRE_*                                                              */
    int32_t in_h, in_w;    /* source frame size                                                       */
    int32_t side;          /* ViT input side in pixels: in_h (= in_w) for REFERENCE, resize target for PROCESSOR */
    int32_t max_frames;    /* largest n a forward call may pass (workspace is sized for it)           */
    float ln_eps;
    int32_t resize_taps_y, resize_taps_x; /* PROCESSOR only */
    int32_t patch;         /* patch size in pixels: 16 (DINOv3; 0 means 16) or 14 (DINOv2-with-registers).  The grid is
                              floor(side / patch) patches per side, like the stride-`patch` Conv2d it replaces        */
} cbas_encoder_cfg;

/* norm1 / norm2 (modeling_dinov3_vit.py:433,445) are fused into the GEMMs around them (csrc/gemm_tcgen05.cuh): the
 * host folds gamma into the weight that follows the norm, W' = W * gamma (per input column), and passes
 *   c1[n] = sum_k W'[n,k]  (of the bf16-rounded W'),   c2[n] = sum_k W[n,k] * beta[k] + bias[n]
 * so that  LN(h) W^T + bias = rstd * (h - mean) W'^T + c2 = rstd * ((h - s) W'^T) - rstd * (mean - s) * c1 + c2. */
typedef struct {
    const void* w_qkv; const void* c1_qkv; const void* b_qkv; /* bf16 [3D,D] (q|k|v rows) with norm1's gamma folded in,
                                                                 f32 [3D] c1, f32 [3D] c2 (= bias + W beta; k bias zero) */
    const void* w_o;   const void* b_o;       /* bf16 [D,D], f32 [D]   - LayerScale lambda1 folded in  */
    const void* w_up;  const void* c1_up; const void* b_up;   /* bf16 [I,D] with norm2's gamma folded in, f32 [I] c1, c2 */
    const void* w_down; const void* b_down;   /* bf16 [D,I], f32 [D]   - LayerScale lambda2 folded in  */
} cbas_layer_weights;

typedef struct {
    const void* w_patch; const void* b_patch; /* bf16 [D,Kp], f32 [D].  Kp = patch^2 (REFERENCE: channels folded) or
                                                 3*patch^2 (PROCESSOR), rounded up to a multiple of 64, zero padded */
    const void* prefix;                       /* f32 [prefix_tokens, D]: cls_token then register_tokens */
    const void* rope_cos; const void* rope_sin; /* f32 [Np, 32], or both null: no rotary embedding (DINOv2)  */
    const void* lnf_g; const void* lnf_b;     /* f32 [D] final norm                                    */
    const cbas_layer_weights* layers;         /* host array [L]                                        */
    /* PROCESSOR mode resize taps (device): first source index + normalised weights per output coordinate */
    const void* rs_ymin; const void* rs_wy; const void* rs_xmin; const void* rs_wx;
    /* learned absolute position embedding of the patch tokens, f32 [Np, D], already interpolated to this grid
     * (Dinov2WithRegistersEmbeddings.interpolate_pos_encoding), or null (DINOv3).  The CLS position is folded into
     * prefix row 0 by the host. */
    const void* pos_embed;
} cbas_encoder_weights;

typedef struct cbas_encoder cbas_encoder;

int cbas_b200_encoder_create(const cbas_encoder_cfg* cfg, const cbas_encoder_weights* w, cbas_encoder** out);
void cbas_b200_encoder_destroy(cbas_encoder* enc);

/* Per-handle test knobs (every setting computes the same function; the parity tests run them against each other). */
#define CBAS_OPT_ATTENTION_IMPL 0   /* 0 = automatic (tcgen05 whenever a tcgen05 kernel covers the token count),
                                       1 = mma.sync kernel, 2 = tcgen05 kernel                                    */
#define CBAS_OPT_PRUNE_LAST_LAYER 1 /* 1 (default) = in the last block compute only what the pooled CLS row needs (K/V
                                       for all tokens, the rest for the CLS rows), 0 = run it on every token       */
#define CBAS_OPT_RESIZE_KERNEL 2    /* PROCESSOR mode: 2 (default) = column-per-thread kernel when the geometry allows,
                                       1 = general shared-memory tiled kernel, 0 = per-pixel kernel                 */
#define CBAS_OPT_LN_FUSION 3        /* 1 = norm1 / norm2 fused into the GEMMs around them (csrc/gemm_tcgen05.cuh),
                                       2 = norm1 only (down GEMM -> next block's QKV GEMM), norm2 stays a kernel,
                                       0 = standalone LayerNorm kernels between the GEMMs (one HBM pass each)       */
#define CBAS_OPT_SERPENTINE 4       /* standalone-LayerNorm path: 1 (default) = consecutive kernels of a block walk the rows
                                       in opposite directions (each starts on what its predecessor left in L2), 0 = all
                                       ascending                                                                    */
int cbas_b200_encoder_set_option(cbas_encoder* enc, int32_t option, int32_t value);

/* frames_dev: uint8 RGB HWC (what decord's get_batch(...).asnumpy() yields, cbas.py:425), n frames,
 * frame_stride / row_stride in bytes.  emb_out_dev: f32 [n, D] pooled CLS embedding after the final norm. */
int cbas_b200_encoder_forward_u8(cbas_encoder* enc, const uint8_t* frames_dev, int32_t n, int64_t frame_stride,
                                 int32_t row_stride, float* emb_out_dev, void* stream);
/* REFERENCE mode only: the frames as ONE uint8 plane each (the green channel, cbas.py:431 `frames_np[:, :, :, 1]`),
 * n planes of in_h x in_w bytes, strides in bytes - a third of the host-to-device traffic of the RGB frames. */
int cbas_b200_encoder_forward_u8_plane(cbas_encoder* enc, const uint8_t* planes_dev, int32_t n, int64_t frame_stride,
                                       int32_t row_stride, float* emb_out_dev, void* stream);
/* DinoEncoder.__call__ compatibility (cbas.py:435,672): planes_dev f32 [n, in_h, in_w] in [0,1] (REFERENCE only). */
int cbas_b200_encoder_forward_f32(cbas_encoder* enc, const float* planes_dev, int32_t n, float* emb_out_dev,
                                  void* stream);
/* Debug/test taps: copy of the residual stream f32 [n*T, D] after `after_layer` blocks (0 = after embeddings). */
int cbas_b200_encoder_debug_hidden(cbas_encoder* enc, const uint8_t* frames_dev, int32_t n, int64_t frame_stride,
                                   int32_t row_stride, int32_t after_layer, float* hidden_out_dev, void* stream);

/* ------------------------------------------------------------------------------------- kernel-level entry points
 * (used by the parity tests and by bench.py's roofline leg; same kernels the encoder launches) */
/* out = epilogue(A[M,K] * W[N,K]^T + bias); epi: 0 bias->bf16, 1 bias+GELU(erf)->bf16, 2 f32 out += , 4 bias->f32,
 * 5 bias+GELU->f32 */
int cbas_b200_gemm_bf16(const void* a_dev, const void* w_dev, const float* bias_dev, void* out_dev, int32_t M,
                        int32_t N, int32_t K, int32_t epi, void* stream);
/* Fused LayerNorm, kernel level (what the encoder chains per block; csrc/gemm_tcgen05.cuh).  Row statistics are
 * 36 floats per row: [0,16) partial sums of y = h - shift, [16,32) partial sums of y^2, [32] shift, [33,36) padding.
 *   ln_stats_init: exact statistics of h f32 [rows, D] (shift = row mean) and hb = bf16(h - shift)
 *   gemm_resid_ln: h += A[M,K] W[N,K]^T + bias in place (N = D <= 1024), hb_out = bf16(h - new shift), stats_out from
 *                  stats_in (the new shift is the exact row mean before the update); stats_in != stats_out
 *   gemm_ln_a:     out bf16 [M,N] = epi(LN(h) W^T + b) computed from hb / stats with the folded weight W' = W * gamma,
 *                  c1[n] = sum_k W'[n,k], c2[n] = sum_k W[n,k] beta[k] + b[n]; K = D; epi 0 (bias) or 1 (bias + GELU) */
int cbas_b200_ln_stats_init(float* h_dev, void* hb_out_dev, float* stats_out_dev, int32_t rows, int32_t D,
                            void* stream);
int cbas_b200_gemm_resid_ln(const void* a_dev, const void* w_dev, const float* bias_dev, float* h_dev, void* hb_out_dev,
                            const float* stats_in_dev, float* stats_out_dev, int32_t M, int32_t N, int32_t K,
                            void* stream);
int cbas_b200_gemm_ln_a(const void* hb_dev, const float* stats_dev, const void* w_folded_dev, const float* c1_dev,
                        const float* c2_dev, void* out_bf16_dev, int32_t M, int32_t N, int32_t K, int32_t epi,
                        float eps, void* stream);
/* Operand format of cbas_b200_attention_tc: 1 = the q and k thirds of the fused QKV activation are IEEE f16 like the
 * v third (frames of at most 256 tokens: the kernel's RoPE prologue and logits run on f16 operands - the reference runs
 * the block under fp16 autocast, cbas.py:433), 0 = q and k are bf16 and only v is f16 (the key-split kernel). */
int cbas_b200_attention_tc_qk_f16(int32_t T);
/* Profiling aid (per host thread): device buffer of 64*16 int64 that CTA 0 of the tcgen05 attention kernel fills with
 * clock64() stamps per pipeline stage (null switches it off). */
int cbas_b200_debug_attention_trace(void* trace_dev);
/* Test knob for cbas_b200_preprocess_resize (per host thread): 2 (default) = column-per-thread resize kernel when the
 * geometry allows (<= 256 output columns, <= 5 taps), 1 = general shared-memory tiled kernel, 0 = per-pixel kernel.
 * All three agree (1 and 2 bitwise). */
int cbas_b200_debug_resize_tiled(int32_t on);
/* Test knob for the GEMM launcher (per host thread): 0 = choose automatically (CTA pairs / tcgen05 cta_group::2 when
 * M >= 4096), 1 or 2 = force. */
int cbas_b200_debug_gemm_cta_group(int32_t cg);
int cbas_b200_layernorm(const float* in_dev, const float* gamma_dev, const float* beta_dev, void* out_bf16_dev,
                        int32_t rows, int32_t D, float eps, void* stream);
int cbas_b200_attention(const void* qkv_bf16_dev, void* out_bf16_dev, const float* rope_cos_dev,
                        const float* rope_sin_dev, int32_t frames, int32_t T, int32_t prefix, int32_t heads,
                        void* stream);
/* 1 when a tcgen05 attention kernel covers frames of T tokens (shared-memory and TMEM budget), else 0: the encoder
 * then uses cbas_b200_attention (mma.sync).  `rope` != 0: with RoPE tables. */
int cbas_b200_attention_tc_supported(int32_t T, int32_t prefix, int32_t rope);
/* tcgen05 attention (frames of <= 256 tokens).  q and k are bf16, the V third of the buffer is IEEE f16 (that is how
 * the encoder's QKV GEMM stores it for this kernel).  RoPE is applied in the kernel's prologue from the given tables
 * ([T - prefix, 32] f32); pass null tables to skip the rotation. */
int cbas_b200_attention_tc(const void* qkv_bf16_dev, void* out_bf16_dev, const float* rope_cos_dev,
                           const float* rope_sin_dev, int32_t frames, int32_t T, int32_t prefix, int32_t heads,
                           void* stream);
int cbas_b200_preprocess_green(const uint8_t* frames_dev, void* a_bf16_dev, int32_t n, int32_t H, int32_t W,
                               int64_t frame_stride, int32_t row_stride, void* stream);
int cbas_b200_preprocess_resize(const uint8_t* frames_dev, void* a_bf16_dev, int32_t n, int32_t H, int32_t W,
                                int64_t frame_stride, int32_t row_stride, int32_t side, const int32_t* ymin_dev,
                                const float* wy_dev, int32_t taps_y, const int32_t* xmin_dev, const float* wx_dev,
                                int32_t taps_x, void* stream);

/* ------------------------------------------------------------------------------------------------ head
 * Replaces ClassifierLSTMDeltas.forward (classifier_head.py:150-172) driven by infer_file's window loop
 * (cbas.py:497-551): for every frame f the window [f-seq_len/2, f+seq_len/2] with replicate padding at the
 * ends of the video, then softmax(logits / max(1e-3, temperature)).  All weights f32 device pointers in the
 * reference state_dict layout (workthreads.py:856 model.pth keys). */
typedef struct {
    int32_t in_features;   /* F (768)                    */
    int32_t out_features;  /* C (number of behaviours)   */
    int32_t seq_len;       /* odd window length (31)     */
    int32_t bottleneck;    /* 128                        */
    int32_t lstm_hidden;   /* 64 or 128                  */
    int32_t center_window; /* sw = 5                     */
    float ema_alpha;       /* 0.3                        */
    int32_t use_acceleration; /* 0: no acceleration stream (acc_* pointers ignored, lin0_w is [256,256]) */
    int32_t lstm_layers;      /* 1 or 2                  */
} cbas_head_cfg;

typedef struct {
    const float* cls_w; const float* cls_b;       /* [128,F], [128] cls_bottleneck.0   */
    const float* delta_w; const float* delta_b;   /* delta_bottleneck.0                */
    const float* acc_w; const float* acc_b;       /* acc_bottleneck.0                  */
    const float* cls_ln_g; const float* cls_ln_b; /* [128]                             */
    const float* delta_ln_g; const float* delta_ln_b;
    const float* acc_ln_g; const float* acc_ln_b;
    const float* lin0_w; const float* lin0_b;     /* [256,384] ([256,256] without acceleration), [256] */
    const float* lin1_w; const float* lin1_b;     /* [C,F], [C]                        */
    const float* lin2_w; const float* lin2_b;     /* [C,2Hs], [C]                      */
    const float* att_w; const float* att_b;       /* [1,2Hs], [1]                      */
    const float* w_ih_f; const float* w_hh_f; const float* b_ih_f; const float* b_hh_f; /* lstm.*_l0         */
    const float* w_ih_r; const float* w_hh_r; const float* b_ih_r; const float* b_hh_r; /* lstm.*_l0_reverse */
    float gate;            /* raw parameter (sigmoid applied inside) */
    float attention_temp;  /* raw parameter (softplus applied inside) */
    /* second LSTM layer (lstm_layers == 2; input width 2*Hs), else null */
    const float* w_ih_f1; const float* w_hh_f1; const float* b_ih_f1; const float* b_hh_f1; /* lstm.*_l1         */
    const float* w_ih_r1; const float* w_hh_r1; const float* b_ih_r1; const float* b_hh_r1; /* lstm.*_l1_reverse */
} cbas_head_weights;

typedef struct cbas_head cbas_head;

int cbas_b200_head_create(const cbas_head_cfg* cfg, const cbas_head_weights* w, cbas_head** out);
void cbas_b200_head_destroy(cbas_head* head);
/* emb_dev: f16 [n_frames, F] (the `cls` dataset as stored, cbas.py:420) ; probs_out_dev: f32 [n_frames, C].
 * logits_out_dev (optional, may be null): f32 [n_frames, C] final_logits before temperature/softmax. */
int cbas_b200_head_infer(cbas_head* head, const void* emb_f16_dev, int64_t n_frames, float temperature,
                         float* probs_out_dev, float* logits_out_dev, void* stream);

/* ClassifierLSTMDeltas.forward on arbitrary windows (classifier_head.py:150-172): x f32 [n_windows, seq_len, F]
 * -> final_logits f32 [n_windows, C] and (optional) rawm f32 [n_windows, 2*lstm_hidden]. */
int cbas_b200_head_forward_windows(cbas_head* head, const float* x_f32_dev, int64_t n_windows, float* logits_out_dev,
                                   float* rawm_out_dev, void* stream);

/* ------------------------------------------------------------------------------------------------ actogram
 * Replaces the numeric part of Actogram.__init__ (cbas.py:969-999): per frame
 * event = (p_b * [max_{b' != b} p_b' < p_b]) >= threshold ; bins[k] = sum of events over bin_frames frames
 * (last partial bin kept).  probs_dev f32 [n, C]; bins_out_dev int32 [ceil(n / bin_frames)]. */
int cbas_b200_actogram_bins(const float* probs_dev, int64_t n, int32_t C, int32_t behavior, float threshold,
                            int64_t bin_frames, int32_t* bins_out_dev, void* stream);
/* The same over float64 probabilities and a float64 threshold: what Actogram.__init__ computes on the tables pandas
 * parsed from the `_outputs.csv` files (cbas.py:989-993), so that a probability that rounds across the threshold in
 * float32 is counted as the reference counts it. */
int cbas_b200_actogram_bins_f64(const double* probs_dev, int64_t n, int32_t C, int32_t behavior, double threshold,
                                int64_t bin_frames, int32_t* bins_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CBAS_B200_H */
