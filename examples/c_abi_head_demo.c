/* Pure C host for libcbas_b200.so: no Python, no torch - the drop-in boundary by itself.
 *
 *   c_abi_head_demo <dir>
 *
 * reads <dir>/cfg.txt (in_features out_features seq_len lstm_hidden lstm_layers use_acceleration n_frames temperature),
 * the head weights <dir>/<name>.f32 in the reference's state_dict tensors (row-major float32, names as in
 * include/cbas_b200.h) and <dir>/emb.f16 (n_frames x in_features IEEE half, the `cls` dataset as stored), runs
 * cbas_b200_head_infer + cbas_b200_actogram_bins on the default stream and writes <dir>/probs.f32 and <dir>/bins.i32.
 * tests/test_c_abi_gpu.py builds it with gcc, runs it on the GPU box and compares with the Python mirror.
 *
 *   gcc -O2 -I include examples/c_abi_head_demo.c -o c_abi_head_demo -L cbas_b200 -lcbas_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/cbas_b200
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cbas_b200.h"

/* the few CUDA runtime entry points this file needs (declared here so that plain gcc can compile it) */
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern int cudaDeviceSynchronize(void);
enum { H2D = 1, D2H = 2 };

static char path[4096];
static const char* in(const char* dir, const char* name) {
    snprintf(path, sizeof path, "%s/%s", dir, name);
    return path;
}

static void* load_to_device(const char* dir, const char* name, size_t bytes) {
    FILE* f = fopen(in(dir, name), "rb");
    if (!f) { fprintf(stderr, "missing %s\n", path); exit(2); }
    void* host = malloc(bytes);
    if (fread(host, 1, bytes, f) != bytes) { fprintf(stderr, "short read %s\n", path); exit(2); }
    fclose(f);
    void* dev = NULL;
    if (cudaMalloc(&dev, bytes) || cudaMemcpy(dev, host, bytes, H2D)) { fprintf(stderr, "cuda copy failed\n"); exit(3); }
    free(host);
    return dev;
}

#define CHECK(call)                                                                       \
    do {                                                                                  \
        if ((call) != 0) {                                                                \
            fprintf(stderr, "%s failed: %s\n", #call, cbas_b200_last_error());            \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int main(int argc, char** argv) {
    if (argc != 2) { fprintf(stderr, "usage: %s <dir>\n", argv[0]); return 2; }
    const char* dir = argv[1];
    int F, C, T, Hs, L, acc;
    long long n;
    float temperature;
    FILE* f = fopen(in(dir, "cfg.txt"), "r");
    if (!f || fscanf(f, "%d %d %d %d %d %d %lld %f", &F, &C, &T, &Hs, &L, &acc, &n, &temperature) != 8) {
        fprintf(stderr, "bad cfg.txt\n");
        return 2;
    }
    fclose(f);
    if (cbas_b200_abi_version() != CBAS_B200_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 2; }

    cbas_head_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.in_features = F; cfg.out_features = C; cfg.seq_len = T; cfg.bottleneck = 128; cfg.lstm_hidden = Hs;
    cfg.center_window = 5; cfg.ema_alpha = 0.3f; cfg.use_acceleration = acc; cfg.lstm_layers = L;

    cbas_head_weights w;
    memset(&w, 0, sizeof w);
    const size_t B = 128, aug = (acc ? 3 : 2) * B;
#define LD(field, name, count) w.field = (const float*)load_to_device(dir, name ".f32", (size_t)(count) * 4)
    LD(cls_w, "cls_bottleneck.0.weight", B * F);     LD(cls_b, "cls_bottleneck.0.bias", B);
    LD(delta_w, "delta_bottleneck.0.weight", B * F); LD(delta_b, "delta_bottleneck.0.bias", B);
    LD(cls_ln_g, "cls_ln.weight", B);     LD(cls_ln_b, "cls_ln.bias", B);
    LD(delta_ln_g, "delta_ln.weight", B); LD(delta_ln_b, "delta_ln.bias", B);
    if (acc) {
        LD(acc_w, "acc_bottleneck.0.weight", B * F); LD(acc_b, "acc_bottleneck.0.bias", B);
        LD(acc_ln_g, "acc_ln.weight", B);            LD(acc_ln_b, "acc_ln.bias", B);
    }
    LD(lin0_w, "lin0.0.weight", 256 * aug); LD(lin0_b, "lin0.0.bias", 256);
    LD(lin1_w, "lin1.weight", C * F);       LD(lin1_b, "lin1.bias", C);
    LD(lin2_w, "lin2.weight", C * 2 * Hs);  LD(lin2_b, "lin2.bias", C);
    LD(att_w, "attention_head.weight", 2 * Hs); LD(att_b, "attention_head.bias", 1);
    LD(w_ih_f, "lstm.weight_ih_l0", 4 * Hs * 256); LD(w_hh_f, "lstm.weight_hh_l0", 4 * Hs * Hs);
    LD(b_ih_f, "lstm.bias_ih_l0", 4 * Hs);         LD(b_hh_f, "lstm.bias_hh_l0", 4 * Hs);
    LD(w_ih_r, "lstm.weight_ih_l0_reverse", 4 * Hs * 256); LD(w_hh_r, "lstm.weight_hh_l0_reverse", 4 * Hs * Hs);
    LD(b_ih_r, "lstm.bias_ih_l0_reverse", 4 * Hs);         LD(b_hh_r, "lstm.bias_hh_l0_reverse", 4 * Hs);
    if (L == 2) {
        LD(w_ih_f1, "lstm.weight_ih_l1", 4 * Hs * 2 * Hs); LD(w_hh_f1, "lstm.weight_hh_l1", 4 * Hs * Hs);
        LD(b_ih_f1, "lstm.bias_ih_l1", 4 * Hs);            LD(b_hh_f1, "lstm.bias_hh_l1", 4 * Hs);
        LD(w_ih_r1, "lstm.weight_ih_l1_reverse", 4 * Hs * 2 * Hs); LD(w_hh_r1, "lstm.weight_hh_l1_reverse", 4 * Hs * Hs);
        LD(b_ih_r1, "lstm.bias_ih_l1_reverse", 4 * Hs);            LD(b_hh_r1, "lstm.bias_hh_l1_reverse", 4 * Hs);
    }
#undef LD
    float scalars[2];
    f = fopen(in(dir, "scalars.f32"), "rb");  /* gate, attention_temp */
    if (!f || fread(scalars, 4, 2, f) != 2) { fprintf(stderr, "bad scalars.f32\n"); return 2; }
    fclose(f);
    w.gate = scalars[0];
    w.attention_temp = scalars[1];

    cbas_head* head = NULL;
    CHECK(cbas_b200_head_create(&cfg, &w, &head));
    void* emb = load_to_device(dir, "emb.f16", (size_t)n * F * 2);
    float* probs_dev = NULL;
    int32_t* bins_dev = NULL;
    const long long bin_frames = 50, n_bins = (n + bin_frames - 1) / bin_frames;
    if (cudaMalloc((void**)&probs_dev, (size_t)n * C * 4) || cudaMalloc((void**)&bins_dev, (size_t)n_bins * 4)) return 3;
    CHECK(cbas_b200_head_infer(head, emb, n, temperature, probs_dev, NULL, NULL /* default stream */));
    CHECK(cbas_b200_actogram_bins(probs_dev, n, C, 0, 0.1f, bin_frames, bins_dev, NULL));
    if (cudaDeviceSynchronize()) { fprintf(stderr, "kernel failure\n"); return 3; }

    float* probs = (float*)malloc((size_t)n * C * 4);
    int32_t* bins = (int32_t*)malloc((size_t)n_bins * 4);
    cudaMemcpy(probs, probs_dev, (size_t)n * C * 4, D2H);
    cudaMemcpy(bins, bins_dev, (size_t)n_bins * 4, D2H);
    f = fopen(in(dir, "probs.f32"), "wb"); fwrite(probs, 4, (size_t)n * C, f); fclose(f);
    f = fopen(in(dir, "bins.i32"), "wb");  fwrite(bins, 4, (size_t)n_bins, f); fclose(f);
    printf("ok: %lld frames, %d behaviours, %lld bins, %lld launches\n", n, C, n_bins, (long long)cbas_b200_launch_count());
    cbas_b200_head_destroy(head);
    cudaFree(emb); cudaFree(probs_dev); cudaFree(bins_dev);
    return 0;
}
