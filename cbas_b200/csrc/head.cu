// LSTM classifier head + actogram binning (placeholder translation unit while the encoder is brought up).
#include "../../include/cbas_b200.h"
#include "common.h"

using namespace cbas;

struct cbas_head {
    cbas_head_cfg cfg;
};

extern "C" {

int cbas_b200_head_create(const cbas_head_cfg*, const cbas_head_weights*, cbas_head**) {
    return fail("head kernels not built yet");
}
void cbas_b200_head_destroy(cbas_head* h) { delete h; }
int cbas_b200_head_infer(cbas_head*, const void*, int64_t, float, float*, float*, void*) {
    return fail("head kernels not built yet");
}
int cbas_b200_actogram_bins(const float*, int64_t, int32_t, int32_t, float, int64_t, int32_t*, void*) {
    return fail("actogram kernel not built yet");
}
}
