// estep3_inst_L8b.cu — warp-group E-step kernel, 12- and 16-warp CTAs for 104 < K <= 208 topics
// (no register batches: the tile lives in Tensor Memory and shared memory).
#include "estep3_kernel.cuh"

namespace mrlda {
const Estep3Variant g_estep3_variants_L8b[] = {
    MRLDA_VARIANT3(8, 26, 0, 12),
    MRLDA_VARIANT3(8, 26, 0, 16),
};
const int g_estep3_variants_L8b_n = sizeof(g_estep3_variants_L8b) / sizeof(g_estep3_variants_L8b[0]);
}  // namespace mrlda
