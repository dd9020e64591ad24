// estep3_inst_L8.cu — instantiations of the warp-group E-step kernel (estep3_kernel.cuh) for
// 104 < K <= 208 topics: 8 lanes per row, 26 columns per lane; 8 warps x 2 register batches.
#include "estep3_kernel.cuh"

namespace mrlda {
const Estep3Variant g_estep3_variants_L8[] = {
    MRLDA_VARIANT3(8, 26, 2, 8),
};
const int g_estep3_variants_L8_n = sizeof(g_estep3_variants_L8) / sizeof(g_estep3_variants_L8[0]);
}  // namespace mrlda
