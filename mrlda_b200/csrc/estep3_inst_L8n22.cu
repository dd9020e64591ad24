// estep3_inst_L8n22.cu — instantiations of the warp-group E-step kernel (estep3_kernel.cuh) for up to
// 176 topics: 8 lanes per row, 22 columns per lane; 8 warps x 2 register batches.
#include "estep3_kernel.cuh"

namespace mrlda {
const Estep3Variant g_estep3_variants_L8n22[] = {
    MRLDA_VARIANT3(8, 22, 2, 8),
};
const int g_estep3_variants_L8n22_n = sizeof(g_estep3_variants_L8n22) / sizeof(g_estep3_variants_L8n22[0]);
}  // namespace mrlda
