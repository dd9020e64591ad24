// estep3_inst_L8n17.cu — instantiations of the warp-group E-step kernel (estep3_kernel.cuh) for up to
// 136 topics: 8 lanes per row, 17 columns per lane; 8 warps x 2 register batches.
#include "estep3_kernel.cuh"

namespace mrlda {
const Estep3Variant g_estep3_variants_L8n17[] = {
    MRLDA_VARIANT3(8, 17, 2, 8),
};
const int g_estep3_variants_L8n17_n = sizeof(g_estep3_variants_L8n17) / sizeof(g_estep3_variants_L8n17[0]);
}  // namespace mrlda
