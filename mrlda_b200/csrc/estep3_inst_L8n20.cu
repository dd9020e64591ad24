// estep3_inst_L8n20.cu — instantiations of the warp-group E-step kernel (estep3_kernel.cuh) for up to
// 160 topics: 8 lanes per row, 20 columns per lane; 8 warps x 2 register batches.
#include "estep3_kernel.cuh"

namespace mrlda {
const Estep3Variant g_estep3_variants_L8n20[] = {
    MRLDA_VARIANT3(8, 20, 2, 8),
};
const int g_estep3_variants_L8n20_n = sizeof(g_estep3_variants_L8n20) / sizeof(g_estep3_variants_L8n20[0]);
}  // namespace mrlda
