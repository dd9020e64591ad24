// estep3c_inst_C5o.cu — instantiations of the cluster warp-group E-step kernel (estep3c_kernel.cuh):
// 5 CTAs of 8 lanes x 25 columns per row, up to 1000 topics (K = 1000 is 5 x 200, no padding column); per CTA
// 8 warps x 2 register batches and the Tensor Memory / shared-memory batches of estep3_kernel.cuh.
#include "estep3c_kernel.cuh"

namespace mrlda {
const Estep3cVariant g_estep3c_variants_C5o[] = {
    MRLDA_VARIANT3C(8, 25, 2, 8, 5),
};
const int g_estep3c_variants_C5o_n = sizeof(g_estep3c_variants_C5o) / sizeof(g_estep3c_variants_C5o[0]);
}  // namespace mrlda
