// estep3c_inst_C2o.cu — instantiations of the cluster warp-group E-step kernel (estep3c_kernel.cuh):
// 2 CTAs of 8 lanes x 25 columns per row, up to 400 topics; per CTA
// 8 warps x 2 register batches and the Tensor Memory / shared-memory batches of estep3_kernel.cuh.
#include "estep3c_kernel.cuh"

namespace mrlda {
const Estep3cVariant g_estep3c_variants_C2o[] = {
    MRLDA_VARIANT3C(8, 25, 2, 8, 2),
};
const int g_estep3c_variants_C2o_n = sizeof(g_estep3c_variants_C2o) / sizeof(g_estep3c_variants_C2o[0]);
}  // namespace mrlda
