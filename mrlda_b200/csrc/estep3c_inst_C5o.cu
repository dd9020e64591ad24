// estep3c_inst_C5o.cu — instantiations of the cluster warp-group E-step kernel (estep3c_kernel.cuh)
// for 800 < K <= 1000 topics: five CTAs of 8 lanes x 25 columns per row (K = 1000 is 5 x 200, no
// padding column); per CTA 8 warps x 2 register batches, five Tensor Memory batches per warp.
#include "estep3c_kernel.cuh"

namespace mrlda {
const Estep3cVariant g_estep3c_variants[] = {
    MRLDA_VARIANT3C(8, 25, 2, 8, 5),
};
const int g_estep3c_variants_n = sizeof(g_estep3c_variants) / sizeof(g_estep3c_variants[0]);
}  // namespace mrlda
