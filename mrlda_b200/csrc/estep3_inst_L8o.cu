// estep3_inst_L8o.cu — instantiations of the warp-group E-step kernel (estep3_kernel.cuh) for
// 192 < K <= 200 topics: 8 lanes per row, 25 columns per lane (12 pairs + one single column, no
// padding column at K = 200); 8 warps x 2 register batches, five Tensor Memory batches per warp.
#include "estep3_kernel.cuh"

namespace mrlda {
const Estep3Variant g_estep3_variants_L8o[] = {
    MRLDA_VARIANT3(8, 25, 2, 8),
};
const int g_estep3_variants_L8o_n = sizeof(g_estep3_variants_L8o) / sizeof(g_estep3_variants_L8o[0]);
}  // namespace mrlda
