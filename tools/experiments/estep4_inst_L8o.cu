// estep4_inst_L8o.cu — instantiation of the paired E-step kernel (estep4_kernel.cuh) for
// 192 < K <= 200 topics: 8 lanes per row, 25 columns per lane, ten batch slots per warp.
#include "estep4_kernel.cuh"

namespace mrlda {
const Estep4Variant g_estep4_variants[] = {
    MRLDA_VARIANT4(8, 25, 8),
};
const int g_estep4_variants_n = sizeof(g_estep4_variants) / sizeof(g_estep4_variants[0]);
}  // namespace mrlda
