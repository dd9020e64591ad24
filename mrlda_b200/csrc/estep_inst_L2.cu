// E-step kernel instantiations for L = 2 lanes per row (see estep_kernel.cuh).
#include "estep_kernel.cuh"

namespace mrlda {
const EstepVariant g_estep_variants_L2[] = {
    MRLDA_VARIANT(2, 18),
    MRLDA_VARIANT(2, 22),
    MRLDA_VARIANT(2, 26),
    MRLDA_VARIANT(2, 30),
};
const int g_estep_variants_L2_n = 4;
}  // namespace mrlda
