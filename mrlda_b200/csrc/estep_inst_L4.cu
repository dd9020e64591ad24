// E-step kernel instantiations for L = 4 lanes per row (see estep_kernel.cuh).
#include "estep_kernel.cuh"

namespace mrlda {
const EstepVariant g_estep_variants_L4[] = {
    MRLDA_VARIANT(4, 18),
    MRLDA_VARIANT(4, 22),
    MRLDA_VARIANT(4, 26),
    MRLDA_VARIANT(4, 30),
};
const int g_estep_variants_L4_n = 4;
}  // namespace mrlda
