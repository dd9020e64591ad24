// E-step kernel instantiations for L = 1 lanes per row (see estep_kernel.cuh).
#include "estep_kernel.cuh"

namespace mrlda {
const EstepVariant g_estep_variants_L1[] = {
    MRLDA_VARIANT(1, 6),
    MRLDA_VARIANT(1, 10),
    MRLDA_VARIANT(1, 14),
    MRLDA_VARIANT(1, 18),
    MRLDA_VARIANT(1, 22),
    MRLDA_VARIANT(1, 26),
    MRLDA_VARIANT(1, 30),
};
const int g_estep_variants_L1_n = 7;
}  // namespace mrlda
