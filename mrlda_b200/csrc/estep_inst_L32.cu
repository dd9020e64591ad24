// E-step kernel instantiations for L = 32 lanes per row (see estep_kernel.cuh).
#include "estep_kernel.cuh"

namespace mrlda {
const EstepVariant g_estep_variants_L32[] = {
    MRLDA_VARIANT(32, 18),
    MRLDA_VARIANT(32, 20),
    MRLDA_VARIANT(32, 22),
    MRLDA_VARIANT(32, 24),
    MRLDA_VARIANT(32, 26),
    MRLDA_VARIANT(32, 28),
    MRLDA_VARIANT(32, 30),
    MRLDA_VARIANT(32, 32),
};
const int g_estep_variants_L32_n = 8;
}  // namespace mrlda
