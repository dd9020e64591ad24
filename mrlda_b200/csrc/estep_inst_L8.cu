// E-step kernel instantiations for L = 8 lanes per row (see estep_kernel.cuh).
#include "estep_kernel.cuh"

namespace mrlda {
const EstepVariant g_estep_variants_L8[] = {
    MRLDA_VARIANT(8, 14),
    MRLDA_VARIANT(8, 16),
    MRLDA_VARIANT(8, 18),
    MRLDA_VARIANT(8, 20),
    MRLDA_VARIANT(8, 22),
    MRLDA_VARIANT(8, 24),
    MRLDA_VARIANT(8, 26),
    MRLDA_VARIANT(8, 28),
    MRLDA_VARIANT(8, 30),
    MRLDA_VARIANT(8, 32),
};
const int g_estep_variants_L8_n = 10;
}  // namespace mrlda
