// E-step kernel instantiations for L = 16 lanes per row (see estep_kernel.cuh).
#include "estep_kernel.cuh"

namespace mrlda {
const EstepVariant g_estep_variants_L16[] = {
    MRLDA_VARIANT(16, 14),
    MRLDA_VARIANT(16, 16),
    MRLDA_VARIANT(16, 18),
    MRLDA_VARIANT(16, 20),
    MRLDA_VARIANT(16, 22),
    MRLDA_VARIANT(16, 24),
    MRLDA_VARIANT(16, 26),
    MRLDA_VARIANT(16, 28),
    MRLDA_VARIANT(16, 30),
    MRLDA_VARIANT(16, 32),
};
const int g_estep_variants_L16_n = 10;
}  // namespace mrlda
