// Register-stationary E-step kernel instantiations for NW = 1 warps per CTA (estep2_kernel.cuh).
#include "estep2_kernel.cuh"

namespace mrlda {
const Estep2Variant g_estep2_variants_W1[] = {
    MRLDA_VARIANT2(1, 4, 8),
    MRLDA_VARIANT2(1, 6, 8),
    MRLDA_VARIANT2(1, 8, 8),
    MRLDA_VARIANT2(1, 10, 8),
    MRLDA_VARIANT2(1, 12, 8),
    MRLDA_VARIANT2(1, 14, 8),
    MRLDA_VARIANT2(1, 16, 8),
    MRLDA_VARIANT2(1, 18, 8),
    MRLDA_VARIANT2(1, 20, 8),
    MRLDA_VARIANT2(1, 22, 8),
    MRLDA_VARIANT2(1, 24, 8),
    MRLDA_VARIANT2(1, 26, 8),
};
const int g_estep2_variants_W1_n = 12;
}  // namespace mrlda
