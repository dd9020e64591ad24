// Register-stationary E-step kernel instantiations for NW = 2 warps per CTA (estep2_kernel.cuh).
#include "estep2_kernel.cuh"

namespace mrlda {
const Estep2Variant g_estep2_variants_W2[] = {
    MRLDA_VARIANT2(2, 14, 8),
    MRLDA_VARIANT2(2, 16, 8),
    MRLDA_VARIANT2(2, 18, 8),
    MRLDA_VARIANT2(2, 20, 8),
    MRLDA_VARIANT2(2, 22, 8),
    MRLDA_VARIANT2(2, 24, 8),
    MRLDA_VARIANT2(2, 26, 8),
};
const int g_estep2_variants_W2_n = 7;
}  // namespace mrlda
