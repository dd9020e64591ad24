// Register-stationary E-step kernel instantiations for NW = 4 warps per CTA (estep2_kernel.cuh).
#include "estep2_kernel.cuh"

namespace mrlda {
const Estep2Variant g_estep2_variants_W4[] = {
    MRLDA_VARIANT2(4, 14, 8),
    MRLDA_VARIANT2(4, 16, 8),
    MRLDA_VARIANT2(4, 18, 8),
    MRLDA_VARIANT2(4, 20, 8),
    MRLDA_VARIANT2(4, 22, 8),
    MRLDA_VARIANT2(4, 24, 8),
    MRLDA_VARIANT2(4, 26, 8),
};
const int g_estep2_variants_W4_n = 7;
}  // namespace mrlda
