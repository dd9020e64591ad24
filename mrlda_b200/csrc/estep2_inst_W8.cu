// Register-stationary E-step kernel instantiations for NW = 8 warps per CTA (estep2_kernel.cuh).
#include "estep2_kernel.cuh"

namespace mrlda {
const Estep2Variant g_estep2_variants_W8[] = {
    MRLDA_VARIANT2(8, 14, 8),
    MRLDA_VARIANT2(8, 16, 8),
    MRLDA_VARIANT2(8, 18, 8),
    MRLDA_VARIANT2(8, 20, 8),
    MRLDA_VARIANT2(8, 22, 8),
    MRLDA_VARIANT2(8, 24, 8),
    MRLDA_VARIANT2(8, 26, 8),
    MRLDA_VARIANT2(8, 28, 8),
    MRLDA_VARIANT2(8, 30, 8),
    MRLDA_VARIANT2(8, 32, 8),
    MRLDA_VARIANT2R(8, 26, 3, 8),  // profiling alternative (MRLDA_ESTEP2_VARIANT=8,26,3)
};
const int g_estep2_variants_W8_n = 11;
}  // namespace mrlda
