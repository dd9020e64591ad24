// Cluster variants of the register/TMEM-stationary E-step: 8 CTAs of 8 warps per document at half the column slice of the 4-CTA variants (longer documents) (estep2_kernel.cuh).
#include "estep2_kernel.cuh"

namespace mrlda {
const Estep2Variant g_estep2_variants_X8[] = {
    MRLDA_VARIANT2CN(8, 10, 8),
    MRLDA_VARIANT2CN(8, 12, 8),
    MRLDA_VARIANT2CN(8, 14, 8),
    MRLDA_VARIANT2CN(8, 16, 8),
};
const int g_estep2_variants_X8_n = 4;
}  // namespace mrlda
