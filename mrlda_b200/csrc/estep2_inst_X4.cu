// Cluster variants of the register/TMEM-stationary E-step: 4 CTAs of 8 warps per document at half the column slice of the 2-CTA variants (longer documents) (estep2_kernel.cuh).
#include "estep2_kernel.cuh"

namespace mrlda {
const Estep2Variant g_estep2_variants_X4[] = {
    MRLDA_VARIANT2CN(8, 10, 4),
    MRLDA_VARIANT2CN(8, 12, 4),
    MRLDA_VARIANT2CN(8, 14, 4),
    MRLDA_VARIANT2CN(8, 16, 4),
};
const int g_estep2_variants_X4_n = 4;
}  // namespace mrlda
