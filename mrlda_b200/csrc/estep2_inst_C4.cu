// Cluster variants of the register/TMEM-stationary E-step: 4 CTAs per document (estep2_kernel.cuh).
#include "estep2_kernel.cuh"

namespace mrlda {
const Estep2Variant g_estep2_variants_C4[] = {
    MRLDA_VARIANT2C(18, 4),
    MRLDA_VARIANT2C(20, 4),
    MRLDA_VARIANT2C(22, 4),
    MRLDA_VARIANT2C(24, 4),
    MRLDA_VARIANT2C(26, 4),
    MRLDA_VARIANT2C(28, 4),
    MRLDA_VARIANT2C(30, 4),
    MRLDA_VARIANT2C(32, 4),
};
const int g_estep2_variants_C4_n = 8;
}  // namespace mrlda
