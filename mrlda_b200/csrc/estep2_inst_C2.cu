// Cluster variants of the register/TMEM-stationary E-step: 2 CTAs per document (estep2_kernel.cuh).
#include "estep2_kernel.cuh"

namespace mrlda {
const Estep2Variant g_estep2_variants_C2[] = {
    MRLDA_VARIANT2C(18, 2),
    MRLDA_VARIANT2C(20, 2),
    MRLDA_VARIANT2C(22, 2),
    MRLDA_VARIANT2C(24, 2),
    MRLDA_VARIANT2C(26, 2),
    MRLDA_VARIANT2C(28, 2),
    MRLDA_VARIANT2C(30, 2),
    MRLDA_VARIANT2C(32, 2),
};
const int g_estep2_variants_C2_n = 8;
}  // namespace mrlda
