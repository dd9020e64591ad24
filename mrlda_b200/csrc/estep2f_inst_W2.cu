// fp32-mode E-step kernel instantiations for NW = 2 warps per CTA (estep2f_kernel.cuh).
#include "estep2f_kernel.cuh"

namespace mrlda {
const Estep2fVariant g_estep2f_variants_W2[] = {
    MRLDA_VARIANT2F(2, 14),
    MRLDA_VARIANT2F(2, 16),
    MRLDA_VARIANT2F(2, 18),
    MRLDA_VARIANT2F(2, 20),
    MRLDA_VARIANT2F(2, 22),
    MRLDA_VARIANT2F(2, 24),
    MRLDA_VARIANT2F(2, 26),
};
const int g_estep2f_variants_W2_n = 7;
}  // namespace mrlda
