// fp32-mode E-step kernel instantiations for NW = 1 warps per CTA (estep2f_kernel.cuh).
#include "estep2f_kernel.cuh"

namespace mrlda {
const Estep2fVariant g_estep2f_variants_W1[] = {
    MRLDA_VARIANT2F(1, 4),
    MRLDA_VARIANT2F(1, 6),
    MRLDA_VARIANT2F(1, 8),
    MRLDA_VARIANT2F(1, 10),
    MRLDA_VARIANT2F(1, 12),
    MRLDA_VARIANT2F(1, 14),
    MRLDA_VARIANT2F(1, 16),
    MRLDA_VARIANT2F(1, 18),
    MRLDA_VARIANT2F(1, 20),
    MRLDA_VARIANT2F(1, 22),
    MRLDA_VARIANT2F(1, 24),
    MRLDA_VARIANT2F(1, 26),
};
const int g_estep2f_variants_W1_n = 12;
}  // namespace mrlda
