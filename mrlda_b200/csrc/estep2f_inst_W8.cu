// fp32-mode E-step kernel instantiations for NW = 8 warps per CTA (estep2f_kernel.cuh).
#include "estep2f_kernel.cuh"

namespace mrlda {
const Estep2fVariant g_estep2f_variants_W8[] = {
    MRLDA_VARIANT2F(8, 14),
    MRLDA_VARIANT2F(8, 16),
    MRLDA_VARIANT2F(8, 18),
    MRLDA_VARIANT2F(8, 20),
    MRLDA_VARIANT2F(8, 22),
    MRLDA_VARIANT2F(8, 24),
    MRLDA_VARIANT2F(8, 26),
    MRLDA_VARIANT2F(8, 28),
    MRLDA_VARIANT2F(8, 30),
    MRLDA_VARIANT2F(8, 32),
};
const int g_estep2f_variants_W8_n = 10;
}  // namespace mrlda
