// fp32-mode E-step kernel instantiations for NW = 4 warps per CTA (estep2f_kernel.cuh).
#include "estep2f_kernel.cuh"

namespace mrlda {
const Estep2fVariant g_estep2f_variants_W4[] = {
    MRLDA_VARIANT2F(4, 14),
    MRLDA_VARIANT2F(4, 16),
    MRLDA_VARIANT2F(4, 18),
    MRLDA_VARIANT2F(4, 20),
    MRLDA_VARIANT2F(4, 22),
    MRLDA_VARIANT2F(4, 24),
    MRLDA_VARIANT2F(4, 26),
};
const int g_estep2f_variants_W4_n = 7;
}  // namespace mrlda
