// Register-stationary E-step kernel instantiations for NW = 16 warps per CTA (estep2_kernel.cuh).
#include "estep2_kernel.cuh"

namespace mrlda {
const Estep2Variant g_estep2_variants_W16[] = {
    MRLDA_VARIANT2(16, 14, 16),  // 16 warps x 128 registers: measured slower than 8 x 255
};
const int g_estep2_variants_W16_n = 1;
}  // namespace mrlda
