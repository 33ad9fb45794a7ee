// tcgen05 (kind::tf32) flavour of the fused KL-NMF pass -- placeholder until the tensor-core
// kernel lands; the ABI refuses the math mode loudly instead of silently falling back.
#include "sal_common.cuh"

int sal_launch_pass_tf32(sal_ctx* c, const PassArgs& a, cudaStream_t st) {
    (void)c, (void)a, (void)st;
    sal_set_error("SAL_MATH_TF32 pass is not built into this library");
    return SAL_EUNSUPPORTED;
}
