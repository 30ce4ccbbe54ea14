// k1_interior.cu — optimised K1 for BF_INTERIOR blocks (placeholder: forwards to the generic fast kernel).
#include "ludwig_internal.h"
namespace ludwig {
void launch_k1_interior(const K1Args& a, cudaStream_t s) { launch_k1_generic_fast(a, s); }
}  // namespace ludwig
