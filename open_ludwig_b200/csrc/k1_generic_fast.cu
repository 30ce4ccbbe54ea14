// K1 generic kernel, fast build (FMA contraction on): runs the non-interior blocks in fast mode.
#define K1_NS k1_fastgen
#define K1_KERNEL_NAME k1_generic_fast_kernel
#define K1_LAUNCH_NAME launch_k1_generic_fast
#include "k1_generic.cuh"
