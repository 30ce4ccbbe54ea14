// K1 generic kernel, parity build: compiled with -fmad=false so that no a*b+c is contracted into an FMA.
#define K1_NS k1_strict
#define K1_KERNEL_NAME k1_generic_strict_kernel
#define K1_LAUNCH_NAME launch_k1_generic_strict
#include "k1_generic.cuh"
