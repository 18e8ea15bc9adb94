// render_fp64.cu -- parity-precision (R = double) instantiation of the render kernels.
// Compiled with -fmad=false: the reference's FP64 arithmetic (g++, no FMA contraction in the golden build) is
// reproduced operation for operation, so images agree with the CPU oracle to the last few float ulps.
#define FRAY_PARITY_UNIT 1
#include "render_kernels.cuh"
namespace fray {
FRAY_DEFINE_LAUNCHERS(double)
}
