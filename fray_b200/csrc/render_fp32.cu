// render_fp32.cu -- fast-precision (R = float) instantiation of the render kernels. Compiled with FMA contraction on.
#include "render_kernels.cuh"
namespace fray {
FRAY_DEFINE_LAUNCHERS(float)
}
