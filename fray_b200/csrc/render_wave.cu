// render_wave.cu -- instantiates the wavefront Whitted kernels (wave_kernels.cuh) for the fast precision.
// Compiled like render_fp32.cu: -prec-div=false -prec-sqrt=false -ftz=true.
#define FRAY_WAVE_IMPL
#include "wave_kernels.cuh"

namespace fray {

cudaError_t launchWaveFrame(const DScene<float>& sc, WaveParams p, int features, WaveLaunch& cfg)
{
	if ((features & ~WaveVariants::kPlain) == 0) return launchWaveFrameT<WaveVariants::kPlain>(sc, p, cfg);
	if ((features & ~WaveVariants::kTextured) == 0) return launchWaveFrameT<WaveVariants::kTextured>(sc, p, cfg);
	return cudaErrorInvalidValue;
}

} // namespace fray
