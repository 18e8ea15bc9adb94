/*
 * ctr_random.cpp -- counter-based drop-in for the reference's src/random_generator.cpp.
 *
 * TEST INFRASTRUCTURE ONLY. Linked into oracle/_ref/fray_ref_ctr INSTEAD of the reference's
 * random_generator.cpp (every other reference translation unit is compiled unmodified from
 * /root/reference/src). It implements the interface declared in the reference's
 * src/random_generator.h:33-57 on top of the contract in oracle/fray_rng.h, so that the
 * reference's own pathtrace / raytrace / shaders / lights / camera draw the same numbers the
 * GPU draws for the same (seed, pixel, sample).
 *
 * All `Random` objects are handles onto ONE thread-local stream: the reference copies the
 * generator in RendMT::entry (src/main.cpp:333) while hemisphereSample, getDOFRay,
 * RectLight::getNthSample and glossy Reflection::shade fetch it again by thread id
 * (src/main.cpp:96, src/camera.cpp:83, src/lights.cpp:60, src/shading.cpp:177); under this file
 * both routes consume the same sequential stream, which is the contract's "branch 0" stream.
 * The driver (ctr_driver.cpp) re-keys the stream before every pixel sample.
 */
#include <math.h>
#include "random_generator.h"   // the reference's header, found via -I/root/reference/src
#include "constants.h"
#include "../fray_rng.h"

static thread_local FrayRng tl_stream;
static unsigned g_seed = 42;

void fray_ctr_rekey(unsigned pixel, unsigned sample)
{
	fray_rng_init(&tl_stream, g_seed, pixel, sample, 0);
}

unsigned fray_ctr_draws(void) { return tl_stream.count; }

Random::Random(unsigned) {}
void Random::seed(unsigned) {}
unsigned Random::_next(void) { return fray_rng_next(&tl_stream); }
int Random::randint(int a, int b) { return fray_rng_int(&tl_stream, a, b); }
float Random::randfloat(void) { return fray_rng_float(&tl_stream); }
double Random::randdouble(void) { return fray_rng_double(&tl_stream); }

double Random::gaussian(double mean, double sigma)
{
	// not on the render path (no caller in src/); Box-Muller for completeness
	double u1 = 1.0 - randdouble(), u2 = randdouble();
	return mean + sigma * sqrt(-2.0 * log(u1)) * cos(2 * PI * u2);
}

void Random::unitDiscSample(double& x, double& y)
{
	// same arithmetic as src/random_generator.cpp:71-80
	double angle = randdouble() * 2 * PI;
	double rad = sqrt(randdouble());
	x = sin(angle) * rad;
	y = cos(angle) * rad;
}

void initRandom(unsigned seed)
{
	g_seed = seed;
	// the scene parser's randfloat()/randint() macros use generator 0 (src/scene.cpp:405):
	// give the parsing thread a stream of its own.
	fray_rng_init(&tl_stream, g_seed, 0xFFFFFFFFu, 0xFFFFFFFFu, 0);
}

Random& getRandomGen(int)
{
	static thread_local Random handle;
	return handle;
}

Random& getRandomGen()
{
	return getRandomGen(0);
}
