/*
 * Minimal stand-in for the three OpenEXR headers the fray reference includes
 * (src/bitmap.cpp:28-30, used at :238-284). TEST INFRASTRUCTURE ONLY.
 *
 * OpenEXR is not installed in this image. Reading: `Imf::RgbaInputFile("x.exr")` opens the
 * side-car "x.exr.f32" = int32 w, int32 h, float32 rgba[h][w][4] that oracle/mirror_data.py
 * produced by decoding the real file with OpenCV's bundled OpenEXR (PIZ + HALF handled there).
 * Values round-trip through `half` in the real library; the side-car already holds the
 * half-precision values widened to float, so the reference sees identical texels.
 * Writing: `Imf::RgbaOutputFile` emits the same raw layout (nothing in the oracle flow needs it).
 */
#ifndef FRAY_ORACLE_EXR_SHIM_H
#define FRAY_ORACLE_EXR_SHIM_H

#include <stdio.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace Iex {
struct BaseExc {
	std::string msg;
	BaseExc(const char* m = ""): msg(m) {}
};
}

namespace Imath {
struct V2i { int x, y; };
struct Box2i { V2i min, max; };
}

namespace Imf {

struct Rgba { float r, g, b, a; };
enum RgbaChannels { WRITE_RGBA = 15 };

template <class T>
class Array2D {
	std::vector<T> d;
	long sx = 0, sy = 0;
public:
	void resizeErase(long sizeX, long sizeY) { sx = sizeX; sy = sizeY; d.assign((size_t) sx * sy, T()); }
	T* operator[](long x) { return d.data() + x * sy; }
	const T* operator[](long x) const { return d.data() + x * sy; }
};

class RgbaInputFile {
	int w = 0, h = 0;
	std::vector<Rgba> px;
	Rgba* base = nullptr;
	size_t xs = 1, ys = 0;
public:
	RgbaInputFile(const char* fn)
	{
		std::string side = std::string(fn) + ".f32";
		FILE* f = fopen(side.c_str(), "rb");
		if (!f) throw Iex::BaseExc("missing EXR side-car");
		int32_t wh[2];
		if (fread(wh, sizeof(wh), 1, f) != 1) { fclose(f); throw Iex::BaseExc("short side-car"); }
		w = wh[0]; h = wh[1];
		px.resize((size_t) w * h);
		size_t got = fread(px.data(), sizeof(Rgba), px.size(), f);
		fclose(f);
		if (got != px.size()) throw Iex::BaseExc("short side-car");
	}
	Imath::Box2i dataWindow() const { return Imath::Box2i{ {0, 0}, {w - 1, h - 1} }; }
	void setFrameBuffer(Rgba* b, size_t xStride, size_t yStride) { base = b; xs = xStride; ys = yStride; }
	void readPixels(int y0, int y1)
	{
		for (int y = y0; y <= y1; y++)
			for (int x = 0; x < w; x++)
				base[x * xs + y * ys] = px[(size_t) y * w + x];
	}
};

class RgbaOutputFile {
	FILE* f;
	int w, h;
	const Rgba* base = nullptr;
	size_t xs = 1, ys = 0;
public:
	RgbaOutputFile(const char* fn, int width, int height, RgbaChannels): w(width), h(height)
	{
		f = fopen(fn, "wb");
		if (!f) throw Iex::BaseExc("cannot open output");
		int32_t wh[2] = { w, h };
		fwrite(wh, sizeof(wh), 1, f);
	}
	~RgbaOutputFile() { if (f) fclose(f); }
	void setFrameBuffer(const Rgba* b, size_t xStride, size_t yStride) { base = b; xs = xStride; ys = yStride; }
	void writePixels(int n)
	{
		for (int y = 0; y < n; y++)
			for (int x = 0; x < w; x++)
				fwrite(&base[x * xs + y * ys], sizeof(Rgba), 1, f);
	}
};

} // namespace Imf

#endif
