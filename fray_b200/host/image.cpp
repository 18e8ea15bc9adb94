// image.cpp -- float RGB bitmaps and BMP file IO of the host layer.
//
// Semantics follow class Bitmap of the reference (/root/reference/src/bitmap.h:30-60, src/bitmap.cpp:36-315):
// texel (0,0) is the top-left corner, reads outside the image return black, BMP rows are stored bottom-up and
// flipped at load, 8-bit files go through their BGRA palette, channel value = byte / 255.0f, the writer emits
// 24-bit BGR rows padded to 4 bytes behind a 54-byte header with the byte-exact field values of saveBMP, and
// channel -> byte conversion is floor(clamp01(x) * 255 + 0.5) (src/color.h:29-34, src/util.h:38).
// EXR lives in exr.cpp.
#include <cstdio>
#include <cstring>
#include <cstdint>

#include "scene.h"

namespace fray {

void Bitmap::generateEmptyImage(int w, int h)
{
	data.clear();
	width = height = -1;
	if (w <= 0 || h <= 0) return;
	width = w;
	height = h;
	data.assign((size_t) w * h, Color(0, 0, 0));
}

Color Bitmap::getPixel(int x, int y) const
{
	if (data.empty() || x < 0 || x >= width || y < 0 || y >= height) return Color(0, 0, 0);
	return data[x + (size_t) y * width];
}

void Bitmap::setPixel(int x, int y, const Color& c)
{
	if (data.empty() || x < 0 || x >= width || y < 0 || y >= height) return;
	data[x + (size_t) y * width] = c;
}

void Bitmap::differentiate()
{
	// forward differences of the channel mean with wrap-around; (dx, dy, 0) replaces the texel
	std::vector<Color> out((size_t) width * height);
	for (int y = 0; y < height; y++)
		for (int x = 0; x < width; x++) {
			float here = getPixel(x, y).intensity();
			float dx = here - getPixel((x + 1) % width, y).intensity();
			float dy = here - getPixel(x, (y + 1) % height).intensity();
			out[x + (size_t) y * width] = Color(dx, dy, 0);
		}
	data.swap(out);
}

namespace {
#pragma pack(push, 1)
struct BmpFileHeader { // after the 2-byte "BM" signature
	int32_t fileSize;
	int32_t reserved;
	int32_t imageOffset;
};
struct BmpInfoHeader {
	int32_t headerSize;
	int32_t width, height;
	uint16_t planes;
	uint16_t bitsPerPixel;
	int32_t compression;
	int32_t imageSize;
	int32_t xPelsPerMeter, yPelsPerMeter;
	int32_t colorsUsed, colorsImportant;
};
#pragma pack(pop)
const uint16_t kBmpMagic = 19778; // "BM"

struct FileCloser {
	FILE* f;
	~FileCloser() { if (f) fclose(f); }
};
} // namespace

bool Bitmap::loadBMP(const char* filename)
{
	data.clear();
	width = height = -1;
	FILE* fp = fopen(filename, "rb");
	if (!fp) {
		printf("loadBMP: Can't open file: `%s'\n", filename);
		return false;
	}
	FileCloser closer{ fp };
	uint16_t sign;
	BmpFileHeader fh;
	BmpInfoHeader ih;
	if (!fread(&sign, 2, 1, fp)) return false;
	if (sign != kBmpMagic) {
		printf("loadBMP: `%s' is not a BMP file.\n", filename);
		return false;
	}
	if (!fread(&fh, sizeof(fh), 1, fp) || !fread(&ih, sizeof(ih), 1, fp)) return false;
	if (!(ih.bitsPerPixel == 8 || ih.bitsPerPixel == 24 || ih.bitsPerPixel == 32)) {
		printf("loadBMP: Cannot handle file format at %d bpp.\n", ih.bitsPerPixel);
		return false;
	}
	if (ih.planes != 1) {
		printf("loadBMP: cannot load multichannel .bmp!\n");
		return false;
	}
	Color palette[256];
	int paletteEntries = 0;
	if (ih.bitsPerPixel <= 8) {
		paletteEntries = ih.colorsUsed ? ih.colorsUsed : (1 << ih.bitsPerPixel);
		if (paletteEntries > 256) return false;
		for (int i = 0; i < paletteEntries; i++) {
			uint32_t bgra;
			if (!fread(&bgra, 1, 4, fp)) return false;
			palette[i] = Color(((bgra >> 16) & 0xff) / 255.0f, ((bgra >> 8) & 0xff) / 255.0f, (bgra & 0xff) / 255.0f);
		}
	}
	fseek(fp, fh.imageOffset - (54 + paletteEntries * 4), SEEK_CUR);
	const int bytesPerPixel = ih.bitsPerPixel / 8;
	int rowSize = ih.width * bytesPerPixel;
	if (rowSize % 4) rowSize = (rowSize / 4 + 1) * 4;
	generateEmptyImage(ih.width, ih.height);
	if (!isOK()) {
		printf("loadBMP: cannot allocate memory for bitmap! Check file integrity!\n");
		return false;
	}
	std::vector<unsigned char> row(rowSize);
	for (int y = ih.height - 1; y >= 0; y--) {
		if (!fread(row.data(), 1, rowSize, fp)) {
			printf("loadBMP: short read while opening `%s', file is probably incomplete!\n", filename);
			data.clear();
			width = height = -1;
			return false;
		}
		for (int x = 0; x < ih.width; x++) {
			const unsigned char* p = &row[(size_t) x * bytesPerPixel];
			if (ih.bitsPerPixel > 8) setPixel(x, y, Color(p[2] / 255.0f, p[1] / 255.0f, p[0] / 255.0f));
			else setPixel(x, y, palette[p[0]]);
		}
	}
	return true;
}

static inline unsigned to8bit(float x)
{
	if (x < 0) x = 0;
	if (x > 1) x = 1;
	return (unsigned) (int) floorf(x * 255.0f + 0.5f);
}

bool Bitmap::saveBMP(const char* filename) const
{
	FILE* fp = fopen(filename, "wb");
	if (!fp) return false;
	FileCloser closer{ fp };
	int rowSize = width * 3;
	if (rowSize % 4) rowSize += 4 - (rowSize % 4);
	BmpFileHeader fh;
	BmpInfoHeader ih;
	fh.fileSize = rowSize * height + 54;
	fh.reserved = 0;
	fh.imageOffset = 54;
	ih.headerSize = 40;
	ih.width = width;
	ih.height = height;
	ih.planes = 1;
	ih.bitsPerPixel = 24;
	ih.compression = ih.imageSize = 0;
	ih.xPelsPerMeter = ih.yPelsPerMeter = 0;
	ih.colorsUsed = ih.colorsImportant = 0;
	fwrite(&kBmpMagic, 2, 1, fp);
	fwrite(&fh, sizeof(fh), 1, fp);
	fwrite(&ih, sizeof(ih), 1, fp);
	std::vector<unsigned char> row(rowSize, 0);
	for (int y = height - 1; y >= 0; y--) {
		for (int x = 0; x < width; x++) {
			Color c = getPixel(x, y);
			row[x * 3] = (unsigned char) to8bit(c.b);
			row[x * 3 + 1] = (unsigned char) to8bit(c.g);
			row[x * 3 + 2] = (unsigned char) to8bit(c.r);
		}
		fwrite(row.data(), rowSize, 1, fp);
	}
	return true;
}

static std::string upperExtension(const char* fn) // src/util.cpp:41-55
{
	const char* dot = strrchr(fn, '.');
	std::string r;
	if (strlen(fn) < 2 || !dot) return r;
	for (const char* p = dot + 1; *p; p++) r += (char) toupper((unsigned char) *p);
	return r;
}

bool Bitmap::loadImage(const char* fn)
{
	std::string ext = upperExtension(fn);
	if (ext == "BMP") return loadBMP(fn);
	if (ext == "EXR") return loadEXR(fn);
	return false;
}

bool Bitmap::saveImage(const char* fn) const
{
	std::string ext = upperExtension(fn);
	if (ext == "BMP") return saveBMP(fn);
	if (ext == "EXR") return saveEXR(fn);
	return false;
}

} // namespace fray
