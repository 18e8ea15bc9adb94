// exr.cpp -- OpenEXR scan-line reader / writer of the host layer (no OpenEXR library in this image).
//
// The reference loads and saves EXR through Imf::RgbaInputFile / RgbaOutputFile
// (/root/reference/src/bitmap.cpp:238-284): R, G, B of every pixel of the data window as float, and
// HALF RGBA with A = 1 on output. The bundled cube maps (data/env/forest/*.exr) are 256x256, channels
// A,B,G,R HALF, PIZ-compressed, increasing-Y scan lines, so a PIZ decoder is required.
// Implemented here from the published file-format description: single-part scan-line files, channel types
// HALF / FLOAT / UINT, compression NONE, ZIPS, ZIP (zlib) and PIZ (bitmap LUT + canonical Huffman + 2-D
// Haar-like wavelet on 16-bit words). Tiled, multi-part and deep files are rejected.
// tests/test_exr.py compares every texel with OpenCV's decode of the same files.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <zlib.h>

#include "scene.h"

namespace fray {
namespace {

float halfToFloat(uint16_t h)
{
	uint32_t sign = (uint32_t) (h & 0x8000u) << 16;
	uint32_t exp = (h >> 10) & 0x1f;
	uint32_t man = h & 0x3ffu;
	uint32_t bits;
	if (exp == 0) {
		if (man == 0) {
			bits = sign;
		} else { // subnormal: renormalise
			int e = -1;
			do {
				e++;
				man <<= 1;
			} while (!(man & 0x400u));
			bits = sign | ((uint32_t) (127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
		}
	} else if (exp == 31) {
		bits = sign | 0x7f800000u | (man << 13);
	} else {
		bits = sign | ((exp + 112) << 23) | (man << 13);
	}
	float f;
	memcpy(&f, &bits, 4);
	return f;
}

uint16_t floatToHalf(float f) // round to nearest even, overflow to infinity
{
	uint32_t x;
	memcpy(&x, &f, 4);
	uint32_t sign = (x >> 16) & 0x8000u;
	int32_t exp = (int32_t) ((x >> 23) & 0xff) - 127 + 15;
	uint32_t man = x & 0x7fffffu;
	if (((x >> 23) & 0xff) == 0xff) return (uint16_t) (sign | 0x7c00u | (man ? 0x200u : 0));
	if (exp >= 31) return (uint16_t) (sign | 0x7c00u);
	if (exp <= 0) {
		if (exp < -10) return (uint16_t) sign;
		man |= 0x800000u;
		int shift = 14 - exp;
		uint32_t half = man >> shift;
		uint32_t rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
		if (rem > mid || (rem == mid && (half & 1))) half++;
		return (uint16_t) (sign | half);
	}
	uint32_t half = ((uint32_t) exp << 10) | (man >> 13);
	uint32_t rem = man & 0x1fffu;
	if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
	return (uint16_t) (sign | half);
}

struct Channel {
	std::string name;
	int type; // 0 UINT, 1 HALF, 2 FLOAT
	int xs, ys;
};

// ---- PIZ: Huffman ----------------------------------------------------------------------------------

const int HUF_ENCBITS = 16, HUF_DECBITS = 14;
const int HUF_ENCSIZE = (1 << HUF_ENCBITS) + 1, HUF_DECSIZE = 1 << HUF_DECBITS, HUF_DECMASK = HUF_DECSIZE - 1;

struct HufDec {
	int len = 0, lit = 0;
	std::vector<int> longs;
};

struct BitReader {
	const uint8_t* p;
	const uint8_t* end;
	uint64_t c = 0;
	int lc = 0;
	bool ok = true;
	uint32_t get(int n)
	{
		while (lc < n) {
			if (p >= end) { ok = false; return 0; }
			c = (c << 8) | *p++;
			lc += 8;
		}
		lc -= n;
		return (uint32_t) ((c >> lc) & ((1u << n) - 1));
	}
};

bool hufUnpackTable(BitReader& br, int im, int iM, std::vector<uint64_t>& code)
{
	for (; im <= iM; im++) {
		uint32_t l = br.get(6);
		if (!br.ok) return false;
		code[im] = l;
		if (l == 63) { // long run of zero-length codes
			int run = (int) br.get(8) + 6;
			if (!br.ok || im + run > iM + 1) return false;
			while (run--) code[im++] = 0;
			im--;
		} else if (l >= 59) { // short run
			int run = (int) l - 59 + 2;
			if (im + run > iM + 1) return false;
			while (run--) code[im++] = 0;
			im--;
		}
	}
	// canonical code assignment: code = length | (value << 6)
	uint64_t n[59] = { 0 };
	for (int i = 0; i < HUF_ENCSIZE; i++) n[code[i]]++;
	uint64_t c = 0;
	for (int i = 58; i > 0; i--) {
		uint64_t nc = (c + n[i]) >> 1;
		n[i] = c;
		c = nc;
	}
	for (int i = 0; i < HUF_ENCSIZE; i++) {
		int l = (int) code[i];
		if (l > 0) code[i] = l | (n[l]++ << 6);
	}
	return true;
}

bool hufBuildDec(const std::vector<uint64_t>& code, int im, int iM, std::vector<HufDec>& dec)
{
	for (; im <= iM; im++) {
		uint64_t c = code[im] >> 6;
		int l = (int) (code[im] & 63);
		if (c >> l) return false;
		if (l > HUF_DECBITS) {
			HufDec& d = dec[c >> (l - HUF_DECBITS)];
			if (d.len) return false;
			d.lit++;
			d.longs.push_back(im);
		} else if (l) {
			size_t base = (size_t) (c << (HUF_DECBITS - l));
			for (size_t i = 0; i < ((size_t) 1 << (HUF_DECBITS - l)); i++) {
				dec[base + i].len = l;
				dec[base + i].lit = im;
			}
		}
	}
	return true;
}

bool hufDecode(const std::vector<uint64_t>& code, const std::vector<HufDec>& dec, const uint8_t* in, int nBits, int rlc,
               size_t no, uint16_t* out)
{
	uint64_t c = 0;
	int lc = 0;
	uint16_t* o = out;
	uint16_t* oe = out + no;
	const uint8_t* ie = in + (nBits + 7) / 8;
	auto emit = [&](int sym) -> bool {
		if (sym == rlc) {
			if (lc < 8) {
				if (in >= ie + 8) return false; // allow the reader to run into the zero padding
				c = (c << 8) | *in++;
				lc += 8;
			}
			lc -= 8;
			uint8_t run = (uint8_t) (c >> lc);
			if (o == out || o + run > oe) return false;
			uint16_t s = o[-1];
			while (run--) *o++ = s;
		} else {
			if (o >= oe) return false;
			*o++ = (uint16_t) sym;
		}
		return true;
	};
	while (in < ie) {
		c = (c << 8) | *in++;
		lc += 8;
		while (lc >= HUF_DECBITS) {
			const HufDec& d = dec[(c >> (lc - HUF_DECBITS)) & HUF_DECMASK];
			if (d.len) {
				lc -= d.len;
				if (!emit(d.lit)) return false;
			} else {
				if (d.longs.empty()) return false;
				size_t j;
				for (j = 0; j < d.longs.size(); j++) {
					int l = (int) (code[d.longs[j]] & 63);
					while (lc < l && in < ie) {
						c = (c << 8) | *in++;
						lc += 8;
					}
					if (lc >= l && (code[d.longs[j]] >> 6) == ((c >> (lc - l)) & (((uint64_t) 1 << l) - 1))) {
						lc -= l;
						if (!emit(d.longs[j])) return false;
						break;
					}
				}
				if (j == d.longs.size()) return false;
			}
		}
	}
	int i = (8 - nBits) & 7;
	c >>= i;
	lc -= i;
	while (lc > 0) {
		const HufDec& d = dec[(c << (HUF_DECBITS - lc)) & HUF_DECMASK];
		if (!d.len) return false;
		lc -= d.len;
		if (!emit(d.lit)) return false;
	}
	return o == oe;
}

bool hufUncompress(const uint8_t* in, size_t nIn, uint16_t* out, size_t nOut)
{
	if (nIn == 0) return nOut == 0;
	if (nIn < 20) return false;
	uint32_t hdr[5];
	memcpy(hdr, in, 20);
	int im = (int) hdr[0], iM = (int) hdr[1], nBits = (int) hdr[3];
	if (im < 0 || im >= HUF_ENCSIZE || iM < 0 || iM >= HUF_ENCSIZE) return false;
	// the compressed buffer is copied with slack so the bit reader may touch a few bytes past the end
	std::vector<uint8_t> buf(in + 20, in + nIn);
	buf.resize(buf.size() + 16, 0);
	BitReader br{ buf.data(), buf.data() + buf.size() };
	std::vector<uint64_t> code(HUF_ENCSIZE, 0);
	if (!hufUnpackTable(br, im, iM, code)) return false;
	const uint8_t* payload = br.p;
	if ((size_t) (payload - buf.data()) + (size_t) (nBits + 7) / 8 > nIn - 20) return false;
	std::vector<HufDec> dec(HUF_DECSIZE);
	if (!hufBuildDec(code, im, iM, dec)) return false;
	return hufDecode(code, dec, payload, nBits, iM, nOut, out);
}

// ---- PIZ: wavelet ----------------------------------------------------------------------------------

inline void wdec14(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
{
	int16_t ls = (int16_t) l, hs = (int16_t) h;
	int hi = hs;
	int ai = ls + (hi & 1) + (hi >> 1);
	a = (uint16_t) (int16_t) ai;
	b = (uint16_t) (int16_t) (ai - hi);
}

inline void wdec16(uint16_t l, uint16_t h, uint16_t& a, uint16_t& b)
{
	int m = l, d = h;
	int bb = (m - (d >> 1)) & 0xffff;
	int aa = (d + bb - 0x8000) & 0xffff;
	b = (uint16_t) bb;
	a = (uint16_t) aa;
}

void wav2Decode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx)
{
	const bool w14 = mx < (1 << 14);
	int n = nx > ny ? ny : nx;
	int p = 1, p2;
	while (p <= n) p <<= 1;
	p >>= 1;
	p2 = p;
	p >>= 1;
	auto dec = [&](uint16_t l, uint16_t h, uint16_t& a, uint16_t& b) { if (w14) wdec14(l, h, a, b); else wdec16(l, h, a, b); };
	while (p >= 1) {
		uint16_t* py = in;
		uint16_t* ey = in + (ptrdiff_t) oy * (ny - p2);
		const int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
		uint16_t i00, i01, i10, i11;
		for (; py <= ey; py += oy2) {
			uint16_t* px = py;
			uint16_t* ex = py + (ptrdiff_t) ox * (nx - p2);
			for (; px <= ex; px += ox2) {
				uint16_t* p01 = px + ox1;
				uint16_t* p10 = px + oy1;
				uint16_t* p11 = p10 + ox1;
				dec(*px, *p10, i00, i10);
				dec(*p01, *p11, i01, i11);
				dec(i00, i01, *px, *p01);
				dec(i10, i11, *p10, *p11);
			}
			if (nx & p) { // odd column
				uint16_t* p10 = px + oy1;
				dec(*px, *p10, i00, *p10);
				*px = i00;
			}
		}
		if (ny & p) { // odd line
			uint16_t* px = py;
			uint16_t* ex = py + (ptrdiff_t) ox * (nx - p2);
			for (; px <= ex; px += ox2) {
				uint16_t* p01 = px + ox1;
				dec(*px, *p01, i00, *p01);
				*px = i00;
			}
		}
		p2 = p;
		p >>= 1;
	}
}

// one PIZ chunk -> `raw` in the uncompressed scan-line layout
bool pizDecode(const uint8_t* in, size_t nIn, std::vector<uint8_t>& raw, const std::vector<Channel>& chans, int width, int lines)
{
	if (nIn < 4) return false;
	const int BITMAP_SIZE = 8192;
	std::vector<uint8_t> bitmap(BITMAP_SIZE, 0);
	uint16_t minNZ, maxNZ;
	memcpy(&minNZ, in, 2);
	memcpy(&maxNZ, in + 2, 2);
	const uint8_t* p = in + 4;
	if (maxNZ >= BITMAP_SIZE) return false;
	if (minNZ <= maxNZ) {
		size_t n = (size_t) maxNZ - minNZ + 1;
		if ((size_t) (p - in) + n > nIn) return false;
		memcpy(&bitmap[minNZ], p, n);
		p += n;
	}
	std::vector<uint16_t> lut(65536, 0);
	int k = 0;
	for (int i = 0; i < 65536; i++)
		if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t) i;
	const uint16_t maxValue = (uint16_t) (k - 1);

	if ((size_t) (p - in) + 4 > nIn) return false;
	int32_t length;
	memcpy(&length, p, 4);
	p += 4;
	if (length < 0 || (size_t) (p - in) + (size_t) length > nIn) return false;

	struct Plane { size_t start; int nx, ny, size; };
	std::vector<Plane> planes;
	size_t total = 0;
	for (const Channel& c: chans) {
		Plane pl;
		pl.start = total;
		pl.nx = width / c.xs;
		pl.ny = lines / c.ys;
		pl.size = c.type == 1 ? 1 : 2;
		total += (size_t) pl.nx * pl.ny * pl.size;
		planes.push_back(pl);
	}
	std::vector<uint16_t> tmp(total);
	if (!hufUncompress(p, (size_t) length, tmp.data(), total)) return false;
	for (const Plane& pl: planes)
		for (int j = 0; j < pl.size; j++)
			wav2Decode(tmp.data() + pl.start + j, pl.nx, pl.size, pl.ny, pl.nx * pl.size, maxValue);
	for (uint16_t& v: tmp) v = lut[v];

	raw.resize(total * 2);
	uint8_t* out = raw.data();
	std::vector<size_t> cursor;
	for (const Plane& pl: planes) cursor.push_back(pl.start);
	for (int y = 0; y < lines; y++)
		for (size_t c = 0; c < planes.size(); c++) {
			if (y % chans[c].ys) continue;
			size_t n = (size_t) planes[c].nx * planes[c].size;
			memcpy(out, &tmp[cursor[c]], n * 2);
			out += n * 2;
			cursor[c] += n;
		}
	return true;
}

// ---- PIZ: encoder ------------------------------------------------------------------------------------
// The inverse of the decoder above, as Imf::RgbaOutputFile applies it by default (the reference's Bitmap::saveEXR,
// /root/reference/src/bitmap.cpp:266-284, leaves the header's compression at its default, PIZ): per chunk of 32 scan lines
// the HALF samples are gathered channel by channel, mapped through the table of values that occur (bitmap + LUT), transformed
// by the 2-D Haar-like wavelet and Huffman-coded with a run-length symbol.

struct BitWriter {
	std::vector<uint8_t>& out;
	uint64_t c = 0;
	int lc = 0;
	void put(int nBits, uint64_t bits)
	{
		c = (c << nBits) | bits;
		lc += nBits;
		while (lc >= 8) {
			lc -= 8;
			out.push_back((uint8_t) (c >> lc));
		}
	}
	void flush()
	{
		if (lc > 0) out.push_back((uint8_t) (c << (8 - lc)));
		lc = 0;
		c = 0;
	}
};

// code lengths of a Huffman code for `freq` (symbols im..iM with non-zero frequency), as code[i] = length; then the canonical
// codes the decoder derives from the lengths (hufUnpackTable): code[i] = length | (value << 6)
bool hufBuildEncTable(const std::vector<uint64_t>& freq, int im, int iM, std::vector<uint64_t>& code)
{
	struct Node { uint64_t f; int left, right; };
	std::vector<Node> nodes;
	std::vector<int> leafOf(HUF_ENCSIZE, -1);
	typedef std::pair<uint64_t, int> Item; // (frequency, node): ties by node number, so the table is deterministic
	std::vector<Item> heap;
	for (int i = im; i <= iM; i++)
		if (freq[i]) {
			leafOf[i] = (int) nodes.size();
			heap.push_back(Item(freq[i], (int) nodes.size()));
			nodes.push_back(Node{ freq[i], -1, -1 });
		}
	if (heap.empty()) return false;
	std::fill(code.begin(), code.end(), 0);
	if (heap.size() == 1) {
		for (int i = im; i <= iM; i++)
			if (leafOf[i] >= 0) code[i] = 1;
	} else {
		auto cmp = [](const Item& a, const Item& b) { return a > b; };
		std::make_heap(heap.begin(), heap.end(), cmp);
		while (heap.size() > 1) {
			std::pop_heap(heap.begin(), heap.end(), cmp);
			const Item a = heap.back();
			heap.pop_back();
			std::pop_heap(heap.begin(), heap.end(), cmp);
			const Item b = heap.back();
			heap.pop_back();
			nodes.push_back(Node{ a.first + b.first, a.second, b.second });
			heap.push_back(Item(a.first + b.first, (int) nodes.size() - 1));
			std::push_heap(heap.begin(), heap.end(), cmp);
		}
		// depth of every leaf
		std::vector<int> depth(nodes.size(), 0);
		for (int n = (int) nodes.size() - 1; n >= 0; n--)
			if (nodes[n].left >= 0) {
				depth[nodes[n].left] = depth[n] + 1;
				depth[nodes[n].right] = depth[n] + 1;
			}
		for (int i = im; i <= iM; i++)
			if (leafOf[i] >= 0) {
				if (depth[leafOf[i]] > 58) return false; // the format stores lengths in 6 bits, 59..63 are run codes
				code[i] = (uint64_t) depth[leafOf[i]];
			}
	}
	uint64_t n[59] = { 0 };
	for (int i = 0; i < HUF_ENCSIZE; i++) n[code[i]]++;
	uint64_t c = 0;
	for (int i = 58; i > 0; i--) {
		const uint64_t nc = (c + n[i]) >> 1;
		n[i] = c;
		c = nc;
	}
	for (int i = 0; i < HUF_ENCSIZE; i++) {
		const int l = (int) code[i];
		if (l > 0) code[i] = l | (n[l]++ << 6);
	}
	return true;
}

void hufPackTable(const std::vector<uint64_t>& code, int im, int iM, BitWriter& bw)
{
	for (; im <= iM; im++) {
		const int l = (int) (code[im] & 63);
		if (l == 0) {
			int run = 1;
			while (im + run <= iM && run < 261 && (code[im + run] & 63) == 0) run++;
			if (run >= 2) {
				if (run >= 6) {
					bw.put(6, 63);
					bw.put(8, (uint64_t) (run - 6));
				} else {
					bw.put(6, (uint64_t) (59 + run - 2));
				}
				im += run - 1;
				continue;
			}
		}
		bw.put(6, (uint64_t) l);
	}
	bw.flush();
}

// Huffman-compress n 16-bit words: header {im, iM, table bytes, data bits, 0}, packed code lengths, code stream
void hufCompress(const uint16_t* raw, size_t n, std::vector<uint8_t>& out)
{
	out.clear();
	if (n == 0) return;
	std::vector<uint64_t> freq(HUF_ENCSIZE, 0);
	for (size_t i = 0; i < n; i++) freq[raw[i]]++;
	int im = 0, iM = HUF_ENCSIZE - 2;
	while (!freq[im]) im++;
	while (!freq[iM]) iM--;
	iM++; // the run-length symbol: one past the largest value in use, frequency 1
	freq[iM] = 1;
	std::vector<uint64_t> code(HUF_ENCSIZE, 0);
	hufBuildEncTable(freq, im, iM, code);
	out.resize(20, 0);
	BitWriter table{ out };
	hufPackTable(code, im, iM, table);
	const uint32_t tableBytes = (uint32_t) (out.size() - 20);
	const size_t dataStart = out.size();
	BitWriter bw{ out };
	auto putCode = [&](uint64_t cde) { bw.put((int) (cde & 63), cde >> 6); };
	const uint64_t runCode = code[iM];
	auto send = [&](uint64_t sCode, int runCount) {
		const int ls = (int) (sCode & 63), lr = (int) (runCode & 63);
		if (ls + lr + 8 < ls * runCount) {
			putCode(sCode);
			putCode(runCode);
			bw.put(8, (uint64_t) runCount);
		} else {
			while (runCount-- >= 0) putCode(sCode);
		}
	};
	uint16_t sym = raw[0];
	int run = 0;
	for (size_t i = 1; i < n; i++) {
		if (raw[i] == sym && run < 255) {
			run++;
		} else {
			send(code[sym], run);
			run = 0;
		}
		sym = raw[i];
	}
	send(code[sym], run);
	const uint32_t nBits = (uint32_t) ((out.size() - dataStart) * 8 + (size_t) bw.lc);
	bw.flush();
	const uint32_t hdr[5] = { (uint32_t) im, (uint32_t) iM, tableBytes, nBits, 0 };
	memcpy(out.data(), hdr, 20);
}

inline void wenc14(uint16_t a, uint16_t b, uint16_t& l, uint16_t& h)
{
	const int as = (int16_t) a, bs = (int16_t) b;
	l = (uint16_t) (int16_t) ((as + bs) >> 1);
	h = (uint16_t) (int16_t) (as - bs);
}

inline void wenc16(uint16_t a, uint16_t b, uint16_t& l, uint16_t& h)
{
	const int ao = (a + 0x8000) & 0xffff;
	int m = (ao + b) >> 1;
	int d = ao - b;
	if (d < 0) m = (m + 0x8000) & 0xffff;
	d &= 0xffff;
	l = (uint16_t) m;
	h = (uint16_t) d;
}

void wav2Encode(uint16_t* in, int nx, int ox, int ny, int oy, uint16_t mx)
{
	const bool w14 = mx < (1 << 14);
	const int n = nx > ny ? ny : nx;
	int p = 1, p2 = 2;
	auto enc = [&](uint16_t a, uint16_t b, uint16_t& l, uint16_t& h) { if (w14) wenc14(a, b, l, h); else wenc16(a, b, l, h); };
	while (p2 <= n) {
		uint16_t* py = in;
		uint16_t* ey = in + (ptrdiff_t) oy * (ny - p2);
		const int oy1 = oy * p, oy2 = oy * p2, ox1 = ox * p, ox2 = ox * p2;
		uint16_t i00, i01, i10, i11;
		for (; py <= ey; py += oy2) {
			uint16_t* px = py;
			uint16_t* ex = py + (ptrdiff_t) ox * (nx - p2);
			for (; px <= ex; px += ox2) {
				uint16_t* p01 = px + ox1;
				uint16_t* p10 = px + oy1;
				uint16_t* p11 = p10 + ox1;
				enc(*px, *p01, i00, i01);
				enc(*p10, *p11, i10, i11);
				enc(i00, i10, *px, *p10);
				enc(i01, i11, *p01, *p11);
			}
			if (nx & p) { // odd column
				uint16_t* p10 = px + oy1;
				enc(*px, *p10, i00, *p10);
				*px = i00;
			}
		}
		if (ny & p) { // odd line
			uint16_t* px = py;
			uint16_t* ex = py + (ptrdiff_t) ox * (nx - p2);
			for (; px <= ex; px += ox2) {
				uint16_t* p01 = px + ox1;
				enc(*px, *p01, i00, *p01);
				*px = i00;
			}
		}
		p = p2;
		p2 <<= 1;
	}
}

// `raw`: `lines` scan lines of `numChans` HALF channels, channel by channel within a line -> one PIZ chunk
void pizEncode(const uint16_t* raw, int width, int lines, int numChans, std::vector<uint8_t>& out)
{
	const size_t plane = (size_t) width * lines, total = plane * numChans;
	std::vector<uint16_t> tmp(total);
	for (int y = 0; y < lines; y++)
		for (int c = 0; c < numChans; c++)
			memcpy(&tmp[(size_t) c * plane + (size_t) y * width], raw + ((size_t) y * numChans + c) * width, (size_t) width * 2);
	const int BITMAP_SIZE = 8192;
	std::vector<uint8_t> bitmap(BITMAP_SIZE, 0);
	for (uint16_t v: tmp) bitmap[v >> 3] |= (uint8_t) (1 << (v & 7));
	bitmap[0] &= ~1; // zero is always in the table, never in the bitmap
	int minNZ = BITMAP_SIZE - 1, maxNZ = 0;
	for (int i = 0; i < BITMAP_SIZE; i++)
		if (bitmap[i]) {
			if (minNZ > i) minNZ = i;
			if (maxNZ < i) maxNZ = i;
		}
	std::vector<uint16_t> lut(65536, 0);
	int k = 0;
	for (int i = 0; i < 65536; i++)
		if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[i] = (uint16_t) k++;
	const uint16_t maxValue = (uint16_t) (k - 1);
	for (uint16_t& v: tmp) v = lut[v];
	for (int c = 0; c < numChans; c++) wav2Encode(tmp.data() + (size_t) c * plane, width, 1, lines, width, maxValue);
	std::vector<uint8_t> huf;
	hufCompress(tmp.data(), total, huf);
	out.clear();
	const uint16_t mn = (uint16_t) minNZ, mx = (uint16_t) maxNZ;
	out.insert(out.end(), (const uint8_t*) &mn, (const uint8_t*) &mn + 2);
	out.insert(out.end(), (const uint8_t*) &mx, (const uint8_t*) &mx + 2);
	if (minNZ <= maxNZ) out.insert(out.end(), bitmap.begin() + minNZ, bitmap.begin() + maxNZ + 1);
	const int32_t length = (int32_t) huf.size();
	out.insert(out.end(), (const uint8_t*) &length, (const uint8_t*) &length + 4);
	out.insert(out.end(), huf.begin(), huf.end());
}

// ---- ZIP -------------------------------------------------------------------------------------------

bool zipDecode(const uint8_t* in, size_t nIn, std::vector<uint8_t>& raw, size_t expected)
{
	std::vector<uint8_t> tmp(expected);
	uLongf n = (uLongf) expected;
	if (uncompress(tmp.data(), &n, in, (uLong) nIn) != Z_OK || n != expected) return false;
	for (size_t i = 1; i < expected; i++) tmp[i] = (uint8_t) (tmp[i - 1] + tmp[i] - 128); // undo the byte predictor
	raw.resize(expected);
	const size_t half = (expected + 1) / 2; // de-interleave: first half = even bytes, second half = odd bytes
	for (size_t i = 0; i < expected; i++) raw[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
	return true;
}

std::string readCString(const uint8_t*& p, const uint8_t* end)
{
	std::string s;
	while (p < end && *p) s += (char) *p++;
	if (p < end) p++;
	return s;
}

} // namespace

bool Bitmap::loadEXR(const char* filename)
{
	data.clear();
	width = height = 0;
	FILE* fp = fopen(filename, "rb");
	if (!fp) return false;
	std::vector<uint8_t> file;
	{
		fseek(fp, 0, SEEK_END);
		long sz = ftell(fp);
		fseek(fp, 0, SEEK_SET);
		file.resize(sz > 0 ? (size_t) sz : 0);
		size_t got = file.empty() ? 0 : fread(file.data(), 1, file.size(), fp);
		fclose(fp);
		if (got != file.size() || file.size() < 16) return false;
	}
	const uint8_t* p = file.data();
	const uint8_t* end = p + file.size();
	uint32_t magic, version;
	memcpy(&magic, p, 4);
	memcpy(&version, p + 4, 4);
	p += 8;
	if (magic != 20000630u || (version & 0xff) != 2 || (version & 0x1a00)) return false; // tiled / deep / multi-part

	std::vector<Channel> chans;
	int compression = -1, lineOrder = 0;
	int32_t dw[4] = { 0, 0, -1, -1 };
	while (p < end && *p) {
		std::string name = readCString(p, end), type = readCString(p, end);
		if (p + 4 > end) return false;
		int32_t size;
		memcpy(&size, p, 4);
		p += 4;
		if (size < 0 || p + size > end) return false;
		const uint8_t* v = p;
		p += size;
		if (name == "channels") {
			const uint8_t* q = v;
			while (q < v + size && *q) {
				Channel c;
				c.name = readCString(q, v + size);
				if (q + 16 > v + size) return false;
				int32_t t, xs, ys;
				memcpy(&t, q, 4);
				memcpy(&xs, q + 8, 4);
				memcpy(&ys, q + 12, 4);
				q += 16;
				c.type = t;
				c.xs = xs;
				c.ys = ys;
				if (t < 0 || t > 2 || xs < 1 || ys < 1) return false;
				chans.push_back(c);
			}
		} else if (name == "compression" && size >= 1) {
			compression = v[0];
		} else if (name == "dataWindow" && size >= 16) {
			memcpy(dw, v, 16);
		} else if (name == "lineOrder" && size >= 1) {
			lineOrder = v[0];
		}
	}
	if (p >= end) return false;
	p++; // end of header
	(void) lineOrder; // chunks carry their own y coordinate
	const int W = dw[2] - dw[0] + 1, H = dw[3] - dw[1] + 1;
	if (W <= 0 || H <= 0 || W > 65536 || H > 65536 || chans.empty()) return false;
	int linesPerBlock;
	switch (compression) {
		case 0: case 2: linesPerBlock = 1; break; // NONE, ZIPS
		case 3: linesPerBlock = 16; break;        // ZIP
		case 4: linesPerBlock = 32; break;        // PIZ
		default: return false;                    // RLE, PXR24, B44, DWA: not needed for the bundled assets
	}
	for (const Channel& c: chans)
		if (c.xs != 1 || c.ys != 1) return false; // RgbaInputFile's luminance/chroma path is not needed either
	const int nBlocks = (H + linesPerBlock - 1) / linesPerBlock;
	if (p + (size_t) nBlocks * 8 > end) return false;
	std::vector<uint64_t> offsets(nBlocks);
	memcpy(offsets.data(), p, (size_t) nBlocks * 8);

	size_t bytesPerLine = 0;
	for (const Channel& c: chans) bytesPerLine += (size_t) W * (c.type == 1 ? 2 : 4);
	int idx[3] = { -1, -1, -1 }; // R, G, B
	for (size_t i = 0; i < chans.size(); i++) {
		if (chans[i].name == "R") idx[0] = (int) i;
		if (chans[i].name == "G") idx[1] = (int) i;
		if (chans[i].name == "B") idx[2] = (int) i;
	}
	std::vector<Color> px((size_t) W * H, Color(0, 0, 0));
	std::vector<uint8_t> raw;
	for (int b = 0; b < nBlocks; b++) {
		if (offsets[b] + 8 > file.size()) return false;
		const uint8_t* c = file.data() + offsets[b];
		int32_t y0, dataSize;
		memcpy(&y0, c, 4);
		memcpy(&dataSize, c + 4, 4);
		c += 8;
		if (dataSize < 0 || c + dataSize > end) return false;
		int lines = std::min(linesPerBlock, dw[3] - y0 + 1);
		if (y0 < dw[1] || lines <= 0) return false;
		size_t expected = bytesPerLine * lines;
		if ((size_t) dataSize >= expected || compression == 0) {
			if ((size_t) dataSize < expected) return false;
			raw.assign(c, c + expected); // stored uncompressed
		} else if (compression == 4) {
			if (!pizDecode(c, (size_t) dataSize, raw, chans, W, lines) || raw.size() != expected) return false;
		} else {
			if (!zipDecode(c, (size_t) dataSize, raw, expected)) return false;
		}
		const uint8_t* r = raw.data();
		for (int ly = 0; ly < lines; ly++) {
			Color* row = &px[(size_t) (y0 - dw[1] + ly) * W];
			for (size_t ci = 0; ci < chans.size(); ci++) {
				int comp = (int) ci == idx[0] ? 0 : ((int) ci == idx[1] ? 1 : ((int) ci == idx[2] ? 2 : -1));
				size_t bpp = chans[ci].type == 1 ? 2 : 4;
				if (comp >= 0)
					for (int x = 0; x < W; x++) {
						float f;
						if (chans[ci].type == 1) {
							uint16_t h;
							memcpy(&h, r + x * 2, 2);
							f = halfToFloat(h);
						} else if (chans[ci].type == 2) {
							memcpy(&f, r + x * 4, 4);
							f = halfToFloat(floatToHalf(f)); // RgbaInputFile hands out half pixels
						} else {
							uint32_t u;
							memcpy(&u, r + x * 4, 4);
							f = halfToFloat(floatToHalf((float) u));
						}
						(comp == 0 ? row[x].r : (comp == 1 ? row[x].g : row[x].b)) = f;
					}
				r += bpp * W;
			}
		}
	}
	width = W;
	height = H;
	data.swap(px);
	return true;
}

bool Bitmap::saveEXR(const char* filename) const
{
	if (!isOK()) return false;
	FILE* fp = fopen(filename, "wb");
	if (!fp) return false;
	std::vector<uint8_t> hdr;
	auto put = [&](const void* p, size_t n) { hdr.insert(hdr.end(), (const uint8_t*) p, (const uint8_t*) p + n); };
	auto putStr = [&](const char* s) { put(s, strlen(s) + 1); };
	auto putI = [&](int32_t v) { put(&v, 4); };
	auto putF = [&](float v) { put(&v, 4); };
	auto attr = [&](const char* name, const char* type, int32_t size) { putStr(name); putStr(type); putI(size); };
	putI(20000630);
	putI(2);
	attr("channels", "chlist", 4 * 18 + 1);
	for (const char* ch: { "A", "B", "G", "R" }) {
		putStr(ch);
		putI(1); // HALF
		putI(0); // pLinear + reserved
		putI(1);
		putI(1);
	}
	hdr.push_back(0);
	attr("compression", "compression", 1);
	hdr.push_back(4); // PIZ_COMPRESSION: what Imf::RgbaOutputFile writes by default (src/bitmap.cpp:266-284)
	attr("dataWindow", "box2i", 16);
	putI(0); putI(0); putI(width - 1); putI(height - 1);
	attr("displayWindow", "box2i", 16);
	putI(0); putI(0); putI(width - 1); putI(height - 1);
	attr("lineOrder", "lineOrder", 1);
	hdr.push_back(0);
	attr("pixelAspectRatio", "float", 4);
	putF(1.0f);
	attr("screenWindowCenter", "v2f", 8);
	putF(0.0f); putF(0.0f);
	attr("screenWindowWidth", "float", 4);
	putF(1.0f);
	hdr.push_back(0);
	// chunks of 32 scan lines, each PIZ-compressed (stored raw where that does not make it smaller, as the library does)
	const int linesPerChunk = 32, numChunks = (height + linesPerChunk - 1) / linesPerChunk;
	const uint16_t one = floatToHalf(1.0f);
	std::vector<std::vector<uint8_t>> chunks(numChunks);
	std::vector<uint16_t> rawLines;
	for (int b = 0; b < numChunks; b++) {
		const int y0 = b * linesPerChunk, lines = std::min(linesPerChunk, height - y0);
		rawLines.assign((size_t) width * 4 * lines, 0);
		for (int ly = 0; ly < lines; ly++) {
			uint16_t* line = &rawLines[(size_t) ly * width * 4];
			for (int x = 0; x < width; x++) {
				const Color& c = data[(size_t) (y0 + ly) * width + x];
				line[x] = one; // A, B, G, R: channels in alphabetical order
				line[width + x] = floatToHalf(c.b);
				line[2 * (size_t) width + x] = floatToHalf(c.g);
				line[3 * (size_t) width + x] = floatToHalf(c.r);
			}
		}
		const size_t rawBytes = rawLines.size() * 2;
		pizEncode(rawLines.data(), width, lines, 4, chunks[b]);
		if (chunks[b].size() >= rawBytes) chunks[b].assign((const uint8_t*) rawLines.data(), (const uint8_t*) rawLines.data() + rawBytes);
	}
	uint64_t off = hdr.size() + (uint64_t) numChunks * 8;
	fwrite(hdr.data(), 1, hdr.size(), fp);
	for (int b = 0; b < numChunks; b++) {
		fwrite(&off, 8, 1, fp);
		off += 8 + chunks[b].size();
	}
	for (int b = 0; b < numChunks; b++) {
		int32_t head[2] = { b * linesPerChunk, (int32_t) chunks[b].size() };
		fwrite(head, 4, 2, fp);
		fwrite(chunks[b].data(), 1, chunks[b].size(), fp);
	}
	fclose(fp);
	return true;
}

} // namespace fray
