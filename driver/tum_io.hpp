// tum_io.hpp -- the I/O side shared by the kernel.cpp-style drivers: a small zlib PNG reader (no OpenCV C++ in this
// image), a PPM writer, and the reference driver's own helpers (read_trajactory utils.cu:62-75, the file-name time
// stamps of kernel.cpp:51-58, cv::glob).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include <zlib.h>

#include "sfm_b200.hpp"

using namespace std;

// ---- minimal PNG reader: 8/16-bit grey, RGB, grey+alpha, RGBA; non-interlaced -------------------
static uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

// Returns rows x cols x channels, 8 or 16 bit (host endian).  want_bgr swaps RGB -> BGR like cv::imread.
static bool read_png(const string &path, sfm::Mat &out, bool want_bgr) {
	ifstream f(path, ios::binary);
	if (!f) return false;
	vector<uint8_t> buf((istreambuf_iterator<char>(f)), istreambuf_iterator<char>());
	static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
	if (buf.size() < 33 || memcmp(buf.data(), sig, 8)) return false;
	uint32_t w = 0, h = 0;
	int depth = 0, ctype = 0, interlace = 0;
	vector<uint8_t> idat;
	for (size_t pos = 8; pos + 12 <= buf.size();) {
		const uint32_t len = be32(&buf[pos]);
		const char *type = (const char *)&buf[pos + 4];
		const uint8_t *data = &buf[pos + 8];
		if (pos + 12 + len > buf.size()) return false;
		if (!memcmp(type, "IHDR", 4)) { w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12]; }
		else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
		else if (!memcmp(type, "IEND", 4)) break;
		pos += 12 + len;
	}
	if (!w || !h || interlace || (depth != 8 && depth != 16)) return false;
	const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	if (!ch) return false;
	const int bpp = ch * depth / 8;
	const size_t stride = (size_t)w * bpp;
	vector<uint8_t> raw((stride + 1) * h);
	uLongf rawlen = raw.size();
	if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) return false;
	vector<uint8_t> img(stride * h);
	for (uint32_t y = 0; y < h; y++) {  // undo the per-row filters
		const uint8_t ft = raw[y * (stride + 1)];
		const uint8_t *src = &raw[y * (stride + 1) + 1];
		uint8_t *dst = &img[y * stride];
		const uint8_t *up = y ? &img[(y - 1) * stride] : nullptr;
		for (size_t i = 0; i < stride; i++) {
			const int a = i >= (size_t)bpp ? dst[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
			int pred = 0;
			switch (ft) {
			case 1: pred = a; break;
			case 2: pred = b; break;
			case 3: pred = (a + b) >> 1; break;
			case 4: { const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); } break;
			default: break;
			}
			dst[i] = (uint8_t)(src[i] + pred);
		}
	}
	const int out_ch = (ch == 2) ? 1 : (ch == 4 ? 3 : ch);
	out = sfm::Mat((int)h, (int)w, out_ch, depth / 8);
	for (size_t p = 0; p < (size_t)w * h; p++)
		for (int c = 0; c < out_ch; c++) {
			const int sc = (want_bgr && out_ch == 3) ? 2 - c : c;
			const uint8_t *s = &img[p * bpp + sc * (depth / 8)];
			if (depth == 8) out.data[p * out_ch + c] = s[0];
			else ((uint16_t *)out.data)[p * out_ch + c] = (uint16_t)(s[0] << 8 | s[1]);  // PNG is big endian
		}
	return true;
}

static bool write_ppm_bgr(const string &path, const sfm::Mat &img) {
	ofstream f(path, ios::binary);
	if (!f) return false;
	f << "P6\n" << img.cols << " " << img.rows << "\n255\n";
	for (size_t p = 0; p < (size_t)img.rows * img.cols; p++) {
		const char rgb[3] = {(char)img.data[p * 3 + 2], (char)img.data[p * 3 + 1], (char)img.data[p * 3]};
		f.write(rgb, 3);
	}
	return true;
}

// ---- the reference driver's helpers ---------------------------------------------------------------
// read_trajactory (utils.cu:62-75): key = fmod(ts, 1e5), value = {tx,ty,tz,qx,qy,qz,qw}
static map<double, vector<double>> read_trajactory(const string &filename) {
	map<double, vector<double>> result;
	string line;
	ifstream infile(filename.c_str());
	while (getline(infile, line)) {
		istringstream iss(line);
		double ts, tx, ty, tz, qx, qy, qz, qw;
		if (!(iss >> ts >> tx >> ty >> tz >> qx >> qy >> qz >> qw)) continue;
		result.insert(make_pair(fmod(ts, 1e5), vector<double>{tx, ty, tz, qx, qy, qz, qw}));
	}
	return result;
}

static vector<string> glob_png(const string &dir) {  // cv::glob(dir/*.png), sorted
	vector<string> out;
	if (DIR *d = opendir(dir.c_str())) {
		while (dirent *e = readdir(d)) {
			const string n = e->d_name;
			if (n.size() > 4 && n.substr(n.size() - 4) == ".png") out.push_back(dir + "/" + n);
		}
		closedir(d);
	}
	sort(out.begin(), out.end());
	return out;
}

// kernel.cpp:51-58: drop the 5 leading digits of the file name, parse the rest; the lambda returns float
static double stamp_of(const string &fn) {
	const size_t s = fn.find_last_of("/");
	return (float)stod(fn.substr(s + 6, fn.find_last_of(".") - s - 6));
}

