// kernel.cpp-style driver for the B200 TSDF path: the calling convention of the reference's
// src/SfM_CUDA/kernel.cpp:37-111 (glob rgb/depth/mask PNGs of a TUM-format sequence, match
// timestamps, look the pose up in groundtruth.txt, TSDF::parse_frame per frame, then spin the
// Viewer), over include/sfm_b200.hpp -> the C-ABI -> libsfm_b200.so.
//
// Differences from the reference driver, all on the I/O side ("next" rows):
//   * no OpenCV in this image: PNGs are decoded by the small zlib-based reader below, cv::imshow /
//     waitKey are dropped, the last rendered view is written as a binary PPM instead;
//   * paths, time window, frame cap, volume size and bin count are command-line arguments whose
//     defaults are the reference's hard-coded values (kernel.cpp:39-44,60-61,74; tsdf.cuh:4,52).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
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

int main(int argc, char **argv) {
	string root = ".", render = "render.ppm", render_color, ply;
	double begin = 68164, end = 68170;  // kernel.cpp:60-61
	int max_frames = 100, dim = 256, bins = MAX_OBJECTS, views = 10;
	bool interp = false;  // --interp: lerp + slerp poses (TSDF_Python front-end) instead of the next entry
	float intr[4] = {520.9f, 521.0f, 325.1f, 249.7f};  // kernel.cpp:39
	for (int i = 1; i < argc; i++) {
		const string a = argv[i];
		auto next = [&]() { return string(i + 1 < argc ? argv[++i] : ""); };
		if (a == "--dim") dim = atoi(next().c_str());
		else if (a == "--bins") bins = atoi(next().c_str());
		else if (a == "--begin") begin = atof(next().c_str());
		else if (a == "--end") end = atof(next().c_str());
		else if (a == "--max-frames") max_frames = atoi(next().c_str());
		else if (a == "--views") views = atoi(next().c_str());
		else if (a == "--render") render = next();
		else if (a == "--interp") interp = true;
		else if (a == "--render-color") render_color = next();
		else if (a == "--ply") ply = next();
		else root = a;
	}
	try {
		auto tsdf = make_shared<TSDF>(intr, dim, bins);
		auto traj = read_trajactory(root + "/groundtruth.txt");
		vector<string> rgb_fn = glob_png(root + "/rgb"), depth_fn = glob_png(root + "/depth"), mask_fn = glob_png(root + "/mask");
		if (traj.empty() || depth_fn.empty() || mask_fn.empty() || rgb_fn.size() != mask_fn.size()) {
			cerr << "need <dir>/groundtruth.txt, depth/*.png, rgb/*.png and one mask/*.png per rgb frame" << endl;
			return 2;
		}
		vector<double> depth_ts, mask_ts;
		for (auto &f : depth_fn) depth_ts.push_back(stamp_of(f));
		for (auto &f : mask_fn) mask_ts.push_back(stamp_of(f));
		size_t j = 0;
		unique_ptr<Viewer> viewer;
		int cnt = 0;
		for (size_t i = 0; i < depth_ts.size(); i++) {  // kernel.cpp:64-100
			if (depth_ts[i] < begin || depth_ts[i] > end) continue;
			while (i < depth_ts.size() && j < mask_ts.size() && depth_ts[i] < mask_ts[j]) i++;
			while (i < depth_ts.size() && j < mask_ts.size() && mask_ts[j] < depth_ts[i]) j++;
			if (i >= depth_ts.size() || j >= mask_ts.size()) break;  // (the reference runs off the end here, kernel.cpp:67)
			sfm::Mat depth_img, mask_img, rgb_img;
			if (!read_png(depth_fn[i], depth_img, false) || !read_png(mask_fn[j], mask_img, false) || !read_png(rgb_fn[j], rgb_img, true)) {
				cerr << "cannot decode " << depth_fn[i] << " / " << mask_fn[j] << " / " << rgb_fn[j] << endl;
				return 2;
			}
			if (depth_img.elem_bytes != 2 || depth_img.channels != 1 || rgb_img.channels != 3 || mask_img.channels != 1) {
				cerr << "unexpected pixel formats (need 16-bit depth, 8-bit RGB, 8-bit labels)" << endl;
				return 2;
			}
			cout << "processing: " << i << ", " << rgb_fn[j] << endl;
			if (++cnt > max_frames) break;  // kernel.cpp:73-74
			if (!viewer) viewer.reset(new Viewer(depth_img.cols, depth_img.rows));
			const float mean = sfm_mean_depth((const uint16_t *)depth_img.data, depth_img.rows * depth_img.cols);  // kernel.cpp:95
			auto low = traj.lower_bound(depth_ts[i]);  // kernel.cpp:97 (no interpolation)
			if (low == traj.end()) --low;
			float extrinsic[16];
			if (interp && low != traj.begin() && low->first > depth_ts[i]) {
				// the TSDF_Python prototype's front-end (main.py:127-138): lerp + slerp between the two entries
				// that bracket the depth timestamp instead of taking the next one
				auto prev = std::prev(low);
				double a8[8] = {prev->first}, b8[8] = {low->first}, pose7[7];
				for (int k = 0; k < 7; k++) { a8[1 + k] = prev->second[k]; b8[1 + k] = low->second[k]; }
				sfm_interpolate_pose(a8, b8, depth_ts[i], pose7);
				sfm_parse_extrinsic(pose7, extrinsic);
			} else
				sfm_parse_extrinsic(low->second.data(), extrinsic);  // kernel.cpp:98
			tsdf->parse_frame(depth_img, rgb_img, mask_img, extrinsic, mean);  // kernel.cpp:99
		}
		const sfm_info info = tsdf->info();
		cout << "fused " << info.n_obs << " frames, num_objs " << info.num_objs << ", voxel " << info.voxel[0] << " m" << endl;
		if (viewer) {  // kernel.cpp:101-107 spins forever; we render `views` steps and keep the last image
			float angle = 0.f;
			sfm::Mat img;
			for (int k = 0; k < views; k++) {
				angle += 0.01f;
				img = viewer->show_tsdf(*tsdf, angle, tsdf->mean_depth_);
			}
			size_t lit = 0;
			for (size_t p = 0; p < (size_t)img.rows * img.cols; p++) lit += (img.data[p * 3] | img.data[p * 3 + 1] | img.data[p * 3 + 2]) != 0;
			write_ppm_bgr(render, img);
			cout << "rendered " << views << " views, last one -> " << render << " (" << lit << " labelled pixels)" << endl;
			if (!render_color.empty()) {  // the colour mode the reference keeps commented out (viewer.cu:68)
				write_ppm_bgr(render_color, viewer->show_tsdf_color(*tsdf, angle, tsdf->mean_depth_));
				cout << "colour view -> " << render_color << endl;
			}
		}
		if (!ply.empty()) {  // surface export: zero-crossing points with colour and label
			vector<float> xyz;
			vector<uint8_t> bgr, label;
			const size_t n = tsdf->extract_surface(xyz, bgr, label);
			ofstream f(ply, ios::binary);
			f << "ply\nformat binary_little_endian 1.0\nelement vertex " << n
			  << "\nproperty float x\nproperty float y\nproperty float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nproperty uchar label\nend_header\n";
			for (size_t i = 0; i < n; i++) {
				f.write((const char *)&xyz[i * 3], 12);
				const uint8_t rec[4] = {bgr[i * 3 + 2], bgr[i * 3 + 1], bgr[i * 3], label[i]};
				f.write((const char *)rec, 4);
			}
			cout << "surface: " << n << " points -> " << ply << endl;
		}
	} catch (const string &e) {  // the reference throws std::string (tsdf.cu:502, viewer.cu:174)
		cerr << e << endl;
		return 1;
	}
	return 0;
}
