// kernel.cpp-style driver for the B200 TSDF path: the calling convention of the reference's
// src/SfM_CUDA/kernel.cpp:37-111 (glob rgb/depth/mask PNGs of a TUM-format sequence, match
// timestamps, look the pose up in groundtruth.txt, TSDF::parse_frame per frame, then spin the
// Viewer), over include/sfm_b200.hpp -> the C-ABI -> libsfm_b200.so.
//
// Differences from the reference driver, all on the I/O side ("next" rows):
//   * no OpenCV in this image: PNGs are decoded by the small zlib-based reader below, cv::imshow /
//     waitKey are dropped, the last rendered view is written as a binary PPM instead;
//   * paths, time window, frame cap, volume size and bin count are command-line arguments whose
//     defaults are the reference's hard-coded values (kernel.cpp:39-44,60-61,74; tsdf.cuh:4,52);
//   * the depth <-> mask association (kernel.cpp:64-74) is done up front and the three PNGs of the next frames are
//     decoded by a pool of threads (--decode-threads, default 4) while frame i is fused, delivered in order (the
//     reference decodes synchronously inside the loop);
//     the volume is created from the first decoded frame's size, like the reference (tsdf.cu:225), and a frame of
//     another size is an error instead of an out-of-bounds read.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "tum_io.hpp"

int main(int argc, char **argv) {
	string root = ".", render = "render.ppm", render_color, ply;
	double begin = 68164, end = 68170;  // kernel.cpp:60-61
	int max_frames = 100, dim = 256, bins = MAX_OBJECTS, views = 10, decode_threads = 4;
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
		else if (a == "--decode-threads") decode_threads = std::max(1, atoi(next().c_str()));
		else if (a == "--render") render = next();
		else if (a == "--interp") interp = true;
		else if (a == "--render-color") render_color = next();
		else if (a == "--ply") ply = next();
		else root = a;
	}
	try {
		shared_ptr<TSDF> tsdf;  // created from the first decoded frame's size
		auto traj = read_trajactory(root + "/groundtruth.txt");
		vector<string> rgb_fn = glob_png(root + "/rgb"), depth_fn = glob_png(root + "/depth"), mask_fn = glob_png(root + "/mask");
		if (traj.empty() || depth_fn.empty() || mask_fn.empty() || rgb_fn.size() != mask_fn.size()) {
			cerr << "need <dir>/groundtruth.txt, depth/*.png, rgb/*.png and one mask/*.png per rgb frame" << endl;
			return 2;
		}
		vector<double> depth_ts, mask_ts;
		for (auto &f : depth_fn) depth_ts.push_back(stamp_of(f));
		for (auto &f : mask_fn) mask_ts.push_back(stamp_of(f));
		// association pass (kernel.cpp:64-74): which depth frame goes with which mask / rgb frame
		vector<pair<size_t, size_t>> pairs;
		size_t j = 0;
		for (size_t i = 0; i < depth_ts.size(); i++) {
			if (depth_ts[i] < begin || depth_ts[i] > end) continue;
			while (i < depth_ts.size() && j < mask_ts.size() && depth_ts[i] < mask_ts[j]) i++;
			while (i < depth_ts.size() && j < mask_ts.size() && mask_ts[j] < depth_ts[i]) j++;
			if (i >= depth_ts.size() || j >= mask_ts.size()) break;  // (the reference runs off the end here, kernel.cpp:67)
			if ((int)pairs.size() >= max_frames) break;            // kernel.cpp:73-74
			pairs.push_back(make_pair(i, j));
		}
		// decode pool: `decode_threads` threads keep up to kAhead decoded frames ready, in order (one frame is three PNGs,
		// ~25 ms of inflate + unfilter on one core, against ~1 ms of GPU work: one thread feeds ~40 frames/s)
		struct Decoded { sfm::Mat depth, mask, rgb; bool ok = false; };
		const size_t kAhead = std::max<size_t>(3, 2 * (size_t)decode_threads);
		vector<Decoded> ring(kAhead);
		vector<size_t> tag(kAhead, 0);  // tag[s] == k + 1: slot s holds frame k
		mutex mu;
		condition_variable cv_full, cv_free;
		size_t next = 0, consumed = 0;
		bool stop = false;
		auto worker = [&]() {
			for (;;) {
				size_t k;
				{
					unique_lock<mutex> lk(mu);
					cv_free.wait(lk, [&] { return stop || next >= pairs.size() || next < consumed + kAhead; });
					if (stop || next >= pairs.size()) return;
					k = next++;  // slot k % kAhead is free: frame k - kAhead was consumed (next < consumed + kAhead)
				}
				Decoded d;
				d.ok = read_png(depth_fn[pairs[k].first], d.depth, false) && read_png(mask_fn[pairs[k].second], d.mask, false) &&
					read_png(rgb_fn[pairs[k].second], d.rgb, true);
				{
					lock_guard<mutex> lk(mu);
					ring[k % kAhead] = std::move(d);
					tag[k % kAhead] = k + 1;
				}
				cv_full.notify_all();
			}
		};
		vector<thread> decoders;
		for (int t = 0; t < decode_threads; t++) decoders.emplace_back(worker);
		struct Joiner {
			vector<thread> &ts; mutex &m; bool &stop; condition_variable &cv;
			~Joiner() { { lock_guard<mutex> lk(m); stop = true; } cv.notify_all(); for (auto &t : ts) if (t.joinable()) t.join(); }
		} joiner{decoders, mu, stop, cv_free};
		unique_ptr<Viewer> viewer;
		int width = 0, height = 0;
		const auto t_start = chrono::steady_clock::now();
		auto t_first = t_start;
		for (size_t k = 0; k < pairs.size(); k++) {
			const size_t i = pairs[k].first, jj = pairs[k].second;
			Decoded d;
			{
				unique_lock<mutex> lk(mu);
				cv_full.wait(lk, [&] { return tag[k % kAhead] == k + 1; });
				d = std::move(ring[k % kAhead]);
				consumed++;
			}
			cv_free.notify_all();
			sfm::Mat &depth_img = d.depth, &mask_img = d.mask, &rgb_img = d.rgb;
			if (!d.ok) {
				cerr << "cannot decode " << depth_fn[i] << " / " << mask_fn[jj] << " / " << rgb_fn[jj] << endl;
				return 2;
			}
			if (depth_img.elem_bytes != 2 || depth_img.channels != 1 || rgb_img.channels != 3 || mask_img.channels != 1) {
				cerr << "unexpected pixel formats (need 16-bit depth, 8-bit RGB, 8-bit labels)" << endl;
				return 2;
			}
			cout << "processing: " << i << ", " << rgb_fn[jj] << endl;
			if (!tsdf) {  // the reference sizes everything from the first depth frame (tsdf.cu:225, kernel.cpp:76-78)
				width = depth_img.cols;
				height = depth_img.rows;
				tsdf = make_shared<TSDF>(intr, dim, bins, width, height);
				viewer.reset(new Viewer(width, height));
			}
			if (depth_img.cols != width || depth_img.rows != height || rgb_img.cols != width || rgb_img.rows != height ||
				mask_img.cols != width || mask_img.rows != height) {
				cerr << "frame " << depth_fn[i] << ": image size differs from the first frame's " << width << "x" << height << endl;
				return 2;
			}
			const float mean = sfm_mean_depth((const uint16_t *)depth_img.data, depth_img.rows * depth_img.cols);  // kernel.cpp:95
			auto low = traj.lower_bound(depth_ts[i]);  // kernel.cpp:97 (no interpolation)
			if (low == traj.end()) --low;
			float extrinsic[16];
			if (interp && low != traj.begin() && low->first > depth_ts[i]) {
				// the TSDF_Python prototype's front-end (main.py:127-138): lerp + slerp between the two entries
				// that bracket the depth timestamp instead of taking the next one
				auto prev = std::prev(low);
				double a8[8] = {prev->first}, b8[8] = {low->first}, pose7[7];
				for (int q = 0; q < 7; q++) { a8[1 + q] = prev->second[q]; b8[1 + q] = low->second[q]; }
				sfm_interpolate_pose(a8, b8, depth_ts[i], pose7);
				sfm_parse_extrinsic(pose7, extrinsic);
			} else
				sfm_parse_extrinsic(low->second.data(), extrinsic);  // kernel.cpp:98
			tsdf->parse_frame(depth_img, rgb_img, mask_img, extrinsic, mean);  // kernel.cpp:99
			if (k == 0) {  // the first frame pays for the context, the volume's allocation and its clearing
				sfm_synchronize(tsdf->handle());
				t_first = chrono::steady_clock::now();
			}
		}
		if (!tsdf) {
			cerr << "no frame inside the time window" << endl;
			return 2;
		}
		sfm_synchronize(tsdf->handle());
		const auto t_end = chrono::steady_clock::now();
		const double secs = chrono::duration<double>(t_end - t_start).count(), steady = chrono::duration<double>(t_end - t_first).count();
		cout << "frames from disk: " << pairs.size() << " in " << secs << " s = " << (pairs.size() / max(secs, 1e-9)) << " frames/s";
		if (pairs.size() > 1) cout << "; after the first frame (context, allocation): " << ((pairs.size() - 1) / max(steady, 1e-9)) << " frames/s";
		cout << " (PNG decode on " << decode_threads << " threads, up to " << kAhead << " frames ahead)" << endl;
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
