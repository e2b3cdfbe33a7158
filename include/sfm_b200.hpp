// sfm_b200.hpp -- header-only C++ mirror of the reference's `class TSDF` (src/SfM_CUDA/tsdf.cuh:7-67)
// and `class Viewer` (viewer.cuh:4-17) on top of the C-ABI (sfm_b200.h).
//
// Same method names, argument meaning and error behaviour as the reference: parse_frame relabels
// `masks` in place, failures throw std::string (tsdf.cu:497-503).  cv::Mat is replaced by the
// 4-field sfm::Mat below; define SFM_WITH_OPENCV before including this header (and have OpenCV on
// the include path) to get overloads taking cv::Mat / cv::Scalar, so that the reference's
// kernel.cpp compiles against this header with only its #include lines changed.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "sfm_b200.h"

#ifdef SFM_WITH_OPENCV
#include <opencv2/core.hpp>
#endif

#ifndef MAX_OBJECTS
#define MAX_OBJECTS 32  // tsdf.cuh:4 -- the default bin count; a constructor argument here
#endif

namespace sfm {

// stand-in for the cv::Mat fields the reference's hot path touches (data, rows, cols)
struct Mat {
	int rows = 0, cols = 0, channels = 1, elem_bytes = 1;
	std::vector<uint8_t> store;
	uint8_t *data = nullptr;
	Mat() {}
	Mat(int r, int c, int ch, int eb) : rows(r), cols(c), channels(ch), elem_bytes(eb), store((size_t)r * c * ch * eb) { data = store.data(); }
	Mat(int r, int c, int ch, int eb, void *ext) : rows(r), cols(c), channels(ch), elem_bytes(eb), data((uint8_t *)ext) {}
	Mat(const Mat &o) { *this = o; }
	Mat(Mat &&o) noexcept { *this = static_cast<Mat &&>(o); }
	Mat &operator=(const Mat &o) {
		rows = o.rows; cols = o.cols; channels = o.channels; elem_bytes = o.elem_bytes;
		store = o.store;
		data = o.store.empty() ? o.data : store.data();  // owning copies re-point, borrowed views stay views
		return *this;
	}
	Mat &operator=(Mat &&o) noexcept {
		rows = o.rows; cols = o.cols; channels = o.channels; elem_bytes = o.elem_bytes;
		const bool owning = !o.store.empty();
		store = static_cast<std::vector<uint8_t> &&>(o.store);
		data = owning ? store.data() : o.data;
		o.data = nullptr;
		return *this;
	}
	bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
#ifdef SFM_WITH_OPENCV
	// `cv::Mat img = viewer->show_tsdf(...)` as in the reference's kernel.cpp:105 (8-bit images only)
	operator cv::Mat() const { return cv::Mat(rows, cols, CV_8UC(channels), data).clone(); }
#endif
};

inline void check(int rc) {
	if (rc != SFM_OK) throw std::string("run_kernel launch failed\n") + sfm_last_error();
}

}  // namespace sfm

class Viewer;

class TSDF {
public:
	// TSDF(cv::Scalar intrinsics)  (tsdf.cu:137-150); dims / bins / frame size are compile-time in the reference
	explicit TSDF(const float intrinsics[4], int dim = 256, int bins = MAX_OBJECTS, int width = 640, int height = 480, int device = 0) {
		sfm_desc d;
		sfm_desc_default(&d);
		d.dims[0] = d.dims[1] = d.dims[2] = dim;
		d.bins = bins;
		d.width = width;
		d.height = height;
		d.device = device;
		d.K[0] = intrinsics[0]; d.K[5] = intrinsics[1]; d.K[2] = intrinsics[2]; d.K[6] = intrinsics[3];
		desc_ = d;
		sfm::check(sfm_create(&d, &vol_));
	}
	~TSDF() { sfm_destroy(vol_); }
	TSDF(const TSDF &) = delete;
	TSDF &operator=(const TSDF &) = delete;

	float mean_depth_ = 0.f;  // tsdf.cuh:11

	// tsdf.cu:171-228.  depth CV_16UC1, color CV_8UC3 (BGR), masks CV_8UC1 (relabelled IN PLACE),
	// extrinsic 4x4 CV_32F world->camera.
	void parse_frame(const sfm::Mat &depth, const sfm::Mat &color, sfm::Mat &masks, const float extrinsic[16], float mean_depth) {
		sfm::check(sfm_parse_frame(vol_, (const uint16_t *)depth.data, color.data, masks.data, extrinsic, mean_depth));
		sfm_info i;
		sfm_get_info(vol_, &i);
		mean_depth_ = i.mean_depth;
	}
	// the getters return live copies of the device planes (the reference's return stale host mirrors, tsdf.cu:492-516)
	std::vector<float> get_tsdf_diff() const { return fetch<float>(SFM_PLANE_SDF); }
	std::vector<uint8_t> get_tsdf_color() const { return fetch<uint8_t>(SFM_PLANE_COLOR); }
	std::vector<uint32_t> get_tsdf_cnt() const { return fetch<uint32_t>(SFM_PLANE_HIST); }
	std::vector<int32_t> get_tsdf_wt() const { return fetch<int32_t>(SFM_PLANE_WEIGHT); }
	void get_dim(int out[3]) const { sfm_info i = info(); memcpy(out, i.dims, sizeof(i.dims)); }
	void get_vol_start(float out[3]) const { sfm_info i = info(); memcpy(out, i.vol_start, 12); }
	void get_vol_end(float out[3]) const { sfm_info i = info(); memcpy(out, i.vol_end, 12); }
	void get_voxel(float out[3]) const { sfm_info i = info(); memcpy(out, i.voxel, 12); }
	const float *get_intrinsic() const { return desc_.K; }
	sfm_info info() const { sfm_info i; sfm::check(sfm_get_info(vol_, &i)); return i; }
	// zero-crossing points of the fused volume (no reference counterpart): xyz[3n], bgr[3n], label[n]
	size_t extract_surface(std::vector<float> &xyz, std::vector<uint8_t> &bgr, std::vector<uint8_t> &label) const {
		uint32_t n = 0;
		sfm::check(sfm_extract_surface(vol_, 0, nullptr, nullptr, nullptr, &n));
		xyz.resize((size_t)n * 3); bgr.resize((size_t)n * 3); label.resize(n);
		if (n) sfm::check(sfm_extract_surface(vol_, n, xyz.data(), bgr.data(), label.data(), &n));
		return n;
	}
	sfm_volume *handle() const { return vol_; }
	// the reference exposes its device pointers as public members (tsdf.cuh:24-43)
	float *tsdf_diff_d() const { return (float *)sfm_plane_device_ptr(vol_, SFM_PLANE_SDF); }
	uint8_t *tsdf_color_d() const { return (uint8_t *)sfm_plane_device_ptr(vol_, SFM_PLANE_COLOR); }
	uint32_t *tsdf_cnt_d() const { return (uint32_t *)sfm_plane_device_ptr(vol_, SFM_PLANE_HIST); }
	int *tsdf_wt_d() const { return (int *)sfm_plane_device_ptr(vol_, SFM_PLANE_WEIGHT); }

#ifdef SFM_WITH_OPENCV
	explicit TSDF(cv::Scalar intrinsics) : TSDF(std::vector<float>{(float)intrinsics[0], (float)intrinsics[1], (float)intrinsics[2], (float)intrinsics[3]}.data()) {}
	void parse_frame(const cv::Mat &depth, const cv::Mat &color, cv::Mat &masks, const cv::Mat &extrinsic, float mean_depth) {
		sfm::check(sfm_parse_frame(vol_, (const uint16_t *)depth.data, color.data, masks.data, (const float *)extrinsic.data, mean_depth));
		mean_depth_ = info().mean_depth;
	}
#endif

private:
	template <typename T> std::vector<T> fetch(int plane) const {
		std::vector<T> out(sfm_plane_bytes(vol_, plane) / sizeof(T));
		sfm::check(sfm_download(vol_, plane, out.data(), out.size() * sizeof(T)));
		return out;
	}
	sfm_volume *vol_ = nullptr;
	sfm_desc desc_;
};

class Viewer {
public:
	int width_, height_;
	explicit Viewer(int width, int height) : width_(width), height_(height) {}  // viewer.cu:88-128
	// viewer.cu:137-179: returns the H x W BGR image (the cv::imshow / waitKey is the caller's business)
	sfm::Mat show_tsdf(const TSDF &tsdf, float angle, float dist) {
		sfm::Mat img(height_, width_, 3, 1);
		sfm::check(sfm_show(tsdf.handle(), angle, dist, width_, height_, img.data));
		return img;
	}
	// the same orbit view in colour mode: interp_tsdf_color at the hit (the call viewer.cu:68 keeps commented out)
	sfm::Mat show_tsdf_color(const TSDF &tsdf, float angle, float dist) {
		sfm::Mat img(height_, width_, 3, 1);
		sfm::check(sfm_show_color(tsdf.handle(), angle, dist, width_, height_, img.data));
		return img;
	}
};
