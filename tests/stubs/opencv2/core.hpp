// TEST STUB -- not OpenCV.  The few cv:: names include/sfm_b200.hpp touches under SFM_WITH_OPENCV, with OpenCV's
// signatures, so that the overloads are at least compiled (OpenCV C++ is absent from this image).  Used by
// tests/test_cpu_abi_and_host.py::test_opencv_overloads_compile only.
#pragma once
#include <cstdint>
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC(n) CV_MAKETYPE(CV_8U, (n))
#define CV_8UC1 CV_8UC(1)
#define CV_8UC3 CV_8UC(3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
namespace cv {
struct Scalar {
	double val[4];
	Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
	double operator[](int i) const { return val[i]; }
};
struct Mat {
	int rows = 0, cols = 0, flags = 0;
	uint8_t *data = nullptr;
	Mat() {}
	Mat(int r, int c, int type, void *ext) : rows(r), cols(c), flags(type), data((uint8_t *)ext) {}
	Mat clone() const { return *this; }
	int type() const { return flags; }
};
}  // namespace cv
