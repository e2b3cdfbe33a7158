#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/ from the reference's own sources.

Compiles the reference's hot-path code VERBATIM, from where it lies under
/root/reference/src/SfM_CUDA, into oracle/_ref/libsfm_ref_L<bins>.so:

  * utils.cu:93-170   mix / interp_tsdf_diff / interp_tsdf_color / interp_tsdf_cnt   (own TU, -dc)
  * tsdf.cu:18-135    tsdf_kernel, back_proj_kernel                                   (own TU, -dc)
  * viewer.cu:17-86   show_tsdf_kernel                                                (own TU, -dc)
  * tsdf.cu:304-416   TSDF::filter_overlaps (CPU) behind a 4-line cv::Mat shim        (own TU)
  * oracle/ref_shim.cu  OUR launch shim (grid/block as tsdf.cu:441,472 / viewer.cu:152)

The reference as shipped does not compile here (every TU includes <opencv2/opencv.hpp>,
tsdf.cuh:2 / utils.cuh:3, and OpenCV C++ is absent), and its build/Makefile links OpenCV, so
the build system itself is not run; only the self-contained line ranges above are compiled,
with the Makefile's flags (nvcc -std=c++11 -dc, default -fmad=true -prec-div=true) plus
-gencode arch=compute_100a,code=sm_100a and -maxrregcount=64 (needed for the reference's own
1024-thread launches to fit the register file; allocation only, arithmetic unchanged).  The extracted text lives only in a temporary
directory; nothing but the .so files is written to oracle/_ref/ (git-ignored, travels to
the GPU box).  No reference source is copied into the repository.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SFM_REF_DIR", "/root/reference/src/SfM_CUDA")
OUT = os.path.join(HERE, "_ref")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
BINS = (16, 32, 80)

PRELUDE = """#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector_functions.h>
#include "helper_math.h"
"""

OVERLAP_PRELUDE = """#include <cstdint>
#include <cstdio>
#include <cmath>
#include <iostream>
#include <sstream>
#include <unordered_map>
#include <cuda_runtime.h>
#include "helper_math.h"
#include "configuration.h"
// minimal stand-ins for the two OpenCV names filter_overlaps touches (tsdf.cu:304-311)
namespace cv {
struct Mat { uint8_t *data; int rows, cols; };
static inline void minMaxLoc(const Mat &m, double *mn, double *mx) {
	double lo = 255, hi = 0;
	for (int i = 0; i < m.rows * m.cols; i++) { if (m.data[i] < lo) lo = m.data[i]; if (m.data[i] > hi) hi = m.data[i]; }
	if (mn) *mn = lo;
	if (mx) *mx = hi;
}
}
struct TSDF {
	uint32_t n_obs_;
	int num_objs;
	void filter_overlaps(float *probs, int width, int height, cv::Mat& mask, bool *box_mask);
};
"""

OVERLAP_EPILOGUE = """
extern "C" int ref_filter_overlaps_impl(float *probs, int width, int height, uint8_t *mask,
	bool *box_mask, uint32_t n_obs, int *num_objs_inout) {
	TSDF t; t.n_obs_ = n_obs; t.num_objs = *num_objs_inout;
	cv::Mat m{mask, height, width};
	t.filter_overlaps(probs, width, height, m, box_mask);
	*num_objs_inout = t.num_objs;
	return 0;
}
"""


def lines(path, lo, hi, first_must_contain, last_must_contain=None):
    with open(os.path.join(REF, path)) as f:
        src = f.read().split("\n")
    seg = src[lo - 1:hi]
    if first_must_contain not in seg[0]:
        raise SystemExit(f"{path}:{lo} does not look like the expected line: {seg[0]!r}")
    if last_must_contain is not None and last_must_contain not in seg[-1]:
        raise SystemExit(f"{path}:{hi} does not look like the expected line: {seg[-1]!r}")
    return "\n".join(seg) + "\n"


def run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise SystemExit("reference oracle build failed")
    return r.stdout


def build(bins_list=BINS, verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"[build_ref] {REF} not present; keeping prebuilt oracle/_ref as is")
        return False
    os.makedirs(OUT, exist_ok=True)
    protos = lines("utils.cuh", 14, 16, "interp_tsdf_diff", "interp_tsdf_cnt")
    tu = {
        "ref_utils.cu": PRELUDE + lines("utils.cu", 93, 170, "template <typename T>", "}"),
        "ref_tsdf.cu": PRELUDE + protos + lines("tsdf.cu", 18, 135, "__global__ void tsdf_kernel", "}"),
        "ref_viewer.cu": PRELUDE + protos + lines("viewer.cu", 17, 86, "__global__ void show_tsdf_kernel", "}"),
        "ref_overlaps.cu": OVERLAP_PRELUDE + lines("tsdf.cu", 304, 416, "void TSDF::filter_overlaps", "}")
        + OVERLAP_EPILOGUE,
    }
    for bins in bins_list:
        tmp = tempfile.mkdtemp(prefix="sfm_ref_")
        try:
            for name, text in tu.items():
                with open(os.path.join(tmp, name), "w") as f:
                    f.write(text)
            shutil.copy(os.path.join(HERE, "ref_shim.cu"), os.path.join(tmp, "ref_shim.cu"))
            # -maxrregcount=64: the reference launches its ray kernels with 32x32 = 1024-thread blocks
            # (tsdf.cu:441, viewer.cu:152), which needs <= 64 registers/thread; with separate
            # compilation nvcc 12.9 allocates 108 for sm_100a and the verbatim launch fails with "too
            # many resources requested".  The cap changes register allocation only, not arithmetic.
            common = ["nvcc", "-std=c++11", "-dc", "-maxrregcount=64", "-Xcompiler", "-fPIC", "-w",
                      f"-DMAX_OBJECTS={bins}", f"-I{REF}"] + ARCH
            objs = []
            for name in list(tu) + ["ref_shim.cu"]:
                run(common + [name, "-o", name + ".o"], tmp)
                objs.append(name + ".o")
            so = os.path.join(OUT, f"libsfm_ref_L{bins}.so")
            run(["nvcc", "-shared", "-Xcompiler", "-fPIC"] + ARCH + objs + ["-o", so], tmp)
            if verbose:
                print(f"[build_ref] built {so}")
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    build_tsdf_python(verbose)
    return True


def build_tsdf_python(verbose=True):
    """Secondary oracle: the TSDF_Python prototype's CUDA module (src/TSDF_Python/tsdf.cu:10-58 kernel, :61-137 host
    wrapper, tsdf.cpp:11-38 pybind11 binding), compiled VERBATIM from where the two files lie into
    oracle/_ref/tsdf_cuda<ext>.  Its CMakeLists only sets -std=c++11 and links CUDA + Python; the build system itself
    is not run (it asks for FindCUDA and a pybind11 CMake package).  nvcc 12.9 needs c++14 for thrust, so the flag is
    -std=c++14; no other option is added (default -fmad=true, as the reference)."""
    import sysconfig
    src = os.path.join(os.path.dirname(REF), "TSDF_Python")
    if not os.path.isfile(os.path.join(src, "tsdf.cu")):
        return False
    try:
        import pybind11
    except ImportError:
        if verbose:
            print("[build_ref] pybind11 missing: TSDF_Python oracle not built")
        return False
    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    so = os.path.join(OUT, "tsdf_cuda" + ext)
    tmp = tempfile.mkdtemp(prefix="sfm_ref_py_")
    try:
        run(["nvcc", "-std=c++14", "-shared", "-Xcompiler", "-fPIC", "-w", f"-I{src}", f"-I{pybind11.get_include()}",
             f"-I{sysconfig.get_paths()['include']}"] + ARCH +
            [os.path.join(src, "tsdf.cu"), os.path.join(src, "tsdf.cpp"), "-o", so], tmp)
        if verbose:
            print(f"[build_ref] built {so}")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return True


if __name__ == "__main__":
    build()
