// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Host-side launch shim around the reference's own device kernels.  The kernel
// bodies are NOT in this repository: oracle/build_ref.py reads them where they
// lie under /root/reference/src/SfM_CUDA (tsdf.cu:18-135, viewer.cu:17-86,
// utils.cu:93-170, tsdf.cu:304-416) into a temporary directory, compiles them
// verbatim with the reference's flags (nvcc -std=c++11 -dc, default fmad) for
// sm_100a and links them with this shim into oracle/_ref/libsfm_ref_L<bins>.so.
//
// This file only re-creates what the reference's host code does around the
// launches (tsdf.cu:441-455, 472-488; viewer.cu:152-166): copy the small
// parameter blocks to device memory and launch with the reference's grid/block.
// Volume planes and frame images are passed as DEVICE pointers so the tests can
// run the reference kernels on bit-identical copies of our own planes.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>

#ifndef MAX_OBJECTS
#error "MAX_OBJECTS must be defined by the build recipe"
#endif

// prototypes of the verbatim reference kernels (defined in the generated TUs)
__global__ void tsdf_kernel(float *tsdf_diff, uint8_t *tsdf_color, uint32_t *tsdf_cnt, int *tsdf_wt,
	int *vol_dim, float *vol_start, float *voxel, float miu, float *intrinsic,
	uint16_t *depth, uint8_t *color, uint8_t *mask, float *extrinsic2init, int width, int height);
__global__ void back_proj_kernel(float *K_inv, float *Rt, float3 *o, float3 *vol_start, float3 *vol_end, float3 *voxel,
	int3 *vol_dim, float *tsdf_diff, uint32_t *tsdf_cnt,
	int width, int height, float *probs, bool *box_mask);
__global__ void show_tsdf_kernel(float *s2w, float3 *c, float3 *vol_start, float3 *vol_end, float3 *voxel,
	int3 *vol_dim, float *tsdf_diff, uchar3 *tsdf_color, uint32_t *tsdf_cnt,
	int width, int height, uchar3 *output, uint8_t *random_colors);

// the reference's colour sampler (utils.cu:121-142): defined in the generated utils TU, unused by the
// reference's own kernels (its only call, viewer.cu:68, is commented out) -- launched from a shim kernel here
__device__ uchar3 interp_tsdf_color(const float3 &pos, const float3 &vol_start, const float3 &voxel, const int3 &vol_dim, uchar3 *tsdf_color);

__global__ void shim_color_at_kernel(const float *xyz, const uint8_t *valid, int n, float3 vol_start, float3 voxel, int3 vol_dim,
	uchar3 *tsdf_color, uchar3 *out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !valid[i]) return;
	out[i] = interp_tsdf_color(make_float3(xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]), vol_start, voxel, vol_dim, tsdf_color);
}

// defined in the generated overlaps TU (verbatim TSDF::filter_overlaps behind a cv::Mat shim)
extern "C" int ref_filter_overlaps_impl(float *probs, int width, int height, uint8_t *mask,
	bool *box_mask, uint32_t n_obs, int *num_objs_inout);

namespace {
struct Scratch {
	float *f = nullptr;   // 256 floats of parameter scratch
	int *i = nullptr;     // 16 ints
	uint8_t *pal = nullptr;
	bool ok = false;
};
Scratch g_s;

int ensure_scratch() {
	if (g_s.ok) return 0;
	if (cudaMalloc(&g_s.f, 256 * sizeof(float)) != cudaSuccess) return -1;
	if (cudaMalloc(&g_s.i, 16 * sizeof(int)) != cudaSuccess) return -1;
	if (cudaMalloc(&g_s.pal, 256 * 3) != cudaSuccess) return -1;
	g_s.ok = true;
	return 0;
}

int finish(const char *what) {
	cudaError_t e = cudaDeviceSynchronize();
	if (e == cudaSuccess) e = cudaGetLastError();
	if (e != cudaSuccess) {
		fprintf(stderr, "[ref_shim] %s failed: %s\n", what, cudaGetErrorString(e));
		return -2;
	}
	return 0;
}
}  // namespace

extern "C" {

int ref_max_objects() { return MAX_OBJECTS; }

// = the launch at tsdf.cu:472-488.  Planes/images: device pointers.  Small blocks: host pointers.
int ref_integrate(float *sdf_d, uint8_t *color_d, uint32_t *cnt_d, int *wt_d,
	const int *dims, const float *vol_start, const float *voxel, float miu, const float *K16,
	uint16_t *depth_d, uint8_t *rgb_d, uint8_t *mask_d, const float *extr16, int width, int height)
{
	if (ensure_scratch()) return -1;
	float hf[64];
	memcpy(hf + 0, vol_start, 12);
	memcpy(hf + 4, voxel, 12);
	memcpy(hf + 8, K16, 64);
	memcpy(hf + 24, extr16, 64);
	cudaMemcpy(g_s.f, hf, sizeof(hf), cudaMemcpyHostToDevice);
	cudaMemcpy(g_s.i, dims, 12, cudaMemcpyHostToDevice);
	dim3 grid((dims[0] - 1) / 8 + 1, (dims[1] - 1) / 8 + 1, (dims[2] - 1) / 8 + 1), block(8, 8, 8);
	tsdf_kernel<<<grid, block>>>(sdf_d, color_d, cnt_d, wt_d, g_s.i, g_s.f + 0, g_s.f + 4, miu,
		g_s.f + 8, depth_d, rgb_d, mask_d, g_s.f + 24, width, height);
	return finish("tsdf_kernel");
}

// = the launch at tsdf.cu:441-455 (probs / box_mask must be zero-filled by the caller,
// as tsdf.cu:428-429 does).
int ref_back_proj(const float *Kinv16, const float *Rt9, const float *o3,
	const float *vol_start, const float *vol_end, const float *voxel, const int *dims,
	float *sdf_d, uint32_t *cnt_d, int width, int height, float *probs_d, bool *box_mask_d)
{
	if (ensure_scratch()) return -1;
	float hf[64];
	memset(hf, 0, sizeof(hf));
	memcpy(hf + 0, Kinv16, 64);
	memcpy(hf + 16, Rt9, 36);
	memcpy(hf + 28, o3, 12);
	memcpy(hf + 32, vol_start, 12);
	memcpy(hf + 36, vol_end, 12);
	memcpy(hf + 40, voxel, 12);
	cudaMemcpy(g_s.f, hf, sizeof(hf), cudaMemcpyHostToDevice);
	cudaMemcpy(g_s.i, dims, 12, cudaMemcpyHostToDevice);
	dim3 grid((width - 1) / 32 + 1, (height - 1) / 32 + 1, 1), block(32, 32, 1);
	back_proj_kernel<<<grid, block>>>(g_s.f + 0, g_s.f + 16, (float3 *)(g_s.f + 28), (float3 *)(g_s.f + 32),
		(float3 *)(g_s.f + 36), (float3 *)(g_s.f + 40), (int3 *)g_s.i, sdf_d, cnt_d, width, height,
		probs_d, box_mask_d);
	return finish("back_proj_kernel");
}

// = the launch at viewer.cu:152-166 (output must be zero-filled by the caller, viewer.cu:150).
int ref_show(const float *s2w16, const float *c3,
	const float *vol_start, const float *vol_end, const float *voxel, const int *dims,
	float *sdf_d, uint8_t *color_d, uint32_t *cnt_d, int width, int height, uint8_t *out_bgr_d,
	const uint8_t *palette /* MAX_OBJECTS*3 host bytes */)
{
	if (ensure_scratch()) return -1;
	float hf[64];
	memset(hf, 0, sizeof(hf));
	memcpy(hf + 0, s2w16, 64);
	memcpy(hf + 16, c3, 12);
	memcpy(hf + 20, vol_start, 12);
	memcpy(hf + 24, vol_end, 12);
	memcpy(hf + 28, voxel, 12);
	cudaMemcpy(g_s.f, hf, sizeof(hf), cudaMemcpyHostToDevice);
	cudaMemcpy(g_s.i, dims, 12, cudaMemcpyHostToDevice);
	cudaMemcpy(g_s.pal, palette, MAX_OBJECTS * 3, cudaMemcpyHostToDevice);
	dim3 grid((width - 1) / 32 + 1, (height - 1) / 32 + 1, 1), block(32, 32, 1);
	show_tsdf_kernel<<<grid, block>>>(g_s.f + 0, (float3 *)(g_s.f + 16), (float3 *)(g_s.f + 20),
		(float3 *)(g_s.f + 24), (float3 *)(g_s.f + 28), (int3 *)g_s.i, sdf_d, (uchar3 *)color_d, cnt_d,
		width, height, (uchar3 *)out_bgr_d, g_s.pal);
	return finish("show_tsdf_kernel");
}

// the reference's interp_tsdf_color at n given positions (device pointers; out must be zero-filled)
int ref_color_at(const float *xyz_d, const uint8_t *valid_d, int n, const float *vol_start, const float *voxel, const int *dims,
	uint8_t *color_d, uint8_t *out_d)
{
	shim_color_at_kernel<<<(n + 127) / 128, 128>>>(xyz_d, valid_d, n, make_float3(vol_start[0], vol_start[1], vol_start[2]),
		make_float3(voxel[0], voxel[1], voxel[2]), make_int3(dims[0], dims[1], dims[2]), (uchar3 *)color_d, (uchar3 *)out_d);
	return finish("interp_tsdf_color");
}

// CPU: verbatim TSDF::filter_overlaps (tsdf.cu:304-416).  All pointers are host pointers.
int ref_filter_overlaps(float *probs, int width, int height, uint8_t *mask, uint8_t *box_mask,
	uint32_t n_obs, int *num_objs_inout)
{
	return ref_filter_overlaps_impl(probs, width, height, mask, (bool *)box_mask, n_obs, num_objs_inout);
}

}  // extern "C"
