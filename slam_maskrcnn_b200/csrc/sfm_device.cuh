// sfm_device.cuh -- shared device-side types and exact-arithmetic helpers (sm_100a).
//
// Bit-exactness contract: every floating-point operation that decides an
// integer result is written with explicit round-to-nearest intrinsics (__fmaf_rn, __fmul_rn,
// __fadd_rn, __fdiv_rn) in the order the reference's compiled code evaluates them, so that this
// translation unit's own compiler flags can never re-contract or re-associate them.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#if defined(SFM_K1_TMA_DEPTH) && SFM_K1_TMA_DEPTH
#include <cuda.h>  // CUtensorMap (A/B build only; the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)
#endif

namespace sfm {

constexpr int kTile = 8;           // depth tile edge (pixels) of the per-frame max-depth grid
constexpr int kStatSlots = 256;    // spread slots for the per-frame U/S counters
constexpr int kMaxBins = 255;

// Internal histogram layout.  The reference stores u32 hist[v*L + label] (tsdf.cu:61): one 4-byte bin per 32-byte
// sector, and the ~10 near-surface voxels of a column (consecutive z) that a frame updates for the same label lie
// 10 sectors -- 10 DRAM row activations -- apart.  Here a column is cut into groups of kHistTZ planes and the bins of
// a group are stored bin-major, 16 bits each:
//     hist[((col*ngz + z/TZ)*L + label)*TZ + z%TZ],   col = x*Dy + y,   ngz = ceil(nz/TZ)
// so the same run of 10 voxels touches 3-4 sectors, the 8-tap gathers of the ray kernels read the same number of
// bytes as before (a z, z+1 pair of taps shares an 8-byte unit three times out of four), and the plane is half the
// size (512^3 x 80 bins: 21.5 GB).  A bin counts the frames that voted for it, at most one per frame, so 16 bits hold
// 65535 frames (enforced by the host; the reference's own frame cap is 100, kernel.cpp:60).  sfm_download / sfm_upload
// / sfm_plane_device_ptr convert to and from the reference layout and type.
typedef uint16_t hist_t;
constexpr int kHistTZ = 4;
__host__ __device__ __forceinline__ size_t hist_index(size_t col, int zl, int ngz, int bins, int label) {
	return ((col * (size_t)ngz + (size_t)(zl / kHistTZ)) * (size_t)bins + (size_t)label) * kHistTZ + (size_t)(zl % kHistTZ);
}

struct VolGeom {
	int Dx, Dy, Dz;  // global volume dimensions (vol_dim_, tsdf.cuh:52)
	int z0, nz;      // z-slab stored by this handle: global planes [z0, z0+nz)
	int own_z0, own_nz;  // planes this handle OWNS in a sharded ray-cast (stored minus the halo)
	float sx, sy, sz;  // vol_start_
	float ex, ey, ez;  // vol_end_
	float vx, vy, vz;  // vol_res_
	float miu;
	int oby, obz;         // strides of the 8x8x8 surface-block map (see Planes::occ)
	int oby2, obz2;       // strides of the coarse 32x32x32 level of the same map
	unsigned occ2_off;    // byte offset of the coarse level inside the map allocation
	int ngz;      // z groups per column of the tiled histogram: ceil(nz / kHistTZ)
	int zl_log2;  // K1 brick shape on the 128-bit path: 2^zl_log2 lanes (4 voxels each) along z per column, i.e. a
	              // brick is (32 >> zl_log2) columns x (4 << zl_log2) planes; 3 = 4 columns x 32 planes.  Thin z-slabs
	              // (a rank that owns only the planes around a fronto-parallel wall) use flatter bricks.
	int fastdiv;  // bit a set: dividing by voxel[a] may use the invariant-divisor sequence (k_raymarch.cuh)
};

struct Planes {
	float *sdf;
	int32_t *wt;
	uint8_t *color;
	hist_t *hist;  // tiled, see hist_index()
	int bins;
	// Surface-block map: one byte per 8x8x8 block of voxels, set (never cleared) when a voxel of the
	// block -- or a voxel one step beyond its low faces, i.e. a trilinear tap of a sample whose floor
	// index lies in the block -- has ever received a near-surface update (diff < near_gate).  A sample
	// whose floor index falls in an unset block only sees SDF values in {miu} U [near_gate, 1], so the
	// ray-marcher may skip gathering it: it can neither be a hit (f < 0) nor trigger the fine step.
	uint8_t *occ;
	int oby, obz;  // block-grid strides: index = (bx*oby + by)*obz + bz
	int oby2, obz2;        // the same for the coarse level of 32x32x32 blocks, stored at occ + occ2_off
	unsigned occ2_off;
};

#ifndef SFM_K1_TMA_DEPTH
#define SFM_K1_TMA_DEPTH 0  // 1: A/B build that stages K1b's depth gathers with cp.async.bulk.tensor.2d (profiles/README.md)
#endif
struct FrameView {
	const uint16_t *depth;
	const uint8_t *rgb;
	const uint8_t *mask;
	const uint16_t *tilemax;  // [TH][TW] max depth per kTile x kTile tile
	const uint16_t *tilemin;  // [TH][TW] min depth per tile, invalid (0) included: > 0 <=> no hole in the tile
	unsigned tile_bytes;      // bytes of [tilemax | tilemin] (contiguous, multiple of 16) for the TMA bulk copy
	const float *depth_m;     // [H][W] depth/depth_scale in f32 (IEEE divide), 0 = invalid
	int W, H, TW, TH;
	float E[12];  // rows 0..2 of extrinsic2init (row-major 3x4)
	float K[9];   // rows 0..2, cols 0..2 of the intrinsic matrix
	float depth_scale, near_gate;
	// conservative-cull constants, computed on the host per frame (see k_integrate.cuh)
	float cull_lin;     // max_r sum_c |E[r][c]|, c<3: amplification of voxel coordinates
	float cull_t;       // max_r |E[r][3]|
	float cull_k2;      // |K20|+|K21|+|K22|
	float cull_slack0;  // constant pixel slack
	int debug;          // ablation switches for profiling (0 in production)
#if SFM_K1_TMA_DEPTH
	const void *depth_tmap;  // host pointer to the CUtensorMap of depth_m (A/B build only; passed to K1b as a __grid_constant__ parameter)
#endif
};

// dot(float4 row,(p,1)) as the reference compiles it (helper_math.h:1249-1252; tsdf.cu:31-33):
//   r3 + fma(pz, r2, fma(px, r0, py*r1)).   `h` is the z-invariant part fma(px,r0,py*r1).
__device__ __forceinline__ float affine_hoist(float px, float py, float r0, float r1) {
	return __fmaf_rn(px, r0, __fmul_rn(py, r1));
}
__device__ __forceinline__ float affine_finish(float h, float pz, float r2, float r3) {
	return __fadd_rn(r3, __fmaf_rn(pz, r2, h));
}
// dot(float3,float3) as compiled (helper_math.h:1245-1248): fma(bz,az, fma(bx,ax, by*ay))
__device__ __forceinline__ float dot3_ref(float a0, float a1, float a2, float bx, float by, float bz) {
	return __fmaf_rn(bz, a2, __fmaf_rn(bx, a0, __fmul_rn(by, a1)));
}
// mix (utils.cu:93-96) as compiled: fma(a, 1-t, t*b)
__device__ __forceinline__ float mix_ref(float a, float b, float t) {
	return __fmaf_rn(a, __fadd_rn(1.f, -t), __fmul_rn(t, b));
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

}  // namespace sfm
