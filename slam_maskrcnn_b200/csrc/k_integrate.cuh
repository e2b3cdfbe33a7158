// k_integrate.cuh -- K0 (frame prep), K1a (brick classification) and K1b (TSDF update) for sm_100a.
//
// K1a + K1b replace tsdf_kernel (reference src/SfM_CUDA/tsdf.cu:18-70).  Design, B200-first:
//
//   * Work item = brick: CPW columns (consecutive y) x 32 consecutive z voxels at one x (thin z-slabs use
//     flatter bricks, VolGeom::zl_log2).  In the reference layout (z fastest) a brick is CPW full 128 B
//     lines of each plane.
//   * K1a (classify_kernel) sorts bricks into three classes with two per-frame 8x8-pixel tile grids (max
//     depth, min depth with invalid = 0) built by K0:
//        CULL   every voxel provably fails the reference's own tests (outside the image, only
//               invalid depth under it, or behind the surface band: cz - depth >= miu);
//        FREE   every voxel provably lands inside the image on a valid pixel and in front of the
//               surface band (depth - cz > miu), i.e. the reference computes diff = miu/miu = 1
//               for all of them: no per-voxel projection is needed at all;
//        MIXED  anything else: per-voxel evaluation, bit-identical to the reference.
//     The tests are conservative with an explicit rounding-error budget, so results never change
//     (SFM_FLAG_NO_CULL lists every brick as MIXED; the parity tests compare both).  Whole boxes of 32
//     bricks (super-blocks: kSbX x-planes x kSbG brick rows x one z chunk) are tested first, one THREAD
//     per box: about half the boxes of a 512^3 frame are CULL there; then one WARP per surviving box, one
//     brick per lane.  Surviving bricks go to a MIXED and a FREE list in global memory.
//   * K1b (integrate_kernel) is a persistent grid whose warps pull bricks from the lists.  One lane owns
//     VEC=4 consecutive z voxels of one column: SDF / weight move as 128-bit loads and stores, a
//     warp request is CPW full lines.  The z-invariant part of the pose transform
//     (fma(px,r0,py*r1), the order the reference's SASS uses) is hoisted per column.
//   * Exact shortcuts: diff clamped to miu gives exactly 1.0f; (1.0f*w + 1.0f)/(w+1) is exactly
//     1.0f; an SDF quad whose bits did not change is not written back.
//   * Per-frame U (weight increments) and S (histogram/colour updates) are folded with warp
//     shuffles + one spread atomic pair per warp -- they define the algorithmic bytes of the step.
//   * K0 and K1a never touch the volume: the host runs them on a second stream, one frame ahead of K1b
//     (sfm_api.cu: enqueue_prepare / enqueue_update).
#pragma once
#include "sfm_device.cuh"

namespace sfm {

// ---------------------------------------------------------------------------------------------
// K0: per-frame prep.  One warp per kTile x kTile tile: max depth, min depth (invalid pixels
// count as 0, so min > 0 <=> the tile has no hole), depth in metres as f32 (the reference's
// depth/5000.f, IEEE divide, tsdf.cu:49) and max label (labels >= bins are a contract violation,
// SURVEY appendix B.2).  Also resets K1's list counters and fetch cursor.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_frame_kernel(const uint16_t *__restrict__ depth,
	const uint8_t *__restrict__ mask, int W, int H, int TW, int TH, int bins, float depth_scale,
	uint16_t *__restrict__ tilemax, uint16_t *__restrict__ tilemin, float *__restrict__ depth_m,
	uint32_t *__restrict__ err, unsigned *__restrict__ work_counter)
{
	const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
	if (gtid < 3) work_counter[gtid] = 0u;  // WorkLists::counts: list sizes and K1b's fetch cursor
	const int warp = gtid >> 5, lane = threadIdx.x & 31;
	if (warp >= TW * TH) return;
	const int ty = warp / TW, tx = warp % TW;
	unsigned dmax = 0, dmin = 0xffffu, lmax = 0;
#pragma unroll
	for (int k = 0; k < 2; k++) {
		const int r = (lane >> 2), c = ((lane & 3) << 1) + k;
		const int y = ty * kTile + r, x = tx * kTile + c;
		if (x < W && y < H) {
			const unsigned d = depth[y * W + x];
			dmax = max(dmax, d);
			dmin = min(dmin, d);
			depth_m[y * W + x] = __fdiv_rn((float)d, depth_scale);
			if (mask) lmax = max(lmax, (unsigned)mask[y * W + x]);
		}
	}
	dmax = __reduce_max_sync(0xffffffffu, dmax);
	dmin = __reduce_min_sync(0xffffffffu, dmin);
	lmax = __reduce_max_sync(0xffffffffu, lmax);
	if (lane == 0) {
		tilemax[ty * TW + tx] = (uint16_t)dmax;
		tilemin[ty * TW + tx] = (uint16_t)dmin;  // invalid pixels count as 0: min > 0 <=> no hole in the tile
		if (bins > 0 && (int)lmax >= bins) atomicOr(err, 1u);
	}
}

// ---------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------
// tsdf.cu:39-44: ix = floor(RN(sx/sz)), iy = floor(RN(sy/sz)).  Only the floors are needed, so the two
// IEEE divides are replaced by one MUFU.RCP and two multiplies: q~ = sx*rcp(sz) is within
// 2.4e-7*|q| of the correctly rounded quotient q (rcp.approx: 1 ulp, the product: 1/2 ulp, q itself:
// 1/2 ulp), hence floor(q~) == floor(q) whenever q~ is farther than that from every integer.  The test
// uses twice that distance; otherwise (about 1 voxel in 10^4, and for NaN/inf/huge values) the exact
// divides decide.  Bit-identical to the reference by construction.
// Rarely executed exact paths are kept out of line: with everything inlined the kernel is 43 KB of
// SASS and the profile shows instruction-fetch stalls (the instruction cache holds 32 KB).
__device__ __noinline__ float exact_div(float a, float b) { return __fdiv_rn(a, b); }

// true when floor(q~) is guaranteed to equal floor(RN(s/sz)) for both coordinates (see pixel_floor)
__device__ __forceinline__ bool pixel_floor_is_safe(float qx, float qy) {
	return fabsf(__fadd_rn(qx, -rintf(qx))) > __fmaf_rn(fabsf(qx), 4.8e-7f, 1e-30f) &&
		fabsf(__fadd_rn(qy, -rintf(qy))) > __fmaf_rn(fabsf(qy), 4.8e-7f, 1e-30f);
}

__device__ __forceinline__ void pixel_floor(float sx, float sy, float sz, int &ix, int &iy) {
	float r;
	asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(sz));
	const float qx = __fmul_rn(sx, r), qy = __fmul_rn(sy, r);
	const bool okx = fabsf(__fadd_rn(qx, -rintf(qx))) > __fmaf_rn(fabsf(qx), 4.8e-7f, 1e-30f);
	const bool oky = fabsf(__fadd_rn(qy, -rintf(qy))) > __fmaf_rn(fabsf(qy), 4.8e-7f, 1e-30f);
	if (okx && oky) {
		ix = __float2int_rd(qx);
		iy = __float2int_rd(qy);
	} else {
		ix = __float2int_rd(exact_div(sx, sz));
		iy = __float2int_rd(exact_div(sy, sz));
	}
}

// tsdf.cu:56  (sdf*w + diff)/(w+1)  -> FFMA, IEEE divide.  (1.0f*w + 1.0f)/(w+1) is exactly 1.0f
// for 0 <= w < 2^24 (both the fma and the quotient are exact), so that case skips the divide.
__device__ __forceinline__ float sdf_update(float s, int w, float diff) {
	if (s == 1.0f && diff == 1.0f && (unsigned)w < (1u << 24)) return 1.0f;
	return exact_div(__fmaf_rn(s, (float)w, diff), (float)(w + 1));
}

template <int VEC> struct VecT;
template <> struct VecT<4> { using F = float4; using I = int4; };
template <> struct VecT<1> { using F = float; using I = int; };

__device__ __forceinline__ bool same_bits(const float4 &a, const float4 &b) {
	return __float_as_uint(a.x) == __float_as_uint(b.x) && __float_as_uint(a.y) == __float_as_uint(b.y) &&
		__float_as_uint(a.z) == __float_as_uint(b.z) && __float_as_uint(a.w) == __float_as_uint(b.w);
}
__device__ __forceinline__ bool same_bits(const float &a, const float &b) {
	return __float_as_uint(a) == __float_as_uint(b);
}

enum BrickClass { kCull = 0, kMixed = 1, kFree = 2 };

#ifndef SFM_K1_THREADS
#define SFM_K1_THREADS 128  // threads per K1b block (7 resident blocks per SM leave room for one K1a block)
#endif
#ifndef SFM_K1_MIN_BLOCKS
#define SFM_K1_MIN_BLOCKS (1024 / SFM_K1_THREADS)  // resident blocks per SM K1b is compiled for (register cap 64)
#endif
constexpr int kK1Threads = SFM_K1_THREADS;

constexpr int kSbX = 8;  // super-block: kSbX x-planes x kSbG brick rows x one 32-z chunk = 32 bricks
constexpr int kSbG = 4;

// s = K[0:3,0:3] * c in the reference's compiled order (tsdf.cu:35-37 -> dot3_ref).  KCANON: the
// intrinsic matrix has the pinhole pattern [[fx,0,cx],[0,fy,cy],[0,0,1]] (kernel.cpp:39-40,
// tsdf.cu:137-150 can build no other).  Then the zero terms drop out bit-exactly for finite
// inputs -- fma(cx,fx, cy*0) == RN(cx*fx), fma(cx,0, cy*fy) == RN(cy*fy), fma(cz,1, 0) == cz --
// up to the sign of an exact zero, which neither floor(s/sz) nor the bounds test can see.
template <bool KCANON>
__device__ __forceinline__ void cam_to_screen(const FrameView &f, float cx, float cy, float cz, float &sx, float &sy, float &sz) {
	if (KCANON) {
		sx = __fmaf_rn(cz, f.K[2], __fmul_rn(cx, f.K[0]));
		sy = __fmaf_rn(cz, f.K[5], __fmul_rn(cy, f.K[4]));
		sz = cz;
	} else {
		sx = dot3_ref(f.K[0], f.K[1], f.K[2], cx, cy, cz);
		sy = dot3_ref(f.K[3], f.K[4], f.K[5], cx, cy, cz);
		sz = dot3_ref(f.K[6], f.K[7], f.K[8], cx, cy, cz);
	}
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2_keep(const void *p) { asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p)); }

// L2 eviction priorities (createpolicy: the 64-bit policy lives in a uniform register, SASS `desc[UR]`).  One integrate
// step streams ~0.25 GB of SDF / weight lines through the 126 MB L2 exactly once (evict_first), while the colour and
// histogram sectors of the near-surface voxels -- ~40 MB per frame, touched by random 4-byte read-modify-writes whose
// DRAM cost is a row activation each -- are the same ones the NEXT frame needs (evict_last): kept in L2 across
// frames, the surface path stops being bound by DRAM activations.
__device__ __forceinline__ unsigned long long l2_policy_stream() {
	unsigned long long p;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ unsigned long long l2_policy_keep() {
	unsigned long long p;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ float4 ld_hint(const float4 *a, unsigned long long pol) {
	float4 v;
	asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(pol));
	return v;
}
__device__ __forceinline__ int4 ld_hint(const int4 *a, unsigned long long pol) {
	int4 v;
	asm volatile("ld.global.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(a), "l"(pol));
	return v;
}
__device__ __forceinline__ float ld_hint(const float *a, unsigned long long pol) {
	float v;
	asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
	return v;
}
__device__ __forceinline__ int ld_hint(const int *a, unsigned long long pol) {
	int v;
	asm volatile("ld.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
	return v;
}
__device__ __forceinline__ void st_hint(float4 *a, const float4 &v, unsigned long long pol) {
	asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(int4 *a, const int4 &v, unsigned long long pol) {
	asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(float *a, const float &v, unsigned long long pol) {
	asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(a), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint(int *a, const int &v, unsigned long long pol) {
	asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ int ld_u8_hint(const uint8_t *a, unsigned long long pol) {
	unsigned v;
	asm volatile("ld.global.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
	return (int)v;
}
__device__ __forceinline__ void st_u8_hint(uint8_t *a, int v, unsigned long long pol) {
	asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_hint(uint32_t *a, unsigned inc, unsigned long long pol) {
	asm volatile("red.global.add.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(a), "r"(inc), "l"(pol) : "memory");
}

constexpr int kTmaBoxW = 32, kTmaBoxH = 16;  // SFM_K1_TMA_DEPTH build: pixels of depth_m staged per brick
constexpr int kQueue = 160;  // per-warp capacity of the deferred near-surface queue (one brick = 128 voxels, drained at >= 32)

// Colour running mean + histogram increment (tsdf.cu:57-62) for the queued near-surface voxels, one
// voxel per lane.  Every voxel appears at most once per frame, so the colour read-modify-write needs no
// atomics (byte stores: neighbouring voxels' bytes are never touched).  The histogram increment is issued
// as a reduction (RED.ADD, no return value): the add happens in L2, the warp never waits for the bin's
// sector to arrive from DRAM -- with a load / store pair every drain stalled on 32 random sectors.
template <bool LABELS>
__device__ __noinline__ unsigned drain_surface_queue(const Planes &p, const VolGeom &g, const FrameView &f, const uint4 *q, int count,
	int lane, uint32_t *err)
{
	__syncwarp();
	unsigned done = 0;
	const unsigned long long keep = l2_policy_keep();
	for (int base = 0; base < count; base += 32) {
		const int i = base + lane;
		if (i < count) {
			const uint4 e = q[i];  // {column, local z, pixel, weight}
			const size_t v = (size_t)e.x * (size_t)g.nz + e.y;
			const int img = (int)e.z, w = (int)e.w;
			const uint8_t *src = f.rgb + (size_t)img * 3;
			uint8_t *dst = p.color + v * 3;
			const int s0 = __ldg(src), s1 = __ldg(src + 1), s2 = __ldg(src + 2);
			const int c0 = ld_u8_hint(dst, keep), c1 = ld_u8_hint(dst + 1, keep), c2 = ld_u8_hint(dst + 2, keep);
			if (LABELS) {
				const unsigned label = __ldg(f.mask + img);
				if ((int)label < p.bins) {
					// 16-bit bin inside an aligned 32-bit word: add 1 to its half (no carry: a bin never exceeds the
					// number of frames, which the host caps at 65535)
					const size_t hi = hist_index(e.x, (int)e.y, g.ngz, p.bins, (int)label);
					red_add_hint(reinterpret_cast<uint32_t *>(p.hist + (hi & ~(size_t)1)), (hi & 1) ? 0x10000u : 1u, keep);
				} else atomicOr(err, 1u);
			}
			st_u8_hint(dst, (c0 * w + s0) / (w + 1), keep);
			st_u8_hint(dst + 1, (c1 * w + s1) / (w + 1), keep);
			st_u8_hint(dst + 2, (c2 * w + s2) / (w + 1), keep);
			done++;
		}
	}
	__syncwarp();
	return done;
}

// Projected bounds of a box of voxels: pixel bounding box of its corners, range of the homogeneous
// and camera-space depth, magnitude bound of the camera-space coordinates (rounding-error budget).
struct BoxBounds {
	float umin, umax, vmin, vmax, szmin, szmax, czmin, czmax, scale_c;
};

// Second half of the conservative classification, shared by bricks and super-blocks (one lane each).
// The projective map sends segments that stay on one side of the camera plane to segments, so with
// all corners of a (convex) box strictly on one side the pixel coordinates of every voxel of the box
// lie inside the bounding box of the projected corners (plus rounding slack), and the camera-space
// depth, being affine, takes its extremes at corners.  Rounding-error budget: |c*| <= scale_c,
// per-voxel error of c*, s* ~ 1e-6*scale; a box is only classified when all corners are at least
// 1e-2*scale_sz away from the camera plane, which bounds the per-voxel pixel error by
// ~1e-4*(Krow/K2row + |u|) -- the slack is 10x that.
__device__ __forceinline__ int classify_bounds(const FrameView &f, const VolGeom &g, const uint16_t *tilemax,
	const uint16_t *tilemin, const BoxBounds &b, int max_tiles)
{
	const float zguard = 1e-2f * f.cull_k2 * b.scale_c;
	const bool one_side = (b.szmin > zguard) || (b.szmax < -zguard);
	if (!one_side) return kMixed;
	const float slack_u = f.cull_slack0 + 1e-3f * fmaxf(fabsf(b.umin), fabsf(b.umax));
	const float slack_v = f.cull_slack0 + 1e-3f * fmaxf(fabsf(b.vmin), fabsf(b.vmax));
	const float ulo = b.umin - slack_u, uhi = b.umax + slack_u;
	const float vlo = b.vmin - slack_v, vhi = b.vmax + slack_v;
	if (uhi < 0.f || ulo >= (float)f.W || vhi < 0.f || vlo >= (float)f.H) return kCull;  // outside the image
	const bool inside = ulo >= 0.f && uhi < (float)f.W && vlo >= 0.f && vhi < (float)f.H;
	const int tx0 = max(0, (int)floorf(ulo)) / kTile, tx1 = min(f.W - 1, (int)floorf(uhi)) / kTile;
	const int ty0 = max(0, (int)floorf(vlo)) / kTile, ty1 = min(f.H - 1, (int)floorf(vhi)) / kTile;
	const int tw = tx1 - tx0 + 1, nt = tw * (ty1 - ty0 + 1);
	if (nt > max_tiles) return kMixed;  // huge footprint (box close to the camera)
	unsigned dmax = 0, dmin = 0xffffu;
	for (int ty = ty0; ty <= ty1; ty++)
		for (int tx = tx0; tx <= tx1; tx++) {
			dmax = max(dmax, (unsigned)tilemax[ty * f.TW + tx]);
			dmin = min(dmin, (unsigned)tilemin[ty * f.TW + tx]);
		}
	if (dmax == 0) return kCull;  // only invalid depth under the box
	if (b.szmin > 0.f) {
		const float dmax_m = __fdiv_rn((float)dmax, f.depth_scale), dmin_m = __fdiv_rn((float)dmin, f.depth_scale);
		const float eps = 1e-4f * (b.scale_c + dmax_m);
		// every voxel: diff = d/scale - cz <= dmax/scale - czmin + eps  =>  behind the surface band
		if (b.czmin - dmax_m >= g.miu + eps) return kCull;
		// every voxel: valid pixel inside the image (no tile under the box has a hole: min > 0) and
		// diff >= dmin/scale - czmax - eps > miu   =>  diff clamps to miu, i.e. exactly 1.0f
		// (An exact per-pixel invalid bitmap was measured here: it lifts the FREE share from 15 % to
		// 23 % of the surviving bricks but costs more instructions than it saves; see DESIGN.md.)
		if (inside && dmin > 0 && dmin_m - b.czmax > g.miu + eps) return kFree;
	}
	return kMixed;
}

// approximate projection of one box corner (classification only: the slack is 1e-3 relative)
__device__ __forceinline__ void project_corner(const FrameView &f, float px, float py, float pz, float &u, float &v,
	float &sz, float &cz, float &scale)
{
	const float cx = affine_finish(affine_hoist(px, py, f.E[0], f.E[1]), pz, f.E[2], f.E[3]);
	const float cy = affine_finish(affine_hoist(px, py, f.E[4], f.E[5]), pz, f.E[6], f.E[7]);
	cz = affine_finish(affine_hoist(px, py, f.E[8], f.E[9]), pz, f.E[10], f.E[11]);
	const float sx = dot3_ref(f.K[0], f.K[1], f.K[2], cx, cy, cz);
	const float sy = dot3_ref(f.K[3], f.K[4], f.K[5], cx, cy, cz);
	sz = dot3_ref(f.K[6], f.K[7], f.K[8], cx, cy, cz);
	u = __fdividef(sx, sz);
	v = __fdividef(sy, sz);
	scale = f.cull_lin * (fabsf(px) + fabsf(py) + fabsf(pz) + 1.f) + f.cull_t;
}

// Stage A: classify one brick (x, columns y0..ylast, local z zc0..zc1), one lane.
__device__ __forceinline__ int classify_brick(const FrameView &f, const VolGeom &g, const uint16_t *tilemax,
	const uint16_t *tilemin, int x, int y0, int ylast, int zc0, int zc1)
{
	const float px = __fmaf_rn((float)x, g.vx, g.sx);
	BoxBounds b{INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY, 0.f};
	bool finite = true;
#pragma unroll
	for (int corner = 0; corner < 4; corner++) {
		const float py = __fmaf_rn((float)((corner & 1) ? ylast : y0), g.vy, g.sy);
		const float pz = __fmaf_rn((float)(g.z0 + ((corner & 2) ? zc1 : zc0)), g.vz, g.sz);
		float u, v, sz, cz, scale;
		project_corner(f, px, py, pz, u, v, sz, cz, scale);
		b.umin = fminf(b.umin, u); b.umax = fmaxf(b.umax, u);
		b.vmin = fminf(b.vmin, v); b.vmax = fmaxf(b.vmax, v);
		b.szmin = fminf(b.szmin, sz); b.szmax = fmaxf(b.szmax, sz);
		b.czmin = fminf(b.czmin, cz); b.czmax = fmaxf(b.czmax, cz);
		b.scale_c = fmaxf(b.scale_c, scale);
		finite &= fabsf(u) < 1e8f && fabsf(v) < 1e8f && fabsf(cz) < 1e30f;  // NaNs fail
	}
	if (!finite) return kMixed;
	return classify_bounds(f, g, tilemax, tilemin, b, 96);
}

// One lane classifies a whole box (8 corners): the super-block pre-pass of K1a.
__device__ __forceinline__ int classify_box_lane(const FrameView &f, const VolGeom &g, const uint16_t *tilemax,
	const uint16_t *tilemin, int x0, int x1, int y0, int y1, int zc0, int zc1)
{
	BoxBounds b{INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY, 0.f};
	bool finite = true;
#pragma unroll 2
	for (int corner = 0; corner < 8; corner++) {
		const float px = __fmaf_rn((float)((corner & 1) ? x1 : x0), g.vx, g.sx);
		const float py = __fmaf_rn((float)((corner & 2) ? y1 : y0), g.vy, g.sy);
		const float pz = __fmaf_rn((float)(g.z0 + ((corner & 4) ? zc1 : zc0)), g.vz, g.sz);
		float u, v, sz, cz, scale;
		project_corner(f, px, py, pz, u, v, sz, cz, scale);
		b.umin = fminf(b.umin, u); b.umax = fmaxf(b.umax, u);
		b.vmin = fminf(b.vmin, v); b.vmax = fmaxf(b.vmax, v);
		b.szmin = fminf(b.szmin, sz); b.szmax = fmaxf(b.szmax, sz);
		b.czmin = fminf(b.czmin, cz); b.czmax = fmaxf(b.czmax, cz);
		b.scale_c = fmaxf(b.scale_c, scale);
		finite &= fabsf(u) < 1e8f && fabsf(v) < 1e8f && fabsf(cz) < 1e30f;  // NaNs fail
	}
	if (!finite) return kMixed;
	return classify_bounds(f, g, tilemax, tilemin, b, 160);
}

// Brick work lists written by K1a and consumed by K1b.  A brick id packs (x, brick row, z chunk).
struct WorkLists {
	uint32_t *mixed;   // bricks that need per-voxel evaluation
	uint32_t *free_;   // bricks whose voxels all get diff == 1.0f
	unsigned *counts;  // [0] = #mixed, [1] = #free, [2] = K1b's fetch cursor   (zeroed by K0)
};
constexpr int kIdXShift = 21, kIdGShift = 10;  // id = x << 21 | row << 10 | chunk  (x, row < 2048, chunk < 1024)
#ifndef SFM_K1_FETCH
#define SFM_K1_FETCH 4
#endif
constexpr int kFetch = SFM_K1_FETCH;           // bricks a K1b warp takes per fetch
#ifndef SFM_K1_PERMUTE
#define SFM_K1_PERMUTE 1
#endif

constexpr int kK1aThreads = 128;               // K1a block size: small enough to run next to a resident wave of K1b
constexpr int kSbPerBlock = 64;                // most super-blocks one K1a block classifies (sizes its shared lists)

// ---------------------------------------------------------------------------------------------
// K1a: classification.  One warp per super-block (static stride over a persistent grid): the whole
// warp classifies the box, then -- unless the box is CULL or FREE as a whole -- each lane classifies
// one brick.  Surviving bricks are appended to the MIXED / FREE lists with one atomic per warp and
// list.  The tile grids are staged into shared memory once per block by a TMA bulk copy.
// ---------------------------------------------------------------------------------------------
template <bool VEC4, bool CULL, bool TMA_TILES>
__global__ void __launch_bounds__(kK1aThreads) classify_kernel(VolGeom g, FrameView f, WorkLists wl)
{
	const int zsh = VEC4 ? g.zl_log2 : 5;   // log2 lanes along z per column
	const int CPW = 32 >> zsh;              // columns per brick
	const int csh = zsh + (VEC4 ? 2 : 0);   // log2 planes per z chunk
	const int lane = threadIdx.x & 31;
	const int nchunks = (g.nz + (1 << csh) - 1) >> csh;
	const int groups_per_x = (g.Dy + CPW - 1) / CPW;
	const int nsby = (groups_per_x + kSbG - 1) / kSbG;
	const unsigned nsb = (unsigned)((g.Dx + kSbX - 1) / kSbX) * (unsigned)nsby * (unsigned)nchunks;
	extern __shared__ __align__(128) unsigned char smem_dyn[];
	const uint16_t *s_tilemax = TMA_TILES ? reinterpret_cast<const uint16_t *>(smem_dyn) : f.tilemax;
	const uint16_t *s_tilemin = TMA_TILES ? s_tilemax + f.TW * f.TH : f.tilemin;
	if (CULL && TMA_TILES) {
		// cp.async.bulk (TMA, SASS: UBLKCP) global -> shared, completion on an mbarrier; the persistent
		// block does this once and then serves every tile query of its super-blocks from shared memory
		// (SFM_FLAG_NO_TMA reads the grids through L1 instead)
		__shared__ __align__(8) unsigned long long tile_bar;
		const unsigned bar = (unsigned)__cvta_generic_to_shared(&tile_bar);
		const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dyn);
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
			asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		}
		__syncthreads();
		if (threadIdx.x == 0) {
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(f.tile_bytes) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				::"r"(dst), "l"(f.tilemax), "r"(f.tile_bytes), "r"(bar) : "memory");
		}
		unsigned done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
				: "=r"(done) : "r"(bar) : "memory");
		}
	}
	// Surviving bricks are collected in shared memory and appended to the global lists with ONE atomic
	// per block and list: tens of thousands of atomics on a single address would serialise in L2
	// (measured: 20 us of a 29 us kernel).  The host sizes the grid so that a block sees <= kSbPerBlock
	// super-blocks.
	__shared__ unsigned s_cnt[2], s_base[2];
	uint32_t *s_mixed = reinterpret_cast<uint32_t *>(smem_dyn + (TMA_TILES ? f.tile_bytes : 0));
	uint32_t *s_free = s_mixed + kSbPerBlock * 32;
	// The block owns the super-blocks blockIdx.x + k * gridDim.x (a sample of the whole volume, so every
	// block gets about the same mix of culled and surviving boxes).
	// Pass 1, one THREAD per super-block: the box test (8 corners, tiles under the bounding box).  About half
	// the boxes of a frame are culled here for ~25 warp instructions each (a whole warp per box costs 170),
	// and a rank whose slab lies behind the surfaces is done after this pass.  Survivors go to a queue.
	// Pass 2, one WARP per surviving super-block, pulled from the queue: one brick per lane.
	__shared__ unsigned s_queue[kSbPerBlock], s_qn, s_qhead;
	if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
	if (threadIdx.x == 2) s_qn = 0;
	if (threadIdx.x == 3) s_qhead = 0;
	__syncthreads();
	auto sb_box = [&](unsigned sb, int &x0, int &gy0, int &sbz, int &zc0, int &zc1) {  // z chunk fastest
		sbz = (int)(sb % (unsigned)nchunks);
		const unsigned sbt = sb / (unsigned)nchunks;
		gy0 = (int)(sbt % (unsigned)nsby) * kSbG;
		x0 = (int)(sbt / (unsigned)nsby) * kSbX;
		zc0 = sbz << csh;
		zc1 = min(zc0 + (1 << csh) - 1, g.nz - 1);
	};
	for (unsigned k = threadIdx.x;; k += blockDim.x) {
		const unsigned long long sb64 = (unsigned long long)blockIdx.x + (unsigned long long)k * gridDim.x;
		if (sb64 >= nsb) break;
		int x0, gy0, sbz, zc0, zc1;
		sb_box((unsigned)sb64, x0, gy0, sbz, zc0, zc1);
		int sbcls = kMixed;
		if (CULL && !(f.debug & 8))
			sbcls = classify_box_lane(f, g, s_tilemax, s_tilemin, x0, min(x0 + kSbX - 1, g.Dx - 1), gy0 * CPW,
				min((gy0 + kSbG) * CPW - 1, g.Dy - 1), zc0, zc1);
		if (sbcls != kCull) s_queue[atomicAdd(&s_qn, 1u)] = (unsigned)sb64 | (sbcls == kFree ? 0x80000000u : 0u);
	}
	__syncthreads();
	const unsigned qn = s_qn;
	for (;;) {
		unsigned qi = 0;
		if (lane == 0) qi = atomicAdd(&s_qhead, 1u);
		qi = __shfl_sync(0xffffffffu, qi, 0);
		if (qi >= qn) break;
		const unsigned entry = s_queue[qi];
		const int sbcls = (entry & 0x80000000u) ? kFree : kMixed;
		int x0, gy0, sbz, zc0, zc1;
		sb_box(entry & 0x7fffffffu, x0, gy0, sbz, zc0, zc1);
		int cls = kCull;
		const int bx = x0 + (lane >> 2), bg = gy0 + (lane & 3), by0 = bg * CPW;
		if (bx < g.Dx && by0 < g.Dy)
			cls = (!CULL) ? kMixed : (sbcls == kFree) ? kFree :
				classify_brick(f, g, s_tilemax, s_tilemin, bx, by0, min(by0 + CPW - 1, g.Dy - 1), zc0, zc1);
		const unsigned mm = __ballot_sync(0xffffffffu, cls == kMixed), fm = __ballot_sync(0xffffffffu, cls == kFree);
		unsigned bm = 0, bf = 0;
		if (lane == 0) {
			if (mm) bm = atomicAdd(&s_cnt[0], (unsigned)__popc(mm));
			if (fm) bf = atomicAdd(&s_cnt[1], (unsigned)__popc(fm));
		}
		bm = __shfl_sync(0xffffffffu, bm, 0);
		bf = __shfl_sync(0xffffffffu, bf, 0);
		const unsigned id = ((unsigned)bx << kIdXShift) | ((unsigned)bg << kIdGShift) | (unsigned)sbz;
		const unsigned below = (1u << lane) - 1u;
		if (cls == kMixed) s_mixed[bm + __popc(mm & below)] = id;
		else if (cls == kFree) s_free[bf + __popc(fm & below)] = id;
	}
	__syncthreads();
	if (threadIdx.x < 2 && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(wl.counts + threadIdx.x, s_cnt[threadIdx.x]);
	__syncthreads();
	for (unsigned i = threadIdx.x; i < s_cnt[0]; i += blockDim.x) wl.mixed[s_base[0] + i] = s_mixed[i];
	for (unsigned i = threadIdx.x; i < s_cnt[1]; i += blockDim.x) wl.free_[s_base[1] + i] = s_free[i];
}

// ---------------------------------------------------------------------------------------------
// K1b: update.  Persistent warps pull kFetch bricks at a time from the lists (MIXED first, the
// cheap FREE bricks last, so the tail of the kernel is made of small items) -- every warp gets the
// same amount of work to within a few bricks, whatever the geometry of the frame.
// ---------------------------------------------------------------------------------------------
template <int VEC, bool LABELS, bool KCANON>
__global__ void __launch_bounds__(SFM_K1_THREADS, SFM_K1_MIN_BLOCKS) integrate_kernel(Planes p, VolGeom g, FrameView f, WorkLists wl,
	unsigned long long *__restrict__ stats, uint32_t *__restrict__ err, const int32_t *__restrict__ gate
#if SFM_K1_TMA_DEPTH
	, const __grid_constant__ CUtensorMap depth_tmap
#endif
	)
{
	// `gate` (nullable): a device flag set by an earlier kernel of the same frame (the merge decision overflowed the
	// histogram's bins): the frame must then leave the volume untouched
	if (gate && *gate) return;
	const int zsh = VEC == 4 ? g.zl_log2 : 5;  // log2 lanes along z per column (3: a brick is 4 columns x 32 planes)
	const int CPW = 32 >> zsh;                 // columns per brick
	const int csh = zsh + (VEC == 4 ? 2 : 0);  // log2 planes per z chunk
	using F = typename VecT<VEC>::F;
	using I = typename VecT<VEC>::I;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int zq = lane & ((1 << zsh) - 1), ci = lane >> zsh;
	unsigned nU = 0, nS = 0;
	// dynamic shared memory: per warp, kQueue deferred near-surface voxels {column, local z, pixel, weight}
	extern __shared__ __align__(128) unsigned char smem_dyn[];
	uint4 *q = reinterpret_cast<uint4 *>(smem_dyn) + warp * kQueue;
	int qcount = 0;  // warp-uniform
#if SFM_K1_TMA_DEPTH
	// A/B build: per warp a kTmaBoxH x kTmaBoxW f32 tile of depth_m, filled by cp.async.bulk.tensor.2d (SASS: UTMALDG)
	// at the top-left corner of the brick's pixel footprint, completion on a per-warp mbarrier
	float *tile = reinterpret_cast<float *>(smem_dyn + (kK1Threads / 32) * kQueue * sizeof(uint4)) + warp * (kTmaBoxW * kTmaBoxH);
	const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
	const unsigned bar_s = (unsigned)__cvta_generic_to_shared(smem_dyn + (kK1Threads / 32) * (kQueue * sizeof(uint4) + kTmaBoxW * kTmaBoxH * 4) + warp * 8);
	unsigned tma_parity = 0;
	if (lane == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
#endif
	const unsigned nmixed = wl.counts[0], total = nmixed + wl.counts[1];
	auto brick_coords = [&](unsigned id, int &x, int &y, int &zl) {  // warp-uniform id + lane offsets
		x = (int)(id >> kIdXShift);
		y = (int)((id >> kIdGShift) & ((1u << (kIdXShift - kIdGShift)) - 1u)) * CPW + ci;
		zl = (int)((id & ((1u << kIdGShift) - 1u)) << csh) + zq * VEC;
	};
	auto brick_voxel = [&](int x, int y, int zl, bool ok) { return ((size_t)x * g.Dy + (ok ? y : 0)) * (size_t)g.nz + (ok ? zl : 0); };
	// Software pipeline: the SDF / weight quads of brick i+1 are requested before brick i is
	// evaluated, so every warp keeps two bricks' worth of 128-bit loads in flight; the atomic that
	// fetches the next group is in flight while this group is evaluated.
	F sv_n{}; I wv_n{};
	auto issue = [&](unsigned id) {
		int x, y, zl;
		brick_coords(id, x, y, zl);
		if ((y < g.Dy) && (zl < g.nz)) {
			const size_t v = brick_voxel(x, y, zl, true);
			sv_n = *reinterpret_cast<const F *>(p.sdf + v);
			wv_n = *reinterpret_cast<const I *>(p.wt + v);
		}
	};
	// list position -> brick id.  SFM_K1_PERMUTE: positions walk each list with a stride coprime to its
	// length, so bricks that sit next to each other in space (K1a appends 32 neighbours at a time) are
	// evaluated far apart in time: warps that chase histogram sectors and warps that do arithmetic mix
	// on every SM instead of the whole machine hitting the same kind of brick at once.
	const unsigned nfree = total - nmixed;
#if SFM_K1_PERMUTE
	const unsigned pm = (nmixed % 4093u) ? 4093u : 4091u, pf = (nfree % 4093u) ? 4093u : 4091u;
#endif
	auto list_at = [&](unsigned i) {
#if SFM_K1_PERMUTE
		if (i < nmixed) return wl.mixed[nmixed < (1u << 20) ? (i * pm) % nmixed : i];
		const unsigned k = i - nmixed;
		return wl.free_[nfree < (1u << 20) ? (k * pf) % nfree : k];
#else
		return i < nmixed ? wl.mixed[i] : wl.free_[i - nmixed];
#endif
	};
	// lane 0 holds the result; it is broadcast when the group is needed, one group of bricks later.  Around an atomic add
	// on a warp-uniform address ptxas builds its warp-aggregation idiom (leader election, ATOMG, shuffle of the result to
	// the participants) even when one lane takes part, and that shuffle waits for the atomic's round trip on the spot:
	// 11 % of K1b's stall samples.  `opaque0` is 0 at run time but not provably uniform, which keeps the plain ATOMG.
	unsigned opaque0;
	asm("{\n\t.reg .u32 a, b;\n\tmov.u32 a, %%laneid;\n\tmov.u32 b, %%smid;\n\tadd.u32 a, a, b;\n\tshr.u32 %0, a, 24;\n\t}" : "=r"(opaque0));
	unsigned *const cursor = wl.counts + 2 + opaque0;
	auto fetch = [&]() {
		unsigned b = 0;
		if (lane == 0) b = atomicAdd(cursor, (unsigned)kFetch);
		return b;
	};
	// prefetch.global.L2 of the SDF / weight lines of bricks 1.. of a freshly fetched group (brick 0 is
	// requested into registers right away): by the time the 128-bit loads ask for them they sit in L2
	auto prefetch_group = [&](unsigned ids, int cnt) {
		// the lanes that start a column segment (zq == 0) request its SDF and weight lines
#pragma unroll
		for (int bi = 1; bi < kFetch; bi++) {
			const unsigned id = __shfl_sync(0xffffffffu, ids, bi);
			if (bi < cnt && zq == 0 && !(f.debug & 32)) {
				int x, y, zl;
				brick_coords(id, x, y, zl);
				if (y < g.Dy) {
					const size_t v = ((size_t)x * g.Dy + y) * (size_t)g.nz + zl;
					prefetch_l2(p.sdf + v);
					prefetch_l2(p.wt + v);
				}
			}
		}
	};
	unsigned base = __shfl_sync(0xffffffffu, fetch(), 0);
	int n = base < total ? (int)min((unsigned)kFetch, total - base) : 0;
	unsigned my_id = lane < n ? list_at(base + lane) : 0u;
	unsigned next_raw = fetch();
	if (n) issue(__shfl_sync(0xffffffffu, my_id, 0));
	if (n > 1) prefetch_group(my_id, n);
	while (n > 0) {
	for (int i = 0; i < n; i++) {
		const unsigned id = __shfl_sync(0xffffffffu, my_id, i);
		const bool is_free = base + i >= nmixed;  // warp-uniform
		const F sv = sv_n;
		I wv = wv_n;
		int x, y, zl;
		brick_coords(id, x, y, zl);
		const bool ok = (y < g.Dy) && (zl < g.nz);
		const size_t v0 = brick_voxel(x, y, zl, ok);
		// prefetch the next brick
		if (i + 1 < n) issue(__shfl_sync(0xffffffffu, my_id, i + 1));
		if (f.debug & 1) continue;  // ablation: fetches and loads only
		F sn = sv;
		float *sp = reinterpret_cast<float *>(&sn);
		int *w = reinterpret_cast<int *>(&wv);
		if (is_free) {
			// FREE brick: every voxel gets diff == 1.0f (> near_gate, so no colour / histogram update)
			if (ok) {
				bool steady = true;
#pragma unroll
				for (int k = 0; k < VEC; k++) steady &= (sp[k] == 1.0f) & ((unsigned)w[k] < (1u << 24));
#pragma unroll
				for (int k = 0; k < VEC; k++) {
					if (!steady) sp[k] = sdf_update(sp[k], w[k], 1.0f);
					w[k] += 1;
				}
				nU += VEC;
				*reinterpret_cast<I *>(p.wt + v0) = wv;
				if (!steady && !same_bits(sv, sn)) *reinterpret_cast<F *>(p.sdf + v0) = sn;
			}
			continue;
		}
		if (f.debug & 4) continue;  // ablation: FREE bricks only
		// MIXED brick: per-voxel evaluation (tsdf.cu:30-68), phased so that the independent loads of
		// the lane's VEC voxels are in flight together instead of one dependent miss after another.
		// All 32 lanes stay converged through this block (the surface queue below uses warp ballots).
		const float px = __fmaf_rn((float)x, g.vx, g.sx);
		const float py = __fmaf_rn((float)y, g.vy, g.sy);
		const float h0 = affine_hoist(px, py, f.E[0], f.E[1]);
		const float h1 = affine_hoist(px, py, f.E[4], f.E[5]);
		const float h2 = affine_hoist(px, py, f.E[8], f.E[9]);
		// The common case of every step is straight-line code; the rare exact re-evaluations are
		// collected in bit masks and handled after the loop by out-of-line helpers, so the hot path
		// carries no per-voxel branches.
		float czv[VEC], diffv[VEC], sxv[VEC], syv[VEC], szv[VEC];
		int img[VEC];
#if SFM_K1_TMA_DEPTH
		int ixv[VEC], iyv[VEC];
#endif
		unsigned inb = 0, inexact = 0;
		// phase 1: projection -> pixel (no memory).  tsdf.cu:30-46
#pragma unroll
		for (int k = 0; k < VEC; k++) {
			const float pz = __fmaf_rn((float)(g.z0 + zl + k), g.vz, g.sz);
			const float cx = affine_finish(h0, pz, f.E[2], f.E[3]);
			const float cy = affine_finish(h1, pz, f.E[6], f.E[7]);
			czv[k] = affine_finish(h2, pz, f.E[10], f.E[11]);
			float sx, sy, sz;
			cam_to_screen<KCANON>(f, cx, cy, czv[k], sx, sy, sz);
			sxv[k] = sx; syv[k] = sy; szv[k] = sz;
			// ix = floor(RN(sx/sz)) without the IEEE divide: see pixel_floor_is_safe()
			float r;
			asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(sz));  // .ftz: one MUFU.RCP, no denormal rescaling
			const float qx = __fmul_rn(sx, r), qy = __fmul_rn(sy, r);
			if (!pixel_floor_is_safe(qx, qy)) inexact |= 1u << k;
			const int ix = __float2int_rd(qx), iy = __float2int_rd(qy);
			if ((unsigned)ix < (unsigned)f.W && (unsigned)iy < (unsigned)f.H) inb |= 1u << k;
			img[k] = iy * f.W + ix;
#if SFM_K1_TMA_DEPTH
			ixv[k] = ix; iyv[k] = iy;
#endif
		}
		if (inexact) {  // ~1 voxel in 10^4: decide with the exact IEEE divides (out-of-line helper)
#pragma unroll
			for (int k = 0; k < VEC; k++)
				if ((inexact >> k) & 1u) {
					const int ix = __float2int_rd(exact_div(sxv[k], szv[k])), iy = __float2int_rd(exact_div(syv[k], szv[k]));
					const bool in = (unsigned)ix < (unsigned)f.W && (unsigned)iy < (unsigned)f.H;
					inb = (inb & ~(1u << k)) | ((in ? 1u : 0u) << k);
					img[k] = iy * f.W + ix;
#if SFM_K1_TMA_DEPTH
					ixv[k] = ix; iyv[k] = iy;
#endif
				}
		}
		if (!ok) inb = 0;
		// phase 2: depth (metres, = depth/5000.f computed once per pixel by K0), all VEC loads at once
		float dm[VEC];
#if SFM_K1_TMA_DEPTH
		{
			int lx = 0x7fffffff, ly = 0x7fffffff;
#pragma unroll
			for (int k = 0; k < VEC; k++)
				if ((inb >> k) & 1u) { lx = min(lx, ixv[k]); ly = min(ly, iyv[k]); }
			// the box must start on a 16-byte boundary of the innermost dimension (an odd start raises 'illegal instruction')
			const int xr = __reduce_min_sync(0xffffffffu, lx), x0 = xr & ~3, y0 = __reduce_min_sync(0xffffffffu, ly);
			if (xr != 0x7fffffff) {  // warp-uniform: some voxel of the brick projects into the image
				__syncwarp();  // every lane is done with the previous tile
				if (lane == 0) {
					asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(kTmaBoxW * kTmaBoxH * 4) : "memory");
					asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
						::"r"(tile_s), "l"(reinterpret_cast<unsigned long long>(&depth_tmap)), "r"(bar_s), "r"(x0), "r"(y0) : "memory");
				}
				unsigned done = 0;
				while (!done)
					asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
						: "=r"(done) : "r"(bar_s), "r"(tma_parity) : "memory");
				tma_parity ^= 1u;
			}
#pragma unroll
			for (int k = 0; k < VEC; k++) {
				const int rx = ixv[k] - x0, ry = iyv[k] - y0;
				const bool in = (inb >> k) & 1u;
				if (in && (unsigned)rx < (unsigned)kTmaBoxW && (unsigned)ry < (unsigned)kTmaBoxH) {
					dm[k] = tile[ry * kTmaBoxW + rx];
					if (f.debug & 64) atomicAdd(err + 2, 1u);  // coverage counter of the A/B
				} else {
					dm[k] = __ldg(f.depth_m + (in ? img[k] : 0));
					if ((f.debug & 64) && in) atomicAdd(err + 3, 1u);
				}
			}
		}
#else
#pragma unroll
		for (int k = 0; k < VEC; k++) dm[k] = __ldg(f.depth_m + (((inb >> k) & 1u) ? img[k] : 0));
#endif
		// phase 3: tsdf.cu:48-52
		float nd[VEC];
		unsigned touched = 0, band = 0, surface = 0;
#pragma unroll
		for (int k = 0; k < VEC; k++) {
			diffv[k] = __fadd_rn(dm[k], -czv[k]);
			// depth == 0 <=> dm == 0;  "diff <= -miu" rejects, NaN survives as in the reference
			if (((inb >> k) & 1u) && dm[k] != 0.f && !(diffv[k] <= -g.miu)) {
				touched |= 1u << k;
				if (!(diffv[k] > g.miu)) band |= 1u << k;  // inside the truncation band: needs diff/miu
			}
			nd[k] = 1.0f;  // miu/miu == 1.0f exactly: the clamped case needs no divide
		}
		if (band) {
#pragma unroll
			for (int k = 0; k < VEC; k++)
				if ((band >> k) & 1u) {
					nd[k] = exact_div(diffv[k], g.miu);
					if (nd[k] < f.near_gate) surface |= 1u << k;  // tsdf.cu:57
				}
		}
		// near-surface voxels (tsdf.cu:57-62: colour running mean + histogram increment) are deferred to
		// a per-warp queue and processed 32 at a time by all lanes (drain_surface_queue), instead of a
		// few lanes chasing label -> histogram loads one voxel after another
		if ((f.debug & 2)) surface = 0;      // ablation: no near-surface updates
		// The queue is drained one brick LATE: the entries it holds were queued -- and their colour and
		// histogram sectors requested with prefetch.global.L2 -- while an earlier brick was evaluated, so
		// the drain's dependent loads hit L2 instead of waiting for DRAM one after another.
		if (qcount >= kQueue - 32 * VEC) {  // not enough room for another brick: drain
			nS += drain_surface_queue<LABELS>(p, g, f, q, qcount, lane, err);
			qcount = 0;
		}
		if (__any_sync(0xffffffffu, surface != 0)) {
			if (surface) {
				// surface-block map (Planes::occ): idempotent byte stores, no atomics.  A lane's VEC voxels
				// share (x, y) and lie in one 8-block along z when VEC == 4; the low-face neighbours are
				// marked too when a voxel sits on a block boundary (it is the +1 tap of the block before).
#pragma unroll
				for (int k = 0; k < VEC; k += (VEC == 4 ? 4 : 1)) {
					const int zz = zl + k;
					const int blk = ((x >> 3) * p.oby + (y >> 3)) * p.obz + (zz >> 3);
					const int fx = ((x & 7) == 0 && x > 0) ? 1 : 0, fy = ((y & 7) == 0 && y > 0) ? 1 : 0;
					const int fz = ((zz & 7) == 0 && zz > 0 && (VEC == 1 || (surface & 1u))) ? 1 : 0;
					for (int dx = 0; dx <= fx; dx++)
						for (int dy = 0; dy <= fy; dy++)
							for (int dz = 0; dz <= fz; dz++)
								p.occ[blk - dx * p.oby * p.obz - dy * p.obz - dz] = 1;
					// coarse level (32^3 blocks), same rule
					uint8_t *occ2 = p.occ + p.occ2_off;
					const int blk2 = ((x >> 5) * p.oby2 + (y >> 5)) * p.obz2 + (zz >> 5);
					const int gx = ((x & 31) == 0 && x > 0) ? 1 : 0, gy = ((y & 31) == 0 && y > 0) ? 1 : 0;
					const int gz = ((zz & 31) == 0 && fz) ? 1 : 0;
					for (int dx = 0; dx <= gx; dx++)
						for (int dy = 0; dy <= gy; dy++)
							for (int dz = 0; dz <= gz; dz++)
								occ2[blk2 - dx * p.oby2 * p.obz2 - dy * p.obz2 - dz] = 1;
				}
			}
#pragma unroll
			for (int k = 0; k < VEC; k++) {
				const bool sf = (surface >> k) & 1u;
				const unsigned m = __ballot_sync(0xffffffffu, sf);
				if (sf) {
					const int slot = qcount + __popc(m & ((1u << lane) - 1u));
					const unsigned col = (unsigned)x * (unsigned)g.Dy + (unsigned)y;
					q[slot] = make_uint4(col, (unsigned)(zl + k), (unsigned)img[k], (unsigned)w[k]);
					if (!(f.debug & 16)) {
						// colour and histogram sectors are requested now and touched by the drain one brick later: the
						// drain's loads hit L2, and its reductions find their sector resident (a RED into a sector that
						// is still in DRAM occupies the L2's atomic unit until the fill arrives and backs up into the
						// SMs: measured 0.194 ms without this prefetch, 0.131 ms with it)
						prefetch_l2_keep(p.color + (v0 + k) * 3);
						if (LABELS) {
							const unsigned label = __ldg(f.mask + img[k]);
							if ((int)label < p.bins) prefetch_l2_keep(p.hist + hist_index(col, zl + k, g.ngz, p.bins, (int)label));
						}
					}
				}
				qcount += __popc(m);
			}
		}
		if (touched) {
			// steady state of free space: sdf == 1.0f, diff == 1.0f, (1*w + 1)/(w+1) == 1.0f exactly -> only w changes
			bool steady = true;
#pragma unroll
			for (int k = 0; k < VEC; k++)
				if ((touched >> k) & 1u) steady &= (sp[k] == 1.0f) & (nd[k] == 1.0f) & ((unsigned)w[k] < (1u << 24));
#pragma unroll
			for (int k = 0; k < VEC; k++)
				if ((touched >> k) & 1u) {
					if (!steady) sp[k] = sdf_update(sp[k], w[k], nd[k]);  // tsdf.cu:56
					w[k] += 1;                                             // tsdf.cu:68
				}
			nU += __popc(touched);
			*reinterpret_cast<I *>(p.wt + v0) = wv;
			if (!steady && !same_bits(sv, sn)) *reinterpret_cast<F *>(p.sdf + v0) = sn;
		}
	}
	base = __shfl_sync(0xffffffffu, next_raw, 0);
	n = base < total ? (int)min((unsigned)kFetch, total - base) : 0;
	my_id = lane < n ? list_at(base + lane) : 0u;
	next_raw = fetch();
	if (n) issue(__shfl_sync(0xffffffffu, my_id, 0));
	if (n > 1) prefetch_group(my_id, n);
	}  // fetch loop
	if (qcount) nS += drain_surface_queue<LABELS>(p, g, f, q, qcount, lane, err);

	// fold U / S: warp shuffle, then one spread atomic pair per warp (no block barrier: warps with
	// little work must not hold their slot waiting for the busiest warp of the block)
	nU = __reduce_add_sync(0xffffffffu, nU);
	nS = __reduce_add_sync(0xffffffffu, nS);
	if (lane == 0 && (nU | nS)) {
		const int slot = (int)((blockIdx.x * (kK1Threads / 32) + warp) % kStatSlots);
		atomicAdd(stats + slot, (unsigned long long)nU);
		if (nS) atomicAdd(stats + kStatSlots + slot, (unsigned long long)nS);
	}
}

// SDF plane := miu (thrust::fill at tsdf.cu:243-244)
__global__ void fill_f32_kernel(float *__restrict__ p, size_t n, float v) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) p[i] = v;
}

}  // namespace sfm
