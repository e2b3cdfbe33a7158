// k_integrate.cuh -- K0 (frame prep) and K1 (TSDF integrate) for sm_100a.
//
// K1 replaces tsdf_kernel (reference src/SfM_CUDA/tsdf.cu:18-70).  Design, B200-first:
//   * one lane owns VEC consecutive z voxels of one (x,y) column; 32/VEC lanes cover a 32-voxel
//     (128 B) run of the column, so every warp request on the SDF / weight planes is made of
//     full 128 B lines moved with 128-bit loads/stores (the reference strides lanes along x,
//     the slowest axis -- 32 lines per request);
//   * the warp marches z in 32-voxel chunks; the z-invariant part of the pose transform
//     (fma(px,r0,py*r1), bit-identical to the reference's evaluation order) is
//     computed once per column and reused;
//   * before touching a chunk ("brick" = CPW columns x 32 z), the warp runs an exact-conservative
//     cull: the brick's 4 corners are projected, and the brick is skipped when every voxel in it
//     provably fails the reference's own tests (outside the image, all-invalid depth, or behind
//     the surface by more than miu against the per-tile max depth).  Skipped voxels would have
//     early-outed in the reference, so results are identical; SFM_FLAG_NO_CULL disables it;
//   * per-frame U (weight increments) and S (histogram/colour updates) are folded with warp
//     shuffles + one spread atomic per block -- they define the algorithmic bytes of the step.
#pragma once
#include "sfm_device.cuh"

namespace sfm {

// ---------------------------------------------------------------------------------------------
// K0: per-frame prep.  One warp per kTile x kTile tile: max depth of the tile (for culling) and
// max label (labels >= bins are a contract violation).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_frame_kernel(const uint16_t *__restrict__ depth,
	const uint8_t *__restrict__ mask, int W, int H, int TW, int TH, int bins,
	uint16_t *__restrict__ tilemax, unsigned long long *__restrict__ stats, uint32_t *__restrict__ err)
{
	const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
	const int warp = gtid >> 5, lane = threadIdx.x & 31;
	if (warp >= TW * TH) return;
	const int ty = warp / TW, tx = warp % TW;
	unsigned dmax = 0, lmax = 0;
#pragma unroll
	for (int k = 0; k < 2; k++) {
		const int r = (lane >> 2), c = ((lane & 3) << 1) + k;
		const int y = ty * kTile + r, x = tx * kTile + c;
		if (x < W && y < H) {
			dmax = max(dmax, (unsigned)depth[y * W + x]);
			if (mask) lmax = max(lmax, (unsigned)mask[y * W + x]);
		}
	}
	dmax = __reduce_max_sync(0xffffffffu, dmax);
	lmax = __reduce_max_sync(0xffffffffu, lmax);
	if (lane == 0) {
		tilemax[ty * TW + tx] = (uint16_t)dmax;
		if (bins > 0 && (int)lmax >= bins) atomicOr(err, 1u);
	}
}

// ---------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------
struct VoxelEval {
	int img;     // pixel index, -1 => voxel rejected before the update
	float diff;  // normalised, clamped SDF sample
};

// tsdf.cu:30-52 for one voxel whose z-invariant parts are hoisted.
__device__ __forceinline__ VoxelEval eval_voxel(const FrameView &f, const VolGeom &g, float h0, float h1,
	float h2, int zglobal)
{
	VoxelEval r;
	r.img = -1;
	r.diff = 0.f;
	const float pz = __fmaf_rn((float)zglobal, g.vz, g.sz);
	const float cx = affine_finish(h0, pz, f.E[2], f.E[3]);
	const float cy = affine_finish(h1, pz, f.E[6], f.E[7]);
	const float cz = affine_finish(h2, pz, f.E[10], f.E[11]);
	float sx = dot3_ref(f.K[0], f.K[1], f.K[2], cx, cy, cz);
	float sy = dot3_ref(f.K[3], f.K[4], f.K[5], cx, cy, cz);
	const float sz = dot3_ref(f.K[6], f.K[7], f.K[8], cx, cy, cz);
	sx = __fdiv_rn(sx, sz);
	sy = __fdiv_rn(sy, sz);
	const int ix = __float2int_rd(sx), iy = __float2int_rd(sy);
	if (ix < 0 || ix >= f.W || iy < 0 || iy >= f.H) return r;
	const int img = iy * f.W + ix;
	const unsigned d = __ldg(f.depth + img);
	if (d == 0) return r;
	float diff = __fadd_rn(__fdiv_rn((float)d, f.depth_scale), -cz);
	if (diff <= -g.miu) return r;  // NaN survives, as in the reference
	if (diff > g.miu) diff = g.miu;
	r.diff = __fdiv_rn(diff, g.miu);
	r.img = img;
	return r;
}

// colour running mean + histogram increment (tsdf.cu:57-62) for one near-surface voxel
template <bool LABELS>
__device__ __forceinline__ void update_surface_voxel(const Planes &p, const FrameView &f, size_t v, int w,
	int img, uint32_t *err)
{
	const uint8_t *src = f.rgb + (size_t)img * 3;
	uint8_t *dst = p.color + v * 3;
#pragma unroll
	for (int c = 0; c < 3; c++) dst[c] = (uint8_t)(((int)dst[c] * w + (int)__ldg(src + c)) / (w + 1));
	if (LABELS) {
		const unsigned label = __ldg(f.mask + img);
		if ((int)label < p.bins) p.hist[v * (size_t)p.bins + label] += 1u;
		else atomicOr(err, 1u);
	}
}

template <int VEC> struct VecT;
template <> struct VecT<4> { using F = float4; using I = int4; };
template <> struct VecT<2> { using F = float2; using I = int2; };
template <> struct VecT<1> { using F = float; using I = int; };

template <int VEC, bool LABELS, bool CULL>
__global__ void __launch_bounds__(256) integrate_kernel(Planes p, VolGeom g, FrameView f,
	unsigned long long *__restrict__ stats, uint32_t *__restrict__ err)
{
	constexpr int LPC = 32 / VEC;  // lanes per column
	constexpr int CPW = 32 / LPC;  // columns per warp
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int zq = lane % LPC, ci = lane / LPC;
	const int groups_per_x = (g.Dy + CPW - 1) / CPW;
	const long long group = (long long)blockIdx.x * 8 + warp;
	unsigned nU = 0, nS = 0;
	if (group < (long long)g.Dx * groups_per_x) {
		const int x = (int)(group / groups_per_x);
		const int y0 = (int)(group % groups_per_x) * CPW;
		const int y = y0 + ci;
		const bool col_ok = y < g.Dy;
		const float px = __fmaf_rn((float)x, g.vx, g.sx);
		const float py = __fmaf_rn((float)y, g.vy, g.sy);
		const float h0 = affine_hoist(px, py, f.E[0], f.E[1]);
		const float h1 = affine_hoist(px, py, f.E[4], f.E[5]);
		const float h2 = affine_hoist(px, py, f.E[8], f.E[9]);
		const size_t colbase = ((size_t)x * g.Dy + (col_ok ? y : 0)) * (size_t)g.nz;

		// corner assignment for the cull test: lane&1 -> y end, lane&2 -> z end
		const int ylast = min(y0 + CPW - 1, g.Dy - 1);
		const float cpy = __fmaf_rn((float)((lane & 1) ? ylast : y0), g.vy, g.sy);
		const float c0 = affine_hoist(px, cpy, f.E[0], f.E[1]);
		const float c1 = affine_hoist(px, cpy, f.E[4], f.E[5]);
		const float c2 = affine_hoist(px, cpy, f.E[8], f.E[9]);

		for (int zc = 0; zc < g.nz; zc += 32) {
			if (CULL) {
				const int zl_end = min(zc + 31, g.nz - 1);
				const float cpz = __fmaf_rn((float)(g.z0 + ((lane & 2) ? zl_end : zc)), g.vz, g.sz);
				const float ccx = affine_finish(c0, cpz, f.E[2], f.E[3]);
				const float ccy = affine_finish(c1, cpz, f.E[6], f.E[7]);
				const float ccz = affine_finish(c2, cpz, f.E[10], f.E[11]);
				const float ssx = dot3_ref(f.K[0], f.K[1], f.K[2], ccx, ccy, ccz);
				const float ssy = dot3_ref(f.K[3], f.K[4], f.K[5], ccx, ccy, ccz);
				const float ssz = dot3_ref(f.K[6], f.K[7], f.K[8], ccx, ccy, ccz);
				const float u = ssx / ssz, vv = ssy / ssz;
				// reduce over the 4 corners (lanes differing in bits 0,1); all groups of 4 are identical
				float umin = u, umax = u, vmin = vv, vmax = vv, szmin = ssz, szmax = ssz, czmin = ccz;
				// magnitude bound of the camera-space coordinates of this brick (rounding-error budget)
				float scale_c = f.cull_lin * (fabsf(px) + fabsf(cpy) + fabsf(cpz) + 1.f) + f.cull_t;
#pragma unroll
				for (int o = 1; o <= 2; o <<= 1) {
					umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o));
					umax = fmaxf(umax, __shfl_xor_sync(0xffffffffu, umax, o));
					vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
					vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
					szmin = fminf(szmin, __shfl_xor_sync(0xffffffffu, szmin, o));
					szmax = fmaxf(szmax, __shfl_xor_sync(0xffffffffu, szmax, o));
					czmin = fminf(czmin, __shfl_xor_sync(0xffffffffu, czmin, o));
					scale_c = fmaxf(scale_c, __shfl_xor_sync(0xffffffffu, scale_c, o));
				}
				// The projective map is monotone along any segment that stays on one side of the
				// camera plane, so with all four corners strictly on one side the pixel coordinates of
				// every voxel of the brick lie inside the corner bounding box (plus rounding slack).
				// Rounding-error budget: |c*| <= scale_c, per-voxel error of c*, s* ~ 1e-6 * scale; a brick
				// is only culled when all corners are at least 1e-2*scale_sz away from the camera plane,
				// which bounds the per-voxel pixel error by ~1e-4*(Krow/K2row + |u|) -- slack is 10x that.
				const float zguard = 1e-2f * f.cull_k2 * scale_c;
				const bool one_side = (szmin > zguard) || (szmax < -zguard);
				const bool finite = (umin == umin) && (umax == umax) && (vmin == vmin) && (vmax == vmax) &&
					fabsf(umin) < 1e8f && fabsf(umax) < 1e8f && fabsf(vmin) < 1e8f && fabsf(vmax) < 1e8f;
				if (one_side && finite) {
					const float slack_u = f.cull_slack0 + 1e-3f * fmaxf(fabsf(umin), fabsf(umax));
					const float slack_v = f.cull_slack0 + 1e-3f * fmaxf(fabsf(vmin), fabsf(vmax));
					const float ulo = umin - slack_u, uhi = umax + slack_u;
					const float vlo = vmin - slack_v, vhi = vmax + slack_v;
					if (uhi < 0.f || ulo >= (float)f.W || vhi < 0.f || vlo >= (float)f.H) continue;  // outside the image
					// pixel bbox -> tile range
					const int tx0 = max(0, (int)floorf(ulo)) / kTile, tx1 = min(f.W - 1, (int)floorf(uhi)) / kTile;
					const int ty0 = max(0, (int)floorf(vlo)) / kTile, ty1 = min(f.H - 1, (int)floorf(vhi)) / kTile;
					const int nx = tx1 - tx0 + 1, ny = ty1 - ty0 + 1;
					const int cnt = nx * ny;
					if (cnt <= 512) {
						unsigned dmax = 0;
						for (int i = lane; i < cnt; i += 32)
							dmax = max(dmax, (unsigned)__ldg(f.tilemax + (ty0 + i / nx) * f.TW + tx0 + i % nx));
						dmax = __reduce_max_sync(0xffffffffu, dmax);
						if (dmax == 0) continue;  // only invalid depth under the brick
						if (szmin > 0.f) {
							// every voxel: diff = d/scale - cz <= dmax/scale - czmin + eps
							const float dmax_m = __fdiv_rn((float)dmax, f.depth_scale);
							const float eps = 1e-4f * (scale_c + dmax_m);
							if (czmin - dmax_m >= g.miu + eps) continue;  // wholly behind the surface band
						}
					}
				}
			}
			const int zl = zc + zq * VEC;
			if (!col_ok || zl >= g.nz) continue;
			VoxelEval ev[VEC];
			bool any = false;
#pragma unroll
			for (int k = 0; k < VEC; k++) {
				ev[k] = eval_voxel(f, g, h0, h1, h2, g.z0 + zl + k);
				any |= ev[k].img >= 0;
			}
			if (!any) continue;
			const size_t v0 = colbase + zl;
			typename VecT<VEC>::F sv = *reinterpret_cast<const typename VecT<VEC>::F *>(p.sdf + v0);
			typename VecT<VEC>::I wv = *reinterpret_cast<const typename VecT<VEC>::I *>(p.wt + v0);
			float *s = reinterpret_cast<float *>(&sv);
			int *w = reinterpret_cast<int *>(&wv);
#pragma unroll
			for (int k = 0; k < VEC; k++) {
				if (ev[k].img < 0) continue;
				const int wk = w[k];
				// tsdf.cu:56  (sdf*w + diff) / (w + 1)   -> FFMA, IEEE divide
				s[k] = __fdiv_rn(__fmaf_rn(s[k], (float)wk, ev[k].diff), (float)(wk + 1));
				if (ev[k].diff < f.near_gate) {  // tsdf.cu:57-62
					update_surface_voxel<LABELS>(p, f, v0 + k, wk, ev[k].img, err);
					nS++;
				}
				w[k] = wk + 1;  // tsdf.cu:68
				nU++;
			}
			*reinterpret_cast<typename VecT<VEC>::F *>(p.sdf + v0) = sv;
			*reinterpret_cast<typename VecT<VEC>::I *>(p.wt + v0) = wv;
		}
	}
	// fold U / S: warp shuffle -> shared -> one spread atomic pair per block
	nU = __reduce_add_sync(0xffffffffu, nU);
	nS = __reduce_add_sync(0xffffffffu, nS);
	__shared__ unsigned sU[8], sS[8];
	if (lane == 0) { sU[warp] = nU; sS[warp] = nS; }
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned tu = 0, ts = 0;
#pragma unroll
		for (int i = 0; i < 8; i++) { tu += sU[i]; ts += sS[i]; }
		if (tu | ts) {
			const int slot = blockIdx.x % kStatSlots;
			atomicAdd(stats + slot, (unsigned long long)tu);
			atomicAdd(stats + kStatSlots + slot, (unsigned long long)ts);
		}
	}
}

// SDF plane := miu (thrust::fill at tsdf.cu:243-244)
__global__ void fill_f32_kernel(float *__restrict__ p, size_t n, float v) {
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) p[i] = v;
}

}  // namespace sfm
