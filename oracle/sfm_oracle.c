/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's TSDF hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's
 * library (oracle/liboracle.so).  The product (slam_maskrcnn_b200/csrc) never links or calls it.
 *
 * Parity pin: the reference ships no tests or golden vectors.  This
 * restatement is pinned instead against outputs of the reference's OWN kernels compiled
 * verbatim (oracle/_ref, see build_ref.py) and run on a B200; those outputs are committed as
 * fixtures under tests/golden/ with the generating script (tests/golden/make_golden.py).
 * orc_filter_overlaps is pinned against the verbatim TSDF::filter_overlaps run on the CPU.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/src/SfM_CUDA).  Floating-point op order follows the SASS nvcc 12.9 emits for
 * the reference: explicit fmaf() where the compiler contracts, plain
 * IEEE ops elsewhere.  Build with -ffp-contract=off so gcc adds no contraction of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* __float2int_rd / F2I.FLOOR semantics: NaN -> 0, saturating (tsdf.cu:43-44). */
static inline int f2i_floor(float v) {
	if (v != v) return 0;
	float f = floorf(v);
	if (f >= 2147483648.0f) return 2147483647;
	if (f <= -2147483648.0f) return (-2147483647 - 1);
	return (int)f;
}

/* helper_math.h:1249-1252 dot(float4,(p,1)) as compiled: r3 + fma(pz,r2, fma(px,r0, py*r1)) */
static inline float dot4_affine(const float *r, float px, float py, float pz) {
	float t0 = py * r[1];
	float t1 = fmaf(px, r[0], t0);
	float t2 = fmaf(pz, r[2], t1);
	return r[3] + t2;
}

/* helper_math.h:1245-1248 dot(float3,float3) as compiled: fma(bz,az, fma(bx,ax, by*ay)) */
static inline float dot3(const float *a, float bx, float by, float bz) {
	float u0 = by * a[1];
	float u1 = fmaf(bx, a[0], u0);
	return fmaf(bz, a[2], u1);
}

/* tsdf.cu:18-70 tsdf_kernel, one call = one frame over voxels with z in [z0,z1).
 * bins == 0 => labels off (no histogram increment; colour gate kept).
 * Layout as tsdf.cu:55,59,61: idx = (x*Dy + y)*Dz + z; colour idx*3+c; hist idx*bins + label,
 * with 64-bit indices (the reference's int32 index overflows above ~406^3 at 32 bins).
 * Returns counts of weight increments (U) and histogram/colour updates (S). */
void orc_integrate(float *sdf, int32_t *wt, uint8_t *color, uint32_t *hist, int bins,
	const int *dims, const float *start, const float *voxel, float miu, const float *K,
	const uint16_t *depth, const uint8_t *rgb, const uint8_t *mask, const float *E,
	int width, int height, int z0, int z1, int64_t *out_U, int64_t *out_S)
{
	const int Dx = dims[0], Dy = dims[1], Dz = dims[2];
	int64_t U = 0, S = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : U, S)
	for (int x = 0; x < Dx; x++) {
		for (int y = 0; y < Dy; y++) {
			/* tsdf.cu:30  pos = start + idx*voxel  -> FFMA(I2F(idx), voxel, start) */
			const float px = fmaf((float)x, voxel[0], start[0]);
			const float py = fmaf((float)y, voxel[1], start[1]);
			for (int z = z0; z < z1; z++) {
				const float pz = fmaf((float)z, voxel[2], start[2]);
				/* tsdf.cu:31-34 */
				const float cx = dot4_affine(E + 0, px, py, pz);
				const float cy = dot4_affine(E + 4, px, py, pz);
				const float cz = dot4_affine(E + 8, px, py, pz);
				/* tsdf.cu:35-40 */
				float sx = dot3(K + 0, cx, cy, cz);
				float sy = dot3(K + 4, cx, cy, cz);
				const float sz = dot3(K + 8, cx, cy, cz);
				sx = sx / sz;
				sy = sy / sz;
				/* tsdf.cu:43-48 */
				const int ix = f2i_floor(sx), iy = f2i_floor(sy);
				if (ix < 0 || ix >= width || iy < 0 || iy >= height) continue;
				const int img = iy * width + ix;
				if (depth[img] == 0) continue;
				/* tsdf.cu:49-52 (note: NaN diff is not rejected by "diff <= -miu") */
				float diff = (float)depth[img] / 5000.f - cz;
				if (diff <= -miu) continue;
				if (diff > miu) diff = miu;
				diff = diff / miu;
				/* tsdf.cu:55-56 */
				const int64_t v = ((int64_t)x * Dy + y) * Dz + z;
				const int w = wt[v];
				sdf[v] = fmaf(sdf[v], (float)w, diff) / (float)(w + 1);
				/* tsdf.cu:57-62 */
				if (diff < 0.99f) {
					for (int c = 0; c < 3; c++)
						color[v * 3 + c] = (uint8_t)(((int)color[v * 3 + c] * w + (int)rgb[(int64_t)img * 3 + c]) / (w + 1));
					if (bins > 0) hist[v * bins + mask[img]]++;
					S++;
				}
				/* tsdf.cu:68 */
				wt[v] = w + 1;
				U++;
			}
		}
	}
	if (out_U) *out_U = U;
	if (out_S) *out_S = S;
}

/* ---- ray-march (tsdf.cu:72-135 == viewer.cu:17-86) ------------------------------------ */

typedef struct {
	const float *sdf;
	const uint32_t *hist;
	int bins;
	int Dx, Dy, Dz;
	float start[3], end[3], voxel[3];
} orc_vol;

static inline float mixf(float a, float b, float t) { /* utils.cu:93-96 -> fma(a, 1-t, t*b) */
	return fmaf(a, 1.f - t, t * b);
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* utils.cu:99-119.  Taps are clamped to the volume (documented contract; identical whenever the
 * reference's own reads are in bounds). */
static float interp_sdf(const orc_vol *V, const float *pos, int *clamped) {
	float idx[3], fr[3];
	int fl[3], i0[3], i1[3];
	const int D[3] = {V->Dx, V->Dy, V->Dz};
	for (int a = 0; a < 3; a++) {
		idx[a] = (pos[a] - V->start[a]) / V->voxel[a];
		fl[a] = f2i_floor(idx[a]);
		fr[a] = idx[a] - (float)fl[a];
		i0[a] = clampi(fl[a], 0, D[a] - 1);
		i1[a] = clampi(fl[a] + 1, 0, D[a] - 1);
		if (clamped && (i0[a] != fl[a] || i1[a] != fl[a] + 1)) *clamped = 1;
	}
	float d[8];
	for (int i = 0; i < 2; i++)
		for (int j = 0; j < 2; j++)
			for (int k = 0; k < 2; k++) {
				int64_t v = ((int64_t)(i ? i1[0] : i0[0]) * V->Dy + (j ? i1[1] : i0[1])) * V->Dz + (k ? i1[2] : i0[2]);
				d[i * 4 + j * 2 + k] = V->sdf[v];
			}
	float low = mixf(mixf(d[0], d[4], fr[0]), mixf(d[2], d[6], fr[0]), fr[1]);
	float high = mixf(mixf(d[1], d[5], fr[0]), mixf(d[3], d[7], fr[0]), fr[1]);
	return mixf(low, high, fr[2]);
}

/* utils.cu:144-170 */
static void interp_hist(const orc_vol *V, const float *pos, float *out, int *clamped) {
	float idx[3], fr[3];
	int fl[3], i0[3], i1[3];
	const int D[3] = {V->Dx, V->Dy, V->Dz};
	for (int a = 0; a < 3; a++) {
		idx[a] = (pos[a] - V->start[a]) / V->voxel[a];
		fl[a] = f2i_floor(idx[a]);
		fr[a] = idx[a] - (float)fl[a];
		i0[a] = clampi(fl[a], 0, D[a] - 1);
		i1[a] = clampi(fl[a] + 1, 0, D[a] - 1);
		if (clamped && (i0[a] != fl[a] || i1[a] != fl[a] + 1)) *clamped = 1;
	}
	int64_t v[8];
	for (int i = 0; i < 2; i++)
		for (int j = 0; j < 2; j++)
			for (int k = 0; k < 2; k++)
				v[i * 4 + j * 2 + k] = ((int64_t)(i ? i1[0] : i0[0]) * V->Dy + (j ? i1[1] : i0[1])) * V->Dz + (k ? i1[2] : i0[2]);
	for (int b = 0; b < V->bins; b++) {
		float d[8];
		for (int c = 0; c < 8; c++) d[c] = (float)V->hist[v[c] * V->bins + b];
		float low = mixf(mixf(d[0], d[4], fr[0]), mixf(d[2], d[6], fr[0]), fr[1]);
		float high = mixf(mixf(d[1], d[5], fr[0]), mixf(d[3], d[7], fr[0]), fr[1]);
		out[b] = mixf(low, high, fr[2]);
	}
}

/* The shared marcher, tsdf.cu:90-124 / viewer.cu:33-67.  dir must already be normalised by the
 * caller.  Returns 1 on a hit and writes the refined t. */
static int march(const orc_vol *V, const float *o, const float *d, float *t_hit, int *clamped) {
	float inv[3], tb[3], tt[3];
	for (int a = 0; a < 3; a++) {
		inv[a] = 1.f / d[a];
		tb[a] = inv[a] * (V->start[a] - o[a]);
		tt[a] = inv[a] * (V->end[a] - o[a]);
	}
	float tnear = fmaxf(fmaxf(fminf(tt[0], tb[0]), fminf(tt[1], tb[1])), fminf(tt[2], tb[2]));
	tnear = fmaxf(tnear, 0.01f);
	float tfar = fminf(fminf(fmaxf(tt[0], tb[0]), fmaxf(tt[1], tb[1])), fmaxf(tt[2], tb[2]));
	tfar = fminf(tfar, 100.f);
	if (tnear > tfar) return 0;
	float t = tnear + 1e-6f;
	tfar -= 1e-6f;
	float f_tt = 0.f;
	float step = V->voxel[0];
	float p[3];
	for (int a = 0; a < 3; a++) p[a] = fmaf(d[a], t, o[a]);
	float f_t = interp_sdf(V, p, clamped);
	if (!(f_t > 0.f)) return 0;
	for (; t < tfar; t += step) {
		for (int a = 0; a < 3; a++) p[a] = fmaf(d[a], t, o[a]);
		f_tt = interp_sdf(V, p, clamped);
		if (f_tt < 0.f) break;
		if (f_tt < V->voxel[0] * 0.5f) step = V->voxel[0] * 0.25f;
		f_t = f_tt;
	}
	if (!(f_tt < 0.f)) return 0;
	t = (step * f_tt) / (f_t - f_tt) + t;
	*t_hit = t;
	return 1;
}

static void fill_vol(orc_vol *V, const float *sdf, const uint32_t *hist, int bins, const int *dims,
	const float *start, const float *end, const float *voxel) {
	V->sdf = sdf; V->hist = hist; V->bins = bins;
	V->Dx = dims[0]; V->Dy = dims[1]; V->Dz = dims[2];
	for (int a = 0; a < 3; a++) { V->start[a] = start[a]; V->end[a] = end[a]; V->voxel[a] = voxel[a]; }
}

/* normalize(): helper_math.h:1306-1310 is v * rsqrtf(dot(v,v)); the GPU's MUFU.RSQ is an
 * approximation, so this CPU restatement (1/sqrtf) is tolerance-level only for ray directions. */
static inline void normalize3(float *v) {
	float s = v[1] * v[1];
	s = fmaf(v[0], v[0], s);
	s = fmaf(v[2], v[2], s);
	float inv = 1.0f / sqrtf(s);
	v[0] *= inv; v[1] *= inv; v[2] *= inv;
}

/* tsdf.cu:72-135 back_proj_kernel.  probs[h*w*bins], box_mask[h*w*bins] are zero-filled here
 * (tsdf.cu:428-429).  t_out (optional) gets the refined hit t or 0. */
void orc_backproject(const float *sdf, const uint32_t *hist, int bins, const int *dims,
	const float *start, const float *end, const float *voxel,
	const float *Kinv, const float *Rt, const float *o, int width, int height,
	float *probs, uint8_t *box_mask, float *t_out, uint8_t *flags_out)
{
	orc_vol V;
	fill_vol(&V, sdf, hist, bins, dims, start, end, voxel);
	memset(probs, 0, sizeof(float) * (size_t)width * height * bins);
	memset(box_mask, 0, (size_t)width * height * bins);
#pragma omp parallel for schedule(dynamic, 4)
	for (int y = 0; y < height; y++)
		for (int x = 0; x < width; x++) {
			const size_t pix = (size_t)y * width + x;
			const float fx = (float)x, fy = (float)y;
			float tgt[3], d[3];
			for (int r = 0; r < 3; r++) /* tsdf.cu:81-85: K[2] + fma(x,K0, y*K1) */
				tgt[r] = Kinv[r * 4 + 2] + fmaf(fx, Kinv[r * 4 + 0], fy * Kinv[r * 4 + 1]);
			for (int r = 0; r < 3; r++) d[r] = dot3(Rt + r * 3, tgt[0], tgt[1], tgt[2]); /* tsdf.cu:87-89 */
			normalize3(d);
			float t = 0.f;
			int clamped = 0;
			int hit = march(&V, o, d, &t, &clamped);
			if (t_out) t_out[pix] = hit ? t : 0.f;
			if (hit) {
				float p[3];
				for (int a = 0; a < 3; a++) p[a] = fmaf(d[a], t, o[a]);
				interp_hist(&V, p, probs + pix * bins, &clamped);
				for (int b = 0; b < bins; b++)
					if (probs[pix * bins + b] > 0.3f) box_mask[pix * bins + b] = 1;
			}
			if (flags_out) flags_out[pix] = (uint8_t)clamped;
		}
}

/* viewer.cu:17-86 show_tsdf_kernel.  out_bgr[h*w*3] zero-filled here (viewer.cu:150). */
void orc_raycast(const float *sdf, const uint32_t *hist, int bins, const int *dims,
	const float *start, const float *end, const float *voxel,
	const float *s2w, const float *c, int width, int height, const uint8_t *palette,
	uint8_t *out_bgr, float *t_out, uint8_t *label_out)
{
	orc_vol V;
	fill_vol(&V, sdf, hist, bins, dims, start, end, voxel);
	memset(out_bgr, 0, (size_t)width * height * 3);
#pragma omp parallel for schedule(dynamic, 4)
	for (int y = 0; y < height; y++) {
		float *cn = (float *)malloc(sizeof(float) * (size_t)bins);
		for (int x = 0; x < width; x++) {
			const size_t pix = (size_t)y * width + x;
			const float fx = (float)x, fy = (float)y;
			float d[3];
			for (int r = 0; r < 3; r++) { /* viewer.cu:26-32: s3 + (s2 + fma(x,s0, y*s1)) - c */
				float v = fmaf(fx, s2w[r * 4 + 0], fy * s2w[r * 4 + 1]);
				v = s2w[r * 4 + 2] + v;
				v = s2w[r * 4 + 3] + v;
				d[r] = v - c[r];
			}
			normalize3(d);
			float t = 0.f;
			int hit = march(&V, c, d, &t, NULL);
			uint8_t lab = 0;
			if (hit) {
				float p[3];
				for (int a = 0; a < 3; a++) p[a] = fmaf(d[a], t, c[a]);
				interp_hist(&V, p, cn, NULL);
				float mx = 0.f; /* viewer.cu:71-79 strict >, ascending k */
				for (int k = 0; k < bins; k++)
					if (cn[k] > mx) { mx = cn[k]; lab = (uint8_t)k; }
				if (lab > 0) { /* viewer.cu:80-83 */
					out_bgr[pix * 3 + 0] = palette[lab * 3 + 2];
					out_bgr[pix * 3 + 1] = palette[lab * 3 + 1];
					out_bgr[pix * 3 + 2] = palette[lab * 3 + 0];
				}
			}
			if (t_out) t_out[pix] = hit ? t : 0.f;
			if (label_out) label_out[pix] = lab;
		}
		free(cn);
	}
}

/* tsdf.cu:304-416 TSDF::filter_overlaps (single thread, float32 sequential sums in raster
 * order exactly as written).  mask is relabelled in place, *num_objs grows.
 * assign_out[bins]: for each current-frame label m, the global id it was mapped to (0 = none).
 * A_out / C_out (optional, bins*bins): the accumulated tables. */
void orc_filter_overlaps(const float *probs, int width, int height, uint8_t *mask,
	const uint8_t *box_mask, int bins, uint32_t n_obs, float prior, float accept_factor,
	int *num_objs, float *A_out, uint32_t *C_out, int *assign_out)
{
	const int n = width * height;
	int mx = 0;
	for (int i = 0; i < n; i++) if (mask[i] > mx) mx = mask[i];
	const int max_obj_now = mx + 1; /* tsdf.cu:305-307 */
	float *A = (float *)calloc((size_t)bins * bins, sizeof(float));
	uint32_t *C = (uint32_t *)calloc((size_t)bins * bins, sizeof(uint32_t));
	for (int i = 0; i < n; i++) { /* tsdf.cu:312-334 */
		const int m0 = mask[i];
		if (m0 > 0)
			for (int j = 1; j < bins; j++) {
				A[m0 * bins + j] += logf(fmaxf(probs[(size_t)i * bins + j] / (float)n_obs, prior));
				C[m0 * bins + j]++;
			}
		for (int b = 1; b < bins; b++)
			if (box_mask[(size_t)i * bins + b])
				for (int m = 1; m < max_obj_now; m++) {
					if (m0 == m) continue;
					A[m * bins + b] += logf(fmaxf(1.f - probs[(size_t)i * bins + b] / (float)n_obs, prior));
					C[m * bins + b]++;
				}
	}
	/* tsdf.cu:335-365: argmax with strict >, first wins; accept if > accept_factor*prior;
	 * collision on the same global id keeps the larger probability */
	int *owner = (int *)calloc(256, sizeof(int));       /* global j -> current m (0 = free) */
	float *owner_p = (float *)calloc(256, sizeof(float));
	int *rev = (int *)calloc(256, sizeof(int));          /* current m -> global j (0 = none) */
	for (int m = 1; m < max_obj_now; m++) {
		int best = -1;
		float bp = 0.f;
		for (int j = 1; j < bins; j++) {
			float p = (C[m * bins + j] == 0) ? 0.f : expf(A[m * bins + j] / (float)C[m * bins + j]);
			if (p > bp) { best = j; bp = p; }
		}
		if (bp > accept_factor * prior) {
			if (owner[best] == 0 || owner_p[best] < bp) { owner[best] = m; owner_p[best] = bp; }
		}
	}
	for (int j = 1; j < 256; j++) if (owner[j]) rev[owner[j]] = j; /* tsdf.cu:366-369 */
	/* tsdf.cu:371-389: relabel; unassigned labels get fresh ids in raster first-appearance order */
	int *extra = (int *)calloc(256, sizeof(int));
	for (int i = 0; i < n; i++) {
		const int m0 = mask[i];
		if (rev[m0]) mask[i] = (uint8_t)rev[m0];
		else if (m0 > 0) {
			if (!extra[m0]) { extra[m0] = *num_objs; (*num_objs)++; }
			mask[i] = (uint8_t)extra[m0];
		}
	}
	if (assign_out) for (int m = 0; m < bins; m++) assign_out[m] = (m < 256) ? (rev[m] ? rev[m] : extra[m]) : 0;
	if (A_out) memcpy(A_out, A, sizeof(float) * (size_t)bins * bins);
	if (C_out) memcpy(C_out, C, sizeof(uint32_t) * (size_t)bins * bins);
	free(A); free(C); free(owner); free(owner_p); free(rev); free(extra);
}

/* utils.cu:77-91 mean_depth */
float orc_mean_depth(const uint16_t *depth, int n) {
	double sum = 0;
	int total = 0;
	for (int i = 0; i < n; i++) {
		if (depth[i] == 0) continue;
		sum += depth[i] / 5000.;
		total++;
	}
	return (float)(sum / total);
}
