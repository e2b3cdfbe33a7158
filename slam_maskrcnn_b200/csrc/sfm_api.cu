// sfm_api.cu -- host side of libsfm_b200.so: the C-ABI declared in include/sfm_b200.h.
//
// Mirrors the per-frame sequencing of the reference's TSDF::parse_frame / launch_kernel
// (src/SfM_CUDA/tsdf.cu:171-228, 418-504) and Viewer::show_tsdf (viewer.cu:137-179), with
// asynchronous copies on one stream per handle, every CUDA call checked, and no CPU fallback:
// without an sm_100 device sfm_create fails with SFM_ERR_NODEVICE.
#include "../../include/sfm_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "k_integrate.cuh"
#include "k_raymarch.cuh"

using namespace sfm;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
	g_err = msg;
	return code;
}

#define CU(call)                                                                                     \
	do {                                                                                             \
		cudaError_t e_ = (call);                                                                     \
		if (e_ != cudaSuccess)                                                                       \
			return fail(e_ == cudaErrorMemoryAllocation ? SFM_ERR_NOMEM : SFM_ERR_CUDA,              \
				std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
	} while (0)

// launch check in the reference's words (tsdf.cu:497-503)
#define LAUNCH_CHECK(v)                                                                              \
	do {                                                                                             \
		(v)->launches++;                                                                             \
		cudaError_t e_ = cudaGetLastError();                                                         \
		if (e_ != cudaSuccess)                                                                       \
			return fail(SFM_ERR_CUDA, std::string("run_kernel launch failed\n") + cudaGetErrorString(e_)); \
	} while (0)

void mat4_mul(const float *a, const float *b, float *out) {  // float inputs, double accumulate
	float r[16];
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++) {
			double s = 0;
			for (int k = 0; k < 4; k++) s += (double)a[i * 4 + k] * (double)b[k * 4 + j];
			r[i * 4 + j] = (float)s;
		}
	memcpy(out, r, sizeof(r));
}

bool mat4_inv(const float *m, float *out) {  // Gauss-Jordan in double, partial pivoting
	double a[4][8];
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++) { a[i][j] = m[i * 4 + j]; a[i][j + 4] = (i == j) ? 1.0 : 0.0; }
	for (int c = 0; c < 4; c++) {
		int piv = c;
		for (int r = c + 1; r < 4; r++) if (fabs(a[r][c]) > fabs(a[piv][c])) piv = r;
		if (fabs(a[piv][c]) < 1e-300) return false;
		if (piv != c) for (int j = 0; j < 8; j++) std::swap(a[piv][j], a[c][j]);
		const double d = a[c][c];
		for (int j = 0; j < 8; j++) a[c][j] /= d;
		for (int r = 0; r < 4; r++) {
			if (r == c) continue;
			const double f = a[r][c];
			if (f != 0.0) for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j];
		}
	}
	for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[i * 4 + j] = (float)a[i][j + 4];
	return true;
}

const uint8_t kPalette16[16 * 3] = {  // Viewer palette values, viewer.cu:93-109 (RGB triplets)
	230, 25, 75, 60, 180, 75, 255, 225, 25, 0, 130, 200, 245, 130, 48, 145, 30, 180, 70, 240, 240, 240, 50, 230,
	210, 245, 60, 250, 190, 190, 0, 128, 128, 230, 190, 255, 170, 110, 40, 255, 250, 200, 128, 0, 0, 170, 255, 195};

struct PinnedFrame {
	uint8_t *buf = nullptr;
	cudaEvent_t free_ev = nullptr;  // recorded after the H2D that read this buffer
};

}  // namespace

struct sfm_volume {
	sfm_desc desc;
	VolGeom g;
	int bins = 0, W = 0, H = 0, TW = 0, TH = 0;
	size_t nvox = 0;
	Planes planes{};
	// frame staging (device): double buffered, filled on a dedicated copy stream so that the H2D
	// copy of frame i+1 overlaps the kernels of frame i
	uint16_t *d_depth = nullptr;                 // = current buffer
	uint8_t *d_rgb = nullptr, *d_mask = nullptr;
	uint8_t *d_frame[2] = {nullptr, nullptr};    // [depth | rgb | mask]
	cudaEvent_t ev_uploaded[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
	int frame_cur = 0;
	bool frame_open = false;
	cudaStream_t copy_stream = nullptr;
	// pipelined read-back of the U/S counters
	static constexpr int kStatRing = 4;
	unsigned long long *h_stat_ring = nullptr;   // pinned, kStatRing x 2*kStatSlots
	cudaEvent_t ev_stat[kStatRing] = {};
	uint64_t stat_tickets = 0;
	// Per-frame preparation contexts (kCtx = 3 of them, rotating): what K0 + K1a produce for K1b.  K0 / K1a of frame i+1
	// run on prep_stream while K1b of frame i still reads the other context on the main stream.
	struct PrepCtx {
		uint16_t *d_tilemax = nullptr, *d_tilemin = nullptr;  // one allocation: [tilemax | tilemin], padded to 16 B
		float *d_depth_m = nullptr;
		alignas(64) unsigned char depth_tmap[128] = {};  // SFM_K1_TMA_DEPTH build: CUtensorMap of d_depth_m
		unsigned *d_work = nullptr;                  // WorkLists::counts (3 counters, zeroed by K0)
		uint32_t *d_list_mixed = nullptr, *d_list_free = nullptr;  // K1a -> K1b brick lists, one slot per brick each
		cudaEvent_t ev_ready = nullptr;  // K1a of the frame that uses this context is done (prep_stream)
		cudaEvent_t ev_free = nullptr;   // K1b of that frame is done (main stream): the context may be rewritten
	} ctx[3];
	static constexpr int kCtx = 3;  // frame i uses ctx[i % 3]: its preparation may start while K1b of frame i-2 still runs
	cudaStream_t prep_stream = nullptr;
	cudaEvent_t ev_call = nullptr;
	size_t tile_bytes = 0;
	unsigned long long *d_stats = nullptr;
	unsigned long long *d_ray_stats = nullptr;  // [0] SDF samples gathered, [1] hits, cumulative (march_kernel)
	unsigned *d_march_work = nullptr;           // march_kernel's tile counter and finished-warp counter (self-resetting)
	unsigned *d_tile_cost = nullptr, *d_tile_order = nullptr;  // per 8x4-pixel tile: cost (time) in the last march; tiles by descending cost
	size_t tile_cap = 0;
	unsigned long long order_key = 0;           // image shape d_tile_order was built for (0: none)
	int march_per_sm = 0;                       // resident march_kernel blocks per SM (occupancy, queried once)
	uint64_t ray_samples_seen = 0, ray_hits_seen = 0;
	uint32_t *d_err = nullptr;
	size_t nbricks = 0;
	int num_sms = 148;
	size_t occ_bytes = 0;
	uint8_t *d_palette = nullptr;
	uint8_t *d_lut = nullptr;
	// pinned staging (host), double buffered
	PinnedFrame pin[2];
	int pin_next = 0;
	size_t frame_bytes = 0;
	// ray-march scratch (device), sized lazily
	float *d_probs = nullptr;
	uint8_t *d_box = nullptr;
	float *d_t = nullptr;
	uint8_t *d_flags = nullptr;
	uint8_t *d_bgr = nullptr, *d_label = nullptr;
	unsigned long long *d_keys = nullptr;
	float4 *d_hits = nullptr;
	size_t ray_px = 0;      // capacity of the per-pixel buffers
	size_t probs_px = 0;    // capacity of probs/box
	// merge scratch
	uint8_t *d_fold = nullptr;
	void *d_merge = nullptr, *h_merge = nullptr;  // MergeOut of the device-side decision (+ pinned mirror)
	cudaEvent_t ev_merge = nullptr;
	uint8_t *h_fold = nullptr;  // pinned mirror
	size_t fold_bytes = 0;
	unsigned long long *h_stats = nullptr;  // pinned
	uint32_t *h_err = nullptr;               // pinned
	sfm_merge_report last_merge{};
	// state (tsdf.cuh:46-61)
	float K[16], Kinv[16];
	float init_extr_inv[16];
	bool init = false;
	uint32_t n_obs = 0;
	int num_objs = 0;
	float mean_depth = 0.f;
	// execution
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
	static constexpr int kRing = 2048;  // per-call event pairs around K1 (integrate kernel only)
	cudaEvent_t ev_k0[kRing] = {}, ev_km[kRing] = {};  // around K1a (prep_stream)
	cudaEvent_t ev_kb[kRing] = {}, ev_k1[kRing] = {};  // around K1b (main stream)
	uint64_t n_integrate = 0;
	uint64_t launches = 0;
	uint64_t stat_U_seen = 0, stat_S_seen = 0;  // cumulative totals already reported by sfm_frame_stats
	// per-handle (= per-device) launch state: cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute
	size_t k1a_smem_set = 0;   // dynamic shared memory K1a has been configured for on this handle's device
	int k1b_per_sm[8] = {};    // resident K1b blocks per SM, per kernel instantiation (0 = not queried yet)
	// histogram: tiled 16-bit plane inside (sfm_device.cuh), reference layout at the boundary
	size_t hist_elems = 0;       // hist_t entries of the tiled plane
	uint32_t *d_hist_ref = nullptr;  // lazily allocated reference-layout mirror behind sfm_plane_device_ptr(HIST)
	uint32_t *d_hist_chunk = nullptr; // staging for sfm_download / sfm_upload (kHistChunk bytes)
	unsigned *d_hist_max = nullptr;
	uint32_t hist_bound = 0;     // upper bound of every bin: frames integrated (+ the largest uploaded value)
	int debug_ablate = 0;      // SFM_DEBUG_ABLATE, read once at creation and only under SFM_FLAG_DEBUG_ABLATE
};

namespace {

int check_device_error(sfm_volume *v) {
	CU(cudaMemcpyAsync(v->h_err, v->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	if (*v->h_err) {
		CU(cudaMemsetAsync(v->d_err, 0, sizeof(uint32_t), v->stream));
		return fail(SFM_ERR_INVALID, "a frame carried a label >= bins (histogram bin out of range); that update was skipped");
	}
	return SFM_OK;
}

size_t plane_elem_bytes(const sfm_volume *v, int plane) {
	switch (plane) {
	case SFM_PLANE_SDF: return 4;
	case SFM_PLANE_WEIGHT: return 4;
	case SFM_PLANE_COLOR: return 3;
	case SFM_PLANE_HIST: return 4 * (size_t)v->bins;
	default: return 0;
	}
}

void *plane_ptr(const sfm_volume *v, int plane) {
	switch (plane) {
	case SFM_PLANE_SDF: return v->planes.sdf;
	case SFM_PLANE_WEIGHT: return v->planes.wt;
	case SFM_PLANE_COLOR: return v->planes.color;
	default: return nullptr;  // the histogram is stored tiled: see hist_export / sfm_plane_device_ptr
	}
}

int reset_planes(sfm_volume *v) {
	// tsdf.cu:242-253: SDF := miu (thrust::fill), colour / histogram / weight := 0
	fill_f32_kernel<<<148 * 8, 256, 0, v->stream>>>(v->planes.sdf, v->nvox, v->g.miu);
	LAUNCH_CHECK(v);
	CU(cudaMemsetAsync(v->planes.wt, 0, v->nvox * 4, v->stream));
	CU(cudaMemsetAsync(v->planes.color, 0, v->nvox * 3, v->stream));
	if (v->bins > 0) CU(cudaMemsetAsync(v->planes.hist, 0, v->hist_elems * sizeof(hist_t), v->stream));
	v->hist_bound = 0;
	CU(cudaMemsetAsync(v->planes.occ, 0, v->occ_bytes, v->stream));
	v->n_obs = 0;
	v->num_objs = 0;
	return SFM_OK;
}

bool is_device_or_pinned(const void *p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Upload one frame's images into the next device frame buffer on the copy stream.  Pinned / device
// sources are copied directly; pageable sources go through a double-buffered pinned bounce buffer.
// The compute stream is made to wait for the upload; release_frame() marks the buffer reusable.
int upload_frame(sfm_volume *v, const uint16_t *depth, const uint8_t *color, const uint8_t *mask) {
	const size_t npx = (size_t)v->W * v->H;
	const size_t bd = npx * 2, bc = npx * 3, bm = npx;
	const int b = v->frame_cur ^ 1;
	v->frame_cur = b;
	uint8_t *base = v->d_frame[b];
	v->d_depth = (uint16_t *)base;
	v->d_rgb = base + bd;
	v->d_mask = base + bd + bc;
	CU(cudaStreamWaitEvent(v->copy_stream, v->ev_consumed[b], 0));  // kernels that read this buffer are done
	const bool direct = (!depth || is_device_or_pinned(depth)) && (!color || is_device_or_pinned(color)) &&
		(!mask || is_device_or_pinned(mask));
	if (direct) {
		if (depth) CU(cudaMemcpyAsync(v->d_depth, depth, bd, cudaMemcpyDefault, v->copy_stream));
		if (color) CU(cudaMemcpyAsync(v->d_rgb, color, bc, cudaMemcpyDefault, v->copy_stream));
		if (mask) CU(cudaMemcpyAsync(v->d_mask, mask, bm, cudaMemcpyDefault, v->copy_stream));
	} else {
		PinnedFrame &pf = v->pin[v->pin_next];
		v->pin_next ^= 1;
		CU(cudaEventSynchronize(pf.free_ev));
		if (depth) { memcpy(pf.buf, depth, bd); CU(cudaMemcpyAsync(v->d_depth, pf.buf, bd, cudaMemcpyHostToDevice, v->copy_stream)); }
		if (color) { memcpy(pf.buf + bd, color, bc); CU(cudaMemcpyAsync(v->d_rgb, pf.buf + bd, bc, cudaMemcpyHostToDevice, v->copy_stream)); }
		if (mask) { memcpy(pf.buf + bd + bc, mask, bm); CU(cudaMemcpyAsync(v->d_mask, pf.buf + bd + bc, bm, cudaMemcpyHostToDevice, v->copy_stream)); }
		CU(cudaEventRecord(pf.free_ev, v->copy_stream));
	}
	CU(cudaEventRecord(v->ev_uploaded[b], v->copy_stream));
	CU(cudaStreamWaitEvent(v->stream, v->ev_uploaded[b], 0));
	v->frame_open = true;
	// Lifetime of the caller's buffers (sfm_b200.h): pageable sources were copied into the bounce buffer above;
	// pinned / device sources are read by the copy engine, so the call waits for the copy unless the caller
	// took that responsibility with SFM_FLAG_ASYNC_SOURCES
	if (direct && !(v->desc.flags & SFM_FLAG_ASYNC_SOURCES)) CU(cudaEventSynchronize(v->ev_uploaded[b]));
	return SFM_OK;
}

// The kernels enqueued so far were the last readers of the current frame buffer.
int release_frame(sfm_volume *v) {
	if (v->frame_open) {
		CU(cudaEventRecord(v->ev_consumed[v->frame_cur], v->stream));
		v->frame_open = false;
	}
	return SFM_OK;
}

FrameView make_frame_view(const sfm_volume *v, const sfm_volume::PrepCtx &c, const void *d_depth, const void *d_rgb, const void *d_mask,
	const float *E16)
{
	FrameView f{};
	f.depth = (const uint16_t *)d_depth;
	f.rgb = (const uint8_t *)d_rgb;
	f.mask = (const uint8_t *)d_mask;
	f.tilemax = c.d_tilemax;
	f.tilemin = c.d_tilemin;
	f.tile_bytes = (unsigned)v->tile_bytes;
	f.depth_m = c.d_depth_m;
#if SFM_K1_TMA_DEPTH
	f.depth_tmap = c.depth_tmap;
#endif
	f.W = v->W; f.H = v->H; f.TW = v->TW; f.TH = v->TH;
	memcpy(f.E, E16, 12 * sizeof(float));
	for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) f.K[r * 3 + c] = v->K[r * 4 + c];
	f.depth_scale = v->desc.depth_scale;
	f.near_gate = v->desc.near_gate;
	float lin = 0.f, tt = 0.f;
	for (int r = 0; r < 3; r++) {
		lin = std::max(lin, fabsf(E16[r * 4 + 0]) + fabsf(E16[r * 4 + 1]) + fabsf(E16[r * 4 + 2]));
		tt = std::max(tt, fabsf(E16[r * 4 + 3]));
	}
	const float k0 = fabsf(f.K[0]) + fabsf(f.K[1]) + fabsf(f.K[2]);
	const float k1 = fabsf(f.K[3]) + fabsf(f.K[4]) + fabsf(f.K[5]);
	const float k2 = fabsf(f.K[6]) + fabsf(f.K[7]) + fabsf(f.K[8]);
	f.cull_lin = lin;
	f.cull_t = tt;
	f.cull_k2 = k2;
	f.cull_slack0 = 1.f + 1e-3f * std::max(k0, k1) / std::max(k2, 1e-20f);
	f.debug = v->debug_ablate;
	return f;
}

// pinhole pattern of K (see cam_to_screen): the only one the reference can build (tsdf.cu:137-150).  The generic
// path stays for arbitrary K and behind SFM_FLAG_GENERIC_K (A/B parity test).
bool canonical_k(const sfm_volume *v, const FrameView &f) {
	const float *K = f.K;
	bool canon = !(v->desc.flags & SFM_FLAG_GENERIC_K) && K[1] == 0.f && K[3] == 0.f && K[6] == 0.f && K[7] == 0.f && K[8] == 1.f;
	for (int i = 0; i < 12 && canon; i++) canon = std::isfinite(f.E[i]) && fabsf(f.E[i]) < 1e15f;
	return canon;
}

// K1a: brick classification into the work lists
template <bool VEC4, bool CULL, bool TMA_TILES>
void launch_classify2(sfm_volume *v, const FrameView &f, const WorkLists &wl, long long nsb) {
	// dynamic shared memory: [TMA-staged tile grids] [block-local MIXED list] [block-local FREE list]
	const size_t smem = (TMA_TILES ? v->tile_bytes : 0) + 2 * (size_t)kSbPerBlock * 32 * sizeof(uint32_t);
	auto kern = classify_kernel<VEC4, CULL, TMA_TILES>;
	if (smem > 48 * 1024 && v->k1a_smem_set != smem) {
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return;  // the launch check reports it
		v->k1a_smem_set = smem;
	}
	// persistent-ish grid: 1024 threads per SM, more blocks when one would otherwise see > kSbPerBlock super-blocks
	constexpr int warps = kK1aThreads / 32;
	const long long want = (nsb + warps - 1) / warps;
	int blocks = (int)std::max<long long>(std::max(1LL, std::min<long long>((1024LL / kK1aThreads) * v->num_sms, want)), (nsb + kSbPerBlock - 1) / kSbPerBlock);
	// without the TMA staging a block has no set-up cost: one super-block per warp, and the hardware
	// block scheduler balances the (very uneven) super-block costs
	if (!TMA_TILES) blocks = (int)std::max(1LL, want);
	kern<<<blocks, kK1aThreads, smem, v->prep_stream>>>(v->g, f, wl);
}

template <bool VEC4>
void launch_classify(sfm_volume *v, const FrameView &f, const WorkLists &wl, long long nsb) {
	// the TMA-staged tile grids must fit a block's shared memory next to its brick lists (frames up to about
	// 1600 x 1200); larger frames read the grids through L1 like SFM_FLAG_NO_TMA does
	const bool fits = v->tile_bytes + 2 * (size_t)kSbPerBlock * 32 * sizeof(uint32_t) <= 160 * 1024;
	const bool cull = !(v->desc.flags & SFM_FLAG_NO_CULL), tma = !(v->desc.flags & SFM_FLAG_NO_TMA) && fits;
	if (!cull) launch_classify2<VEC4, false, false>(v, f, wl, nsb);
	else if (tma) launch_classify2<VEC4, true, true>(v, f, wl, nsb);
	else launch_classify2<VEC4, true, false>(v, f, wl, nsb);
}

// K1b: update of the listed bricks
template <int VEC, bool LABELS, bool KCANON>
void launch_update2(sfm_volume *v, const FrameView &f, const WorkLists &wl, const int32_t *gate) {
	// persistent grid: one resident wave (occupancy x SM count); the warps pull bricks from the lists.
	// dynamic shared memory: the per-warp surface queues
	constexpr int warps = kK1Threads / 32;
	const size_t smem = warps * (kQueue * sizeof(uint4) + (SFM_K1_TMA_DEPTH ? kTmaBoxW * kTmaBoxH * 4 + 8 : 0));
	int &per_sm = v->k1b_per_sm[(VEC == 4 ? 4 : 0) + (LABELS ? 2 : 0) + (KCANON ? 1 : 0)];
	auto kern = integrate_kernel<VEC, LABELS, KCANON>;
	if (!per_sm) {
		if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return;  // the launch check reports it
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kK1Threads, smem) != cudaSuccess || per_sm < 1)
			per_sm = 1;
		// one block slot per SM is left to K1a of the next frame, which runs concurrently on prep_stream: threads
		// and registers of the resident K1b blocks plus one K1a block must fit the SM (registers are allocated per
		// warp in units of 256)
		cudaFuncAttributes ab{}, aa{};
		if (cudaFuncGetAttributes(&ab, kern) == cudaSuccess &&
			cudaFuncGetAttributes(&aa, classify_kernel<true, true, true>) == cudaSuccess) {
			const int rb = ((ab.numRegs + 7) / 8) * 8 * kK1Threads, ra = ((aa.numRegs + 7) / 8) * 8 * kK1aThreads;
			while (per_sm > 1 && (per_sm * kK1Threads + kK1aThreads > 1024 || per_sm * rb + ra > 65536)) per_sm--;
		} else if (per_sm * kK1Threads >= 1024) per_sm = (1024 - kK1aThreads) / kK1Threads;
		if (const char *e = getenv("SFM_K1B_BLOCKS_PER_SM")) per_sm = std::max(1, std::min(per_sm, atoi(e)));
	}
	const long long want = ((long long)v->nbricks + warps - 1) / warps;
	const int blocks = (int)std::max(1LL, std::min<long long>((long long)per_sm * v->num_sms, want));  // persistent: one resident wave
#if SFM_K1_TMA_DEPTH
	kern<<<blocks, kK1Threads, smem, v->stream>>>(v->planes, v->g, f, wl, v->d_stats, v->d_err, gate, *(const CUtensorMap *)f.depth_tmap);
#else
	kern<<<blocks, kK1Threads, smem, v->stream>>>(v->planes, v->g, f, wl, v->d_stats, v->d_err, gate);
#endif
}

template <int VEC, bool LABELS>
void launch_update(sfm_volume *v, const FrameView &f, const WorkLists &wl, const int32_t *gate) {
	const bool canon = VEC == 4 && canonical_k(v, f);
	if (VEC == 4 && canon) launch_update2<VEC, LABELS, VEC == 4>(v, f, wl, gate);
	else launch_update2<VEC, LABELS, false>(v, f, wl, gate);
}

// K0 + K1a + K1b on device-resident frame images.
// K0 and K1a (frame preparation: tile grids, depth in metres, brick lists) only read the frame, never the
// volume, so they run on prep_stream into one of two contexts and overlap K1b of the PREVIOUS frame, which
// still runs on the main stream (K1b is launched with 7 x 128-thread blocks per SM so that one 128-thread
// K1a block fits next to them).  `frame_ready` says when the frame images are valid:
//   kReadyNow      they are valid already (resident frames);
//   kReadyInOrder  they become valid in main-stream order (caller's earlier work on that stream): the
//                  preparation then waits for everything issued on the main stream so far -- correct for any
//                  caller, but serialised behind the previous frame's K1b;
//   an event       recorded by whoever produces the frame (our own upload on copy_stream, a broadcast).
const cudaEvent_t kReadyNow = nullptr;
const cudaEvent_t kReadyInOrder = (cudaEvent_t)(uintptr_t)1;

// first half: K0 + K1a of the next frame on prep_stream (the frame's context is ctx[n_integrate % kCtx])
int enqueue_prepare(sfm_volume *v, const void *d_depth, const void *d_rgb, const void *d_mask, const float *E16, cudaEvent_t frame_ready) {
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set (call sfm_set_bounds / sfm_init_from_frame / sfm_parse_frame first)");
	sfm_volume::PrepCtx &c = v->ctx[v->n_integrate % sfm_volume::kCtx];
	const FrameView f = make_frame_view(v, c, d_depth, d_rgb, d_mask, E16);
	const int slot = (int)(v->n_integrate % sfm_volume::kRing);
	if (frame_ready == kReadyInOrder) {
		CU(cudaEventRecord(v->ev_call, v->stream));
		CU(cudaStreamWaitEvent(v->prep_stream, v->ev_call, 0));
	} else if (frame_ready != kReadyNow) {
		CU(cudaStreamWaitEvent(v->prep_stream, frame_ready, 0));
	}
	CU(cudaStreamWaitEvent(v->prep_stream, c.ev_free, 0));  // K1b of frame i-kCtx has finished with this context
	const int prep_warps = v->TW * v->TH;
	const int prep_blocks = (prep_warps * 32 + 255) / 256;
	prep_frame_kernel<<<prep_blocks, 256, 0, v->prep_stream>>>(f.depth, v->bins > 0 ? f.mask : nullptr, v->W, v->H, v->TW, v->TH,
		v->bins, v->desc.depth_scale, c.d_tilemax, c.d_tilemin, c.d_depth_m, v->d_err, c.d_work);
	LAUNCH_CHECK(v);
	const bool vec4 = (v->g.nz % 4 == 0);
	// K1a work items: super-blocks of kSbX x-planes x kSbG brick rows x one z chunk (k_integrate.cuh)
	const int cpw = vec4 ? (32 >> v->g.zl_log2) : 1, chunk = vec4 ? (4 << v->g.zl_log2) : 32;
	const long long rows = (v->g.Dy + cpw - 1) / cpw;
	const long long nsb = (long long)((v->g.Dx + kSbX - 1) / kSbX) * ((rows + kSbG - 1) / kSbG) * ((v->g.nz + chunk - 1) / chunk);
	if (nsb >= (1LL << 31)) return fail(SFM_ERR_INVALID, "volume too large for the 31-bit super-block ids");
	const WorkLists wl{c.d_list_mixed, c.d_list_free, c.d_work};
	CU(cudaEventRecord(v->ev_k0[slot], v->prep_stream));
	if (vec4) launch_classify<true>(v, f, wl, nsb);
	else launch_classify<false>(v, f, wl, nsb);
	LAUNCH_CHECK(v);

	CU(cudaEventRecord(v->ev_km[slot], v->prep_stream));
	CU(cudaEventRecord(c.ev_ready, v->prep_stream));
	return SFM_OK;
}

// second half: K1b on the main stream, after the context enqueue_prepare filled
int enqueue_update(sfm_volume *v, const void *d_depth, const void *d_rgb, const void *d_mask, const float *E16, const int32_t *gate = nullptr) {
	sfm_volume::PrepCtx &c = v->ctx[v->n_integrate % sfm_volume::kCtx];
	const FrameView f = make_frame_view(v, c, d_depth, d_rgb, d_mask, E16);
	const int slot = (int)(v->n_integrate % sfm_volume::kRing);
	const bool vec4 = (v->g.nz % 4 == 0);
	const WorkLists wl{c.d_list_mixed, c.d_list_free, c.d_work};
	if (v->bins > 0 && v->hist_bound >= 65535u)
		return fail(SFM_ERR_INVALID, "the histogram bins are 16 bits wide inside the library: at most 65535 labelled frames per volume");
	CU(cudaStreamWaitEvent(v->stream, c.ev_ready, 0));
	CU(cudaEventRecord(v->ev_kb[slot], v->stream));
	if (vec4) {
		if (v->bins > 0) launch_update<4, true>(v, f, wl, gate);
		else launch_update<4, false>(v, f, wl, gate);
	} else {
		if (v->bins > 0) launch_update<1, true>(v, f, wl, gate);
		else launch_update<1, false>(v, f, wl, gate);
	}
	LAUNCH_CHECK(v);
	CU(cudaEventRecord(v->ev_k1[slot], v->stream));
	CU(cudaEventRecord(c.ev_free, v->stream));
	v->n_integrate++;
	if (v->bins > 0) v->hist_bound++;
	v->n_obs++;  // tsdf.cu:220 (counted here so the raw / device entry points keep n_obs consistent)
	if (v->desc.flags & SFM_FLAG_SYNC_EVERY_CALL) CU(cudaStreamSynchronize(v->stream));
	return SFM_OK;
}

int integrate_device(sfm_volume *v, const void *d_depth, const void *d_rgb, const void *d_mask, const float *E16,
	cudaEvent_t frame_ready = kReadyInOrder)
{
	int rc = enqueue_prepare(v, d_depth, d_rgb, d_mask, E16, frame_ready);
	if (rc) return rc;
	return enqueue_update(v, d_depth, d_rgb, d_mask, E16);
}

int ensure_ray_buffers(sfm_volume *v, size_t px, bool want_probs) {
	if (px > v->ray_px) {
		cudaFree(v->d_t); cudaFree(v->d_flags); cudaFree(v->d_bgr); cudaFree(v->d_label); cudaFree(v->d_keys); cudaFree(v->d_hits);
		v->d_t = nullptr; v->d_flags = nullptr; v->d_bgr = nullptr; v->d_label = nullptr; v->d_keys = nullptr; v->d_hits = nullptr;
		v->ray_px = 0;
		CU(cudaMalloc(&v->d_hits, px * sizeof(float4)));
		CU(cudaMalloc(&v->d_t, px * 4));
		CU(cudaMalloc(&v->d_flags, px));
		CU(cudaMalloc(&v->d_bgr, px * 3));
		CU(cudaMalloc(&v->d_label, px));
		CU(cudaMalloc(&v->d_keys, px * 8));
		v->ray_px = px;
	}
	if (want_probs && px > v->probs_px) {
		cudaFree(v->d_probs); cudaFree(v->d_box);
		v->d_probs = nullptr; v->d_box = nullptr;
		v->probs_px = 0;
		CU(cudaMalloc(&v->d_probs, px * (size_t)v->bins * 4));
		CU(cudaMalloc(&v->d_box, px * (size_t)v->bins));
		v->probs_px = px;
	}
	return SFM_OK;
}

RayVol make_ray_vol(const sfm_volume *v) {
	RayVol V;
	V.sdf = v->planes.sdf;
	V.hist = v->planes.hist;
	V.bins = v->bins;
	V.g = v->g;
	// skipping is only sound while every value outside the surface blocks stays above the fine-step
	// threshold voxel.x/2 (and positive): those values are miu (unobserved) or >= near_gate
	const bool skip_ok = v->g.miu > v->g.vx && v->desc.near_gate > v->g.vx && v->g.vx > 0.f && !getenv("SFM_NO_SKIP");
	V.occ = skip_ok ? v->planes.occ : nullptr;
	V.oby = v->planes.oby;
	V.obz = v->planes.obz;
	return V;
}

// Rt = R(extrinsic2init)^T, o = -Rt * t   (tsdf.cu:432-435; OpenCV float gemm there)
RayCam make_backproj_cam(const sfm_volume *v, const float *E16) {
	RayCam c{};
	for (int r = 0; r < 3; r++) for (int k = 0; k < 4; k++) c.M[r * 4 + k] = v->Kinv[r * 4 + k];
	for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) c.Rt[r * 3 + k] = E16[k * 4 + r];
	for (int r = 0; r < 3; r++) {
		double s = 0;
		for (int k = 0; k < 3; k++) s += (double)(-c.Rt[r * 3 + k]) * (double)E16[k * 4 + 3];
		c.o[r] = (float)s;
	}
	c.show = 0;
	c.W = v->W;
	c.H = v->H;
	return c;
}

RayCam make_show_cam(const float *s2w16, const float *c3, int w, int h) {
	RayCam c{};
	memcpy(c.M, s2w16, 12 * sizeof(float));
	memcpy(c.o, c3, 3 * sizeof(float));
	c.show = 1;
	c.W = w;
	c.H = h;
	return c;
}

int ray_blocks(int w, int h) {
	const long long tiles = (long long)((w + 7) / 8) * ((h + 3) / 4);
	return (int)((tiles + 3) / 4);  // 4 warps (128 threads) per block
}

// march_kernel on v->stream: a persistent grid (one resident wave) whose warps take the image's 8x4-pixel tiles
// longest-first, by the costs the previous march of this handle measured for the same image shape; the order for the next
// march is built right after this one (order_tiles_kernel, a few microseconds).
int launch_march(sfm_volume *v, const RayVol &V, const RayCam &cam, float4 *d_hits, uint8_t *d_flags, int row0, int rows, int tstride = 1,
	int compact = 0) {
	const long long tiles = (long long)((cam.W + 7) / 8) * ((rows + 3) / 4);
	if (tiles <= 0 || tiles > 0x7fffffffLL) return fail(SFM_ERR_INVALID, "image too large for the ray-marcher");
	if ((size_t)tiles > v->tile_cap) {
		cudaFree(v->d_tile_cost); cudaFree(v->d_tile_order);
		v->d_tile_cost = v->d_tile_order = nullptr;
		v->tile_cap = 0;
		v->order_key = 0;
		CU(cudaMalloc(&v->d_tile_cost, (size_t)tiles * 4));
		CU(cudaMalloc(&v->d_tile_order, (size_t)tiles * 4));
		v->tile_cap = (size_t)tiles;
	}
	if (!v->march_per_sm) {
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v->march_per_sm, march_kernel, kMarchThreads, 0) != cudaSuccess || v->march_per_sm < 1)
			v->march_per_sm = 1;
	}
	const int wpb = kMarchThreads / 32;
	const int blocks = (int)std::max(1LL, std::min<long long>((tiles + wpb - 1) / wpb, (long long)v->march_per_sm * v->num_sms));
	const unsigned long long key = ((unsigned long long)(unsigned)tstride << 52) | ((unsigned long long)(unsigned)cam.W << 36) | ((unsigned long long)(unsigned)rows << 18) |
		(unsigned)row0 | (1ull << 63);
	const unsigned *order = (v->order_key == key && !(v->desc.flags & SFM_FLAG_DEBUG_ABLATE && v->debug_ablate & 128)) ? v->d_tile_order : nullptr;
	march_kernel<<<blocks, kMarchThreads, 0, v->stream>>>(V, cam, d_hits, d_flags, row0, rows, tstride, compact, v->d_ray_stats, v->d_march_work, order, v->d_tile_cost);
	LAUNCH_CHECK(v);
	order_tiles_kernel<<<1, 1024, 0, v->stream>>>(v->d_tile_cost, (unsigned)tiles, v->d_tile_order);
	LAUNCH_CHECK(v);
	v->order_key = key;
	return SFM_OK;
}

int require_full_volume(const sfm_volume *v, const char *what) {
	if (v->g.z0 != 0 || v->g.nz != v->g.Dz)
		return fail(SFM_ERR_INVALID, std::string(what) + ": this handle stores a z-slab; sharded ray-marching goes through sfm_raycast_keys_dev + a min-composite");
	return SFM_OK;
}

struct FoldLayout {
	size_t oPos, oTm, oT, oBm, oB, oCm, oNoHit, oFirst, total;
};
FoldLayout fold_layout(int L) {
	FoldLayout f;
	size_t o = 0;
	f.oPos = o; o += (size_t)L * L * 8;
	f.oTm = o; o += (size_t)L * L * 8;
	f.oT = o; o += (size_t)L * 8;
	f.oBm = o; o += (size_t)L * L * 4;
	f.oB = o; o += (size_t)L * 4;
	f.oCm = o; o += (size_t)L * 4;
	f.oNoHit = o; o += (size_t)L * 4;
	f.oFirst = o; o += (size_t)L * 4;
	f.total = o;
	return f;
}

__global__ void relabel_kernel(uint8_t *__restrict__ mask, int n, const uint8_t *__restrict__ lut, const int32_t *__restrict__ gate);
void combine_tables(const sfm_volume *v, int max_obj_now, double *A, uint32_t *C);
void decide(const sfm_volume *v, const double *A, const uint32_t *C, int max_obj_now, const unsigned *first_pix,
	int *num_objs, sfm_merge_report *rep);

FoldTables fold_tables_at(uint8_t *base, int L) {
	const FoldLayout fl = fold_layout(L);
	FoldTables tb;
	tb.Pos = (long long *)(base + fl.oPos);
	tb.Tm = (long long *)(base + fl.oTm);
	tb.T = (long long *)(base + fl.oT);
	tb.Bm = (unsigned *)(base + fl.oBm);
	tb.B = (unsigned *)(base + fl.oB);
	tb.Cm = (unsigned *)(base + fl.oCm);
	tb.NoHit = (unsigned *)(base + fl.oNoHit);
	tb.FirstPix = (unsigned *)(base + fl.oFirst);
	return tb;
}

// zero the tables at `d_tables` and fold the hits in v->d_hits into them (see fold_kernel)
int launch_fold(sfm_volume *v, uint8_t *d_tables, const uint8_t *d_mask, const unsigned long long *d_gkeys, int do_counts) {
	const int L = v->bins;
	const FoldLayout fl = fold_layout(L);
	CU(cudaMemsetAsync(d_tables, 0, fl.total, v->stream));
	if (do_counts) CU(cudaMemsetAsync(d_tables + fl.oFirst, 0xff, (size_t)L * 4, v->stream));
	const FoldTables tb = fold_tables_at(d_tables, L);
	const RayVol V = make_ray_vol(v);
	const int npix = v->W * v->H;
	const int blocks = (npix + 127) / 128;
	const float n_obs = (float)v->n_obs, prior = v->desc.prior_err_rate, pres = v->desc.presence_thresh;
	const int nb = (L + 31) / 32;
	switch (nb) {
	case 1: fold_kernel<1><<<blocks, 128, 0, v->stream>>>(V, npix, v->d_hits, d_mask, n_obs, prior, pres, tb, d_gkeys, do_counts); break;
	case 2: fold_kernel<2><<<blocks, 128, 0, v->stream>>>(V, npix, v->d_hits, d_mask, n_obs, prior, pres, tb, d_gkeys, do_counts); break;
	case 3: fold_kernel<3><<<blocks, 128, 0, v->stream>>>(V, npix, v->d_hits, d_mask, n_obs, prior, pres, tb, d_gkeys, do_counts); break;
	case 4: fold_kernel<4><<<blocks, 128, 0, v->stream>>>(V, npix, v->d_hits, d_mask, n_obs, prior, pres, tb, d_gkeys, do_counts); break;
	default: fold_kernel<8><<<blocks, 128, 0, v->stream>>>(V, npix, v->d_hits, d_mask, n_obs, prior, pres, tb, d_gkeys, do_counts); break;
	}
	LAUNCH_CHECK(v);
	return SFM_OK;
}

// enqueue the fused back-project + fold on the device mask; tables end up in v->d_fold
int enqueue_march_fold(sfm_volume *v, const float *E16, const uint8_t *d_mask) {
	const RayVol V = make_ray_vol(v);
	const RayCam cam = make_backproj_cam(v, E16);
	const int npix = v->W * v->H;
	int rc = ensure_ray_buffers(v, (size_t)npix, false);
	if (rc) return rc;
	rc = launch_march(v, V, cam, v->d_hits, nullptr, 0, v->H);
	if (rc) return rc;
	return launch_fold(v, v->d_fold, d_mask, nullptr, 1);
}

// the same, then the tables are copied to v->h_fold (pinned) and the stream is synchronised (parity hook)
int run_fold(sfm_volume *v, const float *E16, const uint8_t *d_mask) {
	int rc = enqueue_march_fold(v, E16, d_mask);
	if (rc) return rc;
	CU(cudaMemcpyAsync(v->h_fold, v->d_fold, fold_layout(v->bins).total, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	return SFM_OK;
}

// combine the folded partial tables into the reference's A / C (see k_raymarch.cuh)
void combine_tables(const sfm_volume *v, int max_obj_now, double *A, uint32_t *C) {
	const int L = v->bins;
	const FoldLayout fl = fold_layout(L);
	const long long *Pos = (const long long *)(v->h_fold + fl.oPos);
	const long long *Tm = (const long long *)(v->h_fold + fl.oTm);
	const long long *T = (const long long *)(v->h_fold + fl.oT);
	const unsigned *Bm = (const unsigned *)(v->h_fold + fl.oBm);
	const unsigned *B = (const unsigned *)(v->h_fold + fl.oB);
	const unsigned *Cm = (const unsigned *)(v->h_fold + fl.oCm);
	const unsigned *NoHit = (const unsigned *)(v->h_fold + fl.oNoHit);
	const long long logprior_fx = llrint((double)logf(v->desc.prior_err_rate) * 4294967296.0);
	for (int m = 0; m < L; m++)
		for (int n = 0; n < L; n++) {
			long long a = 0;
			uint32_t c = 0;
			if (m >= 1 && n >= 1 && m < max_obj_now) {
				a = Pos[m * L + n] + (long long)NoHit[m] * logprior_fx + T[n] - Tm[m * L + n];
				c = Cm[m] + B[n] - Bm[m * L + n];
			}
			A[m * L + n] = (double)a / 4294967296.0;
			C[m * L + n] = c;
		}
}

int max_label_host(const uint8_t *mask, size_t n) {
	uint8_t mx = 0;
	for (size_t i = 0; i < n; i++) mx = mask[i] > mx ? mask[i] : mx;
	return mx;
}

// decision half of filter_overlaps (tsdf.cu:335-389) from tables; fills report->assign (a LUT)
void decide(const sfm_volume *v, const double *A, const uint32_t *C, int max_obj_now, const unsigned *first_pix,
	int *num_objs, sfm_merge_report *rep)
{
	const int L = v->bins;
	const float thr = v->desc.accept_factor * v->desc.prior_err_rate;
	memset(rep, 0, sizeof(*rep));
	rep->max_obj_now = max_obj_now;
	float margin = INFINITY;
	std::vector<int> owner(256, 0);
	std::vector<float> owner_p(256, 0.f);
	for (int m = 1; m < max_obj_now && m < L; m++) {
		int best = -1;
		float bp = 0.f, second = 0.f;
		for (int j = 1; j < L; j++) {
			const float p = (C[m * L + j] == 0) ? 0.f : expf((float)A[m * L + j] / (float)C[m * L + j]);
			if (p > bp) { second = bp; best = j; bp = p; }
			else if (p > second) second = p;
		}
		rep->best_prob[m] = bp;
		if (first_pix[m] != 0xffffffffu) {  // only labels present in the frame decide anything visible
			margin = std::min(margin, fabsf(bp - thr));
			if (bp > thr) margin = std::min(margin, bp - second);
		}
		if (bp > thr) {
			if (owner[best] == 0) { owner[best] = m; owner_p[best] = bp; }
			else {
				margin = std::min(margin, fabsf(owner_p[best] - bp));
				if (owner_p[best] < bp) { owner[best] = m; owner_p[best] = bp; }
			}
		}
	}
	for (int j = 1; j < 256; j++) if (owner[j]) rep->assign[owner[j]] = j;
	// unassigned labels that occur get fresh ids in raster order of first appearance (tsdf.cu:378-387)
	std::vector<std::pair<unsigned, int>> fresh;
	for (int m = 1; m < max_obj_now && m < L; m++)
		if (!rep->assign[m] && first_pix[m] != 0xffffffffu) fresh.push_back({first_pix[m], m});
	std::sort(fresh.begin(), fresh.end());
	for (auto &pr : fresh) { rep->assign[pr.second] = *num_objs; (*num_objs)++; }
	rep->num_objs = *num_objs;
	rep->margin = margin;
}

// `gate` (nullable): set when the decision overflowed the bins -- the mask then keeps the caller's labels
__global__ void relabel_kernel(uint8_t *__restrict__ mask, int n, const uint8_t *__restrict__ lut, const int32_t *__restrict__ gate) {
	if (gate && *gate) return;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) mask[i] = lut[mask[i]];
}

// Decision half of filter_overlaps (tsdf.cu:335-389) on the device, from the folded integer tables: the same
// statements as decide() + combine_tables() above, one thread per incoming label for the arg-max over the
// volume's instances (tsdf.cu:340-348), thread 0 for the sequential parts (collisions keep the larger
// probability, tsdf.cu:349-362; fresh ids in raster order of first appearance, tsdf.cu:378-387).  Writes the
// 256-entry label map the relabel kernel uses, so a frame's merge needs no host round trip before its
// integration; the host only reads the 2 KB report back (it relabels the caller's copy of the mask).
struct MergeOut {
	sfm_merge_report rep;
	uint8_t lut[256];
	int32_t overflow;  // a global id reached `bins` (the reference overflows its histogram there)
};

__global__ void __launch_bounds__(256) decide_kernel(FoldTables tb, int L, float prior, float accept_factor, long long logprior_fx,
	int num_objs_in, MergeOut *__restrict__ out)
{
	__shared__ int s_best[256];
	__shared__ float s_bp[256], s_second[256];
	__shared__ unsigned s_first[256], s_cm[256];  // staged in parallel: thread 0's loops below must not chase global loads
	__shared__ int s_max_obj;
	const int m = threadIdx.x;
	s_first[m] = m < L ? tb.FirstPix[m] : 0xffffffffu;
	s_cm[m] = m < L ? tb.Cm[m] : 0u;
	if (m == 0) s_max_obj = 1;
	__syncthreads();
	if (m >= 1 && m < L && s_cm[m]) atomicMax(&s_max_obj, m + 1);
	__syncthreads();
	const int max_obj_now = s_max_obj;
	int best = -1;
	float bp = 0.f, second = 0.f;
	if (m >= 1 && m < max_obj_now && m < L) {
		for (int j = 1; j < L; j++) {
			const long long a = tb.Pos[(size_t)m * L + j] + (long long)tb.NoHit[m] * logprior_fx + tb.T[j] - tb.Tm[(size_t)m * L + j];
			const unsigned c = tb.Cm[m] + tb.B[j] - tb.Bm[(size_t)m * L + j];
			const float A = (float)((double)a / 4294967296.0);
			const float p = (c == 0) ? 0.f : expf(A / (float)c);
			if (p > bp) { second = bp; best = j; bp = p; }
			else if (p > second) second = p;
		}
	}
	__shared__ int s_owner[256], s_assign[256], s_newmax;
	__shared__ float s_owner_p[256];
	s_best[m] = best; s_bp[m] = bp; s_second[m] = second;
	s_owner[m] = 0; s_owner_p[m] = 0.f; s_assign[m] = 0;
	if (m == 0) s_newmax = 0;
	sfm_merge_report &rep = out->rep;
	rep.best_prob[m] = (m >= 1 && m < max_obj_now && m < L) ? bp : 0.f;
	__syncthreads();
	if (m == 0) {  // the sequential parts, on shared memory only
		const float thr = accept_factor * prior;
		float margin = INFINITY;
		for (int k = 1; k < max_obj_now && k < L; k++) {
			const float kbp = s_bp[k];
			const int kb = s_best[k];
			if (s_first[k] != 0xffffffffu) {  // only labels present in the frame decide anything visible
				margin = fminf(margin, fabsf(kbp - thr));
				if (kbp > thr) margin = fminf(margin, kbp - s_second[k]);
			}
			if (kbp > thr) {
				if (s_owner[kb] == 0) { s_owner[kb] = k; s_owner_p[kb] = kbp; }
				else {
					margin = fminf(margin, fabsf(s_owner_p[kb] - kbp));
					if (s_owner_p[kb] < kbp) { s_owner[kb] = k; s_owner_p[kb] = kbp; }
				}
			}
		}
		for (int j = 1; j < 256; j++) if (s_owner[j]) s_assign[s_owner[j]] = j;
		// unassigned labels that occur get fresh ids in raster order of first appearance: repeated selection of
		// the smallest first-pixel index among them (at most 255 labels)
		int num_objs = num_objs_in;
		for (;;) {
			unsigned bestpix = 0xffffffffu;
			int bestm = 0;
			for (int k = 1; k < max_obj_now && k < L; k++)
				if (!s_assign[k] && s_first[k] < bestpix) { bestpix = s_first[k]; bestm = k; }
			if (!bestm) break;
			s_assign[bestm] = num_objs++;
		}
		rep.max_obj_now = max_obj_now;
		rep.num_objs = num_objs;
		rep.margin = margin;
	}
	__syncthreads();
	rep.assign[m] = s_assign[m];
	out->lut[m] = (uint8_t)s_assign[m];
	if (m < max_obj_now) atomicMax(&s_newmax, s_assign[m]);
	__syncthreads();
	if (m == 0) out->overflow = s_newmax >= L ? 1 : 0;
}

static_assert(sizeof(MergeOut) <= 4096, "MergeOut buffer");

// Enqueue: decision on the tables at d_tables -> label map -> relabel of the device mask.  Nothing here
// waits for the GPU; finish_device_decision() later fetches the report.
int enqueue_device_decision(sfm_volume *v, uint8_t *d_tables, uint8_t *d_mask) {
	const int L = v->bins;
	const size_t npx = (size_t)v->W * v->H;
	const FoldTables tb = fold_tables_at(d_tables, L);
	const long long logprior_fx = llrint((double)logf(v->desc.prior_err_rate) * 4294967296.0);
	MergeOut *out = (MergeOut *)v->d_merge;
	decide_kernel<<<1, 256, 0, v->stream>>>(tb, L, v->desc.prior_err_rate, v->desc.accept_factor, logprior_fx, v->num_objs, out);
	LAUNCH_CHECK(v);
	CU(cudaMemcpyAsync(v->h_merge, v->d_merge, sizeof(MergeOut), cudaMemcpyDeviceToHost, v->stream));
	CU(cudaEventRecord(v->ev_merge, v->stream));
	relabel_kernel<<<(int)((npx + 255) / 256), 256, 0, v->stream>>>(d_mask, (int)npx, out->lut, &out->overflow);
	LAUNCH_CHECK(v);
	return SFM_OK;
}

// Wait for the report of the decision enqueued last (not for the kernels behind it), update num_objs.
int finish_device_decision(sfm_volume *v, uint8_t *lut_out) {
	CU(cudaEventSynchronize(v->ev_merge));
	const MergeOut *out = (const MergeOut *)v->h_merge;
	v->last_merge = out->rep;
	if (lut_out) memcpy(lut_out, out->lut, 256);
	if (out->overflow)
		return fail(SFM_ERR_INVALID, "merge produced a global instance id >= bins (num_objs outgrew the histogram; the reference overflows here, tsdf.cu:61,383)");
	v->num_objs = out->rep.num_objs;
	return SFM_OK;
}

}  // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

const char *sfm_last_error(void) { return g_err.c_str(); }
const char *sfm_version(void) { return "sfm_b200 0.1 (sm_100a)"; }

void sfm_desc_default(sfm_desc *d) {
	memset(d, 0, sizeof(*d));
	d->dims[0] = d->dims[1] = d->dims[2] = 256;  // tsdf.cuh:52
	d->bins = 32;                                 // tsdf.cuh:4
	d->width = 640;
	d->height = 480;
	for (int i = 0; i < 4; i++) d->K[i * 4 + i] = 1.f;
	d->K[0] = 520.9f; d->K[5] = 521.0f; d->K[2] = 325.1f; d->K[6] = 249.7f;  // kernel.cpp:39
	d->prior_err_rate = 0.05f;
	d->duplicate_thresh = 0.5f;
	d->presence_thresh = 0.3f;
	d->accept_factor = 3.f;
	d->depth_scale = 5000.f;
	d->trunc_voxels = 5.f;
	d->near_gate = 0.99f;
}

void sfm_palette(uint8_t *rgb, int n) {
	for (int i = 0; i < n; i++) memcpy(rgb + i * 3, kPalette16 + (i % 16) * 3, 3);
}

int sfm_create(const sfm_desc *desc, sfm_volume **out) {
	if (!desc || !out) return fail(SFM_ERR_INVALID, "null argument");
	*out = nullptr;
	if (desc->dims[0] <= 0 || desc->dims[1] <= 0 || desc->dims[2] <= 0 || desc->dims[0] > 65535 ||
		desc->dims[1] > 65535 || desc->dims[2] > 65535)
		return fail(SFM_ERR_INVALID, "dims must be in [1, 65535]");
	if (desc->bins < 0 || desc->bins > kMaxBins) return fail(SFM_ERR_INVALID, "bins must be in [0, 255]");
	if (desc->width <= 0 || desc->height <= 0 || desc->width > 65535 || desc->height > 65535)
		return fail(SFM_ERR_INVALID, "bad frame size");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		return fail(SFM_ERR_NODEVICE, "no CUDA device: this library has no CPU fallback");
	}
	if (desc->device < 0 || desc->device >= ndev) return fail(SFM_ERR_INVALID, "bad device ordinal");
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, desc->device));
	if (prop.major != 10)
		return fail(SFM_ERR_NODEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
			"; the kernels are built for sm_100a only and there is no fallback");
	CU(cudaSetDevice(desc->device));
	if (const char *g = getenv("SFM_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));

	sfm_volume *v = new sfm_volume();
	v->desc = *desc;
	if (desc->flags & SFM_FLAG_DEBUG_ABLATE) {  // profiling only: a stray environment variable alone changes nothing
		if (const char *e = getenv("SFM_DEBUG_ABLATE")) v->debug_ablate = atoi(e) & ~(SFM_K1_TMA_DEPTH ? 0 : 64);
	}
	v->bins = desc->bins;
	v->W = desc->width;
	v->H = desc->height;
	v->TW = (v->W + kTile - 1) / kTile;
	v->TH = (v->H + kTile - 1) / kTile;
	v->g.Dx = desc->dims[0]; v->g.Dy = desc->dims[1]; v->g.Dz = desc->dims[2];
	v->g.z0 = desc->slab_z0;
	v->g.nz = desc->slab_nz > 0 ? desc->slab_nz : desc->dims[2] - desc->slab_z0;
	if (v->g.z0 < 0 || v->g.nz <= 0 || v->g.z0 + v->g.nz > v->g.Dz) { delete v; return fail(SFM_ERR_INVALID, "bad z-slab"); }
	v->g.own_z0 = desc->own_nz > 0 ? desc->own_z0 : v->g.z0;
	v->g.own_nz = desc->own_nz > 0 ? desc->own_nz : v->g.nz;
	if (v->g.own_z0 < v->g.z0 || v->g.own_z0 + v->g.own_nz > v->g.z0 + v->g.nz) { delete v; return fail(SFM_ERR_INVALID, "owned planes must lie inside the stored slab"); }
	v->nvox = (size_t)v->g.Dx * v->g.Dy * v->g.nz;
	memcpy(v->K, desc->K, sizeof(v->K));
	bool kinv_given = false;
	for (int i = 0; i < 16; i++) kinv_given |= desc->Kinv[i] != 0.f;
	if (kinv_given) memcpy(v->Kinv, desc->Kinv, sizeof(v->Kinv));
	else if (!mat4_inv(v->K, v->Kinv)) { delete v; return fail(SFM_ERR_INVALID, "singular intrinsic matrix"); }

#define CU_OR_DESTROY(call)                                                                          \
	do {                                                                                             \
		cudaError_t e_ = (call);                                                                     \
		if (e_ != cudaSuccess) {                                                                     \
			int code_ = fail(e_ == cudaErrorMemoryAllocation ? SFM_ERR_NOMEM : SFM_ERR_CUDA,         \
				std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
			std::string keep_ = g_err;                                                               \
			sfm_destroy(v);                                                                          \
			g_err = keep_;                                                                           \
			return code_;                                                                            \
		}                                                                                            \
	} while (0)

	CU_OR_DESTROY(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
	v->own_stream = true;
	CU_OR_DESTROY(cudaEventCreate(&v->ev_t0));
	CU_OR_DESTROY(cudaEventCreate(&v->ev_t1));
	for (int i = 0; i < sfm_volume::kRing; i++) {
		CU_OR_DESTROY(cudaEventCreate(&v->ev_k0[i]));
		CU_OR_DESTROY(cudaEventCreate(&v->ev_km[i]));
		CU_OR_DESTROY(cudaEventCreate(&v->ev_kb[i]));
		CU_OR_DESTROY(cudaEventCreate(&v->ev_k1[i]));
	}
	CU_OR_DESTROY(cudaMalloc(&v->planes.sdf, v->nvox * 4));
	CU_OR_DESTROY(cudaMalloc(&v->planes.wt, v->nvox * 4));
	CU_OR_DESTROY(cudaMalloc(&v->planes.color, v->nvox * 3));
	v->g.ngz = (v->g.nz + kHistTZ - 1) / kHistTZ;
	v->hist_elems = (size_t)v->g.Dx * v->g.Dy * v->g.ngz * (size_t)std::max(v->bins, 0) * kHistTZ;
	if (v->bins > 0) {
		CU_OR_DESTROY(cudaMalloc(&v->planes.hist, v->hist_elems * sizeof(hist_t)));
		CU_OR_DESTROY(cudaMalloc(&v->d_hist_max, 4));
	}
	v->planes.bins = v->bins;
	{  // surface-block map, see Planes::occ
		const int obx = (v->g.Dx + 7) / 8;
		v->planes.oby = v->g.oby = (v->g.Dy + 7) / 8;
		v->planes.obz = v->g.obz = (v->g.nz + 7) / 8;
		const size_t bytes1 = ((size_t)obx * v->planes.oby * v->planes.obz + 15) / 16 * 16;
		const int obx2 = (v->g.Dx + 31) / 32;
		v->planes.oby2 = v->g.oby2 = (v->g.Dy + 31) / 32;
		v->planes.obz2 = v->g.obz2 = (v->g.nz + 31) / 32;
		v->planes.occ2_off = v->g.occ2_off = (unsigned)bytes1;
		v->occ_bytes = bytes1 + (size_t)obx2 * v->planes.oby2 * v->planes.obz2;
		CU_OR_DESTROY(cudaMalloc(&v->planes.occ, v->occ_bytes));
		CU_OR_DESTROY(cudaMemset(v->planes.occ, 0, v->occ_bytes));
	}
	const size_t npx = (size_t)v->W * v->H;
	CU_OR_DESTROY(cudaStreamCreateWithFlags(&v->copy_stream, cudaStreamNonBlocking));
	for (int i = 0; i < 2; i++) {
		CU_OR_DESTROY(cudaMalloc(&v->d_frame[i], npx * 6));
		CU_OR_DESTROY(cudaMemset(v->d_frame[i], 0, npx * 6));
		CU_OR_DESTROY(cudaEventCreateWithFlags(&v->ev_uploaded[i], cudaEventDisableTiming));
		CU_OR_DESTROY(cudaEventCreateWithFlags(&v->ev_consumed[i], cudaEventDisableTiming));
	}
	v->d_depth = (uint16_t *)v->d_frame[0];
	v->d_rgb = v->d_frame[0] + npx * 2;
	v->d_mask = v->d_frame[0] + npx * 5;
	CU_OR_DESTROY(cudaMallocHost(&v->h_stat_ring, (size_t)sfm_volume::kStatRing * 2 * kStatSlots * 8));
	for (int i = 0; i < sfm_volume::kStatRing; i++) CU_OR_DESTROY(cudaEventCreateWithFlags(&v->ev_stat[i], cudaEventDisableTiming));
	v->tile_bytes = (((size_t)v->TW * v->TH * 4) + 15) / 16 * 16;
	CU_OR_DESTROY(cudaStreamCreateWithFlags(&v->prep_stream, cudaStreamNonBlocking));
	CU_OR_DESTROY(cudaEventCreateWithFlags(&v->ev_call, cudaEventDisableTiming));
	for (auto &c : v->ctx) {
		CU_OR_DESTROY(cudaMalloc(&c.d_tilemax, v->tile_bytes));
		CU_OR_DESTROY(cudaMemset(c.d_tilemax, 0, v->tile_bytes));
		c.d_tilemin = c.d_tilemax + (size_t)v->TW * v->TH;
		CU_OR_DESTROY(cudaMalloc(&c.d_depth_m, npx * 4));
#if SFM_K1_TMA_DEPTH
		{
			typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
				const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
			void *fn = nullptr;
			cudaDriverEntryPointQueryResult qr;
			CU_OR_DESTROY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
			alignas(64) CUtensorMap tm;
			const cuuint64_t gdim[2] = {(cuuint64_t)v->W, (cuuint64_t)v->H}, gstride[1] = {(cuuint64_t)v->W * 4};
			const cuuint32_t box[2] = {(cuuint32_t)kTmaBoxW, (cuuint32_t)kTmaBoxH}, estr[2] = {1, 1};
			if (!fn || ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c.d_depth_m, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
					CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
				sfm_destroy(v);
				return fail(SFM_ERR_CUDA, "cuTensorMapEncodeTiled failed");
			}
			static_assert(sizeof(tm) == sizeof(c.depth_tmap), "CUtensorMap is 128 bytes");
			memcpy(c.depth_tmap, &tm, sizeof(tm));
		}
#endif
		CU_OR_DESTROY(cudaMalloc(&c.d_work, 32));
		CU_OR_DESTROY(cudaMemset(c.d_work, 0, 32));
		CU_OR_DESTROY(cudaEventCreateWithFlags(&c.ev_ready, cudaEventDisableTiming));
		CU_OR_DESTROY(cudaEventCreateWithFlags(&c.ev_free, cudaEventDisableTiming));
	}
	CU_OR_DESTROY(cudaMalloc(&v->d_stats, 2 * kStatSlots * 8));
	CU_OR_DESTROY(cudaMemset(v->d_stats, 0, 2 * kStatSlots * 8));
	CU_OR_DESTROY(cudaMalloc(&v->d_ray_stats, 64));
	CU_OR_DESTROY(cudaMemset(v->d_ray_stats, 0, 64));
	CU_OR_DESTROY(cudaMalloc(&v->d_march_work, 8));
	CU_OR_DESTROY(cudaMemset(v->d_march_work, 0, 8));
	CU_OR_DESTROY(cudaMalloc(&v->d_err, 16));
	{  // brick lists (k_integrate.cuh: WorkLists); ids pack x << 21 | brick row << 10 | z chunk
		// brick shape on the 128-bit path: 32 planes per brick unless the slab is so thin that more than half the lanes
		// of such a brick would fall outside it (see VolGeom::zl_log2); SFM_ZL_LOG2 overrides (experiments)
		const bool vec4 = v->g.nz % 4 == 0;
		int zl = 3;
		if (vec4) {
			double best = 0.0;
			for (int c = 3; c >= 0; c--) {
				const int planes = 4 << c;
				best = std::max(best, (double)v->g.nz / (double)(((v->g.nz + planes - 1) / planes) * planes));
			}
			for (int c = 3; c >= 0; c--) {
				const int planes = 4 << c;
				if ((double)v->g.nz / (double)(((v->g.nz + planes - 1) / planes) * planes) >= 0.5 * best) { zl = c; break; }
			}
			if (const char *e = getenv("SFM_ZL_LOG2")) zl = std::min(3, std::max(0, atoi(e)));
		}
		v->g.zl_log2 = zl;
		const int cpw = vec4 ? (32 >> zl) : 1, planes = vec4 ? (4 << zl) : 32;
		const size_t rows = ((size_t)v->g.Dy + cpw - 1) / cpw, chunks = ((size_t)v->g.nz + planes - 1) / planes;
		if (v->g.Dx > (1 << (32 - kIdXShift)) || rows > (1u << (kIdXShift - kIdGShift)) || chunks > (1u << kIdGShift)) {
			sfm_destroy(v);
			return fail(SFM_ERR_INVALID, "volume too large for the packed brick ids (x <= 2048, y <= 8192 (2048 when nz % 4 != 0), nz <= 32768)");
		}
		v->nbricks = (size_t)v->g.Dx * rows * chunks;
		for (auto &c : v->ctx) {
			CU_OR_DESTROY(cudaMalloc(&c.d_list_mixed, v->nbricks * 4));
			CU_OR_DESTROY(cudaMalloc(&c.d_list_free, v->nbricks * 4));
		}
	}
	v->num_sms = prop.multiProcessorCount;
	CU_OR_DESTROY(cudaMemset(v->d_err, 0, 16));
	CU_OR_DESTROY(cudaMalloc(&v->d_palette, 256 * 3));
	CU_OR_DESTROY(cudaMalloc(&v->d_lut, 256));
	{
		uint8_t pal[256 * 3];
		sfm_palette(pal, 256);
		CU_OR_DESTROY(cudaMemcpy(v->d_palette, pal, sizeof(pal), cudaMemcpyHostToDevice));
	}
	v->frame_bytes = npx * 6;
	for (int i = 0; i < 2; i++) {
		CU_OR_DESTROY(cudaMallocHost(&v->pin[i].buf, v->frame_bytes));
		CU_OR_DESTROY(cudaEventCreateWithFlags(&v->pin[i].free_ev, cudaEventDisableTiming));
	}
	CU_OR_DESTROY(cudaMallocHost(&v->h_stats, 2 * kStatSlots * 8));
	CU_OR_DESTROY(cudaMallocHost(&v->h_err, 4));
	if (v->bins > 0) {
		v->fold_bytes = fold_layout(v->bins).total;
		CU_OR_DESTROY(cudaMalloc(&v->d_fold, v->fold_bytes));
		CU_OR_DESTROY(cudaMallocHost(&v->h_fold, v->fold_bytes));
		CU_OR_DESTROY(cudaMalloc(&v->d_merge, 4096));
		CU_OR_DESTROY(cudaMallocHost(&v->h_merge, 4096));
		CU_OR_DESTROY(cudaEventCreateWithFlags(&v->ev_merge, cudaEventDisableTiming));
	}
#undef CU_OR_DESTROY
	*out = v;
	return SFM_OK;
}

void sfm_destroy(sfm_volume *v) {
	if (!v) return;
	cudaSetDevice(v->desc.device);
	if (v->prep_stream) cudaStreamSynchronize(v->prep_stream);
	if (v->stream) cudaStreamSynchronize(v->stream);
	cudaFree(v->planes.sdf); cudaFree(v->planes.wt); cudaFree(v->planes.color); cudaFree(v->planes.hist); cudaFree(v->planes.occ);
	cudaFree(v->d_hist_ref); cudaFree(v->d_hist_chunk); cudaFree(v->d_hist_max);
	if (v->copy_stream) cudaStreamSynchronize(v->copy_stream);
	for (int i = 0; i < 2; i++) {
		cudaFree(v->d_frame[i]);
		if (v->ev_uploaded[i]) cudaEventDestroy(v->ev_uploaded[i]);
		if (v->ev_consumed[i]) cudaEventDestroy(v->ev_consumed[i]);
	}
	if (v->h_stat_ring) cudaFreeHost(v->h_stat_ring);
	for (int i = 0; i < sfm_volume::kStatRing; i++) if (v->ev_stat[i]) cudaEventDestroy(v->ev_stat[i]);
	if (v->copy_stream) cudaStreamDestroy(v->copy_stream);
	if (v->prep_stream) cudaStreamDestroy(v->prep_stream);
	for (auto &c : v->ctx) {
		cudaFree(c.d_tilemax); cudaFree(c.d_depth_m); cudaFree(c.d_work); cudaFree(c.d_list_mixed); cudaFree(c.d_list_free);
		if (c.ev_ready) cudaEventDestroy(c.ev_ready);
		if (c.ev_free) cudaEventDestroy(c.ev_free);
	}
	if (v->ev_call) cudaEventDestroy(v->ev_call);
	cudaFree(v->d_stats);
	cudaFree(v->d_ray_stats);
	cudaFree(v->d_march_work); cudaFree(v->d_tile_cost); cudaFree(v->d_tile_order);
#if SFM_K1_TMA_DEPTH
	if (v->debug_ablate & 64) {
		uint32_t e[4] = {0, 0, 0, 0};
		if (v->d_err && cudaMemcpy(e, v->d_err, 16, cudaMemcpyDeviceToHost) == cudaSuccess)
			fprintf(stderr, "[sfm A/B] depth reads served by the TMA tile: %u, by the LDG fallback: %u\n", e[2], e[3]);
	}
#endif
	cudaFree(v->d_err); cudaFree(v->d_palette); cudaFree(v->d_lut);
	cudaFree(v->d_probs); cudaFree(v->d_box); cudaFree(v->d_t); cudaFree(v->d_flags); cudaFree(v->d_bgr);
	cudaFree(v->d_label); cudaFree(v->d_keys); cudaFree(v->d_hits); cudaFree(v->d_fold);
	for (int i = 0; i < 2; i++) {
		if (v->pin[i].buf) cudaFreeHost(v->pin[i].buf);
		if (v->pin[i].free_ev) cudaEventDestroy(v->pin[i].free_ev);
	}
	if (v->h_stats) cudaFreeHost(v->h_stats);
	if (v->h_err) cudaFreeHost(v->h_err);
	if (v->h_fold) cudaFreeHost(v->h_fold);
	cudaFree(v->d_merge);
	if (v->h_merge) cudaFreeHost(v->h_merge);
	if (v->ev_merge) cudaEventDestroy(v->ev_merge);
	if (v->ev_t0) cudaEventDestroy(v->ev_t0);
	if (v->ev_t1) cudaEventDestroy(v->ev_t1);
	for (int i = 0; i < sfm_volume::kRing; i++) {
		if (v->ev_k0[i]) cudaEventDestroy(v->ev_k0[i]);
		if (v->ev_km[i]) cudaEventDestroy(v->ev_km[i]);
		if (v->ev_kb[i]) cudaEventDestroy(v->ev_kb[i]);
		if (v->ev_k1[i]) cudaEventDestroy(v->ev_k1[i]);
	}
	if (v->own_stream && v->stream) cudaStreamDestroy(v->stream);
	cudaGetLastError();
	delete v;
}

int sfm_set_bounds(sfm_volume *v, const float *vol_start3, const float *vol_end3, const float *voxel3, float miu) {
	if (!v || !vol_start3 || !vol_end3 || !voxel3) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	v->g.sx = vol_start3[0]; v->g.sy = vol_start3[1]; v->g.sz = vol_start3[2];
	v->g.ex = vol_end3[0]; v->g.ey = vol_end3[1]; v->g.ez = vol_end3[2];
	v->g.vx = voxel3[0]; v->g.vy = voxel3[1]; v->g.vz = voxel3[2];
	v->g.miu = miu;
	v->g.fastdiv = 0;
	for (int a = 0; a < 3; a++) {  // see div_by() in k_raymarch.cuh
		uint32_t bits;
		memcpy(&bits, &voxel3[a], 4);
		const float av = fabsf(voxel3[a]);
		if ((bits & 0x7fffffu) != 0x7fffffu && av > 1e-18f && av < 1e18f) v->g.fastdiv |= 1 << a;
	}
	v->init = true;
	for (int i = 0; i < 16; i++) v->init_extr_inv[i] = (i % 5 == 0) ? 1.f : 0.f;
	return reset_planes(v);
}

int sfm_place_volume(const uint16_t *depth, int width, int height, const float *Kinv16, float mean_depth, const int32_t *dims3,
	float trunc_voxels, float *vol_start3, float *vol_end3, float *voxel3, float *miu)
{
	if (!depth || !Kinv16 || !dims3 || !vol_start3 || !vol_end3 || !voxel3 || !miu || width <= 0 || height <= 0)
		return fail(SFM_ERR_INVALID, "null argument");
	// tsdf.cu:180-182: bounding rectangle of depth != 0 (the saturating cast to u8 keeps nonzero nonzero)
	int x0 = width, y0 = height, x1 = -1, y1 = -1;
	for (int y = 0; y < height; y++)
		for (int x = 0; x < width; x++)
			if (depth[(size_t)y * width + x]) { x0 = std::min(x0, x); x1 = std::max(x1, x); y0 = std::min(y0, y); y1 = std::max(y1, y); }
	if (x1 < 0) return fail(SFM_ERR_INVALID, "first depth frame has no valid pixel");
	// cv::Rect: tl = (x0,y0), br = (x0+width, y0+height) = (x1+1, y1+1)
	const float tlp[4] = {(float)x0, (float)y0, 1.f, 1.f}, brp[4] = {(float)(x1 + 1), (float)(y1 + 1), 1.f, 1.f};
	float tl[4], br[4];
	for (int r = 0; r < 4; r++) {  // tsdf.cu:185-188  Kinv(4x4) * p (cv::Mat float gemm: double accumulation), then * mean_depth
		double a = 0, b = 0;
		for (int k = 0; k < 4; k++) { a += (double)Kinv16[r * 4 + k] * tlp[k]; b += (double)Kinv16[r * 4 + k] * brp[k]; }
		tl[r] = (float)a * mean_depth;
		br[r] = (float)b * mean_depth;
	}
	// tsdf.cu:193: sqrt(pow(dx,2)+pow(dy,2))/2 evaluated in double, stored to float
	const float half_side = (float)(sqrt(pow((double)(tl[0] - br[0]), 2) + pow((double)(tl[1] - br[1]), 2)) / 2);
	for (int a = 0; a < 3; a++) {
		const float center = (tl[a] + br[a]) / 2;      // tsdf.cu:194
		vol_start3[a] = center - half_side;            // tsdf.cu:195
		vol_end3[a] = center + half_side;              // tsdf.cu:196
		voxel3[a] = (vol_end3[a] - vol_start3[a]) / (float)(dims3[a] - 1);  // tsdf.cu:197 (cv::divide in f32)
	}
	*miu = trunc_voxels * voxel3[0];                   // tsdf.cu:199
	return SFM_OK;
}

int sfm_init_from_frame(sfm_volume *v, const uint16_t *depth, const float *extrinsic16, float mean_depth) {
	if (!v || !depth || !extrinsic16) return fail(SFM_ERR_INVALID, "null argument");
	float start[3], end[3], voxel[3], miu = 0.f;
	const int32_t dims[3] = {v->g.Dx, v->g.Dy, v->g.Dz};
	int rc = sfm_place_volume(depth, v->W, v->H, v->Kinv, mean_depth, dims, v->desc.trunc_voxels, start, end, voxel, &miu);
	if (rc) return rc;
	rc = sfm_set_bounds(v, start, end, voxel, miu);
	if (rc) return rc;
	v->mean_depth = mean_depth;                         // tsdf.cu:191
	if (!mat4_inv(extrinsic16, v->init_extr_inv)) return fail(SFM_ERR_INVALID, "singular first extrinsic");  // tsdf.cu:177
	return SFM_OK;
}

int sfm_integrate_dev(sfm_volume *v, const void *d_depth, const void *d_color, const void *d_mask, const float *E16) {
	if (!v || !d_depth || !d_color || !E16 || (v->bins > 0 && !d_mask)) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	return integrate_device(v, d_depth, d_color, d_mask, E16);
}

int sfm_integrate_dev_ready(sfm_volume *v, const void *d_depth, const void *d_color, const void *d_mask, const float *E16, void *ready_event) {
	if (!v || !d_depth || !d_color || !E16 || (v->bins > 0 && !d_mask)) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	int rc = enqueue_prepare(v, d_depth, d_color, d_mask, E16, ready_event ? (cudaEvent_t)ready_event : kReadyNow);
	if (rc) return rc;
	// K1b reads the colour / label images on the main stream: it must see them too
	if (ready_event) CU(cudaStreamWaitEvent(v->stream, (cudaEvent_t)ready_event, 0));
	return enqueue_update(v, d_depth, d_color, d_mask, E16);
}

int sfm_integrate_raw(sfm_volume *v, const uint16_t *depth, const uint8_t *color, const uint8_t *mask, const float *E16) {
	if (!v || !depth || !color || !E16 || (v->bins > 0 && !mask)) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	int rc = upload_frame(v, depth, color, v->bins > 0 ? mask : nullptr);
	if (rc) return rc;
	rc = integrate_device(v, v->d_depth, v->d_rgb, v->d_mask, E16, v->ev_uploaded[v->frame_cur]);
	const std::string keep = g_err;
	const int rc2 = release_frame(v);
	if (rc) { g_err = keep; return rc; }
	return rc2;
}

int sfm_overlap_tables(sfm_volume *v, const float *E16, const uint8_t *mask, double *A, uint32_t *C) {
	if (!v || !E16 || !mask || !A || !C) return fail(SFM_ERR_INVALID, "null argument");
	if (v->bins <= 0) return fail(SFM_ERR_INVALID, "labels are off (bins == 0)");
	if (!v->init || v->n_obs == 0) return fail(SFM_ERR_INVALID, "no observation integrated yet (tsdf.cu:426)");
	CU(cudaSetDevice(v->desc.device));
	int rc = require_full_volume(v, "sfm_overlap_tables");
	if (rc) return rc;
	const size_t npx = (size_t)v->W * v->H;
	const int mx = max_label_host(mask, npx);
	if (mx >= v->bins) return fail(SFM_ERR_INVALID, "mask carries a label >= bins");
	rc = upload_frame(v, nullptr, nullptr, mask);
	if (rc) return rc;
	rc = run_fold(v, E16, v->d_mask);
	{
		const std::string keep = g_err;
		const int rc2 = release_frame(v);
		if (rc) { g_err = keep; return rc; }
		if (rc2) return rc2;
	}
	combine_tables(v, mx + 1, A, C);
	return SFM_OK;
}

int sfm_merge_decide(sfm_volume *v, const double *A, const uint32_t *C, uint8_t *mask_inout, sfm_merge_report *report) {
	if (!v || !A || !C || !mask_inout) return fail(SFM_ERR_INVALID, "null argument");
	const size_t npx = (size_t)v->W * v->H;
	const int mx = max_label_host(mask_inout, npx);
	if (mx >= v->bins) return fail(SFM_ERR_INVALID, "mask carries a label >= bins");
	std::vector<unsigned> first(256, 0xffffffffu);
	for (size_t i = 0; i < npx; i++) if (first[mask_inout[i]] == 0xffffffffu) first[mask_inout[i]] = (unsigned)i;
	sfm_merge_report rep;
	decide(v, A, C, mx + 1, first.data(), &v->num_objs, &rep);
	for (size_t i = 0; i < npx; i++) mask_inout[i] = (uint8_t)rep.assign[mask_inout[i]];
	v->last_merge = rep;
	if (report) *report = rep;
	return SFM_OK;
}

int sfm_set_num_objs(sfm_volume *v, int num_objs) {
	if (!v || num_objs < 0) return fail(SFM_ERR_INVALID, "bad argument");
	v->num_objs = num_objs;
	return SFM_OK;
}

int sfm_last_merge(sfm_volume *v, sfm_merge_report *report) {
	if (!v || !report) return fail(SFM_ERR_INVALID, "null argument");
	*report = v->last_merge;
	return SFM_OK;
}

namespace {
// body of sfm_fuse_frame after the upload; the caller releases the frame buffer whatever this returns
int fuse_uploaded_frame(sfm_volume *v, uint8_t *mask_inout, int mx, const float *E16) {
	const size_t npx = (size_t)v->W * v->H;
	const cudaEvent_t up = v->ev_uploaded[v->frame_cur];
	if (v->bins <= 0 || v->n_obs == 0) {
		if (v->bins > 0) v->num_objs = mx + 1;  // tsdf.cu:464-467
		return integrate_device(v, v->d_depth, v->d_rgb, v->d_mask, E16, up);
	}
	// tsdf.cu:426-461: back-project, fold, decide, relabel
	int rc = require_full_volume(v, "sfm_fuse_frame (merge)");
	if (rc) return rc;
	// everything up to the integration is enqueued without a host round trip: march, fold, decision
	// and relabel on the device; the host only waits for the 2 KB report to relabel ITS copy of the
	// mask, which overlaps the integrate kernels
	// (K0 + K1a of this frame run on prep_stream meanwhile: they only need the uploaded images; the
	// relabel waits for them because K0 range-checks the incoming labels)
	rc = enqueue_prepare(v, v->d_depth, v->d_rgb, v->d_mask, E16, up);
	if (rc) return rc;
	rc = enqueue_march_fold(v, E16, v->d_mask);
	if (rc) return rc;
	CU(cudaStreamWaitEvent(v->stream, v->ctx[v->n_integrate % sfm_volume::kCtx].ev_ready, 0));
	rc = enqueue_device_decision(v, v->d_fold, v->d_mask);
	if (rc) return rc;
	// the decision may overflow the bins (a fresh id >= bins: the reference corrupts its histogram there,
	// tsdf.cu:61,383).  Relabel and update are gated on the device by that flag, so a failing frame leaves
	// the volume, n_obs and num_objs exactly as they were and the caller may retry or skip it.
	const uint32_t n_obs_before = v->n_obs;
	rc = enqueue_update(v, v->d_depth, v->d_rgb, v->d_mask, E16, &((const MergeOut *)v->d_merge)->overflow);
	if (rc) return rc;
	uint8_t lut[256];
	rc = finish_device_decision(v, lut);
	if (rc) {
		v->n_obs = n_obs_before;
		return rc;
	}
	for (size_t i = 0; i < npx; i++) mask_inout[i] = lut[mask_inout[i]];  // tsdf.cu:372-389 does it in place
	return SFM_OK;
}
}  // namespace

int sfm_fuse_frame(sfm_volume *v, const uint16_t *depth, const uint8_t *color, uint8_t *mask_inout, const float *E16) {
	if (!v || !depth || !color || !E16 || (v->bins > 0 && !mask_inout)) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	int mx = 0;
	if (v->bins > 0) {
		mx = max_label_host(mask_inout, (size_t)v->W * v->H);
		if (mx >= v->bins) return fail(SFM_ERR_INVALID, "mask carries a label >= bins");
	}
	int rc = upload_frame(v, depth, color, v->bins > 0 ? mask_inout : nullptr);
	if (rc) return rc;
	rc = fuse_uploaded_frame(v, mask_inout, mx, E16);
	const std::string keep = g_err;
	const int rc2 = release_frame(v);  // on every path: the buffer's consumed event must follow the work enqueued so far
	if (rc) { g_err = keep; return rc; }
	return rc2;
}

int sfm_parse_frame(sfm_volume *v, const uint16_t *depth, const uint8_t *color, uint8_t *mask_inout,
	const float *extrinsic16, float mean_depth)
{
	if (!v || !depth || !extrinsic16) return fail(SFM_ERR_INVALID, "null argument");
	if (!v->init) return sfm_init_from_frame(v, depth, extrinsic16, mean_depth);  // tsdf.cu:173-214
	float e2i[16];
	mat4_mul(extrinsic16, v->init_extr_inv, e2i);  // tsdf.cu:217
	return sfm_fuse_frame(v, depth, color, mask_inout, e2i);
}

int sfm_backproject(sfm_volume *v, const float *E16, float *probs, uint8_t *box_mask, float *t_out, uint8_t *flags_out) {
	if (!v || !E16 || !probs || !box_mask) return fail(SFM_ERR_INVALID, "null argument");
	if (v->bins <= 0) return fail(SFM_ERR_INVALID, "labels are off (bins == 0)");
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	CU(cudaSetDevice(v->desc.device));
	int rc = require_full_volume(v, "sfm_backproject");
	if (rc) return rc;
	const size_t npx = (size_t)v->W * v->H;
	rc = ensure_ray_buffers(v, npx, true);
	if (rc) return rc;
	CU(cudaMemsetAsync(v->d_probs, 0, npx * v->bins * 4, v->stream));  // tsdf.cu:428-429
	CU(cudaMemsetAsync(v->d_box, 0, npx * v->bins, v->stream));
	rc = launch_march(v, make_ray_vol(v), make_backproj_cam(v, E16), v->d_hits, v->d_flags, 0, v->H);
	if (rc) return rc;
	probs_kernel<<<(int)((npx + 127) / 128), 128, 0, v->stream>>>(make_ray_vol(v), (int)npx, v->d_hits, v->desc.presence_thresh,
		v->d_probs, v->d_box, v->d_t, v->d_flags);
	LAUNCH_CHECK(v);
	CU(cudaMemcpyAsync(probs, v->d_probs, npx * v->bins * 4, cudaMemcpyDeviceToHost, v->stream));  // tsdf.cu:457-458
	CU(cudaMemcpyAsync(box_mask, v->d_box, npx * v->bins, cudaMemcpyDeviceToHost, v->stream));
	if (t_out) CU(cudaMemcpyAsync(t_out, v->d_t, npx * 4, cudaMemcpyDeviceToHost, v->stream));
	if (flags_out) CU(cudaMemcpyAsync(flags_out, v->d_flags, npx, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	return SFM_OK;
}

int sfm_raycast_keys_dev(sfm_volume *v, const float *s2w16, const float *c3, int w, int h, void *d_keys) {
	if (!v || !s2w16 || !c3 || !d_keys || w <= 0 || h <= 0) return fail(SFM_ERR_INVALID, "bad argument");
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	CU(cudaSetDevice(v->desc.device));
	int rc = require_full_volume(v, "sfm_raycast_keys_dev");
	if (rc) return rc;
	rc = ensure_ray_buffers(v, (size_t)w * h, false);
	if (rc) return rc;
	rc = launch_march(v, make_ray_vol(v), make_show_cam(s2w16, c3, w, h), v->d_hits, nullptr, 0, h);
	if (rc) return rc;
	shade_kernel<8><<<(w * h + 127) / 128, 128, 0, v->stream>>>(make_ray_vol(v), w * h, v->d_hits, v->d_palette,
		nullptr, nullptr, nullptr, (unsigned long long *)d_keys, nullptr, 0, 0, 0, 0);
	LAUNCH_CHECK(v);
	return SFM_OK;
}

int sfm_raycast(sfm_volume *v, const float *s2w16, const float *c3, int w, int h, uint8_t *bgr, float *t_opt, uint8_t *label_opt) {
	if (!v || !s2w16 || !c3 || !bgr || w <= 0 || h <= 0 || w > 65535 || h > 65535) return fail(SFM_ERR_INVALID, "bad argument");
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	CU(cudaSetDevice(v->desc.device));
	int rc = require_full_volume(v, "sfm_raycast");
	if (rc) return rc;
	const size_t npx = (size_t)w * h;
	rc = ensure_ray_buffers(v, npx, false);
	if (rc) return rc;
	rc = launch_march(v, make_ray_vol(v), make_show_cam(s2w16, c3, w, h), v->d_hits, v->d_flags, 0, h);
	if (rc) return rc;
	shade_kernel<8><<<(int)((npx + 127) / 128), 128, 0, v->stream>>>(make_ray_vol(v), (int)npx, v->d_hits, v->d_palette,
		v->d_bgr, v->d_t, v->d_label, nullptr, v->d_flags, 0, 0, 0, 0);
	LAUNCH_CHECK(v);
	CU(cudaMemcpyAsync(bgr, v->d_bgr, npx * 3, cudaMemcpyDeviceToHost, v->stream));  // viewer.cu:167
	if (t_opt) CU(cudaMemcpyAsync(t_opt, v->d_t, npx * 4, cudaMemcpyDeviceToHost, v->stream));
	if (label_opt) CU(cudaMemcpyAsync(label_opt, v->d_label, npx, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	return SFM_OK;
}

int sfm_raycast_color(sfm_volume *v, const float *s2w16, const float *c3, int w, int h, uint8_t *bgr, float *t_opt, float *xyzt_opt) {
	if (!v || !s2w16 || !c3 || !bgr || w <= 0 || h <= 0 || w > 65535 || h > 65535) return fail(SFM_ERR_INVALID, "bad argument");
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	CU(cudaSetDevice(v->desc.device));
	int rc = require_full_volume(v, "sfm_raycast_color");
	if (rc) return rc;
	const size_t npx = (size_t)w * h;
	rc = ensure_ray_buffers(v, npx, false);
	if (rc) return rc;
	rc = launch_march(v, make_ray_vol(v), make_show_cam(s2w16, c3, w, h), v->d_hits, v->d_flags, 0, h);
	if (rc) return rc;
	shade_color_kernel<<<(int)((npx + 127) / 128), 128, 0, v->stream>>>(make_ray_vol(v), v->planes.color, (int)npx, v->d_hits, v->d_bgr, v->d_t);
	LAUNCH_CHECK(v);
	CU(cudaMemcpyAsync(bgr, v->d_bgr, npx * 3, cudaMemcpyDeviceToHost, v->stream));
	if (t_opt) CU(cudaMemcpyAsync(t_opt, v->d_t, npx * 4, cudaMemcpyDeviceToHost, v->stream));
	if (xyzt_opt) CU(cudaMemcpyAsync(xyzt_opt, v->d_hits, npx * 16, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	return SFM_OK;
}

int sfm_extract_surface(sfm_volume *v, uint32_t max_points, float *xyz, uint8_t *bgr, uint8_t *label, uint32_t *count) {
	if (!v || !count || (max_points && (!xyz || !bgr || !label))) return fail(SFM_ERR_INVALID, "null argument");
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	CU(cudaSetDevice(v->desc.device));
	SurfaceOut out{};
	out.max_points = max_points;
	uint8_t *d_buf = nullptr;
	const size_t bytes = 16 + (size_t)max_points * 16;
	CU(cudaMalloc(&d_buf, bytes));
	out.count = (unsigned *)d_buf;
	out.xyz = (float *)(d_buf + 16);
	out.bgr = d_buf + 16 + (size_t)max_points * 12;
	out.label = d_buf + 16 + (size_t)max_points * 15;
	cudaError_t e = cudaMemsetAsync(d_buf, 0, 16, v->stream);
	if (e == cudaSuccess) {
		extract_surface_kernel<<<v->num_sms * 8, 256, 0, v->stream>>>(v->g, v->planes.sdf, v->planes.wt, v->planes.color, v->planes.hist, v->bins, out);
		v->launches++;
		e = cudaGetLastError();
	}
	unsigned n = 0;
	if (e == cudaSuccess) e = cudaMemcpyAsync(&n, out.count, 4, cudaMemcpyDeviceToHost, v->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(v->stream);
	const size_t m = std::min<size_t>(n, max_points);
	if (e == cudaSuccess && m) e = cudaMemcpy(xyz, out.xyz, m * 12, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess && m) e = cudaMemcpy(bgr, out.bgr, m * 3, cudaMemcpyDeviceToHost);
	if (e == cudaSuccess && m) e = cudaMemcpy(label, out.label, m, cudaMemcpyDeviceToHost);
	cudaFree(d_buf);
	if (e != cudaSuccess) return fail(SFM_ERR_CUDA, cudaGetErrorString(e));
	*count = n;
	return SFM_OK;
}

int sfm_ray_flags(sfm_volume *v, uint8_t *flags, size_t n) {
	if (!v || !flags) return fail(SFM_ERR_INVALID, "null argument");
	if (n > v->ray_px) return fail(SFM_ERR_INVALID, "no ray-march of that size has run");
	CU(cudaSetDevice(v->desc.device));
	CU(cudaMemcpyAsync(flags, v->d_flags, n, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	return SFM_OK;
}

int sfm_shard_halo(const float *voxel3) {
	if (!voxel3 || !(voxel3[2] > 0.f)) return 3;
	return (int)ceilf(voxel3[0] / voxel3[2]) + 2;
}

static int shard_stage(sfm_volume *v, int stage, const RayCam &cam, const void *d_ev1, const void *d_ev2, void *d_out, float4 *hits_out) {
	if ((stage >= 2 && !d_ev1) || (stage == 3 && !d_ev2)) return fail(SFM_ERR_INVALID, "missing reduced events of the previous stage");
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	CU(cudaSetDevice(v->desc.device));
	const float vox[3] = {v->g.vx, v->g.vy, v->g.vz};
	const int halo = sfm_shard_halo(vox);
	const int lo_have = v->g.own_z0 - v->g.z0, hi_have = (v->g.z0 + v->g.nz) - (v->g.own_z0 + v->g.own_nz);
	if ((v->g.own_z0 > 0 && lo_have < halo) || (v->g.own_z0 + v->g.own_nz < v->g.Dz && hi_have < halo))
		return fail(SFM_ERR_INVALID, "sharded ray-cast needs a halo of " + std::to_string(halo) + " stored planes around the owned range");
	const RayVol V = make_ray_vol(v);
	const int blocks = ray_blocks(cam.W, cam.H);
	if (stage == 1) shard_stage1_kernel<<<blocks, 128, 0, v->stream>>>(V, cam, (unsigned long long *)d_out);
	else if (stage == 2) shard_stage2_kernel<<<blocks, 128, 0, v->stream>>>(V, cam, (const unsigned long long *)d_ev1, (unsigned long long *)d_out);
	else shard_stage3_kernel<<<blocks, 128, 0, v->stream>>>(V, cam, (const unsigned long long *)d_ev1, (const unsigned long long *)d_ev2, (unsigned long long *)d_out, hits_out);
	LAUNCH_CHECK(v);
	return SFM_OK;
}

int sfm_shard_raycast_stage(sfm_volume *v, int stage, const float *s2w16, const float *c3, int w, int h,
	const void *d_ev1, const void *d_ev2, void *d_out)
{
	if (!v || !s2w16 || !c3 || !d_out || w <= 0 || h <= 0 || stage < 1 || stage > 3) return fail(SFM_ERR_INVALID, "bad argument");
	return shard_stage(v, stage, make_show_cam(s2w16, c3, w, h), d_ev1, d_ev2, d_out, nullptr);
}

/* ---- duplicate-instance merge over z-slabs (back_proj_kernel + filter_overlaps, tsdf.cu:426-461) ---- */

int sfm_shard_backproj_stage(sfm_volume *v, int stage, const float *E16, const void *d_ev1, const void *d_ev2, void *d_out) {
	if (!v || !E16 || !d_out || stage < 1 || stage > 3) return fail(SFM_ERR_INVALID, "bad argument");
	CU(cudaSetDevice(v->desc.device));
	int rc = ensure_ray_buffers(v, (size_t)v->W * v->H, false);
	if (rc) return rc;
	return shard_stage(v, stage, make_backproj_cam(v, E16), d_ev1, d_ev2, d_out, stage == 3 ? v->d_hits : nullptr);
}

/* first labelled frame of a sharded volume: no merge yet, num_objs = max(mask) + 1 (tsdf.cu:464-467) */
int sfm_shard_first_frame(sfm_volume *v, const void *d_mask) {
	if (!v || !d_mask) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	const size_t npx = (size_t)v->W * v->H;
	std::vector<uint8_t> m(npx);
	CU(cudaMemcpyAsync(m.data(), d_mask, npx, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	const int mx = max_label_host(m.data(), npx);
	if (v->bins > 0 && mx >= v->bins) return fail(SFM_ERR_INVALID, "mask carries a label >= bins");
	v->num_objs = mx + 1;
	return SFM_OK;
}

int sfm_fold_table_bytes(int bins, size_t *bytes_i64, size_t *bytes_total) {
	if (bins <= 0 || bins > kMaxBins || !bytes_i64 || !bytes_total) return fail(SFM_ERR_INVALID, "bad argument");
	const FoldLayout fl = fold_layout(bins);
	*bytes_i64 = fl.oBm;
	*bytes_total = fl.total;
	return SFM_OK;
}

int sfm_shard_fold(sfm_volume *v, const void *d_mask, const void *d_keys_global, int do_counts, void *d_tables) {
	if (!v || !d_mask || !d_keys_global || !d_tables) return fail(SFM_ERR_INVALID, "null argument");
	if (v->bins <= 0) return fail(SFM_ERR_INVALID, "labels are off (bins == 0)");
	if (!v->d_hits) return fail(SFM_ERR_INVALID, "run sfm_shard_backproj_stage(3) first");
	CU(cudaSetDevice(v->desc.device));
	return launch_fold(v, (uint8_t *)d_tables, (const uint8_t *)d_mask, (const unsigned long long *)d_keys_global, do_counts ? 1 : 0);
}

int sfm_shard_merge_finish(sfm_volume *v, const void *d_tables_reduced, void *d_mask_inout, uint8_t *lut256, sfm_merge_report *report) {
	if (!v || !d_tables_reduced || !d_mask_inout) return fail(SFM_ERR_INVALID, "null argument");
	if (v->bins <= 0 || !v->d_merge) return fail(SFM_ERR_INVALID, "labels are off (bins == 0)");
	CU(cudaSetDevice(v->desc.device));
	int rc = enqueue_device_decision(v, (uint8_t *)d_tables_reduced, (uint8_t *)d_mask_inout);
	if (rc) return rc;
	rc = finish_device_decision(v, lut256);
	if (rc) return rc;
	if (report) *report = v->last_merge;
	return SFM_OK;
}

int sfm_keys_to_bgr(sfm_volume *v, const void *d_keys, int w, int h, uint8_t *bgr) {
	if (!v || !d_keys || !bgr || w <= 0 || h <= 0) return fail(SFM_ERR_INVALID, "bad argument");
	CU(cudaSetDevice(v->desc.device));
	const size_t npx = (size_t)w * h;
	int rc = ensure_ray_buffers(v, npx, false);
	if (rc) return rc;
	keys_to_bgr_kernel<<<(int)((npx + 255) / 256), 256, 0, v->stream>>>((const unsigned long long *)d_keys, v->d_palette,
		(int)npx, v->d_bgr);
	LAUNCH_CHECK(v);
	CU(cudaMemcpyAsync(bgr, v->d_bgr, npx * 3, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	return SFM_OK;
}

/* ---- ray-cast of a z-slab-sharded volume with a replicated SDF (viewer.cu:137-179 after the fusion is over) ---- */

int sfm_sdf_planes_dev(sfm_volume *v, int z0, int n, void *d_buf, int to_buffer) {
	if (!v || !d_buf || n <= 0) return fail(SFM_ERR_INVALID, "bad argument");
	if (z0 < v->g.z0 || z0 + n > v->g.z0 + v->g.nz) return fail(SFM_ERR_INVALID, "planes outside the range this handle stores");
	CU(cudaSetDevice(v->desc.device));
	sdf_planes_kernel<<<v->num_sms * 16, 256, 0, v->stream>>>(v->g, v->planes.sdf, z0, n, (float *)d_buf, to_buffer ? 1 : 0);
	LAUNCH_CHECK(v);
	if (!to_buffer) {  // the SDF changed under the skip map and the steady knowledge
		CU(cudaMemsetAsync(v->planes.occ, 1, v->occ_bytes, v->stream));
	}
	return SFM_OK;
}

int sfm_rebuild_skip_map(sfm_volume *v) {
	if (!v) return fail(SFM_ERR_INVALID, "null argument");
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	CU(cudaSetDevice(v->desc.device));
	CU(cudaMemsetAsync(v->planes.occ, 0, v->occ_bytes, v->stream));
	rebuild_skip_map_kernel<<<v->num_sms * 16, 128, 0, v->stream>>>(v->g, v->planes.sdf, v->planes.occ);
	LAUNCH_CHECK(v);
	return SFM_OK;
}

int sfm_raycast_band_dev(sfm_volume *v, const float *s2w16, const float *c3, int w, int h, int row0, int rows, void *d_hits) {
	if (!v || !s2w16 || !c3 || !d_hits || w <= 0 || h <= 0 || row0 < 0 || rows <= 0 || row0 + rows > h) return fail(SFM_ERR_INVALID, "bad argument");
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	CU(cudaSetDevice(v->desc.device));
	int rc = require_full_volume(v, "sfm_raycast_band_dev");
	if (rc) return rc;
	rc = launch_march(v, make_ray_vol(v), make_show_cam(s2w16, c3, w, h), (float4 *)d_hits, nullptr,
		row0, rows);
	if (rc) return rc;
	return SFM_OK;
}

int sfm_part_rows(int h, int n_parts) { return (h > 0 && n_parts > 0) ? 4 * (((h + 3) / 4 + n_parts - 1) / n_parts) : 0; }

int sfm_label_hits_parts_dev(sfm_volume *v, const void *d_hits, int w, int h, int n_parts, void *d_keys) {
	if (!v || !d_hits || !d_keys || w <= 0 || h <= 0 || n_parts < 0) return fail(SFM_ERR_INVALID, "bad argument");
	if (v->bins <= 0) return fail(SFM_ERR_INVALID, "labels are off (bins == 0)");
	CU(cudaSetDevice(v->desc.device));
	const int npx = w * h;
	const int part_tile_rows = n_parts > 0 ? ((h + 3) / 4 + n_parts - 1) / n_parts : 0;
	// a slab that owns less than a third of the planes labels few, clustered hits: 32 lanes per hit there (shade_kernel)
	if (v->g.own_nz * 3 < v->g.Dz)
		shade_kernel<32><<<(npx + 127) / 128, 128, 0, v->stream>>>(make_ray_vol(v), npx, (const float4 *)d_hits, v->d_palette,
			nullptr, nullptr, nullptr, (unsigned long long *)d_keys, nullptr, 1, w, n_parts, part_tile_rows);
	else
		shade_kernel<8><<<(npx + 127) / 128, 128, 0, v->stream>>>(make_ray_vol(v), npx, (const float4 *)d_hits, v->d_palette,
			nullptr, nullptr, nullptr, (unsigned long long *)d_keys, nullptr, 1, w, n_parts, part_tile_rows);
	LAUNCH_CHECK(v);
	return SFM_OK;
}

int sfm_label_hits_dev(sfm_volume *v, const void *d_hits, int w, int h, void *d_keys) {
	return sfm_label_hits_parts_dev(v, d_hits, w, h, 0, d_keys);
}

int sfm_raycast_part_dev(sfm_volume *v, const float *s2w16, const float *c3, int w, int h, int part, int n_parts, void *d_hits_part) {
	if (!v || !s2w16 || !c3 || !d_hits_part || w <= 0 || h <= 0 || n_parts <= 0 || part < 0 || part >= n_parts) return fail(SFM_ERR_INVALID, "bad argument");
	if (!v->init) return fail(SFM_ERR_INVALID, "volume bounds not set");
	CU(cudaSetDevice(v->desc.device));
	int rc = require_full_volume(v, "sfm_raycast_part_dev");
	if (rc) return rc;
	const int part_tile_rows = ((h + 3) / 4 + n_parts - 1) / n_parts;
	return launch_march(v, make_ray_vol(v), make_show_cam(s2w16, c3, w, h), (float4 *)d_hits_part, nullptr, part * 4, part_tile_rows * 4, n_parts, 1);
}

int sfm_ray_stats(sfm_volume *v, uint64_t *samples, uint64_t *hits) {
	if (!v || !samples || !hits) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	unsigned long long h2[2] = {0, 0};
	CU(cudaMemcpyAsync(h2, v->d_ray_stats, 16, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	*samples = h2[0] - v->ray_samples_seen;
	*hits = h2[1] - v->ray_hits_seen;
	v->ray_samples_seen = h2[0];
	v->ray_hits_seen = h2[1];
	return SFM_OK;
}

void sfm_orbit_camera(const float *Kinv16, float angle, float dist, float *s2w16, float *c3) {
	// viewer.cu:140-146
	const float rot[16] = {cosf(angle), 0, -sinf(angle), dist * sinf(angle), 0, 1, 0, 0,
		sinf(angle), 0, cosf(angle), dist - dist * cosf(angle), 0, 0, 0, 1};
	mat4_mul(rot, Kinv16, s2w16);
	c3[0] = (dist + 0.5f) * sinf(angle);
	c3[1] = 0.f;
	c3[2] = (dist + 0.5f) - (dist + 0.5f) * cosf(angle);
}

int sfm_show(sfm_volume *v, float angle, float dist, int w, int h, uint8_t *bgr) {
	if (!v) return fail(SFM_ERR_INVALID, "null argument");
	float s2w[16], c[3];
	sfm_orbit_camera(v->Kinv, angle, dist, s2w, c);
	return sfm_raycast(v, s2w, c, w, h, bgr, nullptr, nullptr);
}

int sfm_show_color(sfm_volume *v, float angle, float dist, int w, int h, uint8_t *bgr) {
	if (!v) return fail(SFM_ERR_INVALID, "null argument");
	float s2w[16], c[3];
	sfm_orbit_camera(v->Kinv, angle, dist, s2w, c);
	return sfm_raycast_color(v, s2w, c, w, h, bgr, nullptr, nullptr);
}

size_t sfm_plane_bytes(sfm_volume *v, int plane) { return v ? v->nvox * plane_elem_bytes(v, plane) : 0; }

namespace {
constexpr size_t kHistChunk = 256u << 20;  // staging bytes for histogram transfers (reference layout)

// reference-layout u32 histogram of voxels [v0, v1) into / from a device buffer
int hist_export(sfm_volume *v, size_t v0, size_t v1, uint32_t *d_ref) {
	hist_export_kernel<<<v->num_sms * 16, 256, 0, v->stream>>>(v->planes.hist, v->g.nz, v->g.ngz, v->bins, v0, v1, d_ref);
	LAUNCH_CHECK(v);
	return SFM_OK;
}
int hist_import(sfm_volume *v, size_t v0, size_t v1, const uint32_t *d_ref) {
	hist_import_kernel<<<v->num_sms * 16, 256, 0, v->stream>>>(v->planes.hist, v->g.nz, v->g.ngz, v->bins, v0, v1, d_ref, v->d_hist_max);
	LAUNCH_CHECK(v);
	return SFM_OK;
}
}  // namespace

void *sfm_plane_device_ptr(sfm_volume *v, int plane) {
	if (!v) return nullptr;
	if (plane != SFM_PLANE_HIST) return plane_ptr(v, plane);
	// the reference exposes tsdf_cnt_d (tsdf.cuh:26); the library's own plane is tiled and 16 bits wide, so this
	// returns a reference-layout SNAPSHOT (u32[v*bins + label]) refreshed by every call
	if (v->bins <= 0 || cudaSetDevice(v->desc.device) != cudaSuccess) return nullptr;
	if (!v->d_hist_ref && cudaMalloc(&v->d_hist_ref, v->nvox * 4 * (size_t)v->bins) != cudaSuccess) {
		cudaGetLastError();
		fail(SFM_ERR_NOMEM, "no memory for the reference-layout histogram snapshot");
		return nullptr;
	}
	if (hist_export(v, 0, v->nvox, v->d_hist_ref) != SFM_OK || cudaStreamSynchronize(v->stream) != cudaSuccess) return nullptr;
	return v->d_hist_ref;
}

int sfm_hist_export_dev(sfm_volume *v, void *d_dst_u32) {
	if (!v || !d_dst_u32) return fail(SFM_ERR_INVALID, "null argument");
	if (v->bins <= 0) return fail(SFM_ERR_INVALID, "labels are off (bins == 0)");
	CU(cudaSetDevice(v->desc.device));
	return hist_export(v, 0, v->nvox, (uint32_t *)d_dst_u32);
}

int sfm_download(sfm_volume *v, int plane, void *dst, size_t bytes) {
	if (!v || !dst) return fail(SFM_ERR_INVALID, "null argument");
	const size_t need = sfm_plane_bytes(v, plane);
	if (!need || bytes != need) return fail(SFM_ERR_INVALID, "plane / size mismatch");
	CU(cudaSetDevice(v->desc.device));
	if (plane != SFM_PLANE_HIST) {
		CU(cudaMemcpyAsync(dst, plane_ptr(v, plane), need, cudaMemcpyDeviceToHost, v->stream));
		return check_device_error(v);
	}
	// histogram: re-emit the reference layout and type (u32[v*bins + label]) chunk by chunk
	if (!v->d_hist_chunk) CU(cudaMalloc(&v->d_hist_chunk, kHistChunk));
	const size_t per_vox = 4 * (size_t)v->bins, vox_per_chunk = std::max<size_t>(1, kHistChunk / per_vox);
	for (size_t v0 = 0; v0 < v->nvox; v0 += vox_per_chunk) {
		const size_t v1 = std::min(v->nvox, v0 + vox_per_chunk);
		int rc = hist_export(v, v0, v1, v->d_hist_chunk);
		if (rc) return rc;
		CU(cudaMemcpyAsync((uint8_t *)dst + v0 * per_vox, v->d_hist_chunk, (v1 - v0) * per_vox, cudaMemcpyDeviceToHost, v->stream));
		CU(cudaStreamSynchronize(v->stream));
	}
	return check_device_error(v);
}

int sfm_upload(sfm_volume *v, int plane, const void *src, size_t bytes) {
	if (!v || !src) return fail(SFM_ERR_INVALID, "null argument");
	const size_t need = sfm_plane_bytes(v, plane);
	if (!need || bytes != need) return fail(SFM_ERR_INVALID, "plane / size mismatch");
	CU(cudaSetDevice(v->desc.device));
	if (plane != SFM_PLANE_HIST) {
		CU(cudaMemcpyAsync(plane_ptr(v, plane), src, need, cudaMemcpyHostToDevice, v->stream));
		if (plane == SFM_PLANE_SDF) CU(cudaMemsetAsync(v->planes.occ, 1, v->occ_bytes, v->stream));  // arbitrary SDF: no block may be skipped
		CU(cudaStreamSynchronize(v->stream));
		return SFM_OK;
	}
	if (!v->d_hist_chunk) CU(cudaMalloc(&v->d_hist_chunk, kHistChunk));
	CU(cudaMemsetAsync(v->d_hist_max, 0, 4, v->stream));
	const size_t per_vox = 4 * (size_t)v->bins, vox_per_chunk = std::max<size_t>(1, kHistChunk / per_vox);
	for (size_t v0 = 0; v0 < v->nvox; v0 += vox_per_chunk) {
		const size_t v1 = std::min(v->nvox, v0 + vox_per_chunk);
		CU(cudaMemcpyAsync(v->d_hist_chunk, (const uint8_t *)src + v0 * per_vox, (v1 - v0) * per_vox, cudaMemcpyHostToDevice, v->stream));
		int rc = hist_import(v, v0, v1, v->d_hist_chunk);
		if (rc) return rc;
		CU(cudaStreamSynchronize(v->stream));
	}
	unsigned mx = 0;
	CU(cudaMemcpy(&mx, v->d_hist_max, 4, cudaMemcpyDeviceToHost));
	v->hist_bound = std::min(mx, 65535u);
	if (mx > 65535u) return fail(SFM_ERR_INVALID, "uploaded histogram holds a count above 65535 (16-bit bins inside the library); it was clamped");
	return SFM_OK;
}

int sfm_get_info(sfm_volume *v, sfm_info *info) {
	if (!v || !info) return fail(SFM_ERR_INVALID, "null argument");
	memset(info, 0, sizeof(*info));
	info->dims[0] = v->g.Dx; info->dims[1] = v->g.Dy; info->dims[2] = v->g.Dz;
	info->bins = v->bins;
	info->width = v->W; info->height = v->H;
	info->slab_z0 = v->g.z0; info->slab_nz = v->g.nz;
	info->vol_start[0] = v->g.sx; info->vol_start[1] = v->g.sy; info->vol_start[2] = v->g.sz;
	info->vol_end[0] = v->g.ex; info->vol_end[1] = v->g.ey; info->vol_end[2] = v->g.ez;
	info->voxel[0] = v->g.vx; info->voxel[1] = v->g.vy; info->voxel[2] = v->g.vz;
	info->miu = v->g.miu;
	info->mean_depth = v->mean_depth;
	info->n_obs = v->n_obs;
	info->num_objs = v->num_objs;
	info->initialised = v->init ? 1 : 0;
	return SFM_OK;
}

int sfm_synchronize(sfm_volume *v) {
	if (!v) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	return check_device_error(v);
}

int sfm_wait_uploads(sfm_volume *v) {
	if (!v) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	CU(cudaStreamSynchronize(v->copy_stream));
	return SFM_OK;
}

int sfm_planes_written(sfm_volume *v) {
	if (!v) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	CU(cudaMemsetAsync(v->planes.occ, 1, v->occ_bytes, v->stream));  // arbitrary SDF: no block may be skipped
	return SFM_OK;
}

int sfm_set_stream(sfm_volume *v, void *cuda_stream) {
	if (!v) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	CU(cudaStreamSynchronize(v->stream));
	if (v->own_stream) cudaStreamDestroy(v->stream);
	v->stream = (cudaStream_t)cuda_stream;
	v->own_stream = false;
	return SFM_OK;
}

int sfm_timer_start(sfm_volume *v) {
	if (!v) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaEventRecord(v->ev_t0, v->stream));
	return SFM_OK;
}

int sfm_timer_stop(sfm_volume *v, float *ms) {
	if (!v || !ms) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaEventRecord(v->ev_t1, v->stream));
	CU(cudaEventSynchronize(v->ev_t1));
	CU(cudaEventElapsedTime(ms, v->ev_t0, v->ev_t1));
	return SFM_OK;
}

uint64_t sfm_launch_count(sfm_volume *v) { return v ? v->launches : 0; }

int sfm_integrate_times(sfm_volume *v, float *ms, int n) {
	if (!v || !ms || n < 0) return fail(SFM_ERR_INVALID, "bad argument");
	if ((uint64_t)n > v->n_integrate || n > sfm_volume::kRing) return fail(SFM_ERR_INVALID, "fewer integrate calls recorded than requested");
	for (int i = 0; i < n; i++) {
		const int slot = (int)((v->n_integrate - n + i) % sfm_volume::kRing);
		CU(cudaEventSynchronize(v->ev_k1[slot]));
		float a = 0.f, b = 0.f;  // K1a (prep_stream) + K1b (main stream): the two may overlap other frames' kernels
		CU(cudaEventElapsedTime(&a, v->ev_k0[slot], v->ev_km[slot]));
		CU(cudaEventElapsedTime(&b, v->ev_kb[slot], v->ev_k1[slot]));
		ms[i] = a + b;
	}
	return SFM_OK;
}

int sfm_last_integrate_ms(sfm_volume *v, float *ms) { return sfm_integrate_times(v, ms, 1); }

int sfm_integrate_times2(sfm_volume *v, float *ms_classify, float *ms_update, int n) {
	if (!v || !ms_classify || !ms_update || n < 0) return fail(SFM_ERR_INVALID, "bad argument");
	if ((uint64_t)n > v->n_integrate || n > sfm_volume::kRing) return fail(SFM_ERR_INVALID, "fewer integrate calls recorded than requested");
	for (int i = 0; i < n; i++) {
		const int slot = (int)((v->n_integrate - n + i) % sfm_volume::kRing);
		CU(cudaEventSynchronize(v->ev_k1[slot]));
		CU(cudaEventElapsedTime(ms_classify + i, v->ev_k0[slot], v->ev_km[slot]));
		CU(cudaEventElapsedTime(ms_update + i, v->ev_kb[slot], v->ev_k1[slot]));
	}
	return SFM_OK;
}

/* U = voxels whose weight was incremented, S = voxels whose colour/histogram was updated, summed
 * over the integrate calls since the previous sfm_frame_stats call. */
int sfm_frame_stats(sfm_volume *v, uint64_t *U, uint64_t *S) {
	if (!v || !U || !S) return fail(SFM_ERR_INVALID, "null argument");
	CU(cudaSetDevice(v->desc.device));
	CU(cudaMemcpyAsync(v->h_stats, v->d_stats, 2 * kStatSlots * 8, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaStreamSynchronize(v->stream));
	uint64_t u = 0, s = 0;
	for (int i = 0; i < kStatSlots; i++) { u += v->h_stats[i]; s += v->h_stats[kStatSlots + i]; }
	*U = u - v->stat_U_seen;
	*S = s - v->stat_S_seen;
	v->stat_U_seen = u;
	v->stat_S_seen = s;
	return SFM_OK;
}

/* Pipelined form: _begin enqueues the read-back of the counters after the work submitted so far and
 * returns a ticket; _end waits for that ticket only and returns the CUMULATIVE totals at that point,
 * so a caller can read step i-1's result while step i is already running. */
int sfm_stats_begin(sfm_volume *v, uint64_t *ticket) {
	if (!v || !ticket) return fail(SFM_ERR_INVALID, "null argument");
	const int slot = (int)(v->stat_tickets % sfm_volume::kStatRing);
	CU(cudaMemcpyAsync(v->h_stat_ring + (size_t)slot * 2 * kStatSlots, v->d_stats, 2 * kStatSlots * 8, cudaMemcpyDeviceToHost, v->stream));
	CU(cudaEventRecord(v->ev_stat[slot], v->stream));
	*ticket = v->stat_tickets++;
	return SFM_OK;
}

int sfm_stats_end(sfm_volume *v, uint64_t ticket, uint64_t *U_total, uint64_t *S_total) {
	if (!v || !U_total || !S_total) return fail(SFM_ERR_INVALID, "null argument");
	if (ticket >= v->stat_tickets || ticket + sfm_volume::kStatRing < v->stat_tickets) return fail(SFM_ERR_INVALID, "stale or unknown stats ticket");
	const int slot = (int)(ticket % sfm_volume::kStatRing);
	CU(cudaEventSynchronize(v->ev_stat[slot]));
	const unsigned long long *h = v->h_stat_ring + (size_t)slot * 2 * kStatSlots;
	uint64_t u = 0, s2 = 0;
	for (int i = 0; i < kStatSlots; i++) { u += h[i]; s2 += h[kStatSlots + i]; }
	*U_total = u;
	*S_total = s2;
	return SFM_OK;
}

/* test hook: mismatches between the invariant-divisor division used by the ray-marcher and the IEEE divide */
int sfm_debug_divcheck(float b, unsigned seed, int blocks, int per_thread, float amax, uint64_t *mismatches) {
	if (!mismatches) return fail(SFM_ERR_INVALID, "null argument");
	unsigned long long *d = nullptr;
	CU(cudaMalloc(&d, 8));
	CU(cudaMemset(d, 0, 8));
	uint32_t bits;
	memcpy(&bits, &b, 4);
	const bool ok = (bits & 0x7fffffu) != 0x7fffffu && fabsf(b) > 1e-18f && fabsf(b) < 1e18f;
	divcheck_kernel<<<blocks, 256>>>(b, ok, seed, per_thread, amax, d);
	unsigned long long h = 0;
	cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
	cudaFree(d);
	if (e != cudaSuccess) return fail(SFM_ERR_CUDA, cudaGetErrorString(e));
	*mismatches = h;
	return SFM_OK;
}

/* The two small host products parse_frame is built from (tsdf.cu:177 extrinsic.inv(), tsdf.cu:217 extrinsic *
 * init_extrinsic_inv): exported so that a caller that shards the volume itself (driver/kernel_mgpu.cpp) feeds every
 * slab the same extrinsic2init bits the single-volume sfm_parse_frame computes. */
int sfm_mat4_inv(const float *m16, float *out16) {
	if (!m16 || !out16) return fail(SFM_ERR_INVALID, "null argument");
	return mat4_inv(m16, out16) ? SFM_OK : fail(SFM_ERR_INVALID, "singular matrix");
}
void sfm_mat4_mul(const float *a16, const float *b16, float *out16) { mat4_mul(a16, b16, out16); }

float sfm_mean_depth(const uint16_t *depth, int n) {  // utils.cu:77-91
	double sum = 0;
	int total = 0;
	for (int i = 0; i < n; i++) {
		if (depth[i] == 0) continue;
		sum += depth[i] / 5000.;
		total++;
	}
	return static_cast<float>(sum / total);
}

void sfm_parse_extrinsic(const double *p, float *extrinsic16) {  // utils.cu:8-24
	const double ax = p[3], ay = p[4], az = p[5];
	const double n = sqrt(ax * ax + ay * ay + az * az);
	const double theta = 2 * atan2(n, p[6]);
	double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	if (n > 0 && theta != 0) {  // Rodrigues(theta * axis)
		const double rx = ax / n, ry = ay / n, rz = az / n, c = cos(theta), s = sin(theta), c1 = 1 - c;
		R[0] = c + c1 * rx * rx;      R[1] = c1 * rx * ry - s * rz; R[2] = c1 * rx * rz + s * ry;
		R[3] = c1 * ry * rx + s * rz; R[4] = c + c1 * ry * ry;      R[5] = c1 * ry * rz - s * rx;
		R[6] = c1 * rz * rx - s * ry; R[7] = c1 * rz * ry + s * rx; R[8] = c + c1 * rz * rz;
	}
	float m[16] = {(float)R[0], (float)R[1], (float)R[2], (float)p[0], (float)R[3], (float)R[4], (float)R[5], (float)p[1],
		(float)R[6], (float)R[7], (float)R[8], (float)p[2], 0, 0, 0, 1};
	if (!mat4_inv(m, extrinsic16)) memcpy(extrinsic16, m, sizeof(m));
}

/* Pose at `timestamp` between two trajectory entries {ts, tx, ty, tz, qx, qy, qz, qw}: translation lerp
 * (src/TSDF_Python/main.py:133-135) and quaternion slerp exactly as src/TSDF_Python/tsdf_utils.py:80-100
 * (normalise, flip to the same hemisphere, linear blend WITHOUT renormalisation above dot 0.9995). */
void sfm_interpolate_pose(const double *a8, const double *b8, double timestamp, double *pose7_out) {
	const double t = (timestamp - a8[0]) / (b8[0] - a8[0]);
	for (int k = 0; k < 3; k++) pose7_out[k] = (b8[1 + k] - a8[1 + k]) * t + a8[1 + k];
	double q1[4], q2[4];
	double n1 = 0, n2 = 0;
	for (int k = 0; k < 4; k++) { n1 += a8[4 + k] * a8[4 + k]; n2 += b8[4 + k] * b8[4 + k]; }
	n1 = sqrt(n1); n2 = sqrt(n2);
	double dot = 0;
	for (int k = 0; k < 4; k++) { q1[k] = a8[4 + k] / n1; q2[k] = b8[4 + k] / n2; }
	for (int k = 0; k < 4; k++) dot += q1[k] * q2[k];
	if (dot < 0) {
		for (int k = 0; k < 4; k++) q1[k] = -q1[k];
		dot = -dot;
	}
	if (dot > 0.9995) {
		for (int k = 0; k < 4; k++) pose7_out[3 + k] = q1[k] + t * (q2[k] - q1[k]);
		return;
	}
	dot = std::max(std::min(dot, 1.0), -1.0);
	const double theta0 = acos(dot), theta = theta0 * t;
	const double s1 = cos(theta) - dot * sin(theta) / sin(theta0), s2 = sin(theta) / sin(theta0);
	for (int k = 0; k < 4; k++) pose7_out[3 + k] = s1 * q1[k] + s2 * q2[k];
}

}  // extern "C"
