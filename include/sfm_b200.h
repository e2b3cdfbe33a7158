/* sfm_b200.h -- C-ABI of the B200-native TSDF fusion + ray-cast path.
 *
 * Drop-in boundary for the hot path of qq456cvb/SLAM-MaskRCNN's src/SfM_CUDA.  The reference
 * has no FFI: its boundary is two C++ classes with OpenCV types (`class TSDF`, tsdf.cuh:7-67;
 * `class Viewer`, viewer.cuh:4-17) called from kernel.cpp:37-111.  Every entry point below
 * names the reference interface it replaces.  include/sfm_b200.hpp re-creates `TSDF` / `Viewer`
 * with the reference's method names on top of this header so a kernel.cpp-style main compiles
 * against it.
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Return 0 on success, a negative
 * sfm_status otherwise (sfm_last_error() gives the text; the reference throws std::string after
 * cudaGetLastError, tsdf.cu:497-503, viewer.cu:169-175).  Host buffers are borrowed for the
 * duration of the call only.  One handle = one volume (or one z-slab of it) on one GPU, one
 * serialised command stream; not thread-safe per handle (neither is the reference).
 *
 * Layouts at the boundary are the reference's (tsdf.cu:55,59,61):
 *   voxel index  v = (x*Dy + y)*Dz + z            (z fastest)
 *   SDF f32[v], weight i32[v], colour u8[v*3+c], histogram u32[v*bins + label]
 *   depth u16[H][W] (metres*5000, 0 = invalid), colour u8[H][W][3] (BGR as loaded), mask u8[H][W]
 *   all 4x4 matrices row-major f32[16].
 */
#ifndef SFM_B200_H
#define SFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sfm_volume sfm_volume;

typedef enum {
	SFM_OK = 0,
	SFM_ERR_INVALID = -1,  /* bad argument / state (e.g. label >= bins, not initialised) */
	SFM_ERR_CUDA = -2,     /* a CUDA runtime call or kernel launch failed                  */
	SFM_ERR_NOMEM = -3,    /* device or host allocation failed                             */
	SFM_ERR_NODEVICE = -4  /* no usable sm_100 device (there is NO CPU fallback)           */
} sfm_status;

typedef enum {
	SFM_PLANE_SDF = 0,    /* f32, tsdf_diff_d  (tsdf.cuh:24) */
	SFM_PLANE_WEIGHT = 1, /* i32, tsdf_wt_d    (tsdf.cuh:27) */
	SFM_PLANE_COLOR = 2,  /* u8x3, tsdf_color_d (tsdf.cuh:25) */
	SFM_PLANE_HIST = 3    /* u32 x bins, tsdf_cnt_d (tsdf.cuh:26) */
} sfm_plane;

enum {
	SFM_FLAG_NO_CULL = 1,      /* visit every voxel (disables exact brick culling; same results) */
	SFM_FLAG_NO_TMA = 2,       /* read the per-frame depth tile grids through L1 instead of staging them into
	                              shared memory with a TMA bulk copy (cp.async.bulk)                      */
	SFM_FLAG_SYNC_EVERY_CALL = 4, /* cudaStreamSynchronize before returning from every call        */
	SFM_FLAG_GENERIC_K = 8,    /* evaluate K*c with all nine terms even when K has the pinhole zero pattern
	                              (the default drops the exact-zero terms; same results)                  */
	SFM_FLAG_DEBUG_ABLATE = 32, /* profiling builds only: honour the SFM_DEBUG_ABLATE environment variable
	                              (switches parts of the integrate kernel OFF; results are then wrong)    */
	SFM_FLAG_ASYNC_SOURCES = 64 /* pinned-host / device frame buffers passed to sfm_integrate_raw,
	                              sfm_fuse_frame and sfm_parse_frame may still be read by the GPU after the call
	                              returns: the caller keeps them unchanged until sfm_wait_uploads() or
	                              sfm_synchronize().  Without it every call returns only after its frame has
	                              been copied, like the reference's blocking cudaMemcpy (tsdf.cu:422-424)   */
};

/* Creation parameters.  Defaults (sfm_desc_default) are the reference's hard-coded values. */
typedef struct {
	int32_t dims[3];        /* vol_dim_, tsdf.cuh:52 (reference: 256^3 fixed)                 */
	int32_t bins;           /* MAX_OBJECTS, tsdf.cuh:4 (reference: 32); 0 = labels off        */
	int32_t width, height;  /* frame size (reference: taken from the first depth frame)        */
	float K[16];            /* intrinsic_, tsdf.cu:143-146: eye(4) with fx,fy,cx,cy            */
	float Kinv[16];         /* intrinsic_inv_, tsdf.cu:147; all-zero => library inverts K      */
	float prior_err_rate;   /* Configuration::prior_mrcnn_err_rate = 0.05, configuration.h:8   */
	float duplicate_thresh; /* Configuration::duplicate_thresh = 0.5, configuration.h:9; the
	                           reference declares and never reads it -- carried, unused        */
	float presence_thresh;  /* 0.3f, tsdf.cu:128                                               */
	float accept_factor;    /* 3 (x prior), tsdf.cu:349                                        */
	float depth_scale;      /* 5000.f, tsdf.cu:49                                              */
	float trunc_voxels;     /* 5 (miu = 5*voxel.x), tsdf.cu:199                                */
	float near_gate;        /* 0.99f, tsdf.cu:57                                               */
	int32_t device;         /* CUDA device ordinal                                             */
	int32_t slab_z0;        /* first global z plane stored by this handle                      */
	int32_t slab_nz;        /* number of z planes stored (0 => all of dims[2])                 */
	int32_t flags;          /* SFM_FLAG_*                                                      */
	int32_t own_z0;         /* sharded ray-cast: first global z plane this handle OWNS ...     */
	int32_t own_nz;         /* ... and how many (0 => all stored planes; stored minus owned = halo) */
	int32_t reserved[6];
} sfm_desc;

typedef struct {
	int32_t dims[3];
	int32_t bins;
	int32_t width, height;
	int32_t slab_z0, slab_nz;
	float vol_start[3], vol_end[3], voxel[3]; /* vol_start_/vol_end_/vol_res_, tsdf.cuh:54-55 */
	float miu;                                /* miu_, tsdf.cuh:51                             */
	float mean_depth;                         /* mean_depth_, tsdf.cuh:11                      */
	uint32_t n_obs;                           /* n_obs_, tsdf.cuh:46                           */
	int32_t num_objs;                         /* num_objs, tsdf.cuh:61                         */
	int32_t initialised;                      /* init_, tsdf.cuh:60                            */
	int32_t reserved[8];
} sfm_info;

/* Result of the last duplicate-instance merge (TSDF::filter_overlaps, tsdf.cu:304-416). */
typedef struct {
	int32_t max_obj_now;    /* max(mask)+1 of the incoming frame, tsdf.cu:305-307              */
	int32_t num_objs;       /* after the merge                                                 */
	int32_t assign[256];    /* current-frame label m -> global id written into the mask        */
	float best_prob[256];   /* max_j exp(A[m][j]/C[m][j]), tsdf.cu:340-348                     */
	float margin;           /* min distance of any decision from flipping (threshold or rival) */
} sfm_merge_report;

void sfm_desc_default(sfm_desc *d);
const char *sfm_last_error(void);
const char *sfm_version(void);

/* TSDF::TSDF(cv::Scalar intrinsics) (tsdf.cu:137-150) + init_cuda_vars (tsdf.cu:230-280). */
int sfm_create(const sfm_desc *desc, sfm_volume **out);
/* TSDF::~TSDF / free_cuda_vars (tsdf.cu:152-168, 282-302). */
void sfm_destroy(sfm_volume *v);

/* Volume placement of the first parse_frame call (tsdf.cu:173-212): bbox of depth != 0,
 * corners through Kinv scaled by mean_depth, cube of half the XY diagonal, voxel =
 * (end-start)/(dim-1), miu = trunc_voxels*voxel.x, SDF := miu, other planes := 0, n_obs := 0,
 * init_extrinsic_inv := extrinsic^-1.  The frame itself is NOT integrated (tsdf.cu:213). */
int sfm_init_from_frame(sfm_volume *v, const uint16_t *depth, const float *extrinsic16, float mean_depth);
/* The placement rule alone (tsdf.cu:180-199), host arithmetic only -- no handle, no device: bounds, voxel size and
 * truncation distance from the first depth frame.  sfm_init_from_frame is this + sfm_set_bounds. */
int sfm_place_volume(const uint16_t *depth, int width, int height, const float *Kinv16, float mean_depth,
	const int32_t *dims3, float trunc_voxels, float *vol_start3, float *vol_end3, float *voxel3, float *miu);
/* Explicit placement (what the parity tests use so both sides get identical bits). */
int sfm_set_bounds(sfm_volume *v, const float *vol_start3, const float *vol_end3, const float *voxel3, float miu);

/* TSDF::parse_frame (tsdf.cu:171-228): first call = sfm_init_from_frame; later calls =
 * extrinsic2init = extrinsic * init_extrinsic^-1, then launch_kernel (tsdf.cu:418-504):
 * back-project + duplicate-instance merge (relabels `mask` IN PLACE) when n_obs > 0, else
 * num_objs = max(mask)+1; integrate; n_obs++. */
int sfm_parse_frame(sfm_volume *v, const uint16_t *depth, const uint8_t *color, uint8_t *mask_inout,
	const float *extrinsic16, float mean_depth);
/* Same as the non-first branch of parse_frame with extrinsic2init given directly. */
int sfm_fuse_frame(sfm_volume *v, const uint16_t *depth, const uint8_t *color, uint8_t *mask_inout,
	const float *extrinsic2init16);

/* tsdf_kernel only (tsdf.cu:18-70, launch tsdf.cu:472-488): parity hook, host frame buffers.
 * Buffer lifetime (this call, sfm_fuse_frame, sfm_parse_frame, sfm_overlap_tables): the frame buffers are the
 * caller's and may be reused as soon as the call returns, as with the reference's blocking cudaMemcpy
 * (tsdf.cu:422-424, 470) -- pageable memory is staged into a pinned bounce buffer, pinned-host / device memory
 * is copied by the copy engine and the call waits for that copy.  A caller that sets SFM_FLAG_ASYNC_SOURCES
 * gets the call back before the copy has finished and must leave pinned / device buffers unchanged until
 * sfm_wait_uploads() (or sfm_synchronize()) returns. */
int sfm_integrate_raw(sfm_volume *v, const uint16_t *depth, const uint8_t *color, const uint8_t *mask,
	const float *extrinsic2init16);
/* Same with frame buffers already resident in device memory (multi-GPU: after the NCCL broadcast). */
int sfm_integrate_dev(sfm_volume *v, const void *d_depth, const void *d_color, const void *d_mask,
	const float *extrinsic2init16);
/* The same with an explicit statement of WHEN the frame images are valid: `ready_event` is a cudaEvent_t the
 * producer of the frame recorded (an upload, an ncclBroadcast), or NULL if the images are valid already.
 * sfm_integrate_dev has to assume they become valid in the order of the handle's stream, which serialises the
 * frame preparation (tile grids, brick classification: K0 + K1a, which never touch the volume) behind the
 * previous frame's update kernel; with this call it overlaps it on a second stream. */
int sfm_integrate_dev_ready(sfm_volume *v, const void *d_depth, const void *d_color, const void *d_mask,
	const float *extrinsic2init16, void *ready_event);

/* back_proj_kernel (tsdf.cu:72-135, launch 441-455), materialised: probs f32[H*W*bins],
 * box_mask u8[H*W*bins] on the HOST (parity hook; the fused path never builds these).
 * t_out f32[H*W] (refined hit t, 0 = no hit) and flags_out u8[H*W] (bit0: a trilinear tap was
 * clamped to the volume, i.e. the reference read out of bounds there) are optional. */
int sfm_backproject(sfm_volume *v, const float *extrinsic2init16, float *probs, uint8_t *box_mask,
	float *t_out, uint8_t *flags_out);
/* Fused back-project + overlap fold: the tables of filter_overlaps (tsdf.cu:309-334) without
 * materialising probs.  A f64[bins*bins] (sum of log terms), C u32[bins*bins]. */
int sfm_overlap_tables(sfm_volume *v, const float *extrinsic2init16, const uint8_t *mask,
	double *A, uint32_t *C);
/* The decision + relabel half of filter_overlaps (tsdf.cu:335-389) on given tables. */
int sfm_merge_decide(sfm_volume *v, const double *A, const uint32_t *C, uint8_t *mask_inout,
	sfm_merge_report *report);
int sfm_last_merge(sfm_volume *v, sfm_merge_report *report);
/* `num_objs` is a public member of the reference's TSDF (tsdf.cuh:58): the count of global instance ids handed out so
 * far, from which the merge numbers new instances (tsdf.cu:383).  sfm_get_info reads it; this sets it (parity hook:
 * the raw integrate entry points do not maintain it). */
int sfm_set_num_objs(sfm_volume *v, int num_objs);

/* Planes in the reference layout / dtypes (the getters get_tsdf_diff/color/cnt, tsdf.cu:506-516,
 * return stale host mirrors in the reference; these copy the live device planes).
 * For a slab handle the region is the slab's z range: [Dx][Dy][slab_nz]. */
int sfm_download(sfm_volume *v, int plane, void *dst, size_t bytes);
int sfm_upload(sfm_volume *v, int plane, const void *src, size_t bytes);
size_t sfm_plane_bytes(sfm_volume *v, int plane);
/* The reference exposes its device pointers as public members (tsdf.cuh:24-43).  A caller that WRITES the SDF
 * plane through this pointer must call sfm_planes_written() afterwards: the ray-marcher keeps a map of the
 * blocks that ever held a near-surface value and skips the others (sfm_upload does this itself). */
void *sfm_plane_device_ptr(sfm_volume *v, int plane);
int sfm_planes_written(sfm_volume *v);
/* Histogram at the boundary vs inside.  sfm_download / sfm_upload / sfm_plane_bytes speak the reference's
 * u32 hist[v*bins + label] (tsdf.cu:61, tsdf.cuh:26).  Inside, the plane is tiled along z and 16 bits wide (a bin
 * counts frames, one vote per frame at most: 65535 labelled frames per volume, checked; the reference's own cap is
 * 100 frames, kernel.cpp:60).  sfm_plane_device_ptr(SFM_PLANE_HIST) therefore returns a reference-layout device
 * SNAPSHOT owned by the library and refreshed by every call (bins x 4 bytes per voxel of extra device memory);
 * sfm_hist_export_dev writes the same into a caller-owned device buffer of sfm_plane_bytes(SFM_PLANE_HIST) bytes,
 * ordered on the handle's stream. */
int sfm_hist_export_dev(sfm_volume *v, void *d_dst_u32);
/* Ray-cast of a z-slab-sharded volume AFTER the fusion is over (the viewer runs after the frame loop, kernel.cpp:101-107):
 * the SDF (4 bytes per voxel) is replicated, the histogram (2 x bins bytes per voxel) stays sharded.
 *   sfm_sdf_planes_dev    copies global planes [z0, z0+n) of the SDF between a handle and a packed device buffer
 *                         [Dx][Dy][n] (to_buffer != 0: export, e.g. a slab's owned planes; 0: import into a replica)
 *   sfm_rebuild_skip_map  recomputes the marcher's surface-block map from the SDF plane as it is now
 *   sfm_raycast_band_dev  marches the rays of image rows [row0, row0+rows) on a full-volume handle (a bins = 0 replica
 *                         will do) and writes their hits {x, y, z, t} (t == 0: no hit) into d_hits f32[h*w*4]
 *   sfm_label_hits_dev    on a slab handle: arg-max label (viewer.cu:69-79) of the hits whose sample lies in the OWNED
 *                         planes -> d_keys u64[h*w] = float_bits(t) << 32 | label, SFM_NO_HIT_KEY elsewhere; a MIN
 *                         over the ranks gives the single-volume keys.
 * The march on the replica is the single-volume march (same kernel, same SDF bits), so the composite equals
 * sfm_raycast_keys_dev on one big volume bit for bit. */
int sfm_sdf_planes_dev(sfm_volume *v, int z0, int n, void *d_buf, int to_buffer);
int sfm_rebuild_skip_map(sfm_volume *v);
int sfm_raycast_band_dev(sfm_volume *v, const float *s2w16, const float *c3, int w, int h, int row0, int rows, void *d_hits);
int sfm_label_hits_dev(sfm_volume *v, const void *d_hits, int w, int h, void *d_keys);
/* The same two steps with the image split over n_parts GPUs by INTERLEAVED 4-row tile rows (part p marches tile rows p,
 * p + n_parts, ...): every part sees the same mix of cheap and expensive image regions, where contiguous bands differ by
 * 3x (rays that graze the floor at the bottom of a view march ten times as many samples as rays into the sky).
 *   sfm_raycast_part_dev       writes this part's hits densely into d_hits_part f32[sfm_part_rows(h, n_parts)][w][4]
 *                              (zero where the image has ended) -- the chunk an all-gather collects from every rank
 *   sfm_label_hits_parts_dev   like sfm_label_hits_dev on the all-gathered chunks (part-major); keys in raster order
 * sfm_part_rows = 4 * ceil(ceil(h / 4) / n_parts). */
int sfm_part_rows(int h, int n_parts);
int sfm_raycast_part_dev(sfm_volume *v, const float *s2w16, const float *c3, int w, int h, int part, int n_parts, void *d_hits_part);
int sfm_label_hits_parts_dev(sfm_volume *v, const void *d_hits_parts, int w, int h, int n_parts, void *d_keys);
/* SDF samples gathered (8 taps x 4 bytes each) and surface hits of the ray marches since the previous call: the
 * algorithmic bytes of the ray kernels. */
int sfm_ray_stats(sfm_volume *v, uint64_t *samples, uint64_t *hits);
/* Blocks until every frame copy issued so far has read its source buffers (see SFM_FLAG_ASYNC_SOURCES). */
int sfm_wait_uploads(sfm_volume *v);

/* show_tsdf_kernel (viewer.cu:17-86, launch 152-166).  bgr u8[h*w*3] host; t_opt f32[h*w],
 * label_opt u8[h*w] optional. */
int sfm_raycast(sfm_volume *v, const float *s2w16, const float *c3, int w, int h,
	uint8_t *bgr, float *t_opt, uint8_t *label_opt);
/* Per-ray flags of the last sfm_raycast / sfm_backproject (bit0: a trilinear tap was clamped to
 * the volume -- the reference reads out of bounds on those rays). */
int sfm_ray_flags(sfm_volume *v, uint8_t *flags, size_t n);
/* Per-ray first-hit keys for the multi-GPU min-composite: d_keys u64[h*w] DEVICE memory,
 * key = (float_bits(t_hit) << 32) | label, or UINT64_MAX when this slab has no hit. */
int sfm_raycast_keys_dev(sfm_volume *v, const float *s2w16, const float *c3, int w, int h, void *d_keys);
/* Sharded ray-cast over z-slab handles, exact first-hit compositing in three MIN reductions (see
 * k_raymarch.cuh).  All buffers are DEVICE u64[h*w]; between the stages the caller all-reduces the
 * output with MIN over the ranks (values are < 2^63, so a signed 64-bit MIN works):
 *   stage 1: out = first coarse-step event of the samples this handle owns
 *   stage 2: in ev1 = reduced stage-1 events; out = first fine-step hit of the owned samples
 *   stage 3: in ev1, ev2 = reduced events; out = float_bits(t_hit) << 32 | label for the rays whose
 *            hit sample this handle owns, SFM_NO_HIT_KEY otherwise
 * The handle must store a halo of ceil(voxel.x/voxel.z)+2 planes beyond its owned range on both
 * sides (except at the volume faces); sfm_shard_halo() returns that number. */
#define SFM_NO_HIT_KEY 0x7fffffffffffffffull
int sfm_shard_raycast_stage(sfm_volume *v, int stage, const float *s2w16, const float *c3, int w, int h,
	const void *d_ev1, const void *d_ev2, void *d_out);
int sfm_shard_halo(const float *voxel3);
/* Colour render mode: the marcher of sfm_raycast with interp_tsdf_color (utils.cu:121-142) at the hit -- the
 * call the reference keeps commented out at viewer.cu:68.  bgr: u8[h][w][3] in the colour plane's channel
 * order, zero where no surface is hit; t_opt (optional): refined t per ray; xyzt_opt (optional): f32[h][w][4]
 * hit position in the volume's frame + t (all zero where no surface is hit). */
int sfm_raycast_color(sfm_volume *v, const float *s2w16, const float *c3, int w, int h, uint8_t *bgr, float *t_opt, float *xyzt_opt);
int sfm_show_color(sfm_volume *v, float angle, float dist, int w, int h, uint8_t *bgr);  /* Viewer::show_tsdf's orbit camera */

/* Surface export (no reference counterpart: the reference can only look at the volume through its viewer
 * window).  One point per voxel edge (+x, +y, +z) between two observed voxels whose SDF values differ in
 * sign, at the linear zero crossing; colour and arg-max label of the end voxel nearer to the crossing.
 * *count receives the number of crossings found (it may exceed max_points; only the first max_points, in
 * arbitrary order, are written -- call with max_points = 0 to size the buffers). */
int sfm_extract_surface(sfm_volume *v, uint32_t max_points, float *xyz, uint8_t *bgr, uint8_t *label, uint32_t *count);

/* Duplicate-instance merge over z-slabs (tsdf.cu:426-461 on a sharded volume).  Per frame and rank:
 *   1. sfm_shard_backproj_stage(v, 1|2|3, extrinsic2init, ...) -- the three stages of the exact sharded march
 *      (see sfm_shard_raycast_stage) from the INCOMING camera; the caller MIN-all-reduces each stage's
 *      output.  Stage 3 also keeps the hit positions this rank owns inside the handle.
 *   2. sfm_shard_fold(v, d_mask, d_keys_global, do_counts, d_tables) -- folds the owned hits into the
 *      overlap tables (fixed-point integers, sfm_fold_table_bytes() gives the size and the length of the
 *      leading int64 part; the rest is int32); exactly one rank passes do_counts = 1.  The caller
 *      SUM-all-reduces d_tables (int64 part and int32 part): integer sums are exact and order-independent,
 *      so the result equals the single-GPU tables bit for bit.
 *   3. sfm_shard_merge_finish(v, d_tables_reduced, d_mask_inout, lut256, report) -- the decision half of
 *      filter_overlaps (tsdf.cu:335-389) on every rank (same inputs, same result), relabels the device mask
 *      in place and returns the 256-entry label map for the caller's host copy.
 *   4. sfm_integrate_dev(v, d_depth, d_color, d_mask_inout, extrinsic2init).
 * All device pointers are caller-owned; the mask is u8[H*W]. */
int sfm_shard_backproj_stage(sfm_volume *v, int stage, const float *extrinsic2init16, const void *d_ev1, const void *d_ev2, void *d_out);
int sfm_shard_first_frame(sfm_volume *v, const void *d_mask);  /* n_obs == 0: num_objs = max(mask)+1, tsdf.cu:464-467 */
int sfm_fold_table_bytes(int bins, size_t *bytes_i64, size_t *bytes_total);
int sfm_shard_fold(sfm_volume *v, const void *d_mask, const void *d_keys_global, int do_counts, void *d_tables);
int sfm_shard_merge_finish(sfm_volume *v, const void *d_tables_reduced, void *d_mask_inout, uint8_t *lut256, sfm_merge_report *report);

/* key image (device) -> BGR image (host) through the palette (viewer.cu:80-83). */
int sfm_keys_to_bgr(sfm_volume *v, const void *d_keys, int w, int h, uint8_t *bgr);
/* Viewer::show_tsdf (viewer.cu:137-179): orbit camera matrices + ray-cast. */
int sfm_show(sfm_volume *v, float angle, float dist, int w, int h, uint8_t *bgr);
/* The camera matrices alone (viewer.cu:140-146). */
void sfm_orbit_camera(const float *Kinv16, float angle, float dist, float *s2w16, float *c3);
/* Viewer palette (viewer.cu:93-126): 16 colours repeated; index = label % 16 beyond 32. */
void sfm_palette(uint8_t *rgb, int n);

int sfm_get_info(sfm_volume *v, sfm_info *info);
int sfm_synchronize(sfm_volume *v);
/* Run the handle's work on a caller-owned CUDA stream (cudaStream_t passed as void*). */
int sfm_set_stream(sfm_volume *v, void *cuda_stream);
/* CUDA-event timer on the handle's stream. */
int sfm_timer_start(sfm_volume *v);
int sfm_timer_stop(sfm_volume *v, float *ms);
/* Number of kernel launches issued by this handle so far. */
uint64_t sfm_launch_count(sfm_volume *v);
/* Device time (ms) of the integrate step (K1a classification + K1b update, tsdf_kernel's work) for the
 * last call / the last n calls, from CUDA events the library records around the launches (ring of 2048
 * calls, oldest first).  _times2 splits it into the two kernels. */
int sfm_last_integrate_ms(sfm_volume *v, float *ms);
int sfm_integrate_times(sfm_volume *v, float *ms, int n);
int sfm_integrate_times2(sfm_volume *v, float *ms_classify, float *ms_update, int n);

/* U = voxels whose weight was incremented, S = voxels whose colour/histogram was updated, summed
 * over the integrate calls since the previous sfm_frame_stats call: the terms of the
 * algorithmic-bytes formula 16*U + 14*S. */
int sfm_frame_stats(sfm_volume *v, uint64_t *U, uint64_t *S);

/* Pipelined read-back of the same counters: _begin enqueues the D2H copy after the work submitted
 * so far and returns a ticket (ring of 4); _end waits for that ticket only and returns the
 * CUMULATIVE totals at that point, so step i-1's result can be read while step i runs. */
int sfm_stats_begin(sfm_volume *v, uint64_t *ticket);
int sfm_stats_end(sfm_volume *v, uint64_t ticket, uint64_t *U_total, uint64_t *S_total);

/* Test hook: number of operands (out of blocks*256*per_thread pseudo-random ones) for which the
 * invariant-divisor division used by the ray-marcher differs from the IEEE divide a/b.  Must be 0. */
int sfm_debug_divcheck(float b, unsigned seed, int blocks, int per_thread, float amax, uint64_t *mismatches);

/* Host-side helpers of the reference's driver (the "next" rows). */
/* extrinsic.inv() (tsdf.cu:177) and extrinsic * init_extrinsic_inv (tsdf.cu:217) exactly as sfm_parse_frame computes
 * them (float inputs, double accumulation): for callers that shard the volume themselves (driver/kernel_mgpu.cpp). */
int sfm_mat4_inv(const float *m16, float *out16);
void sfm_mat4_mul(const float *a16, const float *b16, float *out16);

/* mean_depth (utils.cu:77-91). */
float sfm_mean_depth(const uint16_t *depth, int n);
/* parse_extrinsic (utils.cu:8-24): pose {tx,ty,tz,qx,qy,qz,qw} -> world->camera 4x4 f32. */
void sfm_parse_extrinsic(const double *pose7, float *extrinsic16);
/* The interpolating pose front-end of the TSDF_Python prototype (src/TSDF_Python/main.py:127-138,
 * tsdf_utils.py:80-100): pose at `timestamp` between two groundtruth.txt entries a8, b8 =
 * {ts, tx, ty, tz, qx, qy, qz, qw}; pose7_out = {tx, ty, tz, qx, qy, qz, qw}, ready for sfm_parse_extrinsic.
 * (kernel.cpp takes the next entry without interpolating, kernel.cpp:97-98.) */
void sfm_interpolate_pose(const double *a8, const double *b8, double timestamp, double *pose7_out);

#ifdef __cplusplus
}
#endif
#endif /* SFM_B200_H */
