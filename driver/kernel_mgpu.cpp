// kernel_mgpu.cpp -- the kernel.cpp-style driver over N GPUs of one box, NCCL called directly from C++ (no Python, no
// torch): the reference's frame loop (src/SfM_CUDA/kernel.cpp:64-100) and viewer loop (kernel.cpp:101-107) with the
// volume cut into z-slabs, one C-ABI handle (include/sfm_b200.h) per GPU, one host thread.
//
// Per frame (TSDF::parse_frame / launch_kernel, tsdf.cu:171-228, 418-504, over slabs):
//   H2D of the frame on GPU 0 -> ncclBroadcast to every GPU -> duplicate-instance merge: the exact three-stage
//   sharded march from the incoming camera (three ncclAllReduce MIN), fold of the hits every GPU owns, ncclAllReduce
//   SUM of the integer overlap tables, the same decision on every GPU -> integrate into every slab.
// Viewer (Viewer::show_tsdf, viewer.cu:137-179): the SDF is replicated once (ncclAllGather of the owned planes), every
//   GPU marches a band of image rows, the hits are all-gathered, every GPU labels the hits in its own planes, one
//   ncclAllReduce MIN composites the keys; GPU 0 maps labels to colours.
// With --gpus 1 the same code runs on one slab; tests/test_gpu_driver.py compares the rendered image, the frame count
// and num_objs of --gpus 2 with driver/sfm_driver (one whole volume) on the same sequence: they must be identical.
#include <cstdio>
#include <cstdlib>
#include <memory>

#include <cuda_runtime.h>
#include <nccl.h>

#include "tum_io.hpp"

#define CUDA_OK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw string(#x) + ": " + cudaGetErrorString(e_); } while (0)
#define NCCL_OK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) throw string(#x) + ": " + ncclGetErrorString(r_); } while (0)
#define SFM_OK_(x) do { if ((x) != SFM_OK) throw string(#x) + ": " + sfm_last_error(); } while (0)

struct Rank {
	int dev = 0;
	cudaStream_t stream = nullptr;
	ncclComm_t comm = nullptr;
	sfm_volume *vol = nullptr, *replica = nullptr;
	int own_z0 = 0, own_nz = 0;
	uint8_t *d_frame = nullptr;  // [depth u16 | bgr u8x3 | mask u8]
	unsigned long long *d_ev1 = nullptr, *d_ev2 = nullptr, *d_keys = nullptr;
	uint8_t *d_tables = nullptr;
	float *d_sdf_mine = nullptr, *d_sdf_all = nullptr, *d_hits = nullptr;
};

int main(int argc, char **argv) {
	string root = ".", render = "render_mgpu.ppm";
	double begin = 68164, end = 68170;
	int max_frames = 100, dim = 256, bins = MAX_OBJECTS, views = 10, ngpu = 2;
	float intr[4] = {520.9f, 521.0f, 325.1f, 249.7f};
	for (int i = 1; i < argc; i++) {
		const string a = argv[i];
		auto next = [&]() { return string(i + 1 < argc ? argv[++i] : ""); };
		if (a == "--dim") dim = atoi(next().c_str());
		else if (a == "--bins") bins = atoi(next().c_str());
		else if (a == "--begin") begin = atof(next().c_str());
		else if (a == "--end") end = atof(next().c_str());
		else if (a == "--max-frames") max_frames = atoi(next().c_str());
		else if (a == "--views") views = atoi(next().c_str());
		else if (a == "--render") render = next();
		else if (a == "--gpus") ngpu = atoi(next().c_str());
		else root = a;
	}
	try {
		int have = 0;
		CUDA_OK(cudaGetDeviceCount(&have));
		if (ngpu < 1 || ngpu > have) throw string("--gpus ") + to_string(ngpu) + ": the box has " + to_string(have) + " GPUs";
		auto traj = read_trajactory(root + "/groundtruth.txt");
		vector<string> rgb_fn = glob_png(root + "/rgb"), depth_fn = glob_png(root + "/depth"), mask_fn = glob_png(root + "/mask");
		if (traj.empty() || depth_fn.empty() || mask_fn.empty() || rgb_fn.size() != mask_fn.size()) {
			cerr << "need <dir>/groundtruth.txt, depth/*.png, rgb/*.png and one mask/*.png per rgb frame" << endl;
			return 2;
		}
		vector<double> depth_ts, mask_ts;
		for (auto &f : depth_fn) depth_ts.push_back(stamp_of(f));
		for (auto &f : mask_fn) mask_ts.push_back(stamp_of(f));
		vector<pair<size_t, size_t>> pairs;  // kernel.cpp:64-74
		size_t j = 0;
		for (size_t i = 0; i < depth_ts.size(); i++) {
			if (depth_ts[i] < begin || depth_ts[i] > end) continue;
			while (i < depth_ts.size() && j < mask_ts.size() && depth_ts[i] < mask_ts[j]) i++;
			while (i < depth_ts.size() && j < mask_ts.size() && mask_ts[j] < depth_ts[i]) j++;
			if (i >= depth_ts.size() || j >= mask_ts.size()) break;
			if ((int)pairs.size() >= max_frames) break;
			pairs.push_back(make_pair(i, j));
		}
		if (pairs.empty()) throw string("no frame inside the time window");

		vector<Rank> R(ngpu);
		vector<int> devs(ngpu);
		vector<ncclComm_t> comms(ngpu);
		for (int r = 0; r < ngpu; r++) devs[r] = r;
		NCCL_OK(ncclCommInitAll(comms.data(), ngpu, devs.data()));
		int W = 0, H = 0;
		size_t npx = 0, n64 = 0, ntot = 0;
		bool init = false;
		float init_inv[16], mean_depth0 = 0.f, Kinv[16], K[16];
		int halo = 0;
		uint32_t n_obs = 0;
		int num_objs = 0;
		vector<uint8_t> host_frame;

		for (size_t k = 0; k < pairs.size(); k++) {
			const size_t i = pairs[k].first, jj = pairs[k].second;
			sfm::Mat depth_img, mask_img, rgb_img;
			if (!read_png(depth_fn[i], depth_img, false) || !read_png(mask_fn[jj], mask_img, false) || !read_png(rgb_fn[jj], rgb_img, true))
				throw string("cannot decode ") + depth_fn[i];
			cout << "processing: " << i << ", " << rgb_fn[jj] << endl;
			const float mean = sfm_mean_depth((const uint16_t *)depth_img.data, depth_img.rows * depth_img.cols);
			auto low = traj.lower_bound(depth_ts[i]);
			if (low == traj.end()) --low;
			float extrinsic[16];
			sfm_parse_extrinsic(low->second.data(), extrinsic);
			if (!init) {
				// first frame: placement only (tsdf.cu:173-214), every GPU gets its slab of the same volume
				W = depth_img.cols; H = depth_img.rows; npx = (size_t)W * H;
				sfm_desc d;
				sfm_desc_default(&d);
				d.dims[0] = d.dims[1] = d.dims[2] = dim;
				d.bins = bins; d.width = W; d.height = H;
				d.K[0] = intr[0]; d.K[5] = intr[1]; d.K[2] = intr[2]; d.K[6] = intr[3];
				memcpy(K, d.K, sizeof(K));
				SFM_OK_(sfm_mat4_inv(K, Kinv));
				float start[3], vend[3], voxel[3], miu;
				SFM_OK_(sfm_place_volume((const uint16_t *)depth_img.data, W, H, Kinv, mean, d.dims, d.trunc_voxels, start, vend, voxel, &miu));
				halo = sfm_shard_halo(voxel);
				const int per = ((dim + ngpu - 1) / ngpu + 7) / 8 * 8;  // owned planes per GPU, whole bricks
				for (int r = 0; r < ngpu; r++) {
					Rank &q = R[r];
					q.dev = r; q.comm = comms[r];
					q.own_z0 = min(dim, r * per);
					q.own_nz = min(dim, (r + 1) * per) - q.own_z0;
					if (q.own_nz <= 0) throw string("more GPUs than 8-plane slabs");
					const int lo = max(0, (q.own_z0 - halo) / 8 * 8), hi = min(dim, (q.own_z0 + q.own_nz + halo + 7) / 8 * 8);
					CUDA_OK(cudaSetDevice(r));
					CUDA_OK(cudaStreamCreateWithFlags(&q.stream, cudaStreamNonBlocking));
					sfm_desc dr = d;
					dr.device = r; dr.slab_z0 = lo; dr.slab_nz = hi - lo; dr.own_z0 = q.own_z0; dr.own_nz = q.own_nz;
					SFM_OK_(sfm_create(&dr, &q.vol));
					SFM_OK_(sfm_set_stream(q.vol, q.stream));
					SFM_OK_(sfm_set_bounds(q.vol, start, vend, voxel, miu));
					sfm_desc df = d;  // label-free full-volume replica for the viewer
					df.device = r; df.bins = 0;
					SFM_OK_(sfm_create(&df, &q.replica));
					SFM_OK_(sfm_set_stream(q.replica, q.stream));
					SFM_OK_(sfm_set_bounds(q.replica, start, vend, voxel, miu));
					SFM_OK_(sfm_fold_table_bytes(bins, &n64, &ntot));
					CUDA_OK(cudaMalloc(&q.d_frame, npx * 6));
					CUDA_OK(cudaMalloc(&q.d_ev1, npx * 8));
					CUDA_OK(cudaMalloc(&q.d_ev2, npx * 8));
					CUDA_OK(cudaMalloc(&q.d_keys, npx * 8));
					CUDA_OK(cudaMalloc(&q.d_tables, ntot));
					const size_t cols = (size_t)dim * dim;
					CUDA_OK(cudaMalloc(&q.d_sdf_mine, cols * per * 4));
					CUDA_OK(cudaMalloc(&q.d_sdf_all, cols * per * 4 * ngpu));
					CUDA_OK(cudaMalloc(&q.d_hits, (size_t)ngpu * sfm_part_rows(H, ngpu) * W * 16));
				}
				SFM_OK_(sfm_mat4_inv(extrinsic, init_inv));  // tsdf.cu:177
				mean_depth0 = mean;
				init = true;
				host_frame.resize(npx * 6);
				continue;  // the first frame is not integrated (tsdf.cu:213)
			}
			if (depth_img.cols != W || depth_img.rows != H) throw string("frame size differs from the first frame's");
			float e2i[16];
			sfm_mat4_mul(extrinsic, init_inv, e2i);  // tsdf.cu:217
			memcpy(host_frame.data(), depth_img.data, npx * 2);
			memcpy(host_frame.data() + npx * 2, rgb_img.data, npx * 3);
			memcpy(host_frame.data() + npx * 5, mask_img.data, npx);
			int mx = 0;
			for (size_t p = 0; p < npx; p++) mx = max(mx, (int)mask_img.data[p]);
			if (mx >= bins) throw string("mask carries a label >= bins");
			// frame to GPU 0, broadcast over NVLink
			CUDA_OK(cudaSetDevice(0));
			CUDA_OK(cudaMemcpyAsync(R[0].d_frame, host_frame.data(), npx * 6, cudaMemcpyHostToDevice, R[0].stream));
			CUDA_OK(cudaStreamSynchronize(R[0].stream));  // (host_frame is reused next frame)
			NCCL_OK(ncclGroupStart());
			for (auto &q : R) NCCL_OK(ncclBroadcast(q.d_frame, q.d_frame, npx * 6, ncclUint8, 0, q.comm, q.stream));
			NCCL_OK(ncclGroupEnd());
			auto all_reduce = [&](size_t off_bytes, size_t count, ncclDataType_t ty, ncclRedOp_t op, int which) {
				NCCL_OK(ncclGroupStart());
				for (auto &q : R) {
					void *p = which == 0 ? (void *)q.d_ev1 : which == 1 ? (void *)q.d_ev2 : which == 2 ? (void *)q.d_keys : (void *)(q.d_tables + off_bytes);
					NCCL_OK(ncclAllReduce(p, p, count, ty, op, q.comm, q.stream));
				}
				NCCL_OK(ncclGroupEnd());
			};
			if (n_obs > 0) {
				// tsdf.cu:426-461 over slabs: exact sharded march (keys are < 2^63, so a signed MIN composites them)
				for (auto &q : R) { CUDA_OK(cudaSetDevice(q.dev)); SFM_OK_(sfm_shard_backproj_stage(q.vol, 1, e2i, nullptr, nullptr, q.d_ev1)); }
				all_reduce(0, npx, ncclInt64, ncclMin, 0);
				for (auto &q : R) { CUDA_OK(cudaSetDevice(q.dev)); SFM_OK_(sfm_shard_backproj_stage(q.vol, 2, e2i, q.d_ev1, nullptr, q.d_ev2)); }
				all_reduce(0, npx, ncclInt64, ncclMin, 1);
				for (auto &q : R) { CUDA_OK(cudaSetDevice(q.dev)); SFM_OK_(sfm_shard_backproj_stage(q.vol, 3, e2i, q.d_ev1, q.d_ev2, q.d_keys)); }
				all_reduce(0, npx, ncclInt64, ncclMin, 2);
				for (auto &q : R) { CUDA_OK(cudaSetDevice(q.dev)); SFM_OK_(sfm_shard_fold(q.vol, q.d_frame + npx * 5, q.d_keys, q.dev == 0, q.d_tables)); }
				all_reduce(0, n64 / 8, ncclInt64, ncclSum, 3);
				all_reduce(n64, (ntot - n64) / 4, ncclInt32, ncclSum, 3);
				for (auto &q : R) {
					CUDA_OK(cudaSetDevice(q.dev));
					sfm_merge_report rep;
					SFM_OK_(sfm_shard_merge_finish(q.vol, q.d_tables, q.d_frame + npx * 5, nullptr, &rep));
					num_objs = rep.num_objs;
				}
			} else {
				for (auto &q : R) { CUDA_OK(cudaSetDevice(q.dev)); SFM_OK_(sfm_shard_first_frame(q.vol, q.d_frame + npx * 5)); }
				num_objs = mx + 1;  // tsdf.cu:464-467
			}
			for (auto &q : R) {
				CUDA_OK(cudaSetDevice(q.dev));
				SFM_OK_(sfm_integrate_dev(q.vol, q.d_frame, q.d_frame + npx * 2, q.d_frame + npx * 5, e2i));
			}
			n_obs++;
		}
		for (auto &q : R) { CUDA_OK(cudaSetDevice(q.dev)); SFM_OK_(sfm_synchronize(q.vol)); }
		cout << "fused " << n_obs << " frames, num_objs " << num_objs << ", " << ngpu << " z-slabs" << endl;

		// ---- viewer over a replicated SDF ----
		const size_t cols = (size_t)dim * dim;
		const int per = R[0].own_nz;
		NCCL_OK(ncclGroupStart());
		for (auto &q : R) {
			CUDA_OK(cudaSetDevice(q.dev));
			SFM_OK_(sfm_sdf_planes_dev(q.vol, q.own_z0, q.own_nz, q.d_sdf_mine, 1));
			NCCL_OK(ncclAllGather(q.d_sdf_mine, q.d_sdf_all, cols * per, ncclFloat, q.comm, q.stream));
		}
		NCCL_OK(ncclGroupEnd());
		for (auto &q : R) {
			CUDA_OK(cudaSetDevice(q.dev));
			for (auto &o : R) SFM_OK_(sfm_sdf_planes_dev(q.replica, o.own_z0, o.own_nz, q.d_sdf_all + (size_t)o.dev * cols * per, 0));
			SFM_OK_(sfm_rebuild_skip_map(q.replica));
		}
		// the image is split over the GPUs by interleaved 4-row tile rows (GPU p marches tile rows p, p + ngpu, ...): every GPU
		// gets the same mix of cheap and expensive rows; each writes its share densely at its offset of d_hits
		const size_t part_floats = (size_t)sfm_part_rows(H, ngpu) * W * 4;
		vector<uint8_t> bgr(npx * 3);
		float angle = 0.f;
		for (int v = 0; v < views; v++) {
			angle += 0.01f;  // kernel.cpp:104
			float s2w[16], c3[3];
			sfm_orbit_camera(Kinv, angle, mean_depth0, s2w, c3);
			NCCL_OK(ncclGroupStart());
			for (auto &q : R) {
				CUDA_OK(cudaSetDevice(q.dev));
				SFM_OK_(sfm_raycast_part_dev(q.replica, s2w, c3, W, H, q.dev, ngpu, q.d_hits + (size_t)q.dev * part_floats));
				// in-place all-gather: this GPU's share is already at its offset in d_hits
				NCCL_OK(ncclAllGather(q.d_hits + (size_t)q.dev * part_floats, q.d_hits, part_floats, ncclFloat, q.comm, q.stream));
			}
			NCCL_OK(ncclGroupEnd());
			NCCL_OK(ncclGroupStart());
			for (auto &q : R) {
				CUDA_OK(cudaSetDevice(q.dev));
				SFM_OK_(sfm_label_hits_parts_dev(q.vol, q.d_hits, W, H, ngpu, q.d_keys));
				NCCL_OK(ncclAllReduce(q.d_keys, q.d_keys, npx, ncclInt64, ncclMin, q.comm, q.stream));
			}
			NCCL_OK(ncclGroupEnd());
		}
		CUDA_OK(cudaSetDevice(0));
		SFM_OK_(sfm_keys_to_bgr(R[0].vol, R[0].d_keys, W, H, bgr.data()));
		size_t lit = 0;
		for (size_t p = 0; p < npx; p++) lit += (bgr[p * 3] | bgr[p * 3 + 1] | bgr[p * 3 + 2]) != 0;
		sfm::Mat img(H, W, 3, 1, bgr.data());
		write_ppm_bgr(render, img);
		cout << "rendered " << views << " views, last one -> " << render << " (" << lit << " labelled pixels)" << endl;
		for (auto &q : R) {
			CUDA_OK(cudaSetDevice(q.dev));
			sfm_destroy(q.replica);
			sfm_destroy(q.vol);
			cudaFree(q.d_frame); cudaFree(q.d_ev1); cudaFree(q.d_ev2); cudaFree(q.d_keys); cudaFree(q.d_tables);
			cudaFree(q.d_sdf_mine); cudaFree(q.d_sdf_all); cudaFree(q.d_hits);
			ncclCommDestroy(q.comm);
			cudaStreamDestroy(q.stream);
		}
	} catch (const string &e) {
		cerr << e << endl;
		return 1;
	}
	return 0;
}
