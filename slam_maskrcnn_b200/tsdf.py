"""Host-side mirror of the reference's `class TSDF` / `class Viewer` over the C-ABI.

Same names, argument meaning and error behaviour as src/SfM_CUDA/tsdf.cuh:7-67 and
viewer.cuh:4-17 (parse_frame relabels `masks` in place; failures raise), with NumPy arrays in
place of cv::Mat.  All compute happens in libsfm_b200.so on the GPU; nothing here falls back to
NumPy.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Desc, Info, MergeReport, check

MAX_OBJECTS = 32  # tsdf.cuh:4 (default bin count; a run-time parameter here)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a, n):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    assert a.size == n, f"expected {n} floats, got {a.size}"
    return a


def place_volume(depth, Kinv, mean_depth, dims, trunc_voxels=5.0):
    """sfm_place_volume: the reference's volume placement rule (tsdf.cu:180-199) -> (start, end, voxel, miu), float32.
    Host arithmetic only (no GPU needed)."""
    from . import _lib
    lib = _lib.load()
    depth = np.ascontiguousarray(depth, np.uint16)
    h, w = depth.shape
    start, end, voxel = np.zeros(3, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32)
    miu = C.c_float(0)
    d = np.ascontiguousarray(dims, np.int32)
    check(lib.sfm_place_volume(_ptr(depth), w, h, _ptr(_f32(Kinv, 16)), C.c_float(float(mean_depth)), _ptr(d), C.c_float(trunc_voxels),
                               _ptr(start), _ptr(end), _ptr(voxel), C.byref(miu)))
    return start, end, voxel, np.float32(miu.value)


def intrinsic_matrix(fx, fy, cx, cy):
    """tsdf.cu:137-146: eye(4) float32 with fx, fy, cx, cy."""
    K = np.eye(4, dtype=np.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx, fy, cx, cy
    return K


class Volume:
    """One sfm_volume handle (a whole volume, or one z-slab of it, on one GPU)."""

    def __init__(self, dims=(256, 256, 256), bins=MAX_OBJECTS, width=640, height=480,
                 intrinsics=(520.9, 521.0, 325.1, 249.7), K=None, Kinv=None, device=0,
                 slab=None, own=None, flags=0, **cfg):
        self.lib = _lib.load()
        d = Desc()
        self.lib.sfm_desc_default(C.byref(d))
        d.dims[:] = [int(x) for x in dims]
        d.bins, d.width, d.height = int(bins), int(width), int(height)
        Kmat = intrinsic_matrix(*intrinsics) if K is None else np.asarray(K, np.float32).reshape(4, 4)
        d.K[:] = Kmat.reshape(-1).tolist()
        if Kinv is not None:
            d.Kinv[:] = np.asarray(Kinv, np.float32).reshape(-1).tolist()
        d.device = int(device)
        if slab is not None:
            d.slab_z0, d.slab_nz = int(slab[0]), int(slab[1])
        d.flags = int(flags)
        if own is not None:
            d.own_z0, d.own_nz = int(own[0]), int(own[1])
        for k, val in cfg.items():
            if not hasattr(d, k):
                raise TypeError(f"unknown sfm_desc field {k}")
            setattr(d, k, val)
        self.desc = d
        self._h = C.c_void_p()
        check(self.lib.sfm_create(C.byref(d), C.byref(self._h)))
        self.dims = tuple(d.dims)
        self.bins, self.width, self.height = d.bins, d.width, d.height
        nz = d.slab_nz if d.slab_nz > 0 else d.dims[2] - d.slab_z0
        self.slab = (d.slab_z0, nz)
        self.local_shape = (d.dims[0], d.dims[1], nz)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.sfm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- placement --------------------------------------------------------------------------
    def set_bounds(self, vol_start, vol_end, voxel=None, miu=None):
        s, e = _f32(vol_start, 3), _f32(vol_end, 3)
        if voxel is None:  # tsdf.cu:197, float32 arithmetic
            voxel = (e - s) / (np.array(self.dims, np.float32) - np.float32(1))
        vx = _f32(voxel, 3)
        if miu is None:  # tsdf.cu:199
            miu = np.float32(self.desc.trunc_voxels) * vx[0]
        check(self.lib.sfm_set_bounds(self._h, _ptr(s), _ptr(e), _ptr(vx), C.c_float(float(miu))))

    def init_from_frame(self, depth, extrinsic, mean_depth):
        depth = self._img(depth, np.uint16, 1)
        check(self.lib.sfm_init_from_frame(self._h, _ptr(depth), _ptr(_f32(extrinsic, 16)), C.c_float(mean_depth)))

    # -- frames -----------------------------------------------------------------------------
    def _img(self, a, dtype, ch):
        a = np.ascontiguousarray(a, dtype=dtype)
        assert a.size == self.width * self.height * ch, f"frame size mismatch: {a.shape}"
        return a

    def integrate_raw(self, depth, color, mask, extrinsic2init):
        depth, color = self._img(depth, np.uint16, 1), self._img(color, np.uint8, 3)
        mask = self._img(mask, np.uint8, 1) if mask is not None else None
        check(self.lib.sfm_integrate_raw(self._h, _ptr(depth), _ptr(color), _ptr(mask), _ptr(_f32(extrinsic2init, 16))))

    READY_IN_ORDER = object()

    def integrate_dev(self, d_depth, d_color, d_mask, extrinsic2init, ready=READY_IN_ORDER):
        """Frame images given as device pointers (ints), e.g. torch tensors' data_ptr().
        `ready`: when the images are valid -- READY_IN_ORDER (default: in the order of the handle's stream),
        None (valid already) or a cudaEvent_t handle / torch.cuda.Event the producer recorded; the last two
        let the frame preparation (K0 + K1a) overlap the previous frame's update kernel."""
        m = C.c_void_p(d_mask) if d_mask else None
        if ready is Volume.READY_IN_ORDER:
            check(self.lib.sfm_integrate_dev(self._h, C.c_void_p(d_depth), C.c_void_p(d_color), m, _ptr(_f32(extrinsic2init, 16))))
            return
        ev = getattr(ready, "cuda_event", ready)  # torch.cuda.Event -> raw handle
        check(self.lib.sfm_integrate_dev_ready(self._h, C.c_void_p(d_depth), C.c_void_p(d_color), m, _ptr(_f32(extrinsic2init, 16)),
                                               C.c_void_p(ev) if ev else None))

    def fuse_frame(self, depth, color, mask_inout, extrinsic2init):
        depth, color = self._img(depth, np.uint16, 1), self._img(color, np.uint8, 3)
        if mask_inout is not None:
            assert mask_inout.dtype == np.uint8 and mask_inout.flags.c_contiguous and mask_inout.flags.writeable
        check(self.lib.sfm_fuse_frame(self._h, _ptr(depth), _ptr(color), _ptr(mask_inout), _ptr(_f32(extrinsic2init, 16))))

    def parse_frame(self, depth, color, mask_inout, extrinsic, mean_depth):
        depth = self._img(depth, np.uint16, 1)
        color = self._img(color, np.uint8, 3) if color is not None else None
        if mask_inout is not None:
            assert mask_inout.dtype == np.uint8 and mask_inout.flags.c_contiguous and mask_inout.flags.writeable
        check(self.lib.sfm_parse_frame(self._h, _ptr(depth), _ptr(color), _ptr(mask_inout), _ptr(_f32(extrinsic, 16)),
                                       C.c_float(mean_depth)))

    # -- merge hooks ------------------------------------------------------------------------
    def backproject(self, extrinsic2init, want_t=True, want_flags=True):
        n = self.width * self.height
        probs = np.empty((self.height, self.width, self.bins), np.float32)
        box = np.empty((self.height, self.width, self.bins), np.uint8)
        t = np.empty((self.height, self.width), np.float32) if want_t else None
        fl = np.empty((self.height, self.width), np.uint8) if want_flags else None
        check(self.lib.sfm_backproject(self._h, _ptr(_f32(extrinsic2init, 16)), _ptr(probs), _ptr(box), _ptr(t), _ptr(fl)))
        return probs, box, t, fl

    def overlap_tables(self, extrinsic2init, mask):
        mask = self._img(mask, np.uint8, 1)
        A = np.empty((self.bins, self.bins), np.float64)
        Cn = np.empty((self.bins, self.bins), np.uint32)
        check(self.lib.sfm_overlap_tables(self._h, _ptr(_f32(extrinsic2init, 16)), _ptr(mask), _ptr(A), _ptr(Cn)))
        return A, Cn

    def merge_decide(self, A, Cn, mask_inout):
        A = np.ascontiguousarray(A, np.float64)
        Cn = np.ascontiguousarray(Cn, np.uint32)
        rep = MergeReport()
        check(self.lib.sfm_merge_decide(self._h, _ptr(A), _ptr(Cn), _ptr(mask_inout), C.byref(rep)))
        return rep

    def last_merge(self):
        rep = MergeReport()
        check(self.lib.sfm_last_merge(self._h, C.byref(rep)))
        return rep

    # -- planes -----------------------------------------------------------------------------
    _PLANE = {"sdf": (0, np.float32, ()), "weight": (1, np.int32, ()), "color": (2, np.uint8, (3,)), "hist": (3, np.uint32, None)}

    def download(self, name):
        pid, dt, tail = self._PLANE[name]
        tail = (self.bins,) if tail is None else tail
        out = np.empty(self.local_shape + tail, dt)
        check(self.lib.sfm_download(self._h, pid, _ptr(out), C.c_size_t(out.nbytes)))
        return out

    def upload(self, name, arr):
        pid, dt, tail = self._PLANE[name]
        arr = np.ascontiguousarray(arr, dt)
        check(self.lib.sfm_upload(self._h, pid, _ptr(arr), C.c_size_t(arr.nbytes)))

    def plane_ptr(self, name):
        return self.lib.sfm_plane_device_ptr(self._h, self._PLANE[name][0])

    def plane_bytes(self, name):
        return self.lib.sfm_plane_bytes(self._h, self._PLANE[name][0])

    # -- ray-cast ---------------------------------------------------------------------------
    def raycast(self, s2w, c, w=None, h=None, want_t=False, want_label=False):
        w, h = w or self.width, h or self.height
        bgr = np.empty((h, w, 3), np.uint8)
        t = np.empty((h, w), np.float32) if want_t else None
        lab = np.empty((h, w), np.uint8) if want_label else None
        check(self.lib.sfm_raycast(self._h, _ptr(_f32(s2w, 16)), _ptr(_f32(c, 3)), w, h, _ptr(bgr), _ptr(t), _ptr(lab)))
        return bgr, t, lab

    def ray_flags(self, w=None, h=None):
        w, h = w or self.width, h or self.height
        fl = np.empty((h, w), np.uint8)
        check(self.lib.sfm_ray_flags(self._h, _ptr(fl), C.c_size_t(fl.size)))
        return fl

    def raycast_keys_dev(self, s2w, c, w, h, d_keys):
        check(self.lib.sfm_raycast_keys_dev(self._h, _ptr(_f32(s2w, 16)), _ptr(_f32(c, 3)), w, h, C.c_void_p(d_keys)))

    def shard_raycast_stage(self, stage, s2w, c, w, h, d_ev1, d_ev2, d_out):
        check(self.lib.sfm_shard_raycast_stage(self._h, stage, _ptr(_f32(s2w, 16)), _ptr(_f32(c, 3)), w, h,
                                               C.c_void_p(d_ev1) if d_ev1 else None, C.c_void_p(d_ev2) if d_ev2 else None,
                                               C.c_void_p(d_out)))

    def raycast_color(self, s2w, c, w=None, h=None, want_t=False, want_xyzt=False):
        """Colour render mode (interp_tsdf_color at the hit, the call viewer.cu:68 keeps commented out)."""
        w, h = w or self.width, h or self.height
        bgr = np.empty((h, w, 3), np.uint8)
        t = np.empty((h, w), np.float32) if want_t else None
        xyzt = np.empty((h, w, 4), np.float32) if want_xyzt else None
        check(self.lib.sfm_raycast_color(self._h, _ptr(_f32(s2w, 16)), _ptr(_f32(c, 3)), w, h, _ptr(bgr), _ptr(t), _ptr(xyzt)))
        out = (bgr,) + ((t,) if want_t else ()) + ((xyzt,) if want_xyzt else ())
        return out if len(out) > 1 else bgr

    def extract_surface(self):
        """Zero-crossing points of the fused volume: (xyz float32 [n,3], bgr uint8 [n,3], label uint8 [n]),
        sorted by position so the result is reproducible."""
        n = C.c_uint32()
        check(self.lib.sfm_extract_surface(self._h, 0, None, None, None, C.byref(n)))
        cap = int(n.value)
        xyz, bgr, lab = np.empty((cap, 3), np.float32), np.empty((cap, 3), np.uint8), np.empty(cap, np.uint8)
        if cap:
            check(self.lib.sfm_extract_surface(self._h, cap, _ptr(xyz), _ptr(bgr), _ptr(lab), C.byref(n)))
            assert int(n.value) == cap
            order = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
            xyz, bgr, lab = xyz[order], bgr[order], lab[order]
        return xyz, bgr, lab

    # -- duplicate-instance merge over z-slabs (device pointers; the collectives are the caller's) ----
    def shard_backproj_stage(self, stage, extrinsic2init, d_ev1, d_ev2, d_out):
        check(self.lib.sfm_shard_backproj_stage(self._h, stage, _ptr(_f32(extrinsic2init, 16)),
                                                C.c_void_p(d_ev1) if d_ev1 else None, C.c_void_p(d_ev2) if d_ev2 else None,
                                                C.c_void_p(d_out)))

    def shard_first_frame(self, d_mask):
        check(self.lib.sfm_shard_first_frame(self._h, C.c_void_p(d_mask)))

    def fold_table_bytes(self):
        a, b = C.c_size_t(), C.c_size_t()
        check(self.lib.sfm_fold_table_bytes(self.bins, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def shard_fold(self, d_mask, d_keys_global, do_counts, d_tables):
        check(self.lib.sfm_shard_fold(self._h, C.c_void_p(d_mask), C.c_void_p(d_keys_global), int(bool(do_counts)), C.c_void_p(d_tables)))

    def shard_merge_finish(self, d_tables_reduced, d_mask_inout):
        """-> (lut uint8[256], MergeReport); the device mask is relabelled in place."""
        lut = np.zeros(256, np.uint8)
        rep = MergeReport()
        check(self.lib.sfm_shard_merge_finish(self._h, C.c_void_p(d_tables_reduced), C.c_void_p(d_mask_inout), _ptr(lut), C.byref(rep)))
        return lut, rep

    def keys_to_bgr(self, d_keys, w, h):
        bgr = np.empty((h, w, 3), np.uint8)
        check(self.lib.sfm_keys_to_bgr(self._h, C.c_void_p(d_keys), w, h, _ptr(bgr)))
        return bgr

    def show(self, angle, dist, w=None, h=None):
        w, h = w or self.width, h or self.height
        bgr = np.empty((h, w, 3), np.uint8)
        check(self.lib.sfm_show(self._h, C.c_float(angle), C.c_float(dist), w, h, _ptr(bgr)))
        return bgr

    # -- misc -------------------------------------------------------------------------------
    def info(self):
        i = Info()
        check(self.lib.sfm_get_info(self._h, C.byref(i)))
        return i

    def synchronize(self):
        check(self.lib.sfm_synchronize(self._h))

    def sdf_planes_dev(self, z0, n, d_buf, to_buffer):
        check(self.lib.sfm_sdf_planes_dev(self._h, int(z0), int(n), C.c_void_p(d_buf), 1 if to_buffer else 0))

    def rebuild_skip_map(self):
        check(self.lib.sfm_rebuild_skip_map(self._h))

    def raycast_band_dev(self, s2w, c, w, h, row0, rows, d_hits):
        check(self.lib.sfm_raycast_band_dev(self._h, _ptr(_f32(s2w, 16)), _ptr(_f32(c, 3)), w, h, int(row0), int(rows), C.c_void_p(d_hits)))

    def label_hits_dev(self, d_hits, w, h, d_keys):
        check(self.lib.sfm_label_hits_dev(self._h, C.c_void_p(d_hits), w, h, C.c_void_p(d_keys)))

    def raycast_part_dev(self, s2w, c, w, h, part, n_parts, d_hits_part):
        """This part's share of the image (4-row tile rows part, part + n_parts, ...) marched into a dense buffer of
        part_rows(h, n_parts) x w hits."""
        check(self.lib.sfm_raycast_part_dev(self._h, _ptr(_f32(s2w, 16)), _ptr(_f32(c, 3)), w, h, int(part), int(n_parts), C.c_void_p(d_hits_part)))

    def label_hits_parts_dev(self, d_hits_parts, w, h, n_parts, d_keys):
        check(self.lib.sfm_label_hits_parts_dev(self._h, C.c_void_p(d_hits_parts), w, h, int(n_parts), C.c_void_p(d_keys)))

    def part_rows(self, h, n_parts):
        return int(self.lib.sfm_part_rows(int(h), int(n_parts)))

    def ray_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(self.lib.sfm_ray_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def set_num_objs(self, n):
        check(self.lib.sfm_set_num_objs(self._h, int(n)))

    def wait_uploads(self):
        """Blocks until every frame copy issued so far has read its source buffers (FLAG_ASYNC_SOURCES)."""
        check(self.lib.sfm_wait_uploads(self._h))

    def planes_written(self):
        """Tell the library the SDF plane was written through plane_ptr() (resets the surface-block map)."""
        check(self.lib.sfm_planes_written(self._h))

    def set_stream(self, cuda_stream):
        check(self.lib.sfm_set_stream(self._h, C.c_void_p(cuda_stream)))

    def timer_start(self):
        check(self.lib.sfm_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(self.lib.sfm_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.lib.sfm_launch_count(self._h))

    def last_integrate_ms(self):
        ms = C.c_float()
        check(self.lib.sfm_last_integrate_ms(self._h, C.byref(ms)))
        return ms.value

    def integrate_times(self, n):
        ms = np.empty(n, np.float32)
        check(self.lib.sfm_integrate_times(self._h, _ptr(ms), n))
        return ms

    def integrate_times2(self, n):
        """(K1a classification ms, K1b update ms) of the last n integrate calls."""
        a, b = np.empty(n, np.float32), np.empty(n, np.float32)
        check(self.lib.sfm_integrate_times2(self._h, _ptr(a), _ptr(b), n))
        return a, b

    def stats_begin(self):
        t = C.c_uint64()
        check(self.lib.sfm_stats_begin(self._h, C.byref(t)))
        return int(t.value)

    def stats_end(self, ticket):
        u, s = C.c_uint64(), C.c_uint64()
        check(self.lib.sfm_stats_end(self._h, C.c_uint64(ticket), C.byref(u), C.byref(s)))
        return int(u.value), int(s.value)

    def frame_stats(self):
        u, s = C.c_uint64(), C.c_uint64()
        check(self.lib.sfm_frame_stats(self._h, C.byref(u), C.byref(s)))
        return int(u.value), int(s.value)


def write_ply(path, xyz, bgr, label=None):
    """Binary little-endian PLY point cloud (x y z, red green blue[, label])."""
    n = len(xyz)
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")]
    if label is not None:
        fields.append(("label", "u1"))
    rec = np.empty(n, dtype=fields)
    rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    rec["red"], rec["green"], rec["blue"] = bgr[:, 2], bgr[:, 1], bgr[:, 0]
    if label is not None:
        rec["label"] = label
    names = {"<f4": "float", "u1": "uchar"}
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\n" % n).encode())
        for name, dt in fields:
            f.write(("property %s %s\n" % (names[dt], name)).encode())
        f.write(b"end_header\n")
        f.write(rec.tobytes())


def orbit_camera(Kinv, angle, dist):
    """Viewer::show_tsdf camera (viewer.cu:140-146) -> (s2w 4x4, c 3)."""
    lib = _lib.load()
    s2w = np.empty(16, np.float32)
    c = np.empty(3, np.float32)
    lib.sfm_orbit_camera(_ptr(_f32(Kinv, 16)), C.c_float(angle), C.c_float(dist), _ptr(s2w), _ptr(c))
    return s2w.reshape(4, 4), c


def palette(n=32):
    lib = _lib.load()
    p = np.empty((n, 3), np.uint8)
    lib.sfm_palette(_ptr(p), n)
    return p


def mean_depth(depth):
    """utils.cu:77-91."""
    d = np.ascontiguousarray(depth, np.uint16)
    return float(_lib.load().sfm_mean_depth(_ptr(d), d.size))


def interpolate_pose(a8, b8, timestamp):
    """Pose {tx,ty,tz,qx,qy,qz,qw} at `timestamp` between two trajectory entries {ts, tx..qw}: lerp + slerp
    as the TSDF_Python prototype does (main.py:127-138, tsdf_utils.py:80-100)."""
    lib = _lib.load()
    a = np.ascontiguousarray(a8, np.float64).reshape(8)
    b = np.ascontiguousarray(b8, np.float64).reshape(8)
    out = np.empty(7, np.float64)
    lib.sfm_interpolate_pose(_ptr(a), _ptr(b), C.c_double(float(timestamp)), _ptr(out))
    return out


def parse_extrinsic(pose7):
    """utils.cu:8-24: {tx,ty,tz,qx,qy,qz,qw} -> world->camera 4x4 float32."""
    p = np.ascontiguousarray(pose7, np.float64)
    out = np.empty(16, np.float32)
    _lib.load().sfm_parse_extrinsic(_ptr(p), _ptr(out))
    return out.reshape(4, 4)


class TSDF:
    """Reference-shaped front end: `TSDF(intrinsics)`, `parse_frame(depth, color, masks, extrinsic, mean_depth)`
    (tsdf.cuh:9-22).  The volume size and bin count are construction parameters here; the reference fixes
    them at 256^3 (tsdf.cuh:52) and 32 (tsdf.cuh:4)."""

    def __init__(self, intrinsics, dims=(256, 256, 256), bins=MAX_OBJECTS, width=640, height=480, device=0, **cfg):
        self.vol = Volume(dims=dims, bins=bins, width=width, height=height, intrinsics=tuple(intrinsics)[:4],
                          device=device, **cfg)

    def parse_frame(self, depth, color, masks, extrinsic, mean_depth):
        self.vol.parse_frame(depth, color, masks, extrinsic, mean_depth)

    @property
    def mean_depth_(self):
        return self.vol.info().mean_depth

    def get_tsdf_diff(self):
        return self.vol.download("sdf")

    def get_tsdf_color(self):
        return self.vol.download("color")

    def get_tsdf_cnt(self):
        return self.vol.download("hist")

    def get_dim(self):
        return tuple(self.vol.info().dims)

    def get_vol_start(self):
        return np.array(self.vol.info().vol_start, np.float32)

    def get_vol_end(self):
        return np.array(self.vol.info().vol_end, np.float32)

    def get_voxel(self):
        return np.array(self.vol.info().voxel, np.float32)

    def get_intrinsic(self):
        return np.array(self.vol.desc.K, np.float32).reshape(4, 4)


class Viewer:
    """viewer.cuh:4-17: `Viewer(width, height)`, `show_tsdf(tsdf, angle, dist)` -> HxWx3 BGR uint8
    (the cv::imshow / waitKey of viewer.cu:176-177 is the caller's business)."""

    def __init__(self, width, height):
        self.width_, self.height_ = int(width), int(height)

    def show_tsdf(self, tsdf, angle, dist):
        vol = tsdf.vol if isinstance(tsdf, TSDF) else tsdf
        return vol.show(angle, dist, self.width_, self.height_)
