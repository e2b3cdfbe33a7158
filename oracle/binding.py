"""TEST INFRASTRUCTURE ONLY -- ctypes bindings of the two checkers:

  * oracle/liboracle.so          CPU C restatement (sfm_oracle.c), runs anywhere
  * oracle/_ref/libsfm_ref_L*.so the reference's own kernels compiled verbatim (build_ref.py);
                                 the device kernels need a GPU, filter_overlaps runs on the CPU

Never imported by the product package.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_vp = C.c_void_p


def _p(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def _f32(a, n=None):
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    if n is not None:
        assert a.size == n
    return a


def _i32(a):
    return np.ascontiguousarray(a, np.int32).reshape(-1)


_cpu = None


def cpu():
    global _cpu
    if _cpu is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make oracle/liboracle.so`")
        lib = C.CDLL(path)
        lib.orc_mean_depth.restype = C.c_float
        _cpu = lib
    return _cpu


class CpuVolume:
    """Volume state in the reference layout on the host, driven by the C restatement."""

    def __init__(self, dims, bins, start, end, voxel, miu, slab=None):
        self.dims = tuple(int(d) for d in dims)
        self.bins = int(bins)
        self.start, self.end, self.voxel = _f32(start, 3), _f32(end, 3), _f32(voxel, 3)
        self.miu = np.float32(miu)
        self.z0, self.nz = (0, self.dims[2]) if slab is None else slab
        assert (self.z0, self.nz) == (0, self.dims[2]), "CPU oracle keeps whole volumes; crop the result for slabs"
        n = self.dims[0] * self.dims[1] * self.dims[2]
        self.sdf = np.full(n, self.miu, np.float32)            # tsdf.cu:201-205
        self.wt = np.zeros(n, np.int32)
        self.color = np.zeros(n * 3, np.uint8)
        self.hist = np.zeros(n * max(self.bins, 1), np.uint32)
        self.U = self.S = 0

    def integrate(self, K, depth, rgb, mask, E, width, height, z_range=None):
        lib = cpu()
        z0, z1 = z_range if z_range is not None else (0, self.dims[2])
        U, S = C.c_int64(), C.c_int64()
        depth = np.ascontiguousarray(depth, np.uint16)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        mask = np.ascontiguousarray(mask, np.uint8) if mask is not None else np.zeros(width * height, np.uint8)
        K = _f32(K, 16)
        E = _f32(E, 16)
        lib.orc_integrate(_p(self.sdf), _p(self.wt), _p(self.color), _p(self.hist), C.c_int(self.bins),
                          _p(_i32(self.dims)), _p(self.start), _p(self.voxel), C.c_float(self.miu), _p(K),
                          _p(depth), _p(rgb), _p(mask), _p(E), C.c_int(width), C.c_int(height),
                          C.c_int(z0), C.c_int(z1), C.byref(U), C.byref(S))
        self.U, self.S = U.value, S.value
        return U.value, S.value

    def planes(self):
        sh = self.dims
        return {"sdf": self.sdf.reshape(sh), "weight": self.wt.reshape(sh), "color": self.color.reshape(sh + (3,)),
                "hist": self.hist.reshape(sh + (max(self.bins, 1),))}

    def backproject(self, Kinv, Rt, o, width, height):
        lib = cpu()
        probs = np.empty((height, width, self.bins), np.float32)
        box = np.empty((height, width, self.bins), np.uint8)
        t = np.empty((height, width), np.float32)
        fl = np.empty((height, width), np.uint8)
        lib.orc_backproject(_p(self.sdf), _p(self.hist), C.c_int(self.bins), _p(_i32(self.dims)), _p(self.start),
                            _p(self.end), _p(self.voxel), _p(_f32(Kinv, 16)), _p(_f32(Rt, 9)), _p(_f32(o, 3)),
                            C.c_int(width), C.c_int(height), _p(probs), _p(box), _p(t), _p(fl))
        return probs, box, t, fl

    def raycast(self, s2w, c, width, height, palette):
        lib = cpu()
        bgr = np.empty((height, width, 3), np.uint8)
        t = np.empty((height, width), np.float32)
        lab = np.empty((height, width), np.uint8)
        pal = np.ascontiguousarray(palette, np.uint8)
        lib.orc_raycast(_p(self.sdf), _p(self.hist), C.c_int(self.bins), _p(_i32(self.dims)), _p(self.start),
                        _p(self.end), _p(self.voxel), _p(_f32(s2w, 16)), _p(_f32(c, 3)), C.c_int(width), C.c_int(height),
                        _p(pal), _p(bgr), _p(t), _p(lab))
        return bgr, t, lab


def cpu_filter_overlaps(probs, mask, box_mask, bins, n_obs, num_objs, prior=0.05, accept_factor=3.0):
    """orc_filter_overlaps: returns (relabelled mask, num_objs, A, C, assign)."""
    lib = cpu()
    h, w = mask.shape
    mask = np.ascontiguousarray(mask, np.uint8).copy()
    probs = np.ascontiguousarray(probs, np.float32)
    box = np.ascontiguousarray(box_mask, np.uint8)
    A = np.zeros((bins, bins), np.float32)
    Cn = np.zeros((bins, bins), np.uint32)
    assign = np.zeros(bins, np.int32)
    n = C.c_int(num_objs)
    lib.orc_filter_overlaps(_p(probs), C.c_int(w), C.c_int(h), _p(mask), _p(box), C.c_int(bins), C.c_uint32(n_obs),
                            C.c_float(prior), C.c_float(accept_factor), C.byref(n), _p(A), _p(Cn), _p(assign))
    return mask, n.value, A, Cn, assign


def cpu_mean_depth(depth):
    d = np.ascontiguousarray(depth, np.uint16)
    return float(cpu().orc_mean_depth(_p(d), C.c_int(d.size)))


# ---- the verbatim reference (oracle/_ref) -----------------------------------------------------
_ref = {}


def ref_available(bins):
    return os.path.exists(os.path.join(HERE, "_ref", f"libsfm_ref_L{bins}.so"))


def ref(bins):
    if bins not in _ref:
        path = os.path.join(HERE, "_ref", f"libsfm_ref_L{bins}.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `python oracle/build_ref.py` where /root/reference exists")
        lib = C.CDLL(path)
        assert lib.ref_max_objects() == bins
        _ref[bins] = lib
    return _ref[bins]


def ref_filter_overlaps(probs, mask, box_mask, bins, n_obs, num_objs):
    """The reference's own TSDF::filter_overlaps (tsdf.cu:304-416) on the CPU. Returns (mask, num_objs)."""
    lib = ref(bins)
    h, w = mask.shape
    mask = np.ascontiguousarray(mask, np.uint8).copy()
    probs = np.ascontiguousarray(probs, np.float32)
    box = np.ascontiguousarray(box_mask, np.uint8)
    n = C.c_int(num_objs)
    # the reference prints its decisions to stdout (tsdf.cu:351,368,412); silence fd 1 for the call
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        rc = lib.ref_filter_overlaps(_p(probs), C.c_int(w), C.c_int(h), _p(mask), _p(box), C.c_uint32(n_obs), C.byref(n))
        C.CDLL(None).fflush(None)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    assert rc == 0
    return mask, n.value


def ref_integrate(bins, sdf_d, color_d, cnt_d, wt_d, dims, start, voxel, miu, K, depth_d, rgb_d, mask_d, E, w, h):
    """Reference tsdf_kernel (tsdf.cu:18-70) on device pointers (ints)."""
    lib = ref(bins)
    rc = lib.ref_integrate(_vp(sdf_d), _vp(color_d), _vp(cnt_d), _vp(wt_d), _p(_i32(dims)), _p(_f32(start, 3)),
                           _p(_f32(voxel, 3)), C.c_float(miu), _p(_f32(K, 16)), _vp(depth_d), _vp(rgb_d), _vp(mask_d),
                           _p(_f32(E, 16)), C.c_int(w), C.c_int(h))
    assert rc == 0, "reference tsdf_kernel failed"


def ref_back_proj(bins, Kinv, Rt, o, start, end, voxel, dims, sdf_d, cnt_d, w, h, probs_d, box_d):
    lib = ref(bins)
    rc = lib.ref_back_proj(_p(_f32(Kinv, 16)), _p(_f32(Rt, 9)), _p(_f32(o, 3)), _p(_f32(start, 3)), _p(_f32(end, 3)),
                           _p(_f32(voxel, 3)), _p(_i32(dims)), _vp(sdf_d), _vp(cnt_d), C.c_int(w), C.c_int(h),
                           _vp(probs_d), _vp(box_d))
    assert rc == 0, "reference back_proj_kernel failed"


def ref_color_at(bins, xyz_d, valid_d, n, start, voxel, dims, color_d, out_d):
    """The reference's interp_tsdf_color (utils.cu:121-142) at n positions (device pointers)."""
    lib = ref(bins)
    if not hasattr(lib, "ref_color_at"):
        return False
    rc = lib.ref_color_at(_vp(xyz_d), _vp(valid_d), C.c_int(n), _p(_f32(start, 3)), _p(_f32(voxel, 3)), _p(_i32(dims)),
                          _vp(color_d), _vp(out_d))
    assert rc == 0, "reference interp_tsdf_color failed"
    return True


def ref_show(bins, s2w, c, start, end, voxel, dims, sdf_d, color_d, cnt_d, w, h, out_d, palette):
    lib = ref(bins)
    pal = np.ascontiguousarray(palette, np.uint8)
    assert pal.size >= bins * 3
    rc = lib.ref_show(_p(_f32(s2w, 16)), _p(_f32(c, 3)), _p(_f32(start, 3)), _p(_f32(end, 3)), _p(_f32(voxel, 3)),
                      _p(_i32(dims)), _vp(sdf_d), _vp(color_d), _vp(cnt_d), C.c_int(w), C.c_int(h), _vp(out_d), _p(pal))
    assert rc == 0, "reference show_tsdf_kernel failed"
