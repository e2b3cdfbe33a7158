"""ctypes binding of libsfm_b200.so (the C-ABI in include/sfm_b200.h).

There is no fallback: if the shared library is missing this module raises at import of the
symbol table, and if no sm_100 GPU is present `sfm_create` returns SFM_ERR_NODEVICE.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SFM_B200_LIB selects another BUILD of the same library (A/B experiments); there is still no fallback
LIB_PATH = os.environ.get("SFM_B200_LIB") or os.path.join(_HERE, "libsfm_b200.so")


class SfmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"sfm_b200 error {code}: {msg}")
        self.code = code


class Desc(C.Structure):
    _fields_ = [
        ("dims", C.c_int32 * 3), ("bins", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("K", C.c_float * 16), ("Kinv", C.c_float * 16),
        ("prior_err_rate", C.c_float), ("duplicate_thresh", C.c_float), ("presence_thresh", C.c_float),
        ("accept_factor", C.c_float), ("depth_scale", C.c_float), ("trunc_voxels", C.c_float),
        ("near_gate", C.c_float),
        ("device", C.c_int32), ("slab_z0", C.c_int32), ("slab_nz", C.c_int32), ("flags", C.c_int32),
        ("own_z0", C.c_int32), ("own_nz", C.c_int32), ("reserved", C.c_int32 * 6),
    ]


class Info(C.Structure):
    _fields_ = [
        ("dims", C.c_int32 * 3), ("bins", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("slab_z0", C.c_int32), ("slab_nz", C.c_int32),
        ("vol_start", C.c_float * 3), ("vol_end", C.c_float * 3), ("voxel", C.c_float * 3),
        ("miu", C.c_float), ("mean_depth", C.c_float), ("n_obs", C.c_uint32), ("num_objs", C.c_int32),
        ("initialised", C.c_int32), ("reserved", C.c_int32 * 8),
    ]


class MergeReport(C.Structure):
    _fields_ = [
        ("max_obj_now", C.c_int32), ("num_objs", C.c_int32), ("assign", C.c_int32 * 256),
        ("best_prob", C.c_float * 256), ("margin", C.c_float),
    ]


FLAG_NO_CULL = 1
FLAG_NO_TMA = 2
FLAG_GENERIC_K = 8
FLAG_SYNC_EVERY_CALL = 4
FLAG_DEBUG_ABLATE = 32
FLAG_ASYNC_SOURCES = 64
PLANE_SDF, PLANE_WEIGHT, PLANE_COLOR, PLANE_HIST = 0, 1, 2, 3

# every symbol include/sfm_b200.h declares: name -> (restype, argtypes)
_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
SYMBOLS = {
    "sfm_desc_default": (None, [C.POINTER(Desc)]),
    "sfm_last_error": (C.c_char_p, []),
    "sfm_version": (C.c_char_p, []),
    "sfm_create": (_i, [C.POINTER(Desc), C.POINTER(_vp)]),
    "sfm_destroy": (None, [_vp]),
    "sfm_init_from_frame": (_i, [_vp, _vp, _vp, _f]),
    "sfm_place_volume": (_i, [_vp, _i, _i, _vp, _f, _vp, _f, _vp, _vp, _vp, _vp]),
    "sfm_set_bounds": (_i, [_vp, _vp, _vp, _vp, _f]),
    "sfm_parse_frame": (_i, [_vp, _vp, _vp, _vp, _vp, _f]),
    "sfm_fuse_frame": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sfm_integrate_raw": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sfm_integrate_dev": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sfm_backproject": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "sfm_overlap_tables": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "sfm_merge_decide": (_i, [_vp, _vp, _vp, _vp, C.POINTER(MergeReport)]),
    "sfm_last_merge": (_i, [_vp, C.POINTER(MergeReport)]),
    "sfm_set_num_objs": (_i, [_vp, _i]),
    "sfm_mat4_inv": (_i, [_vp, _vp]),
    "sfm_mat4_mul": (None, [_vp, _vp, _vp]),
    "sfm_download": (_i, [_vp, _i, _vp, _sz]),
    "sfm_upload": (_i, [_vp, _i, _vp, _sz]),
    "sfm_plane_bytes": (_sz, [_vp, _i]),
    "sfm_plane_device_ptr": (_vp, [_vp, _i]),
    "sfm_raycast": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sfm_ray_flags": (_i, [_vp, _vp, _sz]),
    "sfm_raycast_keys_dev": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "sfm_shard_raycast_stage": (_i, [_vp, _i, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sfm_shard_halo": (_i, [_vp]),
    "sfm_keys_to_bgr": (_i, [_vp, _vp, _i, _i, _vp]),
    "sfm_show": (_i, [_vp, _f, _f, _i, _i, _vp]),
    "sfm_orbit_camera": (None, [_vp, _f, _f, _vp, _vp]),
    "sfm_palette": (None, [_vp, _i]),
    "sfm_get_info": (_i, [_vp, C.POINTER(Info)]),
    "sfm_synchronize": (_i, [_vp]),
    "sfm_wait_uploads": (_i, [_vp]),
    "sfm_planes_written": (_i, [_vp]),
    "sfm_hist_export_dev": (_i, [_vp, _vp]),
    "sfm_sdf_planes_dev": (_i, [_vp, _i, _i, _vp, _i]),
    "sfm_rebuild_skip_map": (_i, [_vp]),
    "sfm_raycast_band_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sfm_label_hits_dev": (_i, [_vp, _vp, _i, _i, _vp]),
    "sfm_part_rows": (_i, [_i, _i]),
    "sfm_raycast_part_dev": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "sfm_label_hits_parts_dev": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "sfm_ray_stats": (_i, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "sfm_set_stream": (_i, [_vp, _vp]),
    "sfm_timer_start": (_i, [_vp]),
    "sfm_timer_stop": (_i, [_vp, C.POINTER(_f)]),
    "sfm_launch_count": (C.c_uint64, [_vp]),
    "sfm_last_integrate_ms": (_i, [_vp, C.POINTER(_f)]),
    "sfm_integrate_times": (_i, [_vp, _vp, _i]),
    "sfm_frame_stats": (_i, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "sfm_stats_begin": (_i, [_vp, C.POINTER(C.c_uint64)]),
    "sfm_raycast_color": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "sfm_show_color": (_i, [_vp, _f, _f, _i, _i, _vp]),
    "sfm_extract_surface": (_i, [_vp, C.c_uint32, _vp, _vp, _vp, C.POINTER(C.c_uint32)]),
    "sfm_shard_backproj_stage": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "sfm_shard_first_frame": (_i, [_vp, _vp]),
    "sfm_fold_table_bytes": (_i, [_i, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "sfm_shard_fold": (_i, [_vp, _vp, _vp, _i, _vp]),
    "sfm_shard_merge_finish": (_i, [_vp, _vp, _vp, _vp, C.POINTER(MergeReport)]),
    "sfm_integrate_dev_ready": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "sfm_integrate_times2": (_i, [_vp, _vp, _vp, _i]),
    "sfm_stats_end": (_i, [_vp, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "sfm_debug_divcheck": (_i, [_f, C.c_uint, _i, _i, _f, C.POINTER(C.c_uint64)]),
    "sfm_mean_depth": (_f, [_vp, _i]),
    "sfm_parse_extrinsic": (None, [_vp, _vp]),
    "sfm_interpolate_pose": (None, [_vp, _vp, C.c_double, _vp]),
}

_lib = None


def load():
    """Load libsfm_b200.so and bind every symbol; raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make` or `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU / Python fallback for the CUDA path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SfmError(rc, load().sfm_last_error().decode(errors="replace"))
