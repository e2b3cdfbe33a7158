"""slam-maskrcnn_b200: B200-native TSDF fusion + ray-cast hot path of qq456cvb/SLAM-MaskRCNN.

Only what the path needs: csrc/ (sm_100a kernels + the C-ABI, built into libsfm_b200.so), the
host-side mirror of the reference's `TSDF` / `Viewer` interface (tsdf.py), the z-slab multi-GPU
driver (slabs.py) and the synthetic TUM-shaped input generator (synth.py).
"""
from .tsdf import TSDF, Viewer, Volume, MAX_OBJECTS, place_volume, orbit_camera, palette, mean_depth, parse_extrinsic, interpolate_pose, write_ply  # noqa: F401
from ._lib import SfmError, FLAG_NO_CULL, FLAG_NO_TMA, FLAG_GENERIC_K, FLAG_SYNC_EVERY_CALL, FLAG_DEBUG_ABLATE, FLAG_ASYNC_SOURCES  # noqa: F401
