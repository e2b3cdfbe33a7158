"""Secondary oracle (north_star: "checked against the reference's own tsdf.cu / TSDF_Python results"): the TSDF_Python
prototype's CUDA module, src/TSDF_Python/tsdf.cu + tsdf.cpp compiled verbatim into oracle/_ref/tsdf_cuda*.so
(oracle/build_ref.py).  Its SDF and weight arithmetic is tsdf_kernel's (tsdf.cu:18-42 there == SfM_CUDA/tsdf.cu:30-56:
the only difference is a fourth K-row term K[r][3]*1 == 0), so those two planes must equal ours bit for bit on a cubic
volume with one scalar voxel size, labels off.  Its colour (int32, not gated by diff < 0.99) and its Boyer-Moore label
vote are not tsdf_kernel semantics and are not compared."""
import os
import sys

import numpy as np
import pytest

from tests.common import Scenario, bits

pytestmark = pytest.mark.gpu
REF_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def load_module():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import tsdf_cuda
    except ImportError:
        pytest.skip("oracle/_ref/tsdf_cuda not built")
    return tsdf_cuda


@pytest.mark.parametrize("D,wh,nframes", [(64, (160, 120), 5), (128, (640, 480), 3)])
def test_sdf_and_weight_equal_tsdf_python_module(D, wh, nframes):
    tsdf_cuda = load_module()
    sc = Scenario(dims=(D, D, D), bins=16, width=wh[0], height=wh[1], frames=nframes, yaw_step_deg=2.0)
    v1 = np.float32(sc.voxel[0])                      # TSDF_Python passes ONE voxel size (tsdf.py:63: self.voxel[0])
    voxel = np.array([v1, v1, v1], np.float32)
    miu = np.float32(5) * v1                            # tsdf.py:47
    from slam_maskrcnn_b200 import Volume
    ours = Volume(dims=(D, D, D), bins=0, width=sc.W, height=sc.H, intrinsics=sc.intr, K=sc.K, Kinv=sc.Kinv)
    ours.set_bounds(sc.start, sc.end, voxel, miu)
    n = D ** 3
    diff = np.full(n, miu, np.float32)                  # tsdf.py:48
    color = np.zeros(n * 3, np.int32)
    wt = np.zeros(n, np.int32)
    cls = np.zeros(n, np.int32)
    cnt = np.zeros(n, np.int32)
    start = np.ascontiguousarray(sc.start, np.float32)
    K = np.ascontiguousarray(sc.K, np.float32)
    for fr in sc.frames:
        ours.integrate_raw(fr["depth"], fr["color"], None, fr["extrinsic"])
        tsdf_cuda.tsdf_update(diff, color, wt, cls, cnt, D, start, float(v1), float(miu), K, fr["depth"], fr["color"],
                              np.ascontiguousarray(fr["gt"], np.int32), np.ascontiguousarray(fr["extrinsic"], np.float32), sc.W, sc.H)
    sdf, w = ours.download("sdf").reshape(-1), ours.download("weight").reshape(-1)
    ours.close()
    assert wt.sum() > 0 and (wt > 1).any()
    same_w = w == wt
    assert same_w.all(), f"weights differ from the TSDF_Python module at {int((~same_w).sum())} voxels"
    same_s = bits(sdf) == bits(diff)
    assert same_s.all(), f"SDF bits differ from the TSDF_Python module at {int((~same_s).sum())} voxels"
