"""Volume placement (the init branch of TSDF::parse_frame, tsdf.cu:173-199) through the C-ABI (sfm_place_volume, the
host half of sfm_init_from_frame) against an independent restatement built from the reference's own OpenCV calls with
Python OpenCV 4.13 as the stand-in for the absent C++ OpenCV:
    depth.convertTo(CV_8UC1) (saturating)  ->  cv2.findNonZero  ->  cv2.boundingRect  ->  Kinv * (x, y, 1, 1) as a float
    matrix product  ->  * mean_depth  ->  half the XY diagonal  ->  centre -/+ half  ->  cv2.divide by (dim - 1).
Tolerance: the bounding rectangle (integers) must be equal; start / end within 4 ulp OF THE LARGER BOUND of the axis (they are
centre -/+ half_side: one rounding of the centre or of the half side moves a small sum by several of ITS ulps);
voxel / miu within what that leaves for (end - start)/(dim - 1) plus one rounding.  Not bit-exact by construction: cv::Mat's 4x4 float product may be evaluated with FMA or in
another order depending on the OpenCV build (ours accumulates in double), and the half side is a double sqrt rounded
once; each is at most one rounding of a float operand."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from slam_maskrcnn_b200 import place_volume, synth  # noqa: E402


def cv2_placement(depth, Kinv, mean_depth, dims):
    depth_mask = cv2.convertScaleAbs(depth) if depth.max() < 2 ** 15 else np.clip(depth, 0, 255).astype(np.uint8)  # saturating cast
    assert ((depth_mask != 0) == (depth != 0)).all()
    x, y, w, h = cv2.boundingRect(cv2.findNonZero(depth_mask))
    Kinv = np.asarray(Kinv, np.float32)
    tl = cv2.gemm(Kinv, np.array([[x], [y], [1], [1]], np.float32), 1.0, None, 0.0)
    br = cv2.gemm(Kinv, np.array([[x + w], [y + h], [1], [1]], np.float32), 1.0, None, 0.0)
    tl = (tl * np.float32(mean_depth)).astype(np.float32).reshape(-1)
    br = (br * np.float32(mean_depth)).astype(np.float32).reshape(-1)
    half = np.float32(np.sqrt(float(tl[0] - br[0]) ** 2 + float(tl[1] - br[1]) ** 2) / 2)
    center = ((tl + br) / np.float32(2)).astype(np.float32)[:3]
    start, end = (center - half).astype(np.float32), (center + half).astype(np.float32)
    voxel = cv2.divide((end - start).reshape(1, 3), (np.array(dims, np.float32) - 1).reshape(1, 3)).reshape(-1).astype(np.float32)
    return (x, y, w, h), start, end, voxel, np.float32(5) * voxel[0]


def ulps(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))


@pytest.mark.parametrize("seed", range(24))
def test_placement_matches_opencv_restatement(seed):
    rng = np.random.default_rng(seed)
    if seed % 3 == 0:  # a rendered scene with clustered holes
        sc = synth.SynthScene(n_instances=6, seed=seed, hole_model="tum")
        depth = sc.frame(0)["depth"]
    else:              # random valid region inside the image, random depths (some below 256: the saturating cast matters)
        depth = np.zeros((480, 640), np.uint16)
        x0, y0 = rng.integers(0, 300), rng.integers(0, 200)
        x1, y1 = rng.integers(x0 + 20, 640), rng.integers(y0 + 20, 480)
        depth[y0:y1, x0:x1] = rng.integers(1, 20000, (y1 - y0, x1 - x0))
        depth[rng.random(depth.shape) < 0.3] = 0
        depth[y0, x0] = 300 if seed % 2 else 256  # multiples of 256 vanish under a WRAPPING cast (TSDF_Python), not here
        depth[y1 - 1, x1 - 1] = 512
    K = synth.intrinsic_matrix()
    Kinv = synth.intrinsic_inverse(K)
    md = synth.mean_depth(depth)
    dims = [(256, 256, 256), (512, 512, 512), (128, 96, 200)][seed % 3]
    rect, s_ref, e_ref, v_ref, miu_ref = cv2_placement(depth, Kinv, md, dims)
    start, end, voxel, miu = place_volume(depth, Kinv, md, dims)
    ys, xs = np.nonzero(depth)
    assert rect == (xs.min(), ys.min(), xs.max() - xs.min() + 1, ys.max() - ys.min() + 1)
    tol = 4 * np.spacing(np.maximum(np.abs(s_ref), np.abs(e_ref)).astype(np.float32))  # 4 ulp of the larger bound, per axis
    assert (np.abs(start - s_ref) <= tol).all() and (np.abs(end - e_ref) <= tol).all(), (start, s_ref, end, e_ref)
    vtol = 2 * tol / (np.array(dims, np.float32) - 1) + np.spacing(v_ref)  # what the bounds' tolerance leaves for (end - start)/(dim - 1)
    assert (np.abs(voxel - v_ref) <= vtol).all() and abs(miu - miu_ref) <= 5 * vtol[0] + np.spacing(miu_ref)
    # and the Python mirror used by the parity tests is the same rule
    s2, e2, v2, m2 = synth.place_volume(depth, Kinv, md, dims)
    assert (np.abs(start - s2) <= tol).all() and (np.abs(end - e2) <= tol).all() and (np.abs(voxel - v2) <= vtol).all() and abs(miu - m2) <= 5 * vtol[0] + np.spacing(miu_ref)
