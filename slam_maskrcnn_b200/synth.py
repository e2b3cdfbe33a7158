"""Synthetic TUM-Freiburg-2-shaped sequences: the input side of the path.

There is no network for datasets, so tests and bench.py feed the hot path with analytic scenes that
have the shape of the reference's inputs (kernel.cpp:39-99): 640x480 uint16 depth at 1/5000 m with
invalid (0) pixels, BGR uint8 colour, uint8 Mask R-CNN-style label images whose instance ids are
permuted per frame (so the duplicate-instance merge has work to do) with `prior_mrcnn_err_rate`
(5 %) of instance pixels flipped, and world->camera poses on an arc, also as TUM groundtruth lines.

Scene (metres, first-camera frame: x right, y down, z forward): a room (back wall z=3.5, floor
y=+1.1, ceiling y=-1.3, side walls x=+-1.9) holding K instance spheres.
"""
import numpy as np

FX, FY, CX, CY = 520.9, 521.0, 325.1, 249.7  # kernel.cpp:39
W, H = 640, 480


def intrinsic_matrix(fx=FX, fy=FY, cx=CX, cy=CY):
    K = np.eye(4, dtype=np.float32)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = fx, fy, cx, cy
    return K


def intrinsic_inverse(K):
    """float32 inverse of the 4x4 intrinsic matrix (tsdf.cu:147 uses cv::Mat::inv)."""
    return np.linalg.inv(K.astype(np.float64)).astype(np.float32)


def place_volume(depth0, Kinv, mean_depth, dims, trunc_voxels=5.0):
    """Volume placement rule of the first parse_frame call (tsdf.cu:180-199) in float32.
    Returns (vol_start, vol_end, voxel, miu) as float32 -- fed to BOTH the library
    (sfm_set_bounds) and the oracle so the two sides see identical bits."""
    ys, xs = np.nonzero(depth0)
    x0, x1, y0, y1 = xs.min(), xs.max() + 1, ys.min(), ys.max() + 1  # cv::Rect tl / br
    Kinv = Kinv.astype(np.float32)
    md = np.float32(mean_depth)
    tl = (Kinv @ np.array([x0, y0, 1, 1], np.float32)).astype(np.float32) * md
    br = (Kinv @ np.array([x1, y1, 1, 1], np.float32)).astype(np.float32) * md
    half = np.float32(np.sqrt(float(tl[0] - br[0]) ** 2 + float(tl[1] - br[1]) ** 2) / 2)
    center = ((tl + br) / np.float32(2))[:3]
    start = (center - half).astype(np.float32)
    end = (center + half).astype(np.float32)
    voxel = ((end - start) / (np.array(dims, np.float32) - np.float32(1))).astype(np.float32)
    miu = np.float32(trunc_voxels) * voxel[0]
    return start, end, voxel, np.float32(miu)


def mean_depth(depth):
    """utils.cu:77-91."""
    d = depth[depth != 0].astype(np.float64) / 5000.0
    return np.float32(d.sum() / d.size)


class SynthScene:
    def __init__(self, n_instances=15, seed=0, width=W, height=H, intrinsics=(FX, FY, CX, CY),
                 yaw_step_deg=0.1, hole_frac=0.15, flip_frac=0.05, min_pixels=2000, permute=True,
                 hole_model="tum"):
        self.K_inst = int(n_instances)
        self.W, self.H = int(width), int(height)
        self.fx, self.fy, self.cx, self.cy = [float(v) for v in intrinsics]
        self.yaw_step = np.deg2rad(yaw_step_deg)
        self.hole_frac, self.flip_frac, self.min_pixels = hole_frac, flip_frac, min_pixels
        self.permute = permute
        # "tum": spatially clustered invalid depth (blobs + occlusion shadows + 0.5 % speckle), calibrated
        #        to the two real TUM fr2 depth frames the reference ships (Mask_RCNN/samples/1311871965.993806.png:
        #        15.5 % zeros, 77 % of 8x8 tiles hole-free, 12 % all-hole, 11 % mixed);
        # "salt": independent per-pixel holes (SURVEY 8d's first-cut spec; unlike any real Kinect frame).
        self.hole_model = hole_model
        rng = np.random.default_rng(seed)
        k = self.K_inst
        self.centers = np.stack([rng.uniform(-1.2, 1.2, k), rng.uniform(0.3, 0.9, k), rng.uniform(1.6, 3.1, k)], 1)
        self.radii = rng.uniform(0.2, 0.4, k)
        self.sphere_bgr = rng.integers(40, 256, (k, 3))
        self.plane_bgr = np.array([[200, 200, 190], [90, 110, 130], [230, 230, 230], [150, 170, 150], [150, 150, 175]])
        self.orbit_center = np.array([0.0, 0.0, 2.5])
        u = (np.arange(self.W) + 0.5 - self.cx) / self.fx
        v = (np.arange(self.H) + 0.5 - self.cy) / self.fy
        uu, vv = np.meshgrid(u, v)
        self.rays_cam = np.stack([uu, vv, np.ones_like(uu)], -1)  # z = 1 => t is camera depth

    # -- poses ------------------------------------------------------------------------------
    def cam_to_world(self, f):
        th = self.yaw_step * f
        c, s = np.cos(th), np.sin(th)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        p = self.orbit_center - R @ self.orbit_center
        return R, p

    def extrinsic(self, f):
        """world->camera 4x4 float32 (what parse_extrinsic returns, utils.cu:8-24)."""
        R, p = self.cam_to_world(f)
        E = np.eye(4)
        E[:3, :3] = R.T
        E[:3, 3] = -R.T @ p
        return E.astype(np.float32)

    def tum_pose(self, f, t0=1311868164.0, dt=1.0 / 30):
        """groundtruth.txt line: ts tx ty tz qx qy qz qw (camera-to-world)."""
        _, p = self.cam_to_world(f)
        th = self.yaw_step * f
        return np.array([t0 + f * dt, p[0], p[1], p[2], 0.0, np.sin(th / 2), 0.0, np.cos(th / 2)])

    # -- rendering --------------------------------------------------------------------------
    def frame(self, f):
        R, p = self.cam_to_world(f)
        d = self.rays_cam @ R.T  # world directions, camera z-depth parametrisation
        best = np.full((self.H, self.W), np.inf)
        surf = np.full((self.H, self.W), -1, np.int32)  # >=0 plane id, <-1 sphere
        planes = [(2, 3.5), (1, 1.1), (1, -1.3), (0, 1.9), (0, -1.9)]
        with np.errstate(divide="ignore", invalid="ignore"):
            for pid, (ax, val) in enumerate(planes):
                t = (val - p[ax]) / d[..., ax]
                ok = (t > 1e-6) & (t < best)
                best = np.where(ok, t, best)
                surf = np.where(ok, pid, surf)
            inst = np.zeros((self.H, self.W), np.int32)
            a = (d * d).sum(-1)
            for k in range(self.K_inst):
                oc = p - self.centers[k]
                b = d @ oc
                c = oc @ oc - self.radii[k] ** 2
                disc = b * b - a * c
                t = (-b - np.sqrt(np.maximum(disc, 0))) / a
                ok = (disc > 0) & (t > 1e-6) & (t < best)
                best = np.where(ok, t, best)
                inst = np.where(ok, k + 1, inst)
        rng = np.random.default_rng(1000 + f)
        depth = np.clip(best * 5000.0, 0, 65535).astype(np.uint16)
        depth[~np.isfinite(best)] = 0
        depth[self.hole_mask(rng, inst, surf, best)] = 0
        base = np.where((inst > 0)[..., None], self.sphere_bgr[np.maximum(inst - 1, 0)], self.plane_bgr[np.maximum(surf, 0)])
        color = np.clip(base + rng.integers(-12, 13, base.shape), 0, 255).astype(np.uint8)
        # labels: drop small instances (dmask.py:34-45), permute ids per frame, flip 5 % of instance pixels
        gt = inst.copy()
        ids, counts = np.unique(gt, return_counts=True)
        for i, cnt in zip(ids, counts):
            if i > 0 and cnt < self.min_pixels:
                gt[gt == i] = 0
        mask = gt.copy()
        if self.K_inst > 0:
            perm = np.arange(self.K_inst + 1)
            if self.permute:
                perm[1:] = rng.permutation(self.K_inst) + 1
            mask = perm[mask]
            flip = (gt > 0) & (rng.random((self.H, self.W)) < self.flip_frac)
            wrong = rng.integers(1, self.K_inst + 1, (self.H, self.W))
            mask = np.where(flip, wrong, mask)
        return {
            "depth": depth, "color": color, "mask": mask.astype(np.uint8), "gt": gt.astype(np.uint8),
            "extrinsic": self.extrinsic(f), "pose": self.tum_pose(f),
        }


def _hole_mask(self, rng, inst, surf, best=None):
    if self.hole_frac <= 0:
        return np.zeros((self.H, self.W), bool)
    if self.hole_model == "salt":
        return rng.random((self.H, self.W)) < self.hole_frac
    from scipy import ndimage
    scale = self.W / W
    # occlusion shadows: a thin band of invalid depth on depth discontinuities (> 10 cm jumps)
    edges = np.zeros((self.H, self.W), bool)
    if best is not None:
        z = np.where(np.isfinite(best), best, 0.0)
        edges[:, 1:] |= np.abs(z[:, 1:] - z[:, :-1]) > 0.1
        edges[1:, :] |= np.abs(z[1:, :] - z[:-1, :]) > 0.1
    shadow = ndimage.binary_dilation(edges, iterations=1) if scale > 0.5 else edges
    speckle = rng.random((self.H, self.W)) < 0.001
    blob_frac = max(self.hole_frac - float(shadow.mean()) - 0.001, 0.0)
    field = ndimage.gaussian_filter(rng.standard_normal((self.H, self.W)), sigma=28 * scale)
    blobs = field > np.quantile(field, 1.0 - blob_frac) if blob_frac > 0 else np.zeros_like(shadow)
    return shadow | speckle | blobs


SynthScene.hole_mask = _hole_mask


def small_scene(width=160, height=120, n_instances=6, **kw):
    """A down-scaled camera (same field of view) for fast tests."""
    s = width / W
    return SynthScene(n_instances=n_instances, width=width, height=height,
                      intrinsics=(FX * s, FY * s, CX * s, CY * s), min_pixels=int(2000 * s * s), **kw)
