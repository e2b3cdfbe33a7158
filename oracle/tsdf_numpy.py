"""TEST / BASELINE INFRASTRUCTURE ONLY -- restatement of the reference's NumPy TSDF integrate.

The reference's "NumPy integration" exists only as commented-out code
(/root/reference/src/TSDF_Python/tsdf.py:78-120; init at tsdf.py:32-52), so it cannot be
imported; this file restates it with the two modernisations NumPy 2.x forces
(`np.int` -> int, list-of-arrays indexing -> tuple indexing) and a flat 1-D voxel array instead
of the 2-D `tex_dim` packing (tsdf.py:22 is not exact for 128^3).  Float64, labels off, first
observation handled by the `weight_mask` branch -- all as written in the reference.

Role: the CPU timing baseline (`cpu_baseline.kind == "port"`, bench.py) -- NOT the parity oracle.
It differs from SfM_CUDA/tsdf.cu in precision, rounding and gating.
Parity unpinned: the reference ships no test vectors for this path and the code itself is dead
(commented out), so there is nothing to pin it against except its own text.
"""
import numpy as np


class NumpyTSDF:
    """tsdf.py:10-52 (state) and tsdf.py:78-120 (integrate), flat layout idx = (x*D + y)*D + z."""

    def __init__(self, intrinsics, vol_dim=128):
        self.intrinsic = np.eye(4, dtype=np.float32)                      # tsdf.py:12-13
        self.intrinsic[[0, 1, 0, 1], [0, 1, 2, 2]] = np.array(intrinsics, dtype=np.float32)
        self.vol_dim = int(vol_dim)
        self.init = False

    def set_bounds(self, vol_start, vol_end):
        """tsdf.py:44-52 with explicit bounds (the bbox/mean-depth placement is tsdf.py:35-43)."""
        self.vol_start = np.asarray(vol_start, dtype=np.float64)
        self.vol_end = np.asarray(vol_end, dtype=np.float64)
        self.voxel = (self.vol_end - self.vol_start) / (self.vol_dim - 1)  # tsdf.py:46
        self.mu = 5 * self.voxel[0]                                        # tsdf.py:47
        n = self.vol_dim ** 3
        self.tsdf_diff = np.ones(n, np.float32) * np.float32(self.mu)      # tsdf.py:48
        self.tsdf_wt = np.zeros(n, np.int32)                               # tsdf.py:49
        self.tsdf_color = np.zeros((n, 3), np.int32)                       # tsdf.py:50
        self.init = True

    def integrate(self, depth, color, extrinsic2init, x_range=None):
        """tsdf.py:78-120.  `x_range=(x0,x1)` restricts the pass to a slab of voxel x-planes
        (used only to bound the timing sample at large volumes); None = whole volume."""
        D = self.vol_dim
        x0, x1 = (0, D) if x_range is None else x_range
        flattened_idx = np.arange(x0 * D * D, x1 * D * D)                  # tsdf.py:78-79
        x_idx = flattened_idx // (D * D)                                   # tsdf.py:80
        y_idx = flattened_idx // D - x_idx * D                             # tsdf.py:81
        z_idx = flattened_idx % D                                          # tsdf.py:82
        pos_inhomo = self.vol_start + np.stack([x_idx, y_idx, z_idx], axis=-1) * self.voxel  # :83
        pos_homo = np.concatenate([pos_inhomo, np.ones([pos_inhomo.shape[0], 1])], axis=-1)  # :84
        proj = np.dot(extrinsic2init, pos_homo.transpose())                # tsdf.py:86
        pixel = np.dot(self.intrinsic, proj)                               # tsdf.py:87
        with np.errstate(divide="ignore", invalid="ignore"):
            pixel /= pixel[2, :]                                           # tsdf.py:88
        pixel = pixel.transpose()
        with np.errstate(invalid="ignore"):
            x = np.nan_to_num(pixel[:, 0], nan=-1.0, posinf=-1.0, neginf=-1.0).astype(int)   # :92
            y = np.nan_to_num(pixel[:, 1], nan=-1.0, posinf=-1.0, neginf=-1.0).astype(int)   # :93
        mask = (x >= 0) & (x <= color.shape[1] - 1) & (y >= 0) & (y <= color.shape[0] - 1)   # :97
        idx = (np.minimum(np.maximum(y, 0), color.shape[0] - 1),
               np.minimum(np.maximum(x, 0), color.shape[1] - 1))          # tsdf.py:99-100
        diff = depth[idx] / 5000 - proj[2, :]                              # tsdf.py:101
        mask &= (depth[idx] > 0)                                           # tsdf.py:102
        diff = np.maximum(np.minimum(diff, self.mu), -self.mu) / self.mu   # tsdf.py:104
        mask &= diff > -1                                                  # tsdf.py:105
        weight = 1
        sl = slice(x0 * D * D, x1 * D * D)
        wt = self.tsdf_wt[sl]
        col = self.tsdf_color[sl]
        dif = self.tsdf_diff[sl]
        weight_mask = wt > 0                                               # tsdf.py:110
        a = mask & weight_mask
        dif[a] = (dif[a] * wt[a] + weight * diff[a]) / (wt[a] + weight)    # tsdf.py:111
        col[a] = (col[a] * np.expand_dims(wt[a], -1) + weight * color[idx][a]) \
            / np.expand_dims(wt[a] + weight, -1)                           # tsdf.py:112-113
        b = mask & ~weight_mask
        dif[b] = weight * diff[b]                                          # tsdf.py:115
        col[b] = weight * color[idx][b]                                    # tsdf.py:116
        wt[mask] = wt[mask] + weight                                       # tsdf.py:117
        return int(mask.sum())
