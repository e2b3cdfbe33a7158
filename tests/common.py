"""Shared scenario builders for the parity tests."""
import numpy as np

from slam_maskrcnn_b200 import synth


class Scenario:
    """A small synthetic sequence + volume placement, identical bits for every implementation."""

    def __init__(self, dims=(64, 64, 64), bins=16, width=160, height=120, n_instances=6, frames=5, seed=0,
                 permute=False, yaw_step_deg=1.0):
        self.dims, self.bins = tuple(dims), bins
        self.scene = synth.small_scene(width, height, n_instances, seed=seed, permute=permute, yaw_step_deg=yaw_step_deg) \
            if width != synth.W else synth.SynthScene(n_instances, seed=seed, permute=permute, yaw_step_deg=yaw_step_deg)
        sc = self.scene
        self.W, self.H = sc.W, sc.H
        self.intr = (sc.fx, sc.fy, sc.cx, sc.cy)
        self.K = synth.intrinsic_matrix(*self.intr)
        self.Kinv = synth.intrinsic_inverse(self.K)
        f0 = sc.frame(0)
        self.mean_depth = synth.mean_depth(f0["depth"])
        self.start, self.end, self.voxel, self.miu = synth.place_volume(f0["depth"], self.Kinv, self.mean_depth, self.dims)
        self.frames = [sc.frame(f) for f in range(1, frames + 1)]  # frame 0 only initialises (tsdf.cu:213)

    def make_volume(self, flags=0, slab=None, bins=None, device=0):
        from slam_maskrcnn_b200 import Volume
        v = Volume(dims=self.dims, bins=self.bins if bins is None else bins, width=self.W, height=self.H,
                   intrinsics=self.intr, K=self.K, Kinv=self.Kinv, flags=flags, slab=slab, device=device)
        v.set_bounds(self.start, self.end, self.voxel, self.miu)
        return v

    def make_cpu_volume(self, bins=None):
        from oracle import binding as ob
        return ob.CpuVolume(self.dims, self.bins if bins is None else bins, self.start, self.end, self.voxel, self.miu)


def backproj_camera(E):
    """Rt = R^T, o = -Rt t of extrinsic2init (tsdf.cu:432-435), float32 with double accumulation as the library does."""
    E = np.asarray(E, np.float32).reshape(4, 4)
    Rt = E[:3, :3].T.copy()
    o = (-(Rt.astype(np.float64)) @ E[:3, 3].astype(np.float64)).astype(np.float32)
    return Rt, o


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


class _DevMem:
    """Raw device memory as a __cuda_array_interface__ object (so torch can view a library-owned plane)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_plane(vol, name):
    """torch view (no copy) of one plane of a Volume, shaped like `download(name)`."""
    import torch
    _, dt, tail = vol._PLANE[name]
    tail = (vol.bins,) if tail is None else tail
    t = torch.as_tensor(_DevMem(vol.plane_ptr(name), vol.plane_bytes(name)), device="cuda")
    tdt = {np.float32: torch.float32, np.int32: torch.int32, np.uint8: torch.uint8, np.uint32: torch.int32}[dt]
    return t.view(tdt).reshape(vol.local_shape + tail)
