"""z-slab sharding on one GPU: N slab handles (owned planes + halo) emulate N ranks; the MIN
reductions between the stages are done with torch.minimum.  Every slab must hold bit-identical
planes to the single volume, and the composited ray-cast keys (t bits | label) must equal the
single-volume ray-cast exactly -- the N-GPU == 1-GPU gate of SURVEY.md section 8d/8e."""
import numpy as np
import pytest

from tests.common import Scenario, bits

pytestmark = pytest.mark.gpu


def build(sc, world):
    import torch
    from slam_maskrcnn_b200 import Volume
    from slam_maskrcnn_b200.slabs import slab_range, shard_halo, stored_range
    full = sc.make_volume()
    halo = shard_halo(sc.voxel)
    slabs = []
    for r in range(world):
        z0, nz = slab_range(r, world, sc.dims[2])
        sz0, snz = stored_range(z0, nz, sc.dims[2], halo)
        v = Volume(dims=sc.dims, bins=sc.bins, width=sc.W, height=sc.H, intrinsics=sc.intr, K=sc.K, Kinv=sc.Kinv,
                   slab=(sz0, snz), own=(z0, nz))
        v.set_bounds(sc.start, sc.end, sc.voxel, sc.miu)
        slabs.append((v, sz0, snz, z0, nz))
    for fr in sc.frames:
        full.integrate_raw(fr["depth"], fr["color"], fr["gt"], fr["extrinsic"])
        for v, *_ in slabs:
            v.integrate_raw(fr["depth"], fr["color"], fr["gt"], fr["extrinsic"])
    return full, slabs


@pytest.mark.parametrize("world,dims", [(2, (64, 64, 64)), (4, (48, 56, 96))])
def test_slabs_hold_identical_planes(world, dims):
    sc = Scenario(dims=dims, bins=16, frames=4, yaw_step_deg=3.0)
    full, slabs = build(sc, world)
    ref = {k: full.download(k) for k in ("sdf", "weight", "color", "hist")}
    for v, sz0, snz, z0, nz in slabs:
        for k in ref:
            got = v.download(k)
            want = ref[k][:, :, sz0:sz0 + snz]
            same = (bits(got) == bits(np.ascontiguousarray(want))) if k == "sdf" else (got == want)
            assert same.all(), f"slab [{sz0},{sz0 + snz}) plane {k} differs"
        v.close()
    full.close()


@pytest.mark.parametrize("world,dims,angle", [(2, (64, 64, 64), 0.2), (4, (64, 64, 96), 0.9), (3, (56, 48, 60), 2.6)])
def test_sharded_raycast_equals_single_volume(world, dims, angle):
    import torch
    from slam_maskrcnn_b200 import orbit_camera
    from slam_maskrcnn_b200.slabs import keys_to_int64
    sc = Scenario(dims=dims, bins=16, frames=6, yaw_step_deg=2.0)
    full, slabs = build(sc, world)
    s2w, c = orbit_camera(sc.Kinv, angle, float(sc.mean_depth))
    w, h = sc.W, sc.H
    ref = torch.empty(w * h, dtype=torch.int64, device="cuda")
    full.raycast_keys_dev(s2w, c, w, h, ref.data_ptr())
    full.synchronize()
    ref = keys_to_int64(ref)

    def reduce_stage(stage, ev1, ev2):
        outs = []
        for v, *_ in slabs:
            o = torch.empty(w * h, dtype=torch.int64, device="cuda")
            v.shard_raycast_stage(stage, s2w, c, w, h, ev1.data_ptr() if ev1 is not None else 0,
                                  ev2.data_ptr() if ev2 is not None else 0, o.data_ptr())
            v.synchronize()
            outs.append(o)
        m = outs[0]
        for o in outs[1:]:
            m = torch.minimum(m, o)
        return m, outs

    ev1, _ = reduce_stage(1, None, None)
    ev2, _ = reduce_stage(2, ev1, None)
    keys, per_rank = reduce_stage(3, ev1, ev2)
    same = (keys == ref)
    assert same.all(), f"{int((~same).sum())} of {w * h} rays differ from the single-volume ray-cast"
    hits = ref != np.iinfo(np.int64).max
    assert hits.float().mean() > 0.2
    # exactly one rank owns every hit
    owners = sum((o != np.iinfo(np.int64).max).to(torch.int32) for o in per_rank)
    assert (owners[hits] == 1).all() and (owners[~hits] == 0).all()
    if world >= 4:  # (from behind, angle 2.6, all visible surface happens to lie in one slab)
        assert sum(int((o != np.iinfo(np.int64).max).any()) for o in per_rank) >= 2, "hits should be spread over several slabs"
    for v, *_ in slabs:
        v.close()
    full.close()


def test_sharded_raycast_rejects_missing_halo():
    from slam_maskrcnn_b200 import SfmError, orbit_camera
    import torch
    sc = Scenario(dims=(32, 32, 32), bins=16, frames=1)
    v = sc.make_volume(slab=(8, 8))
    s2w, c = orbit_camera(sc.Kinv, 0.1, float(sc.mean_depth))
    out = torch.empty(sc.W * sc.H, dtype=torch.int64, device="cuda")
    with pytest.raises(SfmError):
        v.shard_raycast_stage(1, s2w, c, sc.W, sc.H, 0, 0, out.data_ptr())
    v.close()
