"""Duplicate-instance merge over z-slabs (tsdf.cu:426-461 on a sharded volume).  N slab handles on one GPU
emulate N ranks; the MIN / SUM all-reduces between the steps are torch.minimum / integer adds.  Every
frame's label map, relabelled mask, num_objs and best probabilities, and in the end every stored plane,
must equal what sfm_fuse_frame produces on the whole volume -- which test_gpu_raymarch.py pins to the
reference's own back_proj_kernel + filter_overlaps."""
import numpy as np
import pytest

from tests.common import Scenario, bits

pytestmark = pytest.mark.gpu


def make_slabs(sc, world, plan=None):
    from slam_maskrcnn_b200 import Volume
    from slam_maskrcnn_b200.slabs import slab_range, shard_halo, stored_range
    halo = shard_halo(sc.voxel)
    out = []
    for r in range(world):
        z0, nz = plan[r] if plan else slab_range(r, world, sc.dims[2])
        sz0, snz = stored_range(z0, nz, sc.dims[2], halo)
        v = Volume(dims=sc.dims, bins=sc.bins, width=sc.W, height=sc.H, intrinsics=sc.intr, K=sc.K, Kinv=sc.Kinv,
                   slab=(sz0, snz), own=(z0, nz))
        v.set_bounds(sc.start, sc.end, sc.voxel, sc.miu)
        out.append((v, sz0, snz))
    return out


@pytest.mark.parametrize("world,dims,bins,ninst,plan", [
    (2, (64, 64, 64), 16, 4, None),
    (3, (64, 64, 96), 32, 6, None),
    (4, (64, 64, 96), 32, 6, [(0, 40), (40, 16), (56, 8), (64, 32)]),  # uneven, thin slabs around the surfaces
])
def test_sharded_merge_equals_single_volume(world, dims, bins, ninst, plan):
    sc = Scenario(dims=dims, bins=bins, n_instances=ninst, frames=6, yaw_step_deg=2.0, permute=True)
    full = sc.make_volume()
    slabs = make_slabs(sc, world, plan)
    assert fuse_whole_and_sharded(sc, full, slabs) == len(sc.frames) - 1
    ref = {k: full.download(k) for k in ("sdf", "weight", "color", "hist")}
    assert ref["hist"].sum() > 0
    for v, sz0, snz in slabs:
        assert v.info().num_objs == full.info().num_objs
        for k in ref:
            got, want = v.download(k), np.ascontiguousarray(ref[k][:, :, sz0:sz0 + snz])
            same = (bits(got) == bits(want)) if k == "sdf" else (got == want)
            assert same.all(), f"slab [{sz0},{sz0 + snz}) plane {k} differs after the sharded merge"
        v.close()
    full.close()


def fuse_whole_and_sharded(sc, full, slabs):
    """Every frame of `sc` through sfm_fuse_frame on `full` and through the sharded sequence on `slabs` (emulated ranks);
    asserts equal relabelled masks and merge reports frame by frame; returns the number of merged frames."""
    import torch
    n = sc.W * sc.H
    n64, ntot = full.fold_table_bytes()
    merged_frames = 0
    for i, fr in enumerate(sc.frames):
        E = fr["extrinsic"]
        mask_full = fr["mask"].copy()
        full.fuse_frame(fr["depth"], fr["color"], mask_full, E)
        d_depth = torch.from_numpy(fr["depth"].view(np.int16)).cuda()
        d_color = torch.from_numpy(fr["color"]).cuda()
        d_masks = [torch.from_numpy(fr["mask"].copy()).cuda() for _ in slabs]
        if i == 0:
            for (v, *_), dm in zip(slabs, d_masks):
                v.shard_first_frame(dm.data_ptr())
        else:
            def stage(k, ev1, ev2):
                outs = []
                for v, *_ in slabs:
                    o = torch.empty(n, dtype=torch.int64, device="cuda")
                    v.shard_backproj_stage(k, E, ev1.data_ptr() if ev1 is not None else 0, ev2.data_ptr() if ev2 is not None else 0, o.data_ptr())
                    v.synchronize()
                    outs.append(o)
                m = outs[0]
                for o in outs[1:]:
                    m = torch.minimum(m, o)
                return m
            ev1 = stage(1, None, None)
            ev2 = stage(2, ev1, None)
            keys = stage(3, ev1, ev2)
            tabs = []
            for r, ((v, *_), dm) in enumerate(zip(slabs, d_masks)):
                t = torch.empty(ntot, dtype=torch.uint8, device="cuda")
                v.shard_fold(dm.data_ptr(), keys.data_ptr(), r == 0, t.data_ptr())
                v.synchronize()
                tabs.append(t)
            red = torch.zeros(ntot, dtype=torch.uint8, device="cuda")
            red[:n64].view(torch.int64).copy_(sum(t[:n64].view(torch.int64) for t in tabs))
            red[n64:].view(torch.int32).copy_(sum(t[n64:].view(torch.int32) for t in tabs))
            ref_rep = full.last_merge()
            for (v, *_), dm in zip(slabs, d_masks):
                lut, rep = v.shard_merge_finish(red.data_ptr(), dm.data_ptr())
                v.synchronize()
                got = dm.cpu().numpy().reshape(mask_full.shape)
                same = got == mask_full
                assert same.all(), f"frame {i}: sharded relabel differs from the single volume at {int((~same).sum())} pixels"
                assert rep.num_objs == ref_rep.num_objs and rep.max_obj_now == ref_rep.max_obj_now
                assert list(rep.assign) == list(ref_rep.assign)
                assert (np.array(rep.best_prob, np.float32).view(np.uint32) == np.array(ref_rep.best_prob, np.float32).view(np.uint32)).all()
                assert (lut[fr["mask"]] == mask_full).all()
            merged_frames += 1
        for (v, *_), dm in zip(slabs, d_masks):
            v.integrate_dev(d_depth.data_ptr(), d_color.data_ptr(), dm.data_ptr(), E)
            v.synchronize()
    return merged_frames
