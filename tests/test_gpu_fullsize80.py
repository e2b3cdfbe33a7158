"""The headline configuration at full size: 512^3 voxels x 80 bins (21.5 GB of histogram), 640x480 labelled frames,
duplicate-instance merge ON.  The reference cannot run here (its int index overflows above 406^3 x 32), so
the checks are the size-independent ones, all on the device:
  * sfm_fuse_frame on the whole volume == the sharded merge over two z-slabs (emulated ranks): relabelled masks, merge
    reports and num_objs frame by frame, and in the end every plane of every slab, byte for byte -- the small-size
    versions of both sides are pinned to the verbatim reference pipeline (test_gpu_raymarch.py, test_gpu_bins80.py);
  * a voxel's label counts never exceed its observation count;
  * the 1280x960 ray-cast of the whole volume == the replicated-SDF composite == the exact three-stage composite,
    key for key (hit, refined t bits, label)."""
import numpy as np
import pytest

from tests.common import Scenario, device_plane
from tests.test_gpu_sharded_merge import fuse_whole_and_sharded, make_slabs

pytestmark = pytest.mark.gpu

DIMS, BINS = (512, 512, 512), 80
NO_HIT = np.iinfo(np.int64).max


def test_full_size_80_bins_merge_and_raycast_whole_vs_slabs():
    import torch
    from slam_maskrcnn_b200 import Volume, orbit_camera
    from slam_maskrcnn_b200.slabs import keys_to_int64, slab_range
    if torch.cuda.mem_get_info()[1] < 150e9:
        pytest.skip("needs ~110 GB of device memory")
    world = 2
    sc = Scenario(dims=DIMS, bins=BINS, width=640, height=480, n_instances=8, frames=4, yaw_step_deg=2.0, permute=True)
    full = sc.make_volume()
    slabs = make_slabs(sc, world)
    assert fuse_whole_and_sharded(sc, full, slabs) == len(sc.frames) - 1
    assert full.info().num_objs >= 9
    own = [slab_range(r, world, DIMS[2]) for r in range(world)]

    # ---- ray-cast: whole volume vs the two multi-GPU composites --------------------------------
    w, h = 1280, 960
    s2w, c = orbit_camera(sc.Kinv, 0.35, float(sc.mean_depth))
    ref = torch.empty(w * h, dtype=torch.int64, device="cuda")
    full.raycast_keys_dev(s2w, c, w, h, ref.data_ptr())
    full.synchronize()
    ref = keys_to_int64(ref)
    hit = ref != NO_HIT
    assert hit.float().mean() > 0.15
    assert len(torch.unique(ref[hit] & 0xff)) >= 5, "several instance labels should be visible"

    rep = Volume(dims=DIMS, bins=0, width=sc.W, height=sc.H, intrinsics=sc.intr, K=sc.K, Kinv=sc.Kinv)
    rep.set_bounds(sc.start, sc.end, sc.voxel, sc.miu)
    cols = DIMS[0] * DIMS[1]
    for (v, _, _), (z0, nz) in zip(slabs, own):
        buf = torch.empty(cols * nz, dtype=torch.float32, device="cuda")
        v.sdf_planes_dev(z0, nz, buf.data_ptr(), True)
        v.synchronize()
        rep.sdf_planes_dev(z0, nz, buf.data_ptr(), False)
        rep.synchronize()
        del buf
    assert torch.equal(device_plane(rep, "sdf").view(torch.int32), device_plane(full, "sdf").view(torch.int32))
    rep.rebuild_skip_map()
    prow = rep.part_rows(h, world)
    hits = torch.zeros(world * prow * w * 4, dtype=torch.float32, device="cuda")
    for r in range(world):
        rep.raycast_part_dev(s2w, c, w, h, r, world, hits[r * prow * w * 4:].data_ptr())
    rep.synchronize()
    keys = None
    for v, *_ in slabs:
        k = torch.empty(w * h, dtype=torch.int64, device="cuda")
        v.label_hits_parts_dev(hits.data_ptr(), w, h, world, k.data_ptr())
        v.synchronize()
        keys = k if keys is None else torch.minimum(keys, k)
    same = keys == ref
    assert bool(same.all()), f"replicated-SDF composite: {int((~same).sum())} of {w * h} rays differ from the whole volume"
    rep.close()

    ev = [None, None]
    for stage in (1, 2, 3):
        red = None
        for v, *_ in slabs:
            o = torch.empty(w * h, dtype=torch.int64, device="cuda")
            v.shard_raycast_stage(stage, s2w, c, w, h, ev[0].data_ptr() if ev[0] is not None else 0,
                                  ev[1].data_ptr() if ev[1] is not None else 0, o.data_ptr())
            v.synchronize()
            red = o if red is None else torch.minimum(red, o)
        if stage < 3:
            ev[stage - 1] = red
    same = red == ref
    assert bool(same.all()), f"three-stage composite: {int((~same).sum())} of {w * h} rays differ from the whole volume"

    # ---- planes: every slab == the same planes of the whole volume, on the device, in x chunks ----
    whole = {k: device_plane(full, k) for k in ("sdf", "weight", "color", "hist")}  # "hist": reference-layout snapshot
    assert int(whole["hist"][:64].sum(dtype=torch.int64)) + int(whole["hist"][256:320].sum(dtype=torch.int64)) > 0
    step = 64
    for xs in range(0, DIMS[0], step):
        ok = whole["hist"][xs:xs + step].sum(dim=3) <= whole["weight"][xs:xs + step]
        assert bool(ok.all()), "a voxel's label counts exceed its observation count"
    for v, sz0, snz in slabs:
        assert v.info().num_objs == full.info().num_objs
        for k, ref_plane in whole.items():
            got = device_plane(v, k)
            for xs in range(0, DIMS[0], step):
                a, b = got[xs:xs + step], ref_plane[xs:xs + step, :, sz0:sz0 + snz]
                if k == "sdf":
                    a, b = a.view(torch.int32), b.contiguous().view(torch.int32)
                assert bool((a == b).all()), f"slab [{sz0},{sz0 + snz}) plane {k} differs from the whole volume (x in [{xs},{xs + step}))"
        v.close()
    full.close()
