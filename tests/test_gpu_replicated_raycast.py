"""Ray-cast of a z-slab-sharded volume with a REPLICATED SDF (BASELINE config 4; the reference's viewer runs after
its frame loop, kernel.cpp:101-107): N slab handles on one GPU emulate N ranks, a label-free full-volume handle
receives their owned SDF planes (the all-gather), its skip map is rebuilt from the SDF, every "rank" marches a band of
image rows on it, every slab labels the hits that fall into its owned planes, and the MIN over the slabs must equal the
single-volume ray-cast keys (float_bits(t) << 32 | label) bit for bit -- hits, t and labels."""
import numpy as np
import pytest

from tests.common import Scenario, bits
from tests.test_gpu_sharded_raycast import build

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,dims,bins,angle,wh", [(2, (64, 64, 64), 16, 0.2, None), (4, (64, 64, 96), 80, 0.9, None),
                                                     (3, (56, 48, 60), 16, 2.6, (200, 150)), (4, (64, 64, 64), 16, 1.3, (161, 122))])
def test_replicated_raycast_equals_single_volume(world, dims, bins, angle, wh):
    import torch
    from slam_maskrcnn_b200 import Volume, orbit_camera
    from slam_maskrcnn_b200.slabs import keys_to_int64
    sc = Scenario(dims=dims, bins=bins, n_instances=40 if bins == 80 else 6, frames=6, yaw_step_deg=2.0)
    full, slabs = build(sc, world)
    s2w, c = orbit_camera(sc.Kinv, angle, float(sc.mean_depth))
    w, h = wh or (sc.W, sc.H)
    ref = torch.empty(w * h, dtype=torch.int64, device="cuda")
    full.raycast_keys_dev(s2w, c, w, h, ref.data_ptr())
    full.synchronize()
    ref = keys_to_int64(ref)

    # the "all-gather": every slab exports its OWNED planes, the replica imports them
    rep = Volume(dims=sc.dims, bins=0, width=sc.W, height=sc.H, intrinsics=sc.intr, K=sc.K, Kinv=sc.Kinv)
    rep.set_bounds(sc.start, sc.end, sc.voxel, sc.miu)
    cols = dims[0] * dims[1]
    for v, sz0, snz, z0, nz in slabs:
        buf = torch.empty(cols * nz, dtype=torch.float32, device="cuda")
        v.sdf_planes_dev(z0, nz, buf.data_ptr(), True)
        v.synchronize()
        rep.sdf_planes_dev(z0, nz, buf.data_ptr(), False)
        rep.synchronize()
    assert (bits(rep.download("sdf")) == bits(full.download("sdf"))).all(), "replica SDF != single-volume SDF"
    rep.rebuild_skip_map()

    # bands of rows, one per "rank" (the last one shorter when h % world != 0)
    rows = (h + world - 1) // world
    hits = torch.zeros(world * rows * w * 4, dtype=torch.float32, device="cuda")
    for r in range(world):
        row0 = r * rows
        n = max(0, min(rows, h - row0))
        if n:
            rep.raycast_band_dev(s2w, c, w, h, row0, n, hits.data_ptr())
    rep.synchronize()
    t_rep = hits[:w * h * 4].view(h, w, 4)[..., 3].cpu().numpy()
    t_ref = np.where(ref.cpu().numpy() == np.iinfo(np.int64).max, 0, ref.cpu().numpy() >> 32).astype(np.uint32).reshape(h, w)
    assert (bits(t_rep) == t_ref).all(), "march on the replica (rebuilt skip map) != single-volume march"

    keys = None
    owners = torch.zeros(w * h, dtype=torch.int32, device="cuda")
    for v, *_ in slabs:
        k = torch.empty(w * h, dtype=torch.int64, device="cuda")
        v.label_hits_dev(hits.data_ptr(), w, h, k.data_ptr())
        v.synchronize()
        owners += (k != np.iinfo(np.int64).max).to(torch.int32)
        keys = k if keys is None else torch.minimum(keys, k)
    same = keys == ref
    assert same.all(), f"{int((~same).sum())} of {w * h} rays differ from the single-volume ray-cast"
    hit = ref != np.iinfo(np.int64).max
    assert hit.float().mean() > 0.15
    assert (owners[hit] == 1).all() and (owners[~hit] == 0).all(), "every hit is labelled by exactly one slab"
    st = rep.ray_stats()
    assert st[1] == int(hit.sum()) and st[0] > st[1]

    # the same with the image split by interleaved 4-row tile rows (what SlabVolume.raycast_replicated and the C++ driver
    # use: every rank gets the same mix of cheap and expensive rows); chunk r of `parts` is what rank r contributes to the
    # all-gather
    prow = rep.part_rows(h, world)
    assert prow % 4 == 0 and prow * world >= h
    parts = torch.full((world * prow * w * 4,), float("nan"), dtype=torch.float32, device="cuda")
    for r in range(world):
        rep.raycast_part_dev(s2w, c, w, h, r, world, parts[r * prow * w * 4:].data_ptr())
    rep.synchronize()
    assert not bool(torch.isnan(parts).any()), "every entry of a part buffer is written (zeros past the end of the image)"
    keys2 = None
    for v, *_ in slabs:
        k = torch.empty(w * h, dtype=torch.int64, device="cuda")
        v.label_hits_parts_dev(parts.data_ptr(), w, h, world, k.data_ptr())
        v.synchronize()
        keys2 = k if keys2 is None else torch.minimum(keys2, k)
    same = keys2 == ref
    assert same.all(), f"interleaved parts: {int((~same).sum())} of {w * h} rays differ from the single-volume ray-cast"
    for v, *_ in slabs:
        v.close()
    rep.close()
    full.close()


def test_rebuilt_skip_map_after_upload_keeps_the_image():
    """sfm_rebuild_skip_map on a volume whose SDF came through sfm_upload (which otherwise disables skipping)."""
    from slam_maskrcnn_b200 import orbit_camera
    sc = Scenario(dims=(64, 64, 64), bins=16, frames=5, yaw_step_deg=2.0)
    v = sc.make_volume()
    for fr in sc.frames:
        v.integrate_raw(fr["depth"], fr["color"], fr["gt"], fr["extrinsic"])
    s2w, c = orbit_camera(sc.Kinv, 0.4, float(sc.mean_depth))
    a = v.raycast(s2w, c, want_t=True, want_label=True)
    v.upload("sdf", v.download("sdf"))  # no block may be skipped now
    b = v.raycast(s2w, c, want_t=True, want_label=True)
    v.rebuild_skip_map()
    d = v.raycast(s2w, c, want_t=True, want_label=True)
    for x, y, z in zip(a, b, d):
        assert (x == y).all() and (x == z).all()
    assert (a[1] > 0).mean() > 0.2
    v.close()
