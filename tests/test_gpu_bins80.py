"""The headline bin count (80-class COCO histogram, BASELINE config 2) against the reference's own kernels
compiled verbatim with MAX_OBJECTS = 80 (oracle/_ref/libsfm_ref_L80.so): integrate (tsdf.cu:18-70) at 128^3 and
256^3 with full 640x480 frames, back-projection (tsdf.cu:72-135), the whole labelled pipeline with the verbatim
filter_overlaps (tsdf.cu:304-416), the ray-cast image (viewer.cu:17-86), and the z-slab forms of merge and
ray-cast.  256^3 x 80 = 1.34e9 histogram entries is the largest volume the reference's 32-bit index reaches
(SURVEY appendix B.1); that case is compared on the device."""
import numpy as np
import pytest

from tests.common import Scenario, bits, device_plane
from tests.test_gpu_integrate import assert_planes_equal, run_ours, run_reference_kernel
from tests.test_gpu_raymarch import FusedPair, reference_backproject, run_reference_pipeline

pytestmark = pytest.mark.gpu
BINS = 80


def need_ref():
    from oracle import binding as ob
    if not ob.ref_available(BINS):
        pytest.skip("oracle/_ref/libsfm_ref_L80.so not built")


def test_integrate_80_bins_128_full_frames():
    need_ref()
    sc = Scenario(dims=(128, 128, 128), bins=BINS, width=640, height=480, n_instances=79, frames=4, yaw_step_deg=2.0)
    # "mask" = the labels with 5 % of the instance pixels flipped to a random id: every one of the 80 bins occurs
    ours, stats = run_ours(sc, mask_key="mask")
    ref = run_reference_kernel(sc, mask_key="mask")
    assert_planes_equal(ours, ref, "128^3 x 80 bins vs reference tsdf_kernel")
    assert ours["hist"].sum() == sum(s for _, s in stats) and ours["weight"].sum() == sum(u for u, _ in stats)
    assert (ours["hist"].reshape(-1, BINS).sum(0) > 0).sum() >= 60, "most of the 80 bins should be in use"


def test_integrate_80_bins_noisy_permuted_labels():
    need_ref()
    sc = Scenario(dims=(96, 80, 72), bins=BINS, width=320, height=240, n_instances=79, frames=4, yaw_step_deg=3.0, permute=True)
    ours, _ = run_ours(sc, mask_key="mask")
    ref = run_reference_kernel(sc, mask_key="mask")
    assert_planes_equal(ours, ref, "noisy permuted labels, 80 bins")
    assert ours["hist"][..., 64:].sum() > 0, "labels beyond 64 (third 32-bin stride) should occur"


def test_integrate_80_bins_256_on_device():
    """256^3 x 80 bins x 640x480: the planes stay on the device (5.4 GB of histogram per side)."""
    import torch
    from oracle import binding as ob
    need_ref()
    sc = Scenario(dims=(256, 256, 256), bins=BINS, width=640, height=480, n_instances=79, frames=3, yaw_step_deg=2.0)
    v = sc.make_volume()
    n = int(np.prod(sc.dims))
    sdf = torch.full((n,), float(sc.miu), dtype=torch.float32, device="cuda")
    wt = torch.zeros(n, dtype=torch.int32, device="cuda")
    col = torch.zeros(n * 3, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(n * BINS, dtype=torch.int32, device="cuda")
    U = S = 0
    for fr in sc.frames:
        v.integrate_raw(fr["depth"], fr["color"], fr["gt"], fr["extrinsic"])
        u, s = v.frame_stats()
        U, S = U + u, S + s
        d = torch.from_numpy(fr["depth"].view(np.int16)).cuda()
        c = torch.from_numpy(fr["color"]).cuda()
        m = torch.from_numpy(fr["gt"]).cuda()
        torch.cuda.synchronize()
        ob.ref_integrate(BINS, sdf.data_ptr(), col.data_ptr(), cnt.data_ptr(), wt.data_ptr(), sc.dims, sc.start,
                         sc.voxel, float(sc.miu), sc.K, d.data_ptr(), c.data_ptr(), m.data_ptr(), fr["extrinsic"], sc.W, sc.H)
    v.synchronize()
    torch.cuda.synchronize()
    assert torch.equal(device_plane(v, "sdf").reshape(-1).view(torch.int32), sdf.view(torch.int32)), "SDF bits differ"
    assert torch.equal(device_plane(v, "weight").reshape(-1), wt)
    assert torch.equal(device_plane(v, "color").reshape(-1), col)
    assert torch.equal(device_plane(v, "hist").reshape(-1), cnt)
    assert int(wt.sum(dtype=torch.int64)) == U and int(cnt.sum(dtype=torch.int64)) == S and S > 0
    v.close()


def test_backproject_80_bins():
    need_ref()
    sc = Scenario(dims=(64, 64, 64), bins=BINS, n_instances=40, frames=6, yaw_step_deg=2.0)
    fp = FusedPair(sc, nframes=5)
    E = sc.frames[5]["extrinsic"]
    probs, box, t, flags = fp.vol.backproject(E)
    rprobs, rbox = reference_backproject(fp, E)
    ok = flags == 0
    assert ok.mean() > 0.9 and (t > 0).mean() > 0.3
    assert (bits(probs)[ok] == bits(rprobs)[ok]).all(), "probs differ on in-bounds rays at 80 bins"
    assert (box[ok] == rbox[ok]).all()
    assert probs[..., 33:].sum() > 0, "bins of the second and third 32-bin stride must carry mass"
    fp.vol.close()


@pytest.mark.parametrize("ninst,nframes", [(6, 7), (10, 6)])
def test_fuse_frame_pipeline_80_bins(ninst, nframes):
    """Labelled fusion with duplicate-instance merge at 80 bins: relabelled masks, num_objs and all planes == the
    reference pipeline (verbatim back_proj_kernel + verbatim filter_overlaps + verbatim tsdf_kernel)."""
    need_ref()
    sc = Scenario(dims=(64, 64, 64), bins=BINS, n_instances=ninst, frames=nframes, yaw_step_deg=2.0, permute=True)
    ref_planes, ref_masks, ref_num = run_reference_pipeline(sc, sc.frames)
    v = sc.make_volume()
    margins = []
    for i, fr in enumerate(sc.frames):
        mask = fr["mask"].copy()
        v.fuse_frame(fr["depth"], fr["color"], mask, fr["extrinsic"])
        if i > 0:
            margins.append(v.last_merge().margin)
        same = mask == ref_masks[i]
        assert same.all(), f"frame {i}: relabelled mask differs at {int((~same).sum())} pixels (margins {margins})"
    assert v.info().num_objs == ref_num
    for k in ("sdf", "weight", "color", "hist"):
        a, b = v.download(k), ref_planes[k]
        same = (bits(a) == bits(b)) if k == "sdf" else (a == b)
        assert same.all(), f"plane {k} differs at {int((~same).sum())} entries"
    assert min(margins) > 1e-4, f"decision margins too small for a meaningful parity claim: {margins}"
    v.close()


@pytest.mark.parametrize("angle", [0.05, 0.6])
def test_raycast_80_bins(angle):
    import torch
    from oracle import binding as ob
    from slam_maskrcnn_b200 import orbit_camera, palette
    need_ref()
    sc = Scenario(dims=(80, 80, 80), bins=BINS, n_instances=40, frames=6, yaw_step_deg=2.0)
    fp = FusedPair(sc)
    s2w, c = orbit_camera(sc.Kinv, angle, float(sc.mean_depth))
    bgr, t, lab = fp.vol.raycast(s2w, c, want_t=True, want_label=True)
    flags = fp.vol.ray_flags()
    out = torch.zeros(sc.H * sc.W * 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ob.ref_show(BINS, s2w, c, sc.start, sc.end, sc.voxel, sc.dims, fp.sdf_ptr, fp.col_ptr, fp.cnt_ptr, sc.W, sc.H,
                out.data_ptr(), palette(BINS))
    ref = out.cpu().numpy().reshape(sc.H, sc.W, 3)
    ok = flags == 0
    assert ok.mean() > 0.9
    assert (lab > 32).sum() > 0, "labels of the upper strides should be visible"
    assert (bgr[ok] == ref[ok]).all(), f"ray-cast image differs on {int((bgr[ok] != ref[ok]).any(-1).sum())} in-bounds rays"
    fp.vol.close()


def test_sharded_merge_80_bins():
    from tests.test_gpu_sharded_merge import test_sharded_merge_equals_single_volume as run
    run(3, (64, 64, 96), BINS, 6, None)


def test_sharded_raycast_80_bins():
    from tests import test_gpu_sharded_raycast as m
    import torch
    from slam_maskrcnn_b200 import orbit_camera
    from slam_maskrcnn_b200.slabs import keys_to_int64
    sc = Scenario(dims=(64, 64, 96), bins=BINS, n_instances=40, frames=6, yaw_step_deg=2.0)
    full, slabs = m.build(sc, 3)
    s2w, c = orbit_camera(sc.Kinv, 0.4, float(sc.mean_depth))
    w, h = sc.W, sc.H
    ref = torch.empty(w * h, dtype=torch.int64, device="cuda")
    full.raycast_keys_dev(s2w, c, w, h, ref.data_ptr())
    full.synchronize()
    ref = keys_to_int64(ref)
    ev = [None, None]
    keys = None
    for stage in (1, 2, 3):
        outs = []
        for v, *_ in slabs:
            o = torch.empty(w * h, dtype=torch.int64, device="cuda")
            v.shard_raycast_stage(stage, s2w, c, w, h, ev[0].data_ptr() if ev[0] is not None else 0,
                                  ev[1].data_ptr() if ev[1] is not None else 0, o.data_ptr())
            v.synchronize()
            outs.append(o)
        red = outs[0]
        for o in outs[1:]:
            red = torch.minimum(red, o)
        if stage < 3:
            ev[stage - 1] = red
        else:
            keys = red
    assert (keys == ref).all()
    hit = ref != np.iinfo(np.int64).max
    assert ((ref[hit] & 0xff) > 32).sum() > 0, "labels of the upper strides should be hit"
    for v, *_ in slabs:
        v.close()
    full.close()
